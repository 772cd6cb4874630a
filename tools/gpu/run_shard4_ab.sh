# 4 GPUs: all-to-all copy-engine rates, then the bench with both forms of the K|V all-gather (same box, back to back).
N=4
mkdir -p gpurun_out
L=gpurun_out/r2_shard_split_${N}gpu.log
: > $L
timeout 120 python tests/gpu_p2p_bandwidth.py 2>&1 | grep p2p >> $L
for ov in 1 0; do
  K5_DIST_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$ov bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ov$ov.json 2>> gpurun_out/r2_bench_${N}gpu_ab.err
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu_ov$ov.json").read().strip().splitlines()[-1])
print("overlap+split=$ov N=$N ms/step", d["ms_per_step"], "tokens/s", d["value"], "attn ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"], "e2e ms", d["e2e"]["ms_per_step"], "|", d["config"]["parallelism"][:90])
PY
done
cat $L
