# 2-GPU check of both all-gather forms + bench; usage: run_shard2.sh <ngpu>
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r2_shard_${N}gpu.log
: > $L
for ov in 0 1; do
  K5_SHARD_VERBOSE=1 K5_DIST_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov tests/gpu_shard_ranks.py 2>&1 | grep -E "shard x|rank [0-9]:|Error|error|Traceback" >> $L
done
for ov in 0 1; do
  K5_DIST_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$ov bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ov$ov.json 2>> $L
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu_ov$ov.json").read().strip().splitlines()[-1])
print("overlap=$ov N=$N ms/step", d["ms_per_step"], "attn ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"])
PY
done
cat $L
