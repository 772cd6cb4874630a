N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r2_shard_${N}gpu_c.log
: > $L
timeout 120 python tests/gpu_p2p_bandwidth.py 2>&1 | grep p2p >> $L
for ns in 3 2 4; do
  K5_DIST_COPY_STREAMS=$ns K5_DIST_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$ns bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_c_ns$ns.json 2>> gpurun_out/r2_bench_${N}gpu_c.err
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu_c_ns$ns.json").read().strip().splitlines()[-1])
print("overlap=1 copy streams=$ns N=$N ms/step", d["ms_per_step"], "tokens/s", d["value"], "attn ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
cat $L
