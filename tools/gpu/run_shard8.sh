# N-GPU verification of the temporal shard: bit-identity of both all-gather forms, bench with both, NVLink counters.
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r2_shard_${N}gpu.log
: > $L
for ov in 0 1; do
  K5_SHARD_VERBOSE=1 K5_DIST_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov tests/gpu_shard_ranks.py 2>&1 | grep -E "shard x|rank [0-9]:|Error|error|Traceback" >> $L
done
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r2_nvlink_before_${N}gpu.txt 2>&1
for ov in 1 0; do
  K5_DIST_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$ov bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ov$ov.json 2>> gpurun_out/r2_bench_${N}gpu.err
  [ $ov = 1 ] && nvidia-smi nvlink -gt d -i 0 > gpurun_out/r2_nvlink_after_${N}gpu.txt 2>&1
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu_ov$ov.json").read().strip().splitlines()[-1])
print("overlap=$ov N=$N ms/step", d["ms_per_step"], "tokens/s", d["value"], "attn ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
cat $L
python - <<PY
import re
def tot(p):
    t=0
    for l in open(p):
        m=re.search(r"Data (Tx|Rx)\d*: (\d+) KiB", l)
        if m: t+=int(m.group(2))
    return t
try:
    a,b=tot("gpurun_out/r2_nvlink_before_${N}gpu.txt"),tot("gpurun_out/r2_nvlink_after_${N}gpu.txt")
    print("GPU 0 NVLink data counters (Tx + Rx, all links) across the overlap bench run: %.1f GB" % ((b-a)/1048576.0))
except Exception as e: print("nvlink counters:", e)
PY
