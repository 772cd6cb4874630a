# 8 GPUs, engine defaults (overlapped all-gather, attention split in three launches): cross-process bit-identity, then the driver's bench command.
N=8
mkdir -p gpurun_out
L=gpurun_out/r2_shard_final_${N}gpu.log
: > $L
K5_SHARD_VERBOSE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/gpu_shard_ranks.py > gpurun_out/r2_shard_final_${N}gpu.raw 2>&1
echo "check exit $?" >> $L
grep -E "shard x|rank [0-9]:|Error|error|Traceback" gpurun_out/r2_shard_final_${N}gpu.raw >> $L
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_final.json 2>> gpurun_out/r2_bench_${N}gpu_final.err
echo "bench exit $?" >> $L
python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu_final.json").read().strip().splitlines()[-1])
print("defaults N=$N ms/step", d["ms_per_step"], "tokens/s", d["value"], "attn ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "|", d["config"]["parallelism"][:100])
PY
cat $L
