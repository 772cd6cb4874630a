N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r2_shard_split_${N}gpu.log
: > $L
for mode in "1 1" "0 0"; do
  set -- $mode
  K5_DIST_OVERLAP=$1 K5_DIST_SPLIT=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$1$2 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_split_$1$2.json 2>> gpurun_out/r2_bench_${N}gpu_split.err
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu_split_$1$2.json").read().strip().splitlines()[-1])
print("overlap=$1 split=$2 N=$N ms/step", d["ms_per_step"], "tokens/s", d["value"], "attn ms", d["roofline"]["avg_launch_ms"], "share", d["roofline"]["share_of_step"], "e2e ms", d["e2e"]["ms_per_step"])
PY
done
cat $L
