# bit-identity check of both all-gather forms only; usage: run_shard_check.sh <ngpu>
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r2_shard_check_${N}gpu.log
: > $L
for ov in 0 1; do
  K5_SHARD_VERBOSE=1 K5_DIST_OVERLAP=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov tests/gpu_shard_ranks.py 2>&1 | grep -E "shard x|rank [0-9]:|Error|error|Traceback" >> $L
done
cat $L
