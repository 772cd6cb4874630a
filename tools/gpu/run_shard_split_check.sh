# Cross-process bit-identity of the overlapped all-gather WITH the split attention launch (the 8-rank default), forced on N ranks.
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r2_shard_split_check_${N}gpu.log
: > $L
K5_SHARD_VERBOSE=1 K5_DIST_OVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/gpu_shard_ranks.py > gpurun_out/r2_shard_split_check_${N}gpu.raw 2>&1
echo "exit $?" >> $L
grep -E "shard x|rank [0-9]:|Error|error|Traceback" gpurun_out/r2_shard_split_check_${N}gpu.raw >> $L
cat $L
