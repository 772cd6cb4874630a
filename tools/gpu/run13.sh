mkdir -p gpurun_out
L=gpurun_out/r2_pair.log
: > $L
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_attn_variants.py tests/test_gpu_kernel_modes.py -m gpu -x -q 2>&1 | tail -4 | tee -a $L
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_forward.py tests/test_gpu_shard.py tests/test_gpu_reference.py -m gpu -x -q 2>&1 | tail -4 | tee -a $L
for round in 1 2; do
for pair in 1 0; do
  K5_ATTN_PAIR=$pair timeout 200 python tests/gpu_attn_variants.py "bounded_pair$pair=K5_VARIANT_BOUND:1" 2>&1 | grep -E "attn S|parity|exit|TIMEOUT" | tee -a $L
done
done
for round in 1 2; do
for pair in 1 0; do
  K5_ATTN_PAIR=$pair timeout 300 python bench.py --steps 6 --warmup 3 --no-vae --no-configs --no-cpu-baseline > gpurun_out/tmp_bench.json 2>/dev/null
  python - "$pair" <<PY | tee -a $L
import json,sys
d=json.loads(open("gpurun_out/tmp_bench.json").read().strip().splitlines()[-1])
print("pair=%s ms/step %.1f  attention %.3f ms in-loop  frac %.4f  sm_mhz %s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
done
