mkdir -p gpurun_out
echo "== nabla selection (bitwise radix)" | tee gpurun_out/r2_nabla_sel2.log
for P in 0.9 0.0; do K5_NABLA_P=$P timeout 200 python tests/gpu_bench_nabla.py 2>&1 | tail -1 | tee -a gpurun_out/r2_nabla_sel2.log; done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_shard.py tests/test_gpu_attn_variants.py tests/test_gpu_kernel_modes.py -m gpu -x -q 2>&1 | grep -v "DeprecationWarning\|warnings.warn" | tail -8 | tee gpurun_out/r2_pytest6.log
timeout 1200 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
tail -c 6000 gpurun_out/r2_bench4.json; tail -5 gpurun_out/r2_bench4.err
