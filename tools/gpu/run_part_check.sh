# After making the key-slab split a template parameter: attention parity / split bit-identity tests, in-process shard tests, then the
# in-loop attention time of the single-GPU step (no VAE / configs / CPU legs).
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_shard.py -m gpu -x -q -k "attention or shard" 2>&1 | tail -3 > gpurun_out/r2_part_check.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-vae --no-configs --no-cpu-baseline > gpurun_out/r2_bench_part.json 2> gpurun_out/r2_bench_part.err
echo "bench exit $?" >> gpurun_out/r2_part_check.log
python - <<'PY' >> gpurun_out/r2_part_check.log
import json
d=json.loads(open("gpurun_out/r2_bench_part.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "attn ms", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "clocks", d.get("clocks"))
PY
cat gpurun_out/r2_part_check.log
