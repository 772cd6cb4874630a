# Final validation of the round on one GPU: the full GPU test suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_pytest_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> gpurun_out/r2_pytest_final.log
timeout 400 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
echo "bench exit $?" >> gpurun_out/r2_pytest_final.log
cat gpurun_out/r2_pytest_final.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "attn ms", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "vae", d.get("vae_decode"), "clocks", d.get("clocks"))
PY
