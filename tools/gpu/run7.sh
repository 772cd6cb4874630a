# GPU tests after the f2 / f4 / VAE changes, then a bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "DeprecationWarning\|warnings.warn" | tail -40 > gpurun_out/r2_pytest3.log
cat gpurun_out/r2_pytest3.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
tail -c 3000 gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
