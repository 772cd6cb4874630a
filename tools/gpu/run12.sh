mkdir -p gpurun_out
L=gpurun_out/r2_poly_inloop.log
: > $L
for round in 1 2; do
for v in "" _m0 _m3 _m5 _m6; do
  K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so timeout 300 python bench.py --steps 6 --warmup 3 --no-vae --no-configs --no-cpu-baseline > gpurun_out/tmp_bench.json 2>/dev/null
  python - "$v" <<PY >> $L
import json,sys
d=json.loads(open("gpurun_out/tmp_bench.json").read().strip().splitlines()[-1])
print("lib%-4s ms/step %.1f  attention %.3f ms in-loop  frac %.4f  sm_mhz %s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
done
cat $L
echo "== nabla selection (vectorised pooling)" | tee gpurun_out/r2_nabla_sel3.log
for P in 0.9 0.0; do K5_NABLA_P=$P timeout 200 python tests/gpu_bench_nabla.py 2>&1 | tail -1 | tee -a gpurun_out/r2_nabla_sel3.log; done
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_shard.py -m gpu -x -q -k "nabla or sta" 2>&1 | tail -3
