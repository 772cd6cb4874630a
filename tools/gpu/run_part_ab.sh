# Same-box A/B inside the sampler step: the dense attention loop as ptxas schedules it with the key-slab split as RUN-TIME
# branches (libk5_rt.so, built from commit 50ee0ac) against the split as a template parameter (libk5.so, the tuned loop).
mkdir -p gpurun_out
L=gpurun_out/r2_attention_part_ab.log
: > $L
for rep in 1 2; do
for v in tmpl rt; do
  lib=""
  [ $v = rt ] && lib=$PWD/kandinsky-5_b200/libk5_rt.so
  K5_LIB_PATH=$lib timeout 120 python bench.py --steps 5 --warmup 3 --no-vae --no-configs --no-cpu-baseline > gpurun_out/tmp_ab.json 2>/dev/null
  python - <<PY >> $L
import json
d=json.loads(open("gpurun_out/tmp_ab.json").read().strip().splitlines()[-1])
print("$v  ms/step %.1f  attention %.3f ms in-loop  frac %.4f  sm_mhz %s" % (d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
done
cat $L
