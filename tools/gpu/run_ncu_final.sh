# ncu evidence for the final build: (1) launch list of two sampler steps (our kernels only), (2) one --set full capture
# of the dense attention kernel at S = 47 616 (kept as profiles/r2_attention_final.ncu-rep)
mkdir -p gpurun_out
K='regex:attention_fwd|gemm_bf16|ln_rows|gemv_f32|patchify|rope|euler|cfg_combine|time_features|pooled_embed|nabla|build_items|bf16_addsub|dist_barrier'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 800 --csv --log-file gpurun_out/r2_launches_5s_nocfg.csv python bench.py --steps 1 --warmup 1 --no-vae --no-configs --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_5s_nocfg.csv 20 | tee gpurun_out/r2_launches_summary.txt
export K5_VARIANT_NAME=ncu K5_VARIANT_BOUND=1 K5_VARIANT_NOCHECK=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 9 -c 1 -f -o gpurun_out/r2_attention_final python tests/gpu_attn_variants.py > gpurun_out/r2_attention_final.log 2>&1
tail -3 gpurun_out/r2_attention_final.log
ls -la gpurun_out/r2_attention_final.ncu-rep
unset K5_VARIANT_NAME K5_VARIANT_BOUND K5_VARIANT_NOCHECK
for sg in 0 1100 0 1100; do
  K5_ATTN_STAGGER=$sg timeout 300 python bench.py --steps 6 --warmup 3 --no-vae --no-configs --no-cpu-baseline > gpurun_out/tmp_bench.json 2>/dev/null
  python - "$sg" <<PY | tee -a gpurun_out/r2_stagger_inloop.log
import json,sys
d=json.loads(open("gpurun_out/tmp_bench.json").read().strip().splitlines()[-1])
print("stagger=%s ms/step %.1f  attention %.3f ms in-loop  frac %.4f  sm_mhz %s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
