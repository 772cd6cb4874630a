mkdir -p gpurun_out
L=gpurun_out/r2_poly_frac.log
: > $L
for round in 1 2; do
for v in "" _p16 _p12 _p8; do
  K5_VARIANT_NOCHECK=$([ $round = 2 ] && echo 1 || echo 0) K5_LIB_PATH=$PWD/kandinsky-5_b200/libk5$v.so timeout 200 python tests/gpu_attn_variants.py "bounded$v=K5_VARIANT_BOUND:1" 2>&1 | grep -E "attn S|parity" >> $L
done
done
cat $L
K5_VAE_T=5 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_vae_tile_launches.csv python tests/gpu_bench_vae.py > gpurun_out/r2_vae_tile_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2_vae_tile_launches.csv 30 | tee gpurun_out/r2_vae_tile_summary.txt
python tests/gpu_bench_vae.py 2>&1 | tail -1
