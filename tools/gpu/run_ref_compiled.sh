# The reference with its own @torch.compile decorators live (max-autotune-no-cudagraphs, dynamic) at the 5 s size, 2 and 4 visual blocks.
mkdir -p gpurun_out
rm -f gpurun_out/r2_reference_compiled.jsonl
K5_REF_COMPILED=${K5_REF_COMPILED:-inner} K5_REF_BLOCKS=${K5_REF_BLOCKS:-2,32} timeout 400 python tests/gpu_reference_timing.py gpurun_out/r2_reference_compiled.jsonl > gpurun_out/r2_reference_compiled.log 2>&1
echo "exit $?" >> gpurun_out/r2_reference_compiled.log
tail -5 gpurun_out/r2_reference_compiled.log
