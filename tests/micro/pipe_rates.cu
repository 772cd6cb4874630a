// Micro-benchmark: issue rate per SM sub-partition of the instructions the attention softmax is made of
// (MUFU.EX2, FFMA, FFMA2, FADD2, FMNMX3, F2FP, IMAD shift-add) and of the MUFU + FMA mixes, with 1 / 2 / 4 warps per
// scheduler.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

template <int KIND>
__global__ void bench(float* out, long long* cycles) {
    float a[8];
    uint64_t q[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = threadIdx.x * 0.001f + i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(q[i]) : "f"(a[i]), "f"(a[i] + 0.5f));
        u[i] = threadIdx.x + i;
    }
    const uint64_t c2 = q[0];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (KIND == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
            if (KIND == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(c2));
            if (KIND == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(c2));
            if (KIND == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
            if (KIND == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
            if (KIND == 6) asm volatile("{.reg .u32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1;}" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            if (KIND == 7) {   // 1 MUFU : 1 FFMA2
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(c2));
            }
            if (KIND == 8) {   // 1 MUFU : 3 FFMA2
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(c2));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[(i + 3) & 7]) : "l"(c2));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[(i + 5) & 7]) : "l"(c2));
            }
            if (KIND == 9) {   // 1 MUFU : 3 scalar FFMA
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 2) & 7]) : "f"(a[(i + 1) & 7]));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 4) & 7]) : "f"(a[(i + 3) & 7]));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 6) & 7]) : "f"(a[(i + 5) & 7]));
            }
            if (KIND == 10) {  // the softmax pair as written: FFMA2, 2 MUFU, FADD2, F2FP
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(c2));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[(i + 3) & 7]) : "l"(c2));
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q[i]));
        s += a[i] + lo + hi + __uint_as_float(u[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char* name, int instr_per_inner) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    for (int warps_per_smsp : {1, 2, 4}) {
        const int threads = 128 * warps_per_smsp;
        bench<KIND><<<148, threads>>>(out, cyc);
        bench<KIND><<<148, threads>>>(out, cyc);
        cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double per_warp_instr = static_cast<double>(h) / (ITERS * 8.0 * instr_per_inner);
        printf("%-28s warps/SMSP=%d  cycles per warp-instruction (one warp) %.2f  -> per SMSP %.2f\n", name, warps_per_smsp,
               per_warp_instr, per_warp_instr / warps_per_smsp);
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("MUFU.EX2", 1);
    run<1>("FFMA", 1);
    run<2>("FFMA2", 1);
    run<3>("FADD2", 1);
    run<4>("FMNMX3", 1);
    run<5>("F2FP.BF16", 1);
    run<6>("SHL+IADD (LEA/IMAD)", 1);
    run<7>("MUFU+FFMA2 (per pair)", 1);
    run<8>("MUFU+2FFMA2+FADD2 (per group)", 1);
    run<9>("MUFU+3FFMA (per group)", 1);
    run<10>("softmax pair (5 instr)", 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
