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
            if (KIND == 11) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));        // 2 x MUFU.EX2.F16 + PRMT in SASS
            if (KIND == 12) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));   // 2 x MUFU.EX2.BF16 + PRMT
            if (KIND == 13) {  // the bounded-score softmax pair (round 2): FMUL2, 2 MUFU, FADD2, F2FP
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(c2));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[(i + 3) & 7]) : "l"(c2));
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
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

// ---------------------------------------------------------------------------------------------------------------
// TMEM read rate and its interaction with the MUFU: `nld` warps per scheduler loop on tcgen05.ld.32x32b.x32 (4 KB per
// warp instruction) while `nmu` warps per scheduler loop on MUFU.EX2.  Answers: (1) bytes / clk / SM a softmax warpgroup
// can pull out of TMEM, (2) whether LDTM traffic slows the MUFU stream that shares the MIO path with it.
__global__ void tmem_bench(int nld_warps, int nmu_warps, long long* cyc_ld, long long* cyc_mu, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            static_cast<uint32_t>(__cvta_generic_to_shared(&slot))) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    const long long t0 = clock64();
    if (warp < nld_warps) {
        uint32_t r[32];
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(base + (it & 3) * 32)
                : "memory");
            if ((it & 3) == 3) {       // four loads (one 128-column score tile) in flight, then the wait the softmax does
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += __uint_as_float(r[0] & 0x3fffffu);
            }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (threadIdx.x == 0 && blockIdx.x == 0) *cyc_ld = clock64() - t0;
    } else if (warp < nld_warps + nmu_warps) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += a[i];
        if ((threadIdx.x & 31) == 0 && warp == nld_warps && blockIdx.x == 0) *cyc_mu = clock64() - t0;
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

void run_tmem() {
    long long *cl, *cm;
    float* sink;
    cudaMalloc(&cl, 8);
    cudaMalloc(&cm, 8);
    cudaMalloc(&sink, 148 * 512 * sizeof(float));
    const int cfg[][2] = {{4, 0}, {8, 0}, {0, 4}, {0, 8}, {4, 4}, {8, 4}, {4, 8}, {8, 8}};
    for (auto& c : cfg) {
        const int nld = c[0], nmu = c[1];
        cudaMemset(cl, 0, 8);
        cudaMemset(cm, 0, 8);
        for (int rep = 0; rep < 2; ++rep) tmem_bench<<<148, 32 * (nld + nmu), 0>>>(nld, nmu, cl, cm, sink);
        cudaDeviceSynchronize();
        long long hl = 0, hm = 0;
        cudaMemcpy(&hl, cl, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hm, cm, 8, cudaMemcpyDeviceToHost);
        printf("TMEM: %d LDTM warps + %d MUFU warps per SM:", nld, nmu);
        if (nld) printf("  LDTM.x32 %.1f cycles per warp instruction -> %.0f B/clk/SM", double(hl) / ITERS,
                        4096.0 * nld * ITERS / double(hl));
        if (nmu) printf("  MUFU %.2f cycles per warp instruction per SMSP", double(hm) / (ITERS * 8.0) / (nmu / 4.0));
        printf("\n");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error (tmem): %s\n", cudaGetErrorString(e));
    cudaFree(cl);
    cudaFree(cm);
    cudaFree(sink);
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
    run<11>("ex2.f16x2 (2 MUFU.F16+PRMT)", 1);
    run<12>("ex2.bf16x2 (2 MUFU.BF16+PRMT)", 1);
    run<13>("bounded softmax pair", 1);
    run_tmem();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
