// Micro-benchmark (round 2): does tensor-core activity slow the MUFU stream of the softmax warps?
// 8 "softmax" warps (2 per scheduler, as in attention.cu) loop on the bounded softmax pair (FMUL2, 2 x MUFU.EX2, FADD2,
// F2FP) while one thread of a ninth warp issues tcgen05.mma 128x128x16 (SS form, bf16, operands in shared memory with
// random bits) at a chosen duty cycle: `gap` = clock cycles the issuer idles after each group of 4 MMAs (a QK^T tile's
// worth: 256 tensor cycles).  gap < 0: no MMA at all.  Prints cycles per softmax pair per warp and the tensor duty.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../kandinsky-5_b200/csrc -o mufu_mma mufu_mma.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace k5;

#define ITERS 4096

__global__ void __launch_bounds__(320, 1) bench(int gap, long long* cyc_mu, long long* n_mma, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    // random operand bits (power draw of the tensor pipe depends on the data)
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < 2 * 16384 / 4; i += blockDim.x) {
        uint32_t x = i * 2654435761u + blockIdx.x * 40503u;
        x ^= x >> 15;
        x *= 2246822519u;
        x ^= x >> 13;
        w[i] = (x & 0xBFFFBFFFu) | 0x30003000u;   // bf16 pairs of moderate magnitude
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        stop = 0;
    }
    if (warp == 8) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    float acc = 0.f;
    if (warp < 8) {
        float a[8];
        uint64_t q[8];
        uint32_t u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = -(threadIdx.x * 0.001f + i);
            q[i] = pack_f32x2(a[i], a[i] - 0.5f);
            u[i] = 0;
        }
        const uint64_t c2 = pack_f32x2(0.999f, 0.999f);
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(c2));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[(i + 3) & 7]) : "l"(c2));
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
            }
        }
        const long long t1 = clock64();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            unpack_f32x2(q[i], lo, hi);
            acc += a[i] + lo + hi + __uint_as_float(u[i]);
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) *cyc_mu = t1 - t0;
        if (threadIdx.x == 0) stop = 1;
    } else if (warp == 9 && gap >= 0) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
            const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 16384;
            long long n = 0;
            uint32_t ph = 0;
            while (!stop) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ss(tmem + (n & 1) * 128, umma_desc_sw128(a0 + k * 32, 0, 1024), umma_desc_sw128(b0 + k * 32, 0, 1024),
                            idesc, k != 0 ? 1u : 0u);
                umma_commit(&bar);
                mbar_wait(&bar, ph);
                ph ^= 1;
                ++n;
                const long long t = clock64();
                while (clock64() - t < gap) {
                }
            }
            if (blockIdx.x == 0) *n_mma = n;
        }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

int main() {
    long long *cm, *nm;
    float* sink;
    cudaMalloc(&cm, 8);
    cudaMalloc(&nm, 8);
    cudaMalloc(&sink, 148 * 320 * sizeof(float));
    const int smem_bytes = 2 * 16384 + 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    for (int gap : {-1, 2000, 768, 256, 0}) {
        cudaMemset(cm, 0, 8);
        cudaMemset(nm, 0, 8);
        for (int rep = 0; rep < 3; ++rep) bench<<<148, 320, smem_bytes>>>(gap, cm, nm, sink);
        cudaDeviceSynchronize();
        long long hc = 0, hn = 0;
        cudaMemcpy(&hc, cm, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hn, nm, 8, cudaMemcpyDeviceToHost);
        printf("gap %5d: %.2f cycles per softmax pair per warp (2 warps / scheduler; 16.0 = MUFU bound), tensor pipe duty %.0f %%\n",
               gap, double(hc) / (ITERS * 8.0), 100.0 * 256.0 * hn / double(hc));
    }
    cudaError_t e = cudaGetLastError();
    printf("CUDA: %s\n", cudaGetErrorString(e));
    return 0;
}
