// HBM-bound passes of the HunyuanVideo VAE decoder (kandinsky/models/vae.py:230-275, 166-205, 343-359, 928-936,
// 1144-1204): GroupNorm statistics, the fused GroupNorm + SiLU + nearest up-sample + replicate-pad gather that
// feeds the implicit-GEMM convolutions, the frame-causal softmax of the mid-block attention and the temporal tile
// blend.  All of them move 8 bf16 channels (16 bytes) per thread access.
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "vae_ops.h"

namespace k5 {

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[2 * j] = bf16_lo(w[j]);
        v[2 * j + 1] = bf16_hi(w[j]);
    }
}
__device__ __forceinline__ uint4 pack8(const float* y) {
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    return o;
}

// Thread = one 8-channel group (fixed for its lifetime) striding over positions; block partials are combined in
// shared memory and written to the block's own row of `sums` ([gridDim.x][2 C] doubles): no atomics, so the
// statistics - and with them the whole decode - are bit-reproducible from run to run.
__global__ void __launch_bounds__(256) gn_sums_kernel(const bf16* __restrict__ x, size_t P, int C, double* __restrict__ sums) {
    extern __shared__ float sh[];                   // [256][16]
    const int cg = C / 8;                            // channel groups per position
    const int slot = threadIdx.x % cg;               // channel group of this thread
    const int lane_pos = threadIdx.x / cg;           // position lane inside the block
    const int pos_per_block = blockDim.x / cg;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    for (size_t p = static_cast<size_t>(blockIdx.x) * pos_per_block + lane_pos; p < P;
         p += static_cast<size_t>(gridDim.x) * pos_per_block) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + p * C) + slot);
        float v[8];
        unpack8(u, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j] += v[j];
            q[j] = fmaf(v[j], v[j], q[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sh[threadIdx.x * 16 + j] = s[j];
        sh[threadIdx.x * 16 + 8 + j] = q[j];
    }
    __syncthreads();
    // thread c < C reduces channel c over the pos_per_block lanes
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int sl = c / 8, j = c % 8;
        double a = 0.0, b = 0.0;
        for (int l = 0; l < pos_per_block; ++l) {
            a += sh[(l * cg + sl) * 16 + j];
            b += sh[(l * cg + sl) * 16 + 8 + j];
        }
        sums[static_cast<size_t>(blockIdx.x) * 2 * C + c] = a;
        sums[static_cast<size_t>(blockIdx.x) * 2 * C + C + c] = b;
    }
}

// One block per group: the (block, channel) partials of the group are summed in a fixed order (strided per thread,
// then a shared-memory tree), in double.
__global__ void __launch_bounds__(256) gn_finalize_kernel(const double* __restrict__ sums, int nblocks, double count, int C,
                                                          int groups, float eps, float* __restrict__ mean_rstd) {
    __shared__ double ra[256], rb[256];
    const int g = blockIdx.x;
    const int cpg = C / groups;
    double a = 0.0, b = 0.0;
    for (int e = threadIdx.x; e < nblocks * cpg; e += 256) {
        const size_t row = static_cast<size_t>(e / cpg) * 2 * C;
        const int c = g * cpg + e % cpg;
        a += sums[row + c];
        b += sums[row + C + c];
    }
    ra[threadIdx.x] = a;
    rb[threadIdx.x] = b;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            ra[threadIdx.x] += ra[threadIdx.x + w];
            rb[threadIdx.x] += rb[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    a = ra[0];
    b = rb[0];
    const double n = count * cpg;
    const double mean = a / n;
    double var = b / n - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[2 * g] = static_cast<float>(mean);
    mean_rstd[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

__device__ __forceinline__ void gn_silu8(float* v, int c0, int cpg, const float* __restrict__ mean_rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, bool silu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c0 + j) / cpg;
        const float mean = __ldg(mean_rstd + 2 * g), rstd = __ldg(mean_rstd + 2 * g + 1);
        float y = (v[j] - mean) * rstd * __ldg(gamma + c0 + j) + __ldg(beta + c0 + j);
        if (silu) y = y / (1.0f + __expf(-y));
        v[j] = y;
    }
}

// One block per padded row (tp, hp): the source row pointer is resolved once, threads sweep (wp, 8-channel slot).
// blockDim (256) is a multiple of the slots per position, so a thread keeps its slot and folds GroupNorm into one
// FMA per element: y = x * a + b with a = rstd * gamma, b = beta - mean * a.
__global__ void __launch_bounds__(256)
pad_gather_kernel(const bf16* __restrict__ x, int Ts, int Hs, int Ws, int C, int ft, int fh, int fw, int T, int H, int W,
                  const float* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  int cpg, int silu, bf16* __restrict__ out) {
    const int cg = C / 8;
    const int hp = blockIdx.x % (H + 2), tp = blockIdx.x / (H + 2);
    int t = tp - 2, h = hp - 1;
    t = t < 0 ? 0 : t;                                        // replicate padding: clamp to the (up-sampled) volume
    h = h < 0 ? 0 : (h >= H ? H - 1 : h);
    const int ts = (ft == 1 || t == 0) ? t : 1 + (t - 1) / ft;  // nearest up-sampling; frame 0 is never repeated
    const int hs = h / fh;
    const uint4* src = reinterpret_cast<const uint4*>(x + (static_cast<size_t>(ts) * Hs + hs) * Ws * C);
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(blockIdx.x) * (W + 2) * C);
    const int slot = threadIdx.x % cg;
    float a[8], b[8];
    if (mean_rstd) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = slot * 8 + j, g = c / cpg;
            a[j] = __ldg(mean_rstd + 2 * g + 1) * __ldg(gamma + c);
            b[j] = __ldg(beta + c) - __ldg(mean_rstd + 2 * g) * a[j];
        }
    }
    const int n = (W + 2) * cg;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int w = i / cg - 1;
        w = w < 0 ? 0 : (w >= W ? W - 1 : w);
        const uint4 u = __ldg(src + static_cast<size_t>(w / fw) * cg + slot);
        uint4 o = u;
        if (mean_rstd) {
            float v[8];
            unpack8(u, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float y = fmaf(v[j], a[j], b[j]);
                if (silu) y = __fdividef(y, 1.0f + __expf(-y));
                v[j] = y;
            }
            o = pack8(v);
        }
        dst[i] = o;
    }
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const bf16* __restrict__ x, size_t P, int C, const float* __restrict__ mean_rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, int cpg, int silu, bf16* __restrict__ out) {
    const int cg = C / 8;
    const size_t total = P * cg;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int slot = static_cast<int>(i % cg);
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
        gn_silu8(v, slot * 8, cpg, mean_rstd, gamma, beta, silu != 0);
        reinterpret_cast<uint4*>(out)[i] = pack8(v);
    }
}

// thread = one padded position: 16x16 mat-vec of the 1x1x1 post_quant_conv in bf16 operands / fp32 accumulate
__global__ void __launch_bounds__(128)
post_quant_pad_kernel(const float* __restrict__ z, int Cz, int Tz, int H, int W, int t0, int T, const float* __restrict__ w,
                      const float* __restrict__ b, bf16* __restrict__ out) {
    __shared__ float sw[16 * 16 + 16];
    for (int i = threadIdx.x; i < Cz * Cz + Cz; i += blockDim.x) sw[i] = i < Cz * Cz ? w[i] : b[i - Cz * Cz];
    __syncthreads();
    const size_t total = static_cast<size_t>(T + 2) * (H + 2) * (W + 2);
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const int wp = static_cast<int>(i % (W + 2));
    const int hp = static_cast<int>((i / (W + 2)) % (H + 2));
    const int tp = static_cast<int>(i / (static_cast<size_t>(W + 2) * (H + 2)));
    int t = tp - 2, h = hp - 1, x = wp - 1;
    t = t < 0 ? 0 : t;
    h = h < 0 ? 0 : (h >= H ? H - 1 : h);
    x = x < 0 ? 0 : (x >= W ? W - 1 : x);
    float in[16];
    const size_t plane = static_cast<size_t>(H) * W;
#pragma unroll
    for (int c = 0; c < 16; ++c)
        in[c] = c < Cz ? bf16_round(z[(static_cast<size_t>(c) * Tz + t0 + t) * plane + static_cast<size_t>(h) * W + x]) : 0.f;
    float y[64];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float acc = 0.f;
        if (o < Cz) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < Cz) acc = fmaf(sw[o * Cz + c], in[c], acc);
            acc += sw[Cz * Cz + o];
        }
        y[o] = acc;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        float q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = (8 * v + j) < 16 ? y[8 * v + j] : 0.f;
        dst[v] = pack8(q);
    }
}

// one block per row; the fp32 score row (up to 5 * 6144 columns) is streamed three times from L2 / HBM and the
// probabilities leave as bf16 (the only rounding: the reference's SDPA keeps scores and softmax in fp32 too)
__global__ void __launch_bounds__(256)
softmax_frame_causal_kernel(const float* __restrict__ s, int lds, bf16* __restrict__ p, int ldp, int hw, float scale_log2,
                            int row0) {
    __shared__ float red[8];
    const int r = row0 + blockIdx.x;
    const int n = (r / hw + 1) * hw;                 // visible columns (multiple of hw; hw is a multiple of 8)
    const float4* row = reinterpret_cast<const float4*>(s + static_cast<size_t>(r) * lds);
    uint4* out = reinterpret_cast<uint4*>(p + static_cast<size_t>(r) * ldp);
    const int nv = n / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto load8 = [&](int i, float* v) {
        const float4 a = row[2 * i], b = row[2 * i + 1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    };
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        float v[8];
        load8(i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) mx = fmaxf(mx, v[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        float v[8];
        load8(i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += exp2f((v[j] - mx) * scale_log2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < 8; ++w) sum += red[w];
    const float inv = 1.0f / sum;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        float v[8];
        load8(i, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = exp2f((v[j] - mx) * scale_log2) * inv;
        out[i] = pack8(v);
    }
}

// thread = one output pixel (all 3 channels): channels-last tile -> planar NCTHW video, with the temporal blend
__global__ void __launch_bounds__(256)
emit_frames_kernel(const bf16* __restrict__ cur, int src0, const bf16* __restrict__ prev, int prev0, int blend, int count,
                   int H, int W, int Fout, int dst0, bf16* __restrict__ out) {
    const size_t plane = static_cast<size_t>(H) * W;
    const size_t total = plane * count;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const int f = static_cast<int>(i / plane);
    const size_t px = i % plane;
    const bf16* c = cur + (static_cast<size_t>(src0 + f) * plane + px) * 3;
    float v[3] = {__bfloat162float(c[0]), __bfloat162float(c[1]), __bfloat162float(c[2])};
    if (f < blend) {
        // Python-float weights applied to bf16 tensors: each product and the sum are rounded to bf16 (vae.py:932-935)
        const float wb = static_cast<float>(static_cast<double>(f) / blend);
        const float wa = static_cast<float>(1.0 - static_cast<double>(f) / blend);
        const bf16* a = prev + (static_cast<size_t>(prev0 + f) * plane + px) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            v[k] = bf16_round(__fadd_rn(bf16_round(__fmul_rn(__bfloat162float(a[k]), wa)), bf16_round(__fmul_rn(v[k], wb))));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out[(static_cast<size_t>(k) * Fout + dst0 + f) * plane + px] = __float2bfloat16_rn(v[k]);
}

__global__ void transpose_kernel(const bf16* __restrict__ in, int R, int C, int ldi, bf16* __restrict__ out, int ldo) {
    __shared__ bf16 tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = in[static_cast<size_t>(r) * ldi + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) out[static_cast<size_t>(c) * ldo + r] = tile[threadIdx.x][i];
    }
}

__global__ void repack_conv_kernel(const void* __restrict__ src, int dtype, int Cout, int Cin, int taps, int Cin_pad,
                                   bf16* __restrict__ dst) {
    const size_t total = static_cast<size_t>(Cout) * Cin * taps;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    const int tap = static_cast<int>(i % taps);
    const int ci = static_cast<int>((i / taps) % Cin);
    const int co = static_cast<int>(i / (static_cast<size_t>(taps) * Cin));
    float v;
    if (dtype == 0) v = static_cast<const float*>(src)[i];
    else if (dtype == 1) v = __bfloat162float(static_cast<const bf16*>(src)[i]);
    else v = __half2float(static_cast<const __half*>(src)[i]);
    dst[(static_cast<size_t>(co) * taps + tap) * Cin_pad + ci] = __float2bfloat16_rn(v);
}

int blocks_for(size_t work, int threads) {
    size_t b = (work + threads - 1) / threads;
    const size_t cap = static_cast<size_t>(sm_count()) * 16;
    return static_cast<int>(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

int gn_channel_sums(const bf16* x, size_t P, int C, double* sums, int* nblocks, cudaStream_t st) {
    K5_REQUIRE(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "GroupNorm: channel count must be 8 * a divisor of 256");
    K5_REQUIRE(nblocks != nullptr, "GroupNorm: null block count");
    const int pos_per_block = 256 / (C / 8);
    size_t want = (P + pos_per_block * 8 - 1) / (static_cast<size_t>(pos_per_block) * 8);
    const size_t cap = GN_MAX_BLOCKS;
    const int grid = static_cast<int>(want < cap ? (want ? want : 1) : cap);
    gn_sums_kernel<<<grid, 256, 256 * 16 * sizeof(float), st>>>(x, P, C, sums);
    K5_CHECK_CUDA(cudaGetLastError());
    *nblocks = grid;
    return K5_OK;
}

int gn_finalize(const double* sums, int nblocks, size_t P, int C, int groups, float eps, float* mean_rstd, cudaStream_t st) {
    K5_REQUIRE(groups > 0 && groups <= 1024 && C % groups == 0 && nblocks > 0, "GroupNorm: bad group count");
    gn_finalize_kernel<<<groups, 256, 0, st>>>(sums, nblocks, static_cast<double>(P), C, groups, eps, mean_rstd);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int pad_gather(const bf16* x, int Ts, int Hs, int Ws, int C, int ft, int fh, int fw, const float* mean_rstd,
               const float* gamma, const float* beta, int groups, bool silu, bf16* out, cudaStream_t st) {
    K5_REQUIRE(C % 8 == 0 && (ft == 1 || ft == 2) && fh >= 1 && fw >= 1, "pad_gather: bad arguments");
    K5_REQUIRE(!mean_rstd || (gamma && beta && groups > 0 && C % groups == 0), "pad_gather: GroupNorm needs gamma / beta");
    const int T = ft == 1 ? Ts : 1 + (Ts - 1) * ft, H = Hs * fh, W = Ws * fw;
    K5_REQUIRE(256 % (C / 8) == 0, "pad_gather: channel count must be 8 * a divisor of 256");
    pad_gather_kernel<<<(T + 2) * (H + 2), 256, 0, st>>>(x, Ts, Hs, Ws, C, ft, fh, fw, T, H, W, mean_rstd, gamma, beta,
                                                         mean_rstd ? C / groups : 1, silu ? 1 : 0, out);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int gn_apply(const bf16* x, size_t P, int C, const float* mean_rstd, const float* gamma, const float* beta, int groups,
             bool silu, bf16* out, cudaStream_t st) {
    K5_REQUIRE(C % 8 == 0 && groups > 0 && C % groups == 0, "gn_apply: bad arguments");
    gn_apply_kernel<<<blocks_for(P * (C / 8), 256), 256, 0, st>>>(x, P, C, mean_rstd, gamma, beta, C / groups, silu ? 1 : 0, out);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int post_quant_pad(const float* z, int Cz, int T, int H, int W, int t0, int Tz, const float* w, const float* b, bf16* out,
                   cudaStream_t st) {
    K5_REQUIRE(Cz > 0 && Cz <= 16, "post_quant_conv: at most 16 latent channels");
    K5_REQUIRE(t0 >= 0 && t0 + T <= Tz, "post_quant_conv: tile outside the latent");
    const size_t total = static_cast<size_t>(T + 2) * (H + 2) * (W + 2);
    post_quant_pad_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, st>>>(z, Cz, Tz, H, W, t0, T, w, b, out);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int softmax_frame_causal(const float* s, int N, int lds, bf16* p, int ldp, int hw, float scale, int row0, int rows,
                         cudaStream_t st) {
    K5_REQUIRE(hw % 8 == 0 && lds % 8 == 0 && ldp % 8 == 0 && N % hw == 0 && row0 >= 0 && row0 + rows <= N && s && p,
               "softmax: bad arguments");
    if (rows <= 0) return K5_OK;
    softmax_frame_causal_kernel<<<rows, 256, 0, st>>>(s, lds, p, ldp, hw, scale * 1.4426950408889634f, row0);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int emit_frames(const bf16* cur, int src0, const bf16* prev, int prev0, int blend, int count, int H, int W, int Fout,
                int dst0, bf16* out, cudaStream_t st) {
    K5_REQUIRE(count > 0 && blend >= 0 && blend <= count && (blend == 0 || prev), "emit_frames: bad arguments");
    K5_REQUIRE(dst0 >= 0 && dst0 + count <= Fout, "emit_frames: frames outside the output video");
    const size_t total = static_cast<size_t>(H) * W * count;
    emit_frames_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(cur, src0, prev, prev0, blend, count, H, W,
                                                                                   Fout, dst0, out);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

}  // namespace k5

namespace k5 {

int transpose_bf16(const bf16* in, int R, int C, int ldi, bf16* out, int ldo, cudaStream_t st) {
    K5_REQUIRE(R > 0 && C > 0, "transpose: empty matrix");
    transpose_kernel<<<dim3((C + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, st>>>(in, R, C, ldi, out, ldo);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int repack_conv_weight(const void* src, int dtype, int Cout, int Cin, int taps, int Cin_pad, bf16* dst, cudaStream_t st) {
    K5_REQUIRE(dtype >= 0 && dtype <= 2 && Cin_pad >= Cin, "repack_conv_weight: bad arguments");
    const size_t total = static_cast<size_t>(Cout) * Cin * taps;
    repack_conv_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(src, dtype, Cout, Cin, taps, Cin_pad, dst);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

}  // namespace k5
