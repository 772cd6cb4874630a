// HBM-bound row / element kernels of the DiT path: LayerNorm(+modulation), the fp32 modulation and
// time-embedding GEMVs, patchify / un-patchify, RoPE table expansion, CFG combine and the Euler update.
// Each cites the reference op it replaces; rounding points follow SURVEY.md Appendix A.
#include "common.h"
#include "ptx.cuh"
#include "rowops.h"
#include <cuda_fp16.h>

namespace k5 {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// ----------------------------------------------------------------------------- LayerNorm rows
// apply_scale_shift_norm (nn.py:25-28): bf16(LN(x) * (scale + 1) + shift), and the affine LayerNorm of
// TextEmbeddings (nn.py:70-72): bf16(LN(x) * w + b).  One warp per row, the row is held in registers
// (D <= 2048, D % 256 == 0), two-pass statistics in fp32.
constexpr int LN_MAXV = 8;
constexpr int LN_BLOCKS_PER_SM = 2;     // 127 registers x 256 threads: two blocks are resident per SM
// Persistent: a warp walks rows warp_id, warp_id + n_warps, ... and issues the loads of its NEXT row before it reduces
// and writes the current one, so HBM reads stay in flight across the dependent reduce -> normalise -> store chain.
__global__ void __launch_bounds__(256) ln_rows_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ out, int ldo,
                                                      int S, int D, const float* __restrict__ mul,
                                                      const float* __restrict__ add, int plus_one, float eps) {
    const int lane = threadIdx.x & 31;
    const int nv = D >> 8;
    const int n_warps = gridDim.x * 8;
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= S) return;
    const float one = plus_one ? 1.0f : 0.0f;
    uint4 nxt[LN_MAXV];
    {
        const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx);
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i)
            if (i < nv) nxt[i] = __ldg(xr + i * 32 + lane);
    }
    for (; row < S; row += n_warps) {
        float v[LN_MAXV][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (i < nv) {
                const uint32_t w[4] = {nxt[i].x, nxt[i].y, nxt[i].z, nxt[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[i][2 * j] = bf16_lo(w[j]);
                    v[i][2 * j + 1] = bf16_hi(w[j]);
                    sum += v[i][2 * j] + v[i][2 * j + 1];
                }
            }
        }
        if (row + n_warps < S) {
            const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row + n_warps) * ldx);
#pragma unroll
            for (int i = 0; i < LN_MAXV; ++i)
                if (i < nv) nxt[i] = __ldg(xr + i * 32 + lane);
        }
        const float mean = warp_sum(sum) / static_cast<float>(D);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (i < nv) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = v[i][j] - mean;
                    sq += d * d;
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(D) + eps);
        uint4* orow = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * ldo);
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (i < nv) {
                const int c0 = (i * 32 + lane) * 8;
                const float4 m0 = __ldg(reinterpret_cast<const float4*>(mul + c0));
                const float4 m1 = __ldg(reinterpret_cast<const float4*>(mul + c0 + 4));
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(add + c0));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(add + c0 + 4));
                const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                const float aa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float y[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float n = (v[i][j] - mean) * rstd;
                    y[j] = __fadd_rn(__fmul_rn(n, __fadd_rn(mm[j], one)), aa[j]);
                }
                uint4 o;
                o.x = pack_bf16x2(y[0], y[1]);
                o.y = pack_bf16x2(y[2], y[3]);
                o.z = pack_bf16x2(y[4], y[5]);
                o.w = pack_bf16x2(y[6], y[7]);
                orow[i * 32 + lane] = o;
            }
        }
    }
}

// ----------------------------------------------------------------------------- fp32 GEMV
// Modulation (nn.py:153-164) and TimeEmbeddings (nn.py:56-61) run in fp32: out = act_out(W . act_in(x) + b).
// One warp per output row; K % 128 == 0.  All 35 modulation layers of a forward are one launch over
// their concatenated weight rows (the input SiLU(time_embed) is the same for all of them).
__global__ void __launch_bounds__(256) gemv_f32_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                       const float* __restrict__ x, float* __restrict__ out, int N, int K,
                                                       int silu_in, int silu_out) {
    extern __shared__ float xs[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float v = x[k];
        xs[k] = silu_in ? silu_f(v) : v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    for (int n = blockIdx.x * warps + (threadIdx.x >> 5); n < N; n += gridDim.x * warps) {
        const float4* wr = reinterpret_cast<const float4*>(W + static_cast<size_t>(n) * K);
        float acc = 0.f;
        for (int k4 = lane; k4 < (K >> 2); k4 += 32) {
            const float4 w = __ldg(wr + k4);
            const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * k4);
            acc = fmaf(w.x, xv.x, acc);
            acc = fmaf(w.y, xv.y, acc);
            acc = fmaf(w.z, xv.z, acc);
            acc = fmaf(w.w, xv.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float r = acc + (bias ? bias[n] : 0.f);
            out[n] = silu_out ? silu_f(r) : r;
        }
    }
}

// sinusoidal features of TimeEmbeddings (nn.py:57-58): [cos(t * f), sin(t * f)]
__global__ void time_features_kernel(const float* __restrict__ freqs, float t, float* __restrict__ out, int half) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < half) {
        const float a = __fmul_rn(t, freqs[i]);
        out[i] = cosf(a);
        out[half + i] = sinf(a);
    }
}

// pooled_text_embeddings (nn.py:64-72 via dit.py:134): bf16 Linear(768 -> time_dim) -> LayerNorm(affine, fp32)
// -> bf16, added to the fp32 time embedding.  One block, blockDim = time_dim.
__global__ void pooled_embed_kernel(const bf16* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ lnw,
                                    const float* __restrict__ lnb, const bf16* __restrict__ pooled, int Kin,
                                    float* __restrict__ time_embed /* in/out [time_dim] */, float eps) {
    extern __shared__ float sh[];           // [Kin] input, then [N] values, then 32 scratch
    const int N = blockDim.x;
    float* xin = sh;
    float* val = sh + Kin;
    float* red = val + N;
    for (int k = threadIdx.x; k < Kin; k += N) xin[k] = __bfloat162float(pooled[k]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = N >> 5;
    for (int n = warp; n < N; n += warps) {
        const bf16* wr = W + static_cast<size_t>(n) * Kin;
        float acc = 0.f;
        for (int k = lane; k < Kin; k += 32) acc = fmaf(__bfloat162float(wr[k]), xin[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) val[n] = bf16_round(acc + bias[n]);
    }
    __syncthreads();
    const float v = val[threadIdx.x];
    float s = warp_sum(v);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < warps; ++w) tot += red[w];
    const float mean = tot / N;
    __syncthreads();
    const float d = v - mean;
    s = warp_sum(d * d);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    tot = 0.f;
    for (int w = 0; w < warps; ++w) tot += red[w];
    const float rstd = rsqrtf(tot / N + eps);
    const float y = bf16_round(d * rstd * lnw[threadIdx.x] + lnb[threadIdx.x]);
    time_embed[threadIdx.x] += y;
}

// ----------------------------------------------------------------------------- token order helpers
struct Grid3 {
    int T, Hp, Wp, fractal;
};
// token index -> (t, h, w) on the patch grid.  Plain order (t,h,w) or the reference's "fractal" order
// (t, hb, wb, hi, wi) with 8x8 tiles (models/utils.py:31-41,54-78).
__device__ __forceinline__ void token_coords(const Grid3& g, int s, int& t, int& h, int& w) {
    if (!g.fractal) {
        w = s % g.Wp;
        h = (s / g.Wp) % g.Hp;
        t = s / (g.Wp * g.Hp);
    } else {
        const int wi = s & 7, hi = (s >> 3) & 7;
        const int blk = s >> 6;
        const int wb = blk % (g.Wp >> 3);
        const int hb = (blk / (g.Wp >> 3)) % (g.Hp >> 3);
        t = blk / ((g.Wp >> 3) * (g.Hp >> 3));
        h = hb * 8 + hi;
        w = wb * 8 + wi;
    }
}

// VisualEmbeddings patchify (nn.py:81-95) for patch (1,2,2): A[s, (ph*2+pw)*C + c] = bf16(x[t, 2h+ph, 2w+pw, c]),
// zero padded to KP columns, rows written in engine token order.  If x has only Cimg < C channels the rest are
// the zero visual_cond / mask channels of generation_utils.py:107-112.
__global__ void patchify_kernel(const float* __restrict__ x, int Cx, int C, Grid3 g, bf16* __restrict__ A, int KP) {
    const int s = blockIdx.x;
    int t, h, w;
    token_coords(g, s, t, h, w);
    const int H = g.Hp * 2, W = g.Wp * 2;
    for (int k = threadIdx.x; k < KP; k += blockDim.x) {
        float v = 0.f;
        if (k < 4 * C) {
            const int c = k % C, pp = k / C;
            const int ph = pp >> 1, pw = pp & 1;
            if (c < Cx) v = x[((static_cast<size_t>(t) * H + 2 * h + ph) * W + 2 * w + pw) * Cx + c];
        }
        A[static_cast<size_t>(s) * KP + k] = __float2bfloat16_rn(v);
    }
}

// OutLayer un-patchify (nn.py:385-399): y[s, c*4 + ph*2 + pw] -> out[t, 2h+ph, 2w+pw, c]
__global__ void unpatchify_kernel(const bf16* __restrict__ y, int ldy, Grid3 g, int Cout, bf16* __restrict__ out) {
    const int s = blockIdx.x;
    int t, h, w;
    token_coords(g, s, t, h, w);
    const int H = g.Hp * 2, W = g.Wp * 2;
    for (int o = threadIdx.x; o < 4 * Cout; o += blockDim.x) {
        const int c = o >> 2, ph = (o >> 1) & 1, pw = o & 1;
        out[((static_cast<size_t>(t) * H + 2 * h + ph) * W + 2 * w + pw) * Cout + c] = y[static_cast<size_t>(s) * ldy + o];
    }
}

// RoPE3D (nn.py:132-150) expanded to a per-token table of (cos, sin) for the 32 rotation pairs, in engine
// token order.  args_* are the reference's buffers pos x freq; the division by scale_factor and the
// cos / sin happen here exactly as in the reference's forward.
__global__ void rope3d_kernel(const float* __restrict__ at, const float* __restrict__ ah, const float* __restrict__ aw,
                              int nt, int nh, int nw, const int* __restrict__ pt, const int* __restrict__ ph,
                              const int* __restrict__ pw, float sf0, float sf1, float sf2, Grid3 g,
                              float2* __restrict__ table) {
    const int s = blockIdx.x;
    int t, h, w;
    token_coords(g, s, t, h, w);
    const int i = threadIdx.x;
    if (i >= nt + nh + nw) return;
    float a;
    if (i < nt) a = __fdiv_rn(at[pt[t] * nt + i], sf0);
    else if (i < nt + nh) a = __fdiv_rn(ah[ph[h] * nh + (i - nt)], sf1);
    else a = __fdiv_rn(aw[pw[w] * nw + (i - nt - nh)], sf2);
    table[static_cast<size_t>(s) * (nt + nh + nw) + i] = make_float2(cosf(a), sinf(a));
}

// RoPE1D (nn.py:110-116)
__global__ void rope1d_kernel(const float* __restrict__ args, int np, const int* __restrict__ pos, int L,
                              float2* __restrict__ table) {
    const int s = blockIdx.x, i = threadIdx.x;
    if (s < L && i < np) {
        const float a = args[pos[s] * np + i];
        table[static_cast<size_t>(s) * np + i] = make_float2(cosf(a), sinf(a));
    }
}

// ----------------------------------------------------------------------------- sampler element ops
// CFG combine (generation_utils.py:74-76), every op rounded to bf16: v = vu + w * (vc - vu)
__global__ void cfg_combine_kernel(const bf16* __restrict__ vc, const bf16* __restrict__ vu, float w, bf16* __restrict__ out,
                                   size_t n) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) {
        const float c = __bfloat162float(vc[i]), u = __bfloat162float(vu[i]);
        const float d = bf16_round(c - u);
        const float m = bf16_round(w * d);
        out[i] = __float2bfloat16_rn(u + m);
    }
}
// Euler step (generation_utils.py:128): img(fp32) += bf16(dt * v)
__global__ void euler_kernel(float* __restrict__ img, const bf16* __restrict__ v, float dt, size_t n) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) img[i] = __fadd_rn(img[i], bf16_round(__fmul_rn(dt, __bfloat162float(v[i]))));
}

// MagCache residual arithmetic (magcache_utils.py:82-88), bf16 tensors, 8 elements per thread:
//   sub: out = bf16(a - b)   (residual = visual_embed_out - visual_embed_in)
//   add: out = bf16(a + b)   (visual_embed_in + cached residual)
__global__ void __launch_bounds__(256)
bf16_addsub_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, size_t n8, float sign) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 x = a[i], y = b[i];
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[j] = pack_bf16x2(fmaf(sign, bf16_lo(ys[j]), bf16_lo(xs[j])), fmaf(sign, bf16_hi(ys[j]), bf16_hi(xs[j])));
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void convert_to_bf16_kernel(const void* __restrict__ src, int src_dtype, bf16* __restrict__ dst, size_t n,
                                       int rows, int cols, int ld_dst) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float v;
    if (src_dtype == 0) v = static_cast<const float*>(src)[i];
    else if (src_dtype == 1) v = __bfloat162float(static_cast<const bf16*>(src)[i]);
    else v = __half2float(static_cast<const __half*>(src)[i]);
    const size_t r = i / cols, c = i % cols;
    dst[r * ld_dst + c] = __float2bfloat16_rn(v);
}
__global__ void convert_to_f32_kernel(const void* __restrict__ src, int src_dtype, float* __restrict__ dst, size_t n,
                                      int round_bf16) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float v;
    if (src_dtype == 0) v = static_cast<const float*>(src)[i];
    else if (src_dtype == 1) v = __bfloat162float(static_cast<const bf16*>(src)[i]);
    else v = __half2float(static_cast<const __half*>(src)[i]);
    dst[i] = round_bf16 ? bf16_round(v) : v;
}

}  // namespace

int ln_rows(const bf16* x, int ldx, bf16* out, int ldo, int S, int D, const float* mul, const float* add, bool plus_one,
            float eps, cudaStream_t st) {
    K5_REQUIRE(D % 256 == 0 && D <= 256 * LN_MAXV, "LayerNorm: model_dim must be a multiple of 256, <= 2048");
    K5_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0, "LayerNorm: pitches must be x8");
    if (S <= 0) return K5_OK;
    int blocks = (S + 7) / 8;
    const int cap = sm_count() * LN_BLOCKS_PER_SM;
    if (blocks > cap) blocks = cap;
    ln_rows_kernel<<<blocks, 256, 0, st>>>(x, ldx, out, ldo, S, D, mul, add, plus_one ? 1 : 0, eps);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int gemv_f32(const float* W, const float* bias, const float* x, float* out, int N, int K, bool silu_in, bool silu_out,
             cudaStream_t st) {
    K5_REQUIRE(K % 4 == 0 && K <= 8192, "GEMV: K must be a multiple of 4 and <= 8192");
    int blocks = (N + 7) / 8;
    const int cap = sm_count() * 16;
    if (blocks > cap) blocks = cap;
    gemv_f32_kernel<<<blocks, 256, K * sizeof(float), st>>>(W, bias, x, out, N, K, silu_in ? 1 : 0, silu_out ? 1 : 0);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int time_features(const float* freqs, float t, float* out, int half, cudaStream_t st) {
    time_features_kernel<<<(half + 255) / 256, 256, 0, st>>>(freqs, t, out, half);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int pooled_embed(const bf16* W, const float* bias, const float* lnw, const float* lnb, const bf16* pooled, int Kin,
                 int time_dim, float* time_embed, float eps, cudaStream_t st) {
    K5_REQUIRE(time_dim % 32 == 0 && time_dim <= 1024, "pooled embed: time_dim must be x32 and <= 1024");
    const size_t sh = (Kin + time_dim + 32) * sizeof(float);
    pooled_embed_kernel<<<1, time_dim, sh, st>>>(W, bias, lnw, lnb, pooled, Kin, time_embed, eps);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int patchify(const float* x, int Cx, int C, int T, int Hp, int Wp, bool fractal, bf16* A, int KP, cudaStream_t st) {
    K5_REQUIRE(4 * C <= KP && KP % 8 == 0, "patchify: padded K too small");
    K5_REQUIRE(!fractal || (Hp % 8 == 0 && Wp % 8 == 0), "fractal token order needs patch-grid H, W divisible by 8");
    Grid3 g{T, Hp, Wp, fractal ? 1 : 0};
    patchify_kernel<<<T * Hp * Wp, 64, 0, st>>>(x, Cx, C, g, A, KP);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int unpatchify(const bf16* y, int ldy, int T, int Hp, int Wp, bool fractal, int Cout, bf16* out, cudaStream_t st) {
    Grid3 g{T, Hp, Wp, fractal ? 1 : 0};
    unpatchify_kernel<<<T * Hp * Wp, 64, 0, st>>>(y, ldy, g, Cout, out);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int rope3d_table(const float* at, const float* ah, const float* aw, int nt, int nh, int nw, const int* pt, const int* ph,
                 const int* pw, const float sf[3], int T, int Hp, int Wp, bool fractal, float2* table, cudaStream_t st) {
    K5_REQUIRE(nt + nh + nw == 32, "RoPE: axes_dims must sum to head_dim 64");
    Grid3 g{T, Hp, Wp, fractal ? 1 : 0};
    rope3d_kernel<<<T * Hp * Wp, 32, 0, st>>>(at, ah, aw, nt, nh, nw, pt, ph, pw, sf[0], sf[1], sf[2], g, table);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int rope1d_table(const float* args, int np, const int* pos, int L, float2* table, cudaStream_t st) {
    K5_REQUIRE(np == 32, "RoPE1D: head_dim must be 64");
    rope1d_kernel<<<L, 32, 0, st>>>(args, np, pos, L, table);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int cfg_combine(const bf16* vc, const bf16* vu, float w, bf16* out, size_t n, cudaStream_t st) {
    cfg_combine_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(vc, vu, w, out, n);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int euler_step(float* img, const bf16* v, float dt, size_t n, cudaStream_t st) {
    euler_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(img, v, dt, n);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int bf16_addsub(const bf16* a, const bf16* b, bf16* out, size_t n, bool subtract, cudaStream_t st) {
    K5_REQUIRE(n % 8 == 0, "bf16_addsub: element count must be a multiple of 8");
    if (n == 0) return K5_OK;
    size_t blocks = (n / 8 + 255) / 256;
    const size_t cap = static_cast<size_t>(sm_count()) * 16;
    bf16_addsub_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, st>>>(
        reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(out), n / 8,
        subtract ? -1.0f : 1.0f);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int convert_to_bf16(const void* src, int src_dtype, bf16* dst, int rows, int cols, int ld_dst, cudaStream_t st) {
    const size_t n = static_cast<size_t>(rows) * cols;
    if (n == 0) return K5_OK;
    convert_to_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, src_dtype, dst, n, rows, cols, ld_dst);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int convert_to_f32(const void* src, int src_dtype, float* dst, size_t n, bool round_bf16, cudaStream_t st) {
    if (n == 0) return K5_OK;
    convert_to_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, src_dtype, dst, n, round_bf16 ? 1 : 0);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

}  // namespace k5
