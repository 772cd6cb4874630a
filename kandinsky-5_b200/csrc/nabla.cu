// NABLA adaptive block selection (nablaT_v2, kandinsky/models/utils.py:136-163) and the Sliding-Tile-Attention
// block mask (fast_sta_nabla, models/utils.py:108-133), as HBM/L2-bound CUDA kernels.
//
// Reference algorithm, per head and per 64-token query block i (S/64 = nb blocks):
//   qa = mean_64(q), ka = mean_64(k)            (bf16 in, fp32 accumulate, bf16 out: torch.mean on bf16)
//   map[i, :] = softmax_fp32( bf16(qa_i . ka_j) / sqrt(64) )
//   sort ascending, cumulative sum, keep the entries whose cumulative mass is >= 1 - P, OR the STA mask
//   kv_nb = number kept, kv_inds = their block ids
// Here: one kernel pools q and k, one kernel does a whole map row per thread block (scores, softmax, bitonic
// sort by value in shared memory, block scan, threshold, OR with STA, ordered compaction).  The map is never
// written to HBM (the reference materialises [1,28,1464,1464] fp32 = 240 MB per layer).
#include "nabla.h"
#include "ptx.cuh"

namespace k5 {

namespace {

constexpr int NB_MAX = 2048;          // max 64-token blocks per sequence (131 072 tokens)
constexpr int SEL_THREADS = 1024;

// pooled[b, c] = bf16( mean over the 64 rows of block b of x[:, c] )
__global__ void pool64_kernel(const bf16* __restrict__ x, int ld, int cols, bf16* __restrict__ pooled) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float acc = 0.f;
        const bf16* p = x + static_cast<size_t>(b) * 64 * ld + c;
#pragma unroll 8
        for (int r = 0; r < 64; ++r) acc += __bfloat162float(p[static_cast<size_t>(r) * ld]);
        pooled[static_cast<size_t>(b) * cols + c] = __float2bfloat16_rn(acc * (1.0f / 64.0f));
    }
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

// One (head, query block) row per thread block.
__global__ void __launch_bounds__(SEL_THREADS)
nabla_row_kernel(const bf16* __restrict__ qa, const bf16* __restrict__ ka, int nbq, int nb, int heads, float need,
                 const uint8_t* __restrict__ sta, int sta_row0, int32_t* __restrict__ kv_count,
                 int32_t* __restrict__ kv_index, float* __restrict__ density_acc) {
    __shared__ float key[NB_MAX];
    __shared__ uint16_t idx[NB_MAX];
    __shared__ float scan[NB_MAX];
    __shared__ uint8_t keep[NB_MAX];
    __shared__ float qrow[64];
    __shared__ float red[32];
    __shared__ int s_cut;
    __shared__ int warp_counts[32];
    const int i = blockIdx.x, h = blockIdx.y;
    const int cols = heads * 64;
    const int tid = threadIdx.x;
    if (tid < 64) qrow[tid] = __bfloat162float(qa[static_cast<size_t>(i) * cols + h * 64 + tid]);
    __syncthreads();
    // scores: bf16(q . k) / 8 kept in bf16 (the reference's matmul and division run in bf16), then fp32 softmax
    float mx = -INFINITY;
    for (int j = tid; j < NB_MAX; j += SEL_THREADS) {
        float s = -INFINITY;
        if (j < nb) {
            const uint4* kr = reinterpret_cast<const uint4*>(ka + static_cast<size_t>(j) * cols + h * 64);
            float acc = 0.f;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint4 u = __ldg(kr + v);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    acc = fmaf(qrow[8 * v + 2 * t], bf16_lo(w[t]), acc);
                    acc = fmaf(qrow[8 * v + 2 * t + 1], bf16_hi(w[t]), acc);
                }
            }
            s = bf16_round(bf16_round(acc) * 0.125f);
        }
        key[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int j = tid; j < NB_MAX; j += SEL_THREADS) {
        const float e = j < nb ? expf(key[j] - mx) : 0.f;
        key[j] = e;
        sum += e;
    }
    sum = block_reduce(sum, red, false);
    const float inv = 1.0f / sum;
    for (int j = tid; j < NB_MAX; j += SEL_THREADS) {
        key[j] = j < nb ? key[j] * inv : INFINITY;       // padding sorts to the end
        idx[j] = static_cast<uint16_t>(j);
    }
    __syncthreads();
    // bitonic sort ascending by probability (ties by index so that the result is deterministic)
    int n2 = 1;
    while (n2 < nb) n2 <<= 1;
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < n2; t += SEL_THREADS) {
                const int p = t ^ j;
                if (p > t) {
                    const bool up = (t & k) == 0;
                    const float a = key[t], b = key[p];
                    const uint16_t ia = idx[t], ib = idx[p];
                    const bool gt = (a > b) || (a == b && ia > ib);
                    if (gt == up) {
                        key[t] = b;
                        key[p] = a;
                        idx[t] = ib;
                        idx[p] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    // inclusive scan of the sorted probabilities (Hillis-Steele, double buffered through `scan`)
    for (int t = tid; t < n2; t += SEL_THREADS) scan[t] = t < nb ? key[t] : 0.f;
    __syncthreads();
    for (int off = 1; off < n2; off <<= 1) {
        float v[NB_MAX / SEL_THREADS];
        int c = 0;
        for (int t = tid; t < n2; t += SEL_THREADS, ++c) v[c] = scan[t] + (t >= off ? scan[t - off] : 0.f);
        __syncthreads();
        c = 0;
        for (int t = tid; t < n2; t += SEL_THREADS, ++c) scan[t] = v[c];
        __syncthreads();
    }
    if (tid == 0) s_cut = nb;
    __syncthreads();
    for (int t = tid; t < nb; t += SEL_THREADS)
        if (scan[t] >= need && (t == 0 || scan[t - 1] < need)) atomicMin(&s_cut, t);
    __syncthreads();
    const int cut = s_cut;
    for (int t = tid; t < n2; t += SEL_THREADS) keep[t] = 0;
    __syncthreads();
    for (int t = tid; t < nb; t += SEL_THREADS)
        if (t >= cut) keep[idx[t]] = 1;
    __syncthreads();
    if (sta)
        for (int j = tid; j < nb; j += SEL_THREADS)
            if (sta[static_cast<size_t>(sta_row0 + i) * nb + j]) keep[j] = 1;
    __syncthreads();
    // ordered compaction (ascending block id)
    int32_t* out = kv_index + (static_cast<size_t>(h) * nbq + i) * nb;
    int base = 0;
    const int lane = tid & 31, warp = tid >> 5;
    for (int j0 = 0; j0 < nb; j0 += SEL_THREADS) {
        const int j = j0 + tid;
        const bool k1 = j < nb && keep[j];
        const unsigned bal = __ballot_sync(0xffffffffu, k1);
        if (lane == 0) warp_counts[warp] = __popc(bal);
        __syncthreads();
        int pre = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < warp) pre += warp_counts[w];
            tot += warp_counts[w];
        }
        if (k1) out[base + pre + __popc(bal & ((1u << lane) - 1u))] = j;
        base += tot;
        __syncthreads();
    }
    if (tid == 0) {
        kv_count[static_cast<size_t>(h) * nbq + i] = base;
        if (density_acc) {
            atomicAdd(density_acc, static_cast<float>(base));
            atomicAdd(density_acc + 1, static_cast<float>(nb));
        }
    }
}

__global__ void sta_mask_kernel(int T, int Hb, int Wb, int wT, int wH, int wW, uint8_t* __restrict__ out) {
    const int n = T * Hb * Wb;
    const size_t id = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (id >= static_cast<size_t>(n) * n) return;
    const int i = static_cast<int>(id / n), j = static_cast<int>(id % n);
    const int ti = i / (Hb * Wb), hi = (i / Wb) % Hb, wi = i % Wb;
    const int tj = j / (Hb * Wb), hj = (j / Wb) % Hb, wj = j % Wb;
    out[id] = (abs(ti - tj) <= wT / 2 && abs(hi - hj) <= wH / 2 && abs(wi - wj) <= wW / 2) ? 1 : 0;
}

}  // namespace

size_t nabla_workspace_floats(int S, int heads) {
    const size_t nb = S / 64;
    return static_cast<size_t>(heads) * nb * nb + 2 * nb * static_cast<size_t>(heads) * 64;
}
int nabla_select_launches() { return 3; }

int nabla_select(const bf16* q, int ldq, int Sq, const bf16* k, int ldk, int Sk, int heads, float P, const uint8_t* sta,
                 int sta_row0, int32_t* kv_count, int32_t* kv_index, float* workspace, float* density_acc, cudaStream_t st) {
    K5_REQUIRE(Sq % 64 == 0 && Sq >= 64 && Sk % 64 == 0 && Sk >= 64, "NABLA: token counts must be multiples of 64");
    const int nbq = Sq / 64, nbk = Sk / 64;
    K5_REQUIRE(nbk <= NB_MAX, "NABLA: at most 2048 key blocks (131072 tokens)");
    K5_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0, "NABLA: pitches must be x8");
    const int cols = heads * 64;
    bf16* qa = reinterpret_cast<bf16*>(workspace);
    bf16* ka = qa + static_cast<size_t>(nbq) * cols;
    pool64_kernel<<<nbq, 256, 0, st>>>(q, ldq, cols, qa);
    pool64_kernel<<<nbk, 256, 0, st>>>(k, ldk, cols, ka);
    // the reference compares against the Python double 1 - P; undo the float round trip of P first
    const double Pd = nearbyint(static_cast<double>(P) * 1e6) / 1e6;
    const float need = static_cast<float>(1.0 - Pd);
    nabla_row_kernel<<<dim3(nbq, heads), SEL_THREADS, 0, st>>>(qa, ka, nbq, nbk, heads, need, sta, sta_row0, kv_count, kv_index,
                                                               density_acc);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int sta_mask(int T, int Hb, int Wb, int wT, int wH, int wW, uint8_t* out, cudaStream_t st) {
    K5_REQUIRE(T > 0 && Hb > 0 && Wb > 0, "STA: empty block grid");
    const size_t n = static_cast<size_t>(T) * Hb * Wb;
    const size_t tot = n * n;
    sta_mask_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, st>>>(T, Hb, Wb, wT, wH, wW, out);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

}  // namespace k5
