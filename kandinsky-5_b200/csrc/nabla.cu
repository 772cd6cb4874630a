// NABLA adaptive block selection (nablaT_v2, kandinsky/models/utils.py:136-163) and the Sliding-Tile-Attention
// block mask (fast_sta_nabla, models/utils.py:108-133), as HBM/L2-bound CUDA kernels.
//
// Reference algorithm, per head and per 64-token query block i (S/64 = nb blocks):
//   qa = mean_64(q), ka = mean_64(k)            (bf16 in, fp32 accumulate, bf16 out: torch.mean on bf16)
//   map[i, :] = softmax_fp32( bf16(qa_i . ka_j) / sqrt(64) )
//   sort ascending, cumulative sum, keep the entries whose cumulative mass is >= 1 - P, OR the STA mask
//   kv_nb = number kept, kv_inds = their block ids
// Here: one kernel pools q and k, one kernel does a whole map row per thread block (scores, softmax, a radix search
// for the probability of the cut entry instead of a sort, OR with STA, ordered compaction).  The map is never
// written to HBM (the reference materialises [1,28,1464,1464] fp32 = 240 MB per layer).
#include "nabla.h"
#include "ptx.cuh"

namespace k5 {

namespace {

constexpr int NB_MAX = 2048;          // max 64-token blocks per sequence (131 072 tokens)
constexpr int SEL_THREADS = 256;
constexpr int SEL_WARPS = SEL_THREADS / 32;
constexpr int EPT = NB_MAX / SEL_THREADS;      // map entries per thread (entry j = tid + e * SEL_THREADS)

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

// Deterministic block reduction (warp tree, then the warps in order); every thread gets the result.
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
#pragma unroll
    for (int w = 1; w < SEL_WARPS; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

// One (head, query block) row of the map per thread block.
//
// The reference sorts the row ascending, takes the cumulative sum and keeps every entry from the first position whose
// cumulative mass reaches 1 - P.  Only that cut matters, not the order, so no sort is done: with A(v) = sum of the
// probabilities whose bit pattern is below v (non-negative floats order like their bit patterns), the probability
// tau of the cut entry is the LARGEST pattern with A(tau) < 1 - P.  tau is built four bits at a time (eight rounds, each
// evaluating the 15 candidates of the next hex digit with one block reduction); entries above tau are kept, entries
// below are dropped, and of the c entries equal to tau (frequent: the scores are bf16) the first m in index order are
// dropped - the tie order of the reference's stable sort - where m is the number of copies of tau that still fit
// under 1 - P.  (A 2 048-wide bitonic sort + scan took 9.2 ms per layer at the 10 s size, 64 % of it in the sort.)
__global__ void __launch_bounds__(SEL_THREADS)
nabla_row_kernel(const bf16* __restrict__ qa, const bf16* __restrict__ ka, int nbq, int nb, int heads, float need,
                 const uint8_t* __restrict__ sta, int sta_row0, int32_t* __restrict__ kv_count,
                 int32_t* __restrict__ kv_index, float* __restrict__ density_acc) {
    __shared__ float qrow[64];
    __shared__ float red[SEL_WARPS];
    __shared__ float cand[SEL_WARPS][16];
    __shared__ float total[16];
    __shared__ int warp_counts[2][SEL_WARPS];
    const int i = blockIdx.x, h = blockIdx.y;
    const int cols = heads * 64;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid < 64) qrow[tid] = __bfloat162float(qa[static_cast<size_t>(i) * cols + h * 64 + tid]);
    __syncthreads();
    // scores: bf16(q . k) / 8 kept in bf16 (the reference's matmul and division run in bf16), then fp32 softmax
    float pv[EPT];
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int j = tid + e * SEL_THREADS;
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
        pv[e] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        pv[e] = (tid + e * SEL_THREADS < nb) ? expf(pv[e] - mx) : 0.f;
        sum += pv[e];
    }
    sum = block_reduce(sum, red, false);
    const float inv = 1.0f / sum;
    uint32_t bits[EPT];                       // bit patterns of the probabilities; entries past nb never match anything
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        pv[e] *= inv;
        bits[e] = __float_as_uint(pv[e]);
    }
    // ---- tau: the largest pattern v with A(v) < need, one hex digit per round
    uint32_t v = 0;
    float a_v = 0.f;                          // A(v)
    for (int shift = 28; shift >= 0; shift -= 4) {
        const uint32_t hv = v >> shift;       // low digit is 0
        float part[15];
#pragma unroll
        for (int d = 0; d < 15; ++d) part[d] = 0.f;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            if (tid + e * SEL_THREADS < nb) {
                const uint32_t hb = bits[e] >> shift;
#pragma unroll
                for (int d = 0; d < 15; ++d)
                    if (hb < hv + d + 1) part[d] += pv[e];          // below candidate v | ((d + 1) << shift)
            }
        }
#pragma unroll
        for (int d = 0; d < 15; ++d) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part[d] += __shfl_xor_sync(0xffffffffu, part[d], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < 15; ++d) cand[warp][d] = part[d];
        }
        __syncthreads();
        if (tid < 15) {
            float t = cand[0][tid];
#pragma unroll
            for (int w = 1; w < SEL_WARPS; ++w) t += cand[w][tid];
            total[tid] = t;
        }
        __syncthreads();
        int best = 0;                         // largest digit whose candidate still has A < need (A is monotone)
#pragma unroll
        for (int d = 0; d < 15; ++d)
            if (total[d] < need) best = d + 1;
        if (best > 0) {
            v |= static_cast<uint32_t>(best) << shift;
            a_v = total[best - 1];
        }
    }
    // ---- ties: entries equal to tau, the first m of them in index order are dropped
    const float tau = __uint_as_float(v);
    int m = NB_MAX;                           // tau == 0: zero-probability entries never reach the mass
    if (need <= 0.f) {
        m = 0;                                // P >= 1: the reference keeps the whole row
    } else if (tau > 0.f && v != 0x7f800000u) {
        const double left = static_cast<double>(need) - static_cast<double>(a_v);
        const double q = ceil(left / static_cast<double>(tau)) - 1.0;
        m = q < 0.0 ? 0 : (q > static_cast<double>(NB_MAX) ? NB_MAX : static_cast<int>(q));
    }
    // ---- keep flags, OR with the STA row, ordered compaction (ascending block id)
    int32_t* out = kv_index + (static_cast<size_t>(h) * nbq + i) * nb;
    const uint8_t* sta_row = sta ? sta + static_cast<size_t>(sta_row0 + i) * nb : nullptr;
    int base = 0, tie_base = 0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int j = tid + e * SEL_THREADS;
        if (e * SEL_THREADS >= nb) break;                          // block-uniform
        const bool in = j < nb;
        const bool tie = in && bits[e] == v;
        const unsigned bal_t = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) warp_counts[0][warp] = __popc(bal_t);
        __syncthreads();
        int pre_t = 0, tot_t = 0;
#pragma unroll
        for (int w = 0; w < SEL_WARPS; ++w) {
            if (w < warp) pre_t += warp_counts[0][w];
            tot_t += warp_counts[0][w];
        }
        const int rank = tie_base + pre_t + __popc(bal_t & ((1u << lane) - 1u));
        tie_base += tot_t;
        bool k1 = in && (bits[e] > v || (tie && rank >= m));
        if (in && sta_row && sta_row[j]) k1 = true;
        const unsigned bal = __ballot_sync(0xffffffffu, k1);
        if (lane == 0) warp_counts[1][warp] = __popc(bal);
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < SEL_WARPS; ++w) {
            if (w < warp) pre += warp_counts[1][w];
            tot += warp_counts[1][w];
        }
        if (k1) out[base + pre + __popc(bal & ((1u << lane) - 1u))] = j;
        base += tot;
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
