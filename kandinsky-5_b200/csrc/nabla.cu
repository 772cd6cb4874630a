// NABLA adaptive block selection (nablaT_v2, kandinsky/models/utils.py:136-163) and the Sliding-Tile-Attention
// block mask (fast_sta_nabla, models/utils.py:108-133), as HBM/L2-bound CUDA kernels.
//
// Reference algorithm, per head and per 64-token query block i (S/64 = nb blocks):
//   qa = mean_64(q), ka = mean_64(k)            (bf16 in, fp32 accumulate, bf16 out: torch.mean on bf16)
//   map[i, :] = softmax_fp32( bf16(qa_i . ka_j) / sqrt(64) )
//   sort ascending, cumulative sum, keep the entries whose cumulative mass is >= 1 - P, OR the STA mask
//   kv_nb = number kept, kv_inds = their block ids
// Here: one kernel pools q and k, one kernel does eight map rows per thread block (scores against the pooled keys
// shared by the eight rows, then one warp per row: softmax, a radix search for the probability of the cut entry instead
// of a sort, OR with STA, ordered compaction).  The map is never written to HBM (the reference materialises [1,28,1464,1464] fp32 = 240 MB per layer).
#include "nabla.h"
#include "ptx.cuh"

namespace k5 {

namespace {

constexpr int NB_MAX = 2048;          // max 64-token blocks per sequence (131 072 tokens)
constexpr int SEL_THREADS = 256;
constexpr int SEL_ROWS = SEL_THREADS / 32;     // map rows (query blocks) per thread block: one per warp
constexpr int EPL = NB_MAX / 32;               // map entries per lane (entry j = lane + 32 e)

// pooled[b, c] = bf16( mean over the 64 rows of block b of x[:, c] ), fp32 sum in row order.  One launch pools q
// (blocks [0, nbq)) and k (blocks [nbq, nbq + nbk)); a thread owns 8 adjacent columns (16-byte loads, 8 rows in flight).
__global__ void __launch_bounds__(256)
pool64_kernel(const bf16* __restrict__ q, int ldq, int nbq, const bf16* __restrict__ k, int ldk, int cols,
              bf16* __restrict__ qa, bf16* __restrict__ ka) {
    const bool is_q = static_cast<int>(blockIdx.x) < nbq;
    const int b = is_q ? blockIdx.x : blockIdx.x - nbq;
    const bf16* x = is_q ? q : k;
    const int ld = is_q ? ldq : ldk;
    bf16* pooled = is_q ? qa : ka;
    for (int c8 = threadIdx.x; c8 < cols / 8; c8 += blockDim.x) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const bf16* p = x + static_cast<size_t>(b) * 64 * ld + c8 * 8;
#pragma unroll
        for (int r0 = 0; r0 < 64; r0 += 8) {
            uint4 u[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) u[r] = __ldg(reinterpret_cast<const uint4*>(p + static_cast<size_t>(r0 + r) * ld));
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t w[4] = {u[r].x, u[r].y, u[r].z, u[r].w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    acc[2 * t] += bf16_lo(w[t]);
                    acc[2 * t + 1] += bf16_hi(w[t]);
                }
            }
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0] * (1.0f / 64.0f), acc[1] * (1.0f / 64.0f));
        o.y = pack_bf16x2(acc[2] * (1.0f / 64.0f), acc[3] * (1.0f / 64.0f));
        o.z = pack_bf16x2(acc[4] * (1.0f / 64.0f), acc[5] * (1.0f / 64.0f));
        o.w = pack_bf16x2(acc[6] * (1.0f / 64.0f), acc[7] * (1.0f / 64.0f));
        *reinterpret_cast<uint4*>(pooled + static_cast<size_t>(b) * cols + c8 * 8) = o;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Eight map rows (query blocks i0 .. i0 + 7 of one head) per thread block.
//
// Phase 1, all 256 threads: the scores of the eight rows against every pooled key.  A key row (128 B) is fetched once
// and used for all eight query rows, so the pooled K of a head is read from L2 once per EIGHT map rows (one map row per
// block re-read it for every row: 7.7 GB of L2 traffic per layer at the 10 s size, which bounded that kernel).
// Phase 2, one warp per row, no block-wide synchronisation: fp32 softmax, the cut, OR with the STA row, compaction.
//
// The reference sorts the row ascending, takes the cumulative sum and keeps every entry from the first position whose
// cumulative mass reaches 1 - P.  Only that cut matters, not the order, so no sort is done: with A(v) = sum of the
// probabilities whose bit pattern is below v (non-negative floats order like their bit patterns), the probability
// tau of the cut entry is the LARGEST pattern with A(tau) < 1 - P.  tau is built bit by bit from the top (31 rounds, each
// one masked sum over the row reduced with warp shuffles); entries above tau are kept, entries
// below are dropped, and of the c entries equal to tau (frequent: the scores are bf16) the first m in index order are
// dropped - the tie order of the reference's stable sort - where m is the number of copies of tau that still fit
// under 1 - P.  (History at the 10 s size, per layer: 2 048-wide bitonic sort + scan 9.4 ms; radix search with one map
// row per block 3.25 ms; eight rows per block, hex digits 1.83 ms.)
__global__ void __launch_bounds__(SEL_THREADS)
nabla_rows_kernel(const bf16* __restrict__ qa, const bf16* __restrict__ ka, int nbq, int nb, int heads, float need,
                  const uint8_t* __restrict__ sta, int sta_row0, int32_t* __restrict__ kv_count,
                  int32_t* __restrict__ kv_index, float* __restrict__ density_acc) {
    extern __shared__ float sm[];
    float* qrows = sm;                          // [SEL_ROWS][64]
    float* sc = sm + SEL_ROWS * 64;             // [SEL_ROWS][nb]
    const int i0 = blockIdx.x * SEL_ROWS, h = blockIdx.y;
    const int cols = heads * 64;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < SEL_ROWS * 64; t += SEL_THREADS) {
        const int r = t >> 6;
        qrows[t] = (i0 + r < nbq) ? __bfloat162float(qa[static_cast<size_t>(i0 + r) * cols + h * 64 + (t & 63)]) : 0.f;
    }
    __syncthreads();
    // ---- phase 1: scores, bf16(q . k) / 8 kept in bf16 (the reference's matmul and division run in bf16)
    for (int j = tid; j < nb; j += SEL_THREADS) {
        const uint4* kr = reinterpret_cast<const uint4*>(ka + static_cast<size_t>(j) * cols + h * 64);
        float kf[64];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const uint4 u = __ldg(kr + v);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                kf[8 * v + 2 * t] = bf16_lo(w[t]);
                kf[8 * v + 2 * t + 1] = bf16_hi(w[t]);
            }
        }
#pragma unroll
        for (int r = 0; r < SEL_ROWS; ++r) {
            const float4* q4 = reinterpret_cast<const float4*>(qrows + r * 64);
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
                const float4 q = q4[d];          // same address in every lane: a broadcast
                acc = fmaf(q.x, kf[4 * d], acc);
                acc = fmaf(q.y, kf[4 * d + 1], acc);
                acc = fmaf(q.z, kf[4 * d + 2], acc);
                acc = fmaf(q.w, kf[4 * d + 3], acc);
            }
            sc[r * nb + j] = bf16_round(bf16_round(acc) * 0.125f);
        }
    }
    __syncthreads();
    // ---- phase 2: one warp per map row
    const int i = i0 + warp;
    if (i >= nbq) return;
    const float* srow = sc + warp * nb;
    float pv[EPL];
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int j = lane + 32 * e;
        pv[e] = j < nb ? srow[j] : -INFINITY;
        mx = fmaxf(mx, pv[e]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        pv[e] = (lane + 32 * e < nb) ? expf(pv[e] - mx) : 0.f;
        sum += pv[e];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int e = 0; e < EPL; ++e) pv[e] *= inv;          // entries past nb are +0.0: never above a positive tau, never a tie
    // ---- tau: the largest pattern v with A(v) < need, one bit per round (A is monotone in v, so the greedy choice of
    // every bit from the top is exact; 31 rounds of one compare + one predicated add per entry cost a quarter of
    // eight rounds with the 15 candidates of a hex digit)
    uint32_t v = 0;
    float a_v = 0.f;                          // A(v)
    for (int bit = 30; bit >= 0; --bit) {     // probabilities are non-negative: the sign bit stays 0
        const uint32_t c = v | (1u << bit);
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            if (32 * e < nb) {                                          // warp-uniform; entries past nb hold +0.0
                if (__float_as_uint(pv[e]) < c) part += pv[e];
            }
        }
        const float t = warp_sum(part);
        if (t < need) {
            v = c;
            a_v = t;
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
    const unsigned lt = (1u << lane) - 1u;
    int base = 0, tie_base = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        if (32 * e < nb) {                                              // warp-uniform
            const int j = lane + 32 * e;
            const bool in = j < nb;
            const uint32_t bits = __float_as_uint(pv[e]);
            const bool tie = in && bits == v;
            const unsigned bal_t = __ballot_sync(0xffffffffu, tie);
            const int rank = tie_base + __popc(bal_t & lt);
            tie_base += __popc(bal_t);
            bool k1 = in && (bits > v || (tie && rank >= m));
            if (in && sta_row && sta_row[j]) k1 = true;
            const unsigned bal = __ballot_sync(0xffffffffu, k1);
            if (k1) out[base + __popc(bal & lt)] = j;
            base += __popc(bal);
        }
    }
    if (lane == 0) {
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
int nabla_select_launches() { return 2; }

int nabla_select(const bf16* q, int ldq, int Sq, const bf16* k, int ldk, int Sk, int heads, float P, const uint8_t* sta,
                 int sta_row0, int32_t* kv_count, int32_t* kv_index, float* workspace, float* density_acc, cudaStream_t st) {
    K5_REQUIRE(Sq % 64 == 0 && Sq >= 64 && Sk % 64 == 0 && Sk >= 64, "NABLA: token counts must be multiples of 64");
    const int nbq = Sq / 64, nbk = Sk / 64;
    K5_REQUIRE(nbk <= NB_MAX, "NABLA: at most 2048 key blocks (131072 tokens)");
    K5_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0, "NABLA: pitches must be x8");
    const int cols = heads * 64;
    bf16* qa = reinterpret_cast<bf16*>(workspace);
    bf16* ka = qa + static_cast<size_t>(nbq) * cols;
    K5_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "NABLA: q, k and the workspace must be 16-byte aligned");
    pool64_kernel<<<nbq + nbk, 256, 0, st>>>(q, ldq, nbq, k, ldk, cols, qa, ka);
    // the reference compares against the Python double 1 - P; undo the float round trip of P first
    const double Pd = nearbyint(static_cast<double>(P) * 1e6) / 1e6;
    const float need = static_cast<float>(1.0 - Pd);
    const size_t smem = (static_cast<size_t>(SEL_ROWS) * 64 + static_cast<size_t>(SEL_ROWS) * nbk) * sizeof(float);
    {
        static PerDevice<int> pd;                 // > 48 KB of dynamic shared memory needs the opt-in, per device
        const int dev = current_device();
        std::lock_guard<std::mutex> lk(pd.m);
        if (!pd.set[dev]) {
            K5_CHECK_CUDA(cudaFuncSetAttribute(nabla_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>((SEL_ROWS * 64 + SEL_ROWS * NB_MAX) * sizeof(float))));
            pd.set[dev] = true;
        }
    }
    nabla_rows_kernel<<<dim3((nbq + SEL_ROWS - 1) / SEL_ROWS, heads), SEL_THREADS, smem, st>>>(qa, ka, nbq, nbk, heads, need, sta,
                                                                                                sta_row0, kv_count, kv_index,
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
