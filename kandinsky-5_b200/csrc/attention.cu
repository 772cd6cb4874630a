// Non-causal multi-head attention forward for head_dim 64 on tcgen05 / TMEM:  O = softmax(Q K^T * scale) V.
//
// Replaces flash_attn_func at kandinsky/models/nn.py:201,254,336 (visual self-attention, cross-attention
// to the text tokens, text self-attention) and, with a per-(head, q-block) KV block list,
// flex_attention at nn.py:257-280 (NABLA, K2).  bf16 in / out, fp32 scores, softmax and accumulation.
//
// One persistent CTA per SM works on (head, 256-query-row) items, heads outermost so that the K/V of
// one head (12 MB at S = 47 616) stay L2 resident while all CTAs sweep its query tiles.  Roles:
//   warps 0-3 / 4-7 : softmax warpgroups for query tile 0 / 1 (thread = one query row, so row max / sum
//                     need no shuffles); they pull S from TMEM into registers (and hand the S buffer
//                     straight back to the tensor pipe), compute P = exp2(.) and write it as packed bf16
//                     into a separate TMEM buffer, rescale O lazily (only when the running max grew by
//                     > 2^8), and normalise + store O at the end;
//   warp 8          : TMA producer (Q tiles once per item, K and V tiles through a 5-stage ring);
//   warp 9 / 11     : tcgen05.mma issuers for query tile 0 / 1:  S_a = Q_a K_j^T (SS form, 128x128x64) as soon as
//                     the softmax warpgroup has S_a(j-1) in registers, O_a += P_a V_j (TS form: A = P from
//                     TMEM, B = V MN-major from shared memory, 128x64x128) as soon as P_a(j) is written;
//   warp 10         : TMEM allocation (S0 S1 O0 O1 P0 P1 = all 512 columns).
// Because P does not alias S, QK(j+1) never waits for the exponentials of tile j: the next scores are
// ready long before a warpgroup finishes P(j), so the MUFU / FMA pipes of the two warpgroups stay busy.
// The exponentials are the bottleneck at head_dim 64 (16 384 per 128x128 tile = 1024 MUFU cycles against 512
// tensor cycles), therefore a compile-time fraction of them is evaluated on the FMA pipe instead
// (Cody-Waite split + degree-3 polynomial, rel. error 9e-5 < bf16 rounding of P), and the surrounding
// arithmetic uses the packed fp32x2 forms (FFMA2 / FADD2) and the 3-input max to save issue slots.
#include <cstdlib>

#include "attention.h"
#include "common.h"
// Spacer instructions per MUFU pair in the softmax stream (see the loop): 2 measured best (profiles/r1_attention_v3_v4.md)
#ifndef K5_ATTN_PAD
#define K5_ATTN_PAD 2
#endif
#include "ptx.cuh"

namespace k5 {

#ifdef K5_ATTN_TRACE
// Debug build only (-DK5_ATTN_TRACE): clock64 stamps of the softmax warps 0 and 4 of CTA 0 (one per query tile, same
// scheduler) at four points of every KV tile, [2][4][512][8] long long, set through k5_debug_attn_trace().
__device__ long long* g_attn_trace = nullptr;
#define K5_TRACE(k)                                                                                       \
    do {                                                                                                  \
        if (g_attn_trace && blockIdx.x == 0 && lane == 0 && cnt < 512)                                    \
            g_attn_trace[((a * 4 + wq) * 512 + cnt) * 8 + (k)] = clock64();                                \
    } while (0)
#define K5_TRACE_V(k, v)                                                                                  \
    do {                                                                                                  \
        if (g_attn_trace && blockIdx.x == 0 && lane == 0 && cnt < 512)                                    \
            g_attn_trace[((a * 4 + wq) * 512 + cnt) * 8 + (k)] = (v);                                      \
    } while (0)
#define K5_TRACE_ISSUER(k)                                                                                \
    do {                                                                                                  \
        if (g_attn_trace && blockIdx.x == 0 && g < 512) g_attn_trace[(a * 4 * 512 + g) * 8 + (k)] = clock64(); \
    } while (0)
#else
#define K5_TRACE(k)
#define K5_TRACE_V(k, v)
#define K5_TRACE_ISSUER(k)
#endif

namespace {

constexpr int QT = 128;            // query rows per tile (2 tiles per CTA item)
constexpr int KT = 128;            // kv rows per tile
constexpr int HD = 64;             // head dim
constexpr int TILE_BYTES = 128 * 64 * 2;
constexpr int KV_STAGES = 5;
constexpr int ATT_SMEM = 2 * TILE_BYTES + KV_STAGES * 2 * TILE_BYTES + 1024 + 512 + 2048;
constexpr int ATT_THREADS = 384;             // general kernel: 8 softmax warps + 4 role warps
constexpr int ATT_THREADS_B = 640;           // BOUNDED kernel: 16 softmax warps (two per query row half) + 4 role warps
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 320, TM_P0 = 384, TM_P1 = 448;
#ifndef K5_ATTN_PROBE_PAIR
#define K5_ATTN_PROBE_PAIR 56
#endif
#ifndef K5_ATTN_PV_PROBE_PAIR
#define K5_ATTN_PV_PROBE_PAIR 12
#endif
// Ping-pong of the two softmax warpgroups (dense BOUNDED kernel): warp w of query tile 0 and warp w of query tile 1 sit
// on the same scheduler and share its MUFU.  Left alone they run in phase - both stream exponentials at half rate,
// then both load / store / synchronise while the MUFU idles (ncu: 16.9 cycles per MUFU inside the stream, 400 cycles
// per tile outside it).  Two named barriers per warp pair hold them half a tile apart: a warp may start the
// exponentials of a tile only after its partner has passed the middle of its own current tile.
#ifndef K5_ATTN_PINGPONG
#define K5_ATTN_PINGPONG 0
#endif
// Suspend-time hints (ns) of the single-thread roles' barrier waits; 0 = plain polling.  The TMA producer runs five
// stages ahead and is not latency critical; the issuers' waits sit on the P -> PV -> pv_done chain.
#ifndef K5_ATTN_PROD_HINT
#define K5_ATTN_PROD_HINT 2000
#endif
#ifndef K5_ATTN_ISS_HINT
#define K5_ATTN_ISS_HINT 0
#endif
// BOUNDED kernel: pair (of 64) after which the first half of P is stored (needs pv_done of the previous tile)
#ifndef K5_ATTN_HALF_AT
#define K5_ATTN_HALF_AT 48
#endif
#ifndef K5_ATTN_SPLITST
#define K5_ATTN_SPLITST 1
#endif
constexpr int PROBE_PAIR = K5_ATTN_PROBE_PAIR;
constexpr int PV_PROBE_PAIR = K5_ATTN_PV_PROBE_PAIR;   // BOUNDED: pair (of 64) after which pv_done of the previous tile is probed
constexpr int PROBE_PAIR16 = 28, PV_PROBE_PAIR16 = 20;   // the same for the two-threads-per-row kernel (32 pairs per thread)
constexpr float RESCALE_THRESHOLD = 8.0f;   // in log2 units: P stays <= 2^8 before a rescale is forced

struct Bars {
    uint64_t q_full[2], q_empty[2];
    uint64_t k_full[KV_STAGES], k_empty[KV_STAGES], v_full[KV_STAGES], v_empty[KV_STAGES];
    uint64_t s_full[2], s_free[2], p_ready[2], pv_done[2], o_free[2];
    uint32_t tmem_slot;
    float half_sum[2][2][128];       // BOUNDED: row sums of the two column halves, exchanged once per item
};

// exp2 of a pair on the FMA / ALU pipes: x = floor(x) + f, 2^f by a degree-3 polynomial on [0,1), the integer
// part goes straight into the exponent field.  Valid for x in [-126, 126]; smaller inputs are clamped (result
// ~1e-38, i.e. zero for the purposes of a softmax).
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
    const float MAGIC = 12582912.0f;                    // 1.5 * 2^23: low mantissa bits hold floor(x)
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const uint64_t x = pack_f32x2(x0, x1);
    const uint64_t xr = add_rm_f32x2(x, pack_f32x2(MAGIC, MAGIC));
    const uint64_t xi = sub_f32x2(xr, pack_f32x2(MAGIC, MAGIC));
    const uint64_t f = sub_f32x2(x, xi);
    uint64_t p = fma_f32x2(f, pack_f32x2(0.077119089663028717f, 0.077119089663028717f),
                           pack_f32x2(0.227564394474029541f, 0.227564394474029541f));
    p = fma_f32x2(p, f, pack_f32x2(0.695146143436431885f, 0.695146143436431885f));
    p = fma_f32x2(p, f, pack_f32x2(1.0f, 1.0f));
    float r0, r1, q0, q1;
    unpack_f32x2(xr, r0, r1);
    unpack_f32x2(p, q0, q1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}

// Same without the clamp, for inputs known to lie in [-126, 126]: the dense BOUNDED kernel (|x| <= score_bound <= 60;
// the columns of a ragged KV tail are masked with the finite score -126 / scale_log2 there, not with -inf).
__device__ __forceinline__ void exp2_poly2_nc(float x0, float x1, float& p0, float& p1) {
    const float MAGIC = 12582912.0f;
    const uint64_t x = pack_f32x2(x0, x1);
    const uint64_t xr = add_rm_f32x2(x, pack_f32x2(MAGIC, MAGIC));
    const uint64_t xi = sub_f32x2(xr, pack_f32x2(MAGIC, MAGIC));
    const uint64_t f = sub_f32x2(x, xi);
    uint64_t p = fma_f32x2(f, pack_f32x2(0.077119089663028717f, 0.077119089663028717f),
                           pack_f32x2(0.227564394474029541f, 0.227564394474029541f));
    p = fma_f32x2(p, f, pack_f32x2(0.695146143436431885f, 0.695146143436431885f));
    p = fma_f32x2(p, f, pack_f32x2(1.0f, 1.0f));
    float r0, r1, q0, q1;
    unpack_f32x2(xr, r0, r1);
    unpack_f32x2(p, q0, q1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}
// Dense BOUNDED kernel: pair q (of the 64 per tile and thread) takes the polynomial when bit (q mod 16) of
// K5_ATTN_POLY_MASK is set (0 = none).  Bits 0, 8 and 12 are left alone: pairs 16 / 32 / 48 gate the s_free arrive, pairs
// 56 and 12 carry the barrier probes.
// Measured (isolated, S = 47 616, two boxes, gpurun_out/r2_poly_nc.log): none 18.31 / 18.35 ms, every 8th 17.96 / 18.06,
// every 6th 17.94 / 17.96, every 5th 18.01 / 18.02, every 4th 17.84 / 17.78 (pair 1 of 4) and 17.90 / 17.81 (pair 3 of 4);
// with the clamp of the general path every 4th gave 17.96 / 18.11 and every 3rd 18.53 / 18.61 (slower than none).
#ifndef K5_ATTN_POLY_MASK
#define K5_ATTN_POLY_MASK 0x2222
#endif

// NPOLY of every 8 element pairs take the polynomial path, the rest the MUFU.
//
// BOUNDED: the caller guarantees |q . k| * scale * log2(e) <= p.score_bound <= 60 for every query / key pair (the DiT
// RMS-normalises q and k per head, nn.py:246-250, so |q| <= 8 max|w_q|, |k| <= 8 max|w_k| and the bound follows from
// the two norm weight vectors alone).  Softmax is shift invariant and bf16 / fp32 carry the same relative precision at
// every magnitude, so with such a bound NO running row max is needed: P = exp2(s * scale*log2e) can neither overflow
// (P <= 2^60, row sums <= 2^77) nor underflow, and O / l is the same number.  What disappears from the softmax
// thread's serial stream: the 64 three-input maxima per tile, the rescale decision, the O correction path (and its
// registers), the per-tile subtraction.  The two waits that used to sit between tiles (s_full of the next scores,
// pv_done of the previous P) are probed without blocking from inside the exponential stream, half a tile before
// their result is needed, so their ~100-cycle round trips overlap MUFU work (profiles/r2_attention_bounded.md).
// W16 (with BOUNDED): 16 softmax warps, two threads per query row (key columns 0-63 / 64-127 of every KV tile).  Without
// a row maximum the halves of a row share nothing per tile.  Dense attention gains nothing from it (the two halves
// wait on the same barriers, so they are not independent streams: 19.1 against 18.7 ms) but block-sparse attention does:
// a warp then owns exactly one 64 x 64 block of the tile and skips all work for a block that is not selected
// (5.5 / 26.3 / 47.6 ms against 6.4 / 30.4 / 57.3 at densities 0.05 / 0.14 / 0.33, profiles/r2_attention.md).
// PAIR (dense only): the kernel runs as thread-block clusters of two CTAs that work on adjacent query items of the
// SAME head, i.e. on the same sequence of K / V tiles.  Each producer fetches one 64-row half of every K and V tile and
// TMA-multicasts it into both CTAs' rings, so a tile crosses the L2 -> SM fabric once per pair instead of once per CTA
// (148 CTAs x 12 MB of K|V per head otherwise).  A ring stage is handed back when the issuers of BOTH CTAs are done with
// it (their commits arrive on both CTAs' barriers).  Everything else - Q, TMEM, the softmax warps - is per CTA as before.
// The kernel sits at the 1 kW power cap, so what the fabric does not burn comes back as clock.
// PART (dense BOUNDED only): the launch is one of several over key slabs of the same attention (AttnParams::part_mode,
// slab_skip).  A template parameter, not a run-time branch: the instantiation without it is the code ptxas scheduled before
// the split existed - run-time branches around the exponential stream changed its interleaving (SASS of the hot loop 25 %
// similar) and cost 3 % in the single-GPU step (19.17 against 18.45 - 18.7 ms per launch).
template <bool SPARSE, int NPOLY, bool BOUNDED, bool W16 = false, bool PAIR = false, bool PART = false>
__global__ void __launch_bounds__(W16 ? ATT_THREADS_B : ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                  // 2 tiles
    uint8_t* sKV = smem + 2 * TILE_BYTES;                // stage s: K at s*2*TILE, V at +TILE
    Bars* B = reinterpret_cast<Bars*>(smem + 2 * TILE_BYTES + KV_STAGES * 2 * TILE_BYTES);

    const int warp = threadIdx.x >> 5;
    // softmax warps 0 .. NSW-1, then producer, issuer of query tile 0, TMEM allocator, issuer of query tile 1
    static_assert(!W16 || BOUNDED, "two threads per row need the fixed-offset softmax");
    static_assert(!PAIR || !SPARSE, "CTA pairs share dense K / V streams only");
    static_assert(!PART || (BOUNDED && !W16 && !SPARSE), "launches split by key slab need the dense fixed-offset kernel");
    [[maybe_unused]] const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    constexpr int NSW = W16 ? 16 : 8;
    constexpr int W_PROD = NSW, W_ISS0 = NSW + 1, W_ALLOC = NSW + 2, W_ISS1 = NSW + 3;
    constexpr uint32_t SM_THREADS = NSW * 16;            // softmax threads per query tile
    const int n_qpairs = (p.Sq + 2 * QT - 1) / (2 * QT);
    const int n_items = n_qpairs * p.heads;
    const int nkv_dense = (p.Sk + KT - 1) / KT;
    const int kv_rem = p.Sk - (nkv_dense - 1) * KT;     // valid kv rows in the last tile (1..128)
    // Static schedule: unit u of this CTA is item blockIdx.x + u * gridDim.x with both query tiles.  The last, partial
    // round leaves most CTAs idle for a whole item; when it holds at most gridDim.x / 2 items (dense mode) each of them
    // is split into its two query tiles and given to two CTAs - rows are independent, so the results do not change,
    // and a CTA running one softmax warpgroup needs ~52 % of the time of a full item.
    const int G = static_cast<int>(gridDim.x);
    const int full_rounds = n_items / G, tail_items = n_items % G;
    const bool split_tail = !SPARSE && p.split_tail && tail_items > 0 && 2 * tail_items <= G;
    const int n_units = full_rounds + ((split_tail ? static_cast<int>(blockIdx.x) < 2 * tail_items
                                                   : static_cast<int>(blockIdx.x) < tail_items) ? 1 : 0);
    auto unit_item = [&](int u, int& mask) -> int {
        if (u < full_rounds || !split_tail) {
            mask = 3;
            return static_cast<int>(blockIdx.x) + u * G;
        }
        mask = 1 << (blockIdx.x & 1);
        return full_rounds * G + (static_cast<int>(blockIdx.x) >> 1);
    };

    if (warp == W_PROD && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == W_ISS0 && elect_one()) {
        for (int a = 0; a < 2; ++a) {
            mbar_init(&B->q_full[a], 1);
            mbar_init(&B->q_empty[a], 1);
            mbar_init(&B->s_full[a], 1);
            mbar_init(&B->s_free[a], SM_THREADS);
            mbar_init(&B->p_ready[a], SM_THREADS);
            mbar_init(&B->pv_done[a], 1);
            mbar_init(&B->o_free[a], SM_THREADS);
        }
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(&B->k_full[s], 1);
            mbar_init(&B->k_empty[s], PAIR ? 4 : 2);  // both issuers (of both CTAs of a pair) commit on a stage before it is refilled
            mbar_init(&B->v_full[s], 1);
            mbar_init(&B->v_empty[s], PAIR ? 4 : 2);
        }
        fence_barrier_init();
    }
    if (warp == W_ALLOC) tmem_alloc<512>(&B->tmem_slot);
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();          // the peer's barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = B->tmem_slot;

    if (warp >= NSW) {
        // register pool: 384 x 168 = 8 warps x 216 + 4 x 72 (general), 640 x 96 = 16 warps x 104 + 4 x 64 (W16);
        // setmaxnreg can only redistribute what the CTA was launched with - asking for more blocks forever
        if constexpr (W16) reg_dec<64>();
        else reg_dec<72>();
        if (warp == W_PROD) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                int st = 0;
                uint32_t ph = 0;
                uint32_t qi[2] = {0, 0};             // Q loads so far per query tile (a tile skips half units of the other)
                for (int u = 0; u < n_units; ++u) {
                    int mask;
                    const int item = unit_item(u, mask);
                    const int h = item / n_qpairs;
                    const int q0 = (item % n_qpairs) * 2 * QT;
                    for (int a = 0; a < 2; ++a) {
                        if (!((mask >> a) & 1)) continue;
                        mbar_wait(&B->q_empty[a], (qi[a] & 1) ^ 1);
                        ++qi[a];
                        mbar_expect_tx(&B->q_full[a], TILE_BYTES);
                        tma_load_2d(sQ + a * TILE_BYTES, &tmQ, &B->q_full[a], h * HD, q0 + a * QT);
                    }
                    const int nkv = SPARSE ? p.item_count[item] : nkv_dense;
                    const int32_t* pairs = SPARSE ? p.item_pairs + static_cast<size_t>(item) * p.max_pairs : nullptr;
                    [[maybe_unused]] int slab = p.slab_first, slab_left = 0, slab_tile = 0;   // dense, overlapped gather
                    if constexpr (!SPARSE) {
                        if (p.n_slabs > 0 && slab < 0) {
                            // debug order of a single engine: every query row starts at the slab that holds it, as the
                            // rank owning that row would (slab boundaries are multiples of 256 rows here)
                            slab = 0;
                            while (slab + 1 < p.n_slabs && p.slab_tile0[slab + 1] * KT <= q0) ++slab;
                        }
                    }
                    [[maybe_unused]] const int slab_own = slab;
                    if constexpr (!SPARSE) {
                        if constexpr (PART) {
                            if (p.n_slabs > 0)   // slabs of the rotated order that earlier launches of this attention consumed
                                for (int sk = 0; sk < p.slab_skip; ++sk) slab = slab + 1 == p.n_slabs ? 0 : slab + 1;
                        }
                    }
                    for (int j = 0; j < nkv; ++j) {
                        int tile = SPARSE ? pairs[j] : j;
                        if constexpr (!SPARSE) {
                            if (p.n_slabs > 0) {
                                // own slab first, then the slabs in the order their owners push them (rank+1, rank+2, ...)
                                if (slab_left == 0) {
                                    if (j > 0) slab = slab + 1 == p.n_slabs ? 0 : slab + 1;
                                    slab_tile = p.slab_tile0[slab];
                                    slab_left = p.slab_tile0[slab + 1] - slab_tile;
                                    if (p.slab_flags && slab != slab_own) {
                                        unsigned long long t0;
                                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                                        for (uint32_t spins = 0;; ++spins) {
                                            uint32_t v;
                                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.slab_flags + slab) : "memory");
                                            if (static_cast<int32_t>(v - p.slab_epoch) >= 0) break;
                                            if ((spins & 1023u) == 1023u) {       // a peer died or fell far behind: report, do not trap
                                                unsigned long long t1;
                                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                                                if (*p.slab_err != 0u) break;
                                                if (t1 - t0 > p.slab_timeout_ns) {
                                                    *p.slab_err = K5_DIST_ERR_SLAB;
                                                    __threadfence_system();
                                                    break;
                                                }
                                            }
                                        }
                                        asm volatile("fence.proxy.async;" ::: "memory");     // the TMA reads what a peer's copy wrote
                                    }
                                }
                                tile = slab_tile++;
                                --slab_left;
                            }
                        }
                        const int kv0 = tile * KT;
                        uint8_t* sk = sKV + st * 2 * TILE_BYTES;
                        mbar_wait_lean_a<K5_ATTN_PROD_HINT>(smem_u32(&B->k_empty[st]), ph ^ 1);
                        mbar_expect_tx(&B->k_full[st], TILE_BYTES);
                        if constexpr (PAIR) {
                            // this CTA's 64-row half (tmK / tmV have 64-row boxes here) goes to both rings; the other
                            // half arrives from the peer's producer on the same barrier
                            tma_load_2d_mc(sk + crank * (TILE_BYTES / 2), &tmK, &B->k_full[st], h * HD, kv0 + crank * (KT / 2), 0x3);
                        } else {
                            tma_load_2d(sk, &tmK, &B->k_full[st], h * HD, kv0);
                        }
                        mbar_wait_lean_a<K5_ATTN_PROD_HINT>(smem_u32(&B->v_empty[st]), ph ^ 1);
                        mbar_expect_tx(&B->v_full[st], TILE_BYTES);
                        if constexpr (PAIR) {
                            tma_load_2d_mc(sk + TILE_BYTES + crank * (TILE_BYTES / 2), &tmV, &B->v_full[st], h * HD,
                                           kv0 + crank * (KT / 2), 0x3);
                        } else {
                            tma_load_2d(sk + TILE_BYTES, &tmV, &B->v_full[st], h * HD, kv0);
                        }
                        if (++st == KV_STAGES) {
                            st = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        } else if (warp == W_ISS0 || warp == W_ISS1) {
            // ===================== MMA issuers: one warp per query tile =====================
            // Per query tile a the order of events is fixed:  S_a(g) pulled into registers (s_free) -> QK_a(g+1);
            // P_a(g) written (p_ready) -> PV_a(g); so each issuer simply blocks on the next event of its own tile.
            // The two issuers share only the K / V ring: a stage is released when BOTH have committed their MMA on
            // it (k_empty / v_empty are initialised with count 2).
            if (elect_one()) {
                const int a = warp == W_ISS0 ? 0 : 1;
                constexpr uint32_t idesc_qk = umma_idesc_bf16(QT, KT, 0, 0);
                constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);
                const uint32_t tS = tmem_base + (a == 0 ? TM_S0 : TM_S1);
                const uint32_t tO = tmem_base + (a == 0 ? TM_O0 : TM_O1);
                const uint32_t tP = tmem_base + (a == 0 ? TM_P0 : TM_P1);
                const uint32_t qa = smem_u32(sQ) + a * TILE_BYTES;
                const uint32_t skv_addr = smem_u32(sKV);
                int kst = 0, vst = 0;
                uint32_t kph = 0, vph = 0;
                uint32_t g = 0;                     // KV tiles handled so far (over all items)
                int i = 0;
                [[maybe_unused]] uint32_t sq = 0, sw = 0;   // block-sparse: QKs issued / s_free phases consumed
                [[maybe_unused]] int io = 0;                // block-sparse: units in which this query tile had any KV tile

                // Shared-memory descriptors advance by a constant per K step (the start-address field holds addr >> 4 and
                // never carries out of its 14 bits), so one descriptor per operand tile is built BEFORE the wait that
                // gates the MMA and the k-th step only adds k * stride: what is left between the barrier flipping and
                // the last tcgen05.mma is a fence and ~2 instructions per MMA (measured before: ~450 cycles from
                // p_ready to the PV commit, tools/attn_trace.py).
                const uint64_t qdesc0 = umma_desc_sw128(qa, 0, 1024);
#ifdef K5_ATTN_ISS_SLEEP
                auto wait_iss = [&](uint64_t* bar, uint32_t parity) { mbar_wait_sleep_a<K5_ATTN_ISS_SLEEP>(smem_u32(bar), parity); };
#else
                auto wait_iss = [&](uint64_t* bar, uint32_t parity) { mbar_wait_lean_a<K5_ATTN_ISS_HINT>(smem_u32(bar), parity); };
#endif
                auto issue_qk = [&](bool last_of_item, uint64_t* gate, uint32_t gate_parity) {
                    wait_iss(&B->k_full[kst], kph);
                    const uint64_t kdesc0 = umma_desc_sw128(skv_addr + kst * 2 * TILE_BYTES, 0, 1024);
                    if (gate) wait_iss(gate, gate_parity);           // S_a free again (the scores of the tile before are in registers)
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) umma_ss(tS, qdesc0 + 2 * k, kdesc0 + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
                    umma_commit(&B->s_full[a]);
                    if constexpr (PAIR) umma_commit_mc(&B->k_empty[kst], 0x3);
                    else umma_commit(&B->k_empty[kst]);
                    if (last_of_item) umma_commit(&B->q_empty[a]);
                    if (++kst == KV_STAGES) {
                        kst = 0;
                        kph ^= 1;
                    }
                };
                for (int u = 0; u < n_units; ++u) {
                    int mask;
                    const int item = unit_item(u, mask);
                    const int nkv = SPARSE ? p.item_count[item] : nkv_dense;
                    if (!((mask >> a) & 1)) {
                        // half unit of the other query tile: nothing to multiply, but the K / V ring expects this
                        // issuer's release of every stage
                        for (int j = 0; j < nkv; ++j) {
                            mbar_wait(&B->k_full[kst], kph);
                            mbar_arrive(&B->k_empty[kst]);
                            if constexpr (PAIR) mbar_arrive_cluster(&B->k_empty[kst], crank ^ 1u);
                            if (++kst == KV_STAGES) {
                                kst = 0;
                                kph ^= 1;
                            }
                            mbar_wait(&B->v_full[vst], vph);
                            mbar_arrive(&B->v_empty[vst]);
                            if constexpr (PAIR) mbar_arrive_cluster(&B->v_empty[vst], crank ^ 1u);
                            if (++vst == KV_STAGES) {
                                vst = 0;
                                vph ^= 1;
                            }
                        }
                        continue;
                    }
                    if constexpr (SPARSE) {
                        // Block-sparse: a KV tile none of this query tile's two 64-row blocks selected is skipped
                        // altogether - the issuer only hands its ring stages back, the softmax warpgroup does not take
                        // part - so the tiles of the item's UNION are not all multiplied for both query tiles.  The
                        // K cursor stays one tile ahead of the V cursor exactly as in the dense loop below.
                        const uint8_t* masks = p.item_mask + static_cast<size_t>(item) * p.max_pairs;
                        auto act = [&](int j) { return ((masks[j] >> (4 * a)) & 0xFu) != 0u; };
                        mbar_wait(&B->q_full[a], i & 1);
                        bool first = true;
                        auto step_k = [&](int j) {
                            if (act(j)) {
                                if (sq > sw) {                  // the S buffer is free once the last scores were pulled
                                    mbar_wait(&B->s_free[a], sw & 1);
                                    ++sw;
                                }
                                issue_qk(false, nullptr, 0);
                                ++sq;
                            } else {
                                mbar_wait(&B->k_full[kst], kph);
                                mbar_arrive(&B->k_empty[kst]);
                                if (++kst == KV_STAGES) {
                                    kst = 0;
                                    kph ^= 1;
                                }
                            }
                        };
                        step_k(0);
                        if (nkv == 1) umma_commit(&B->q_empty[a]);      // all QKs issued: Q_a may be replaced when they are done
                        for (int j = 0; j < nkv; ++j) {
                            if (j + 1 < nkv) {
                                step_k(j + 1);
                                if (j + 2 == nkv) umma_commit(&B->q_empty[a]);
                            }
                            if (act(j)) {
                                if (first) mbar_wait(&B->o_free[a], (io & 1) ^ 1);
                                mbar_wait(&B->p_ready[a], g & 1);
                                mbar_wait(&B->v_full[vst], vph);
                                tc_fence_after();
                                const uint32_t va = skv_addr + vst * 2 * TILE_BYTES + TILE_BYTES;
#pragma unroll
                                for (int k = 0; k < KT / 16; ++k)
                                    umma_ts(tO, tP + k * 8, umma_desc_sw128(va + k * 2048, 16384, 1024), idesc_pv,
                                            (!first || k != 0) ? 1u : 0u);
                                umma_commit(&B->pv_done[a]);
                                umma_commit(&B->v_empty[vst]);
                                ++g;
                                first = false;
                            } else {
                                mbar_wait(&B->v_full[vst], vph);
                                mbar_arrive(&B->v_empty[vst]);
                            }
                            if (++vst == KV_STAGES) {
                                vst = 0;
                                vph ^= 1;
                            }
                        }
                        if (!first) ++io;
                        ++i;
                        continue;
                    }
                    // S_a(0) = Q_a K_0^T: needs Q_a and the S buffer (drained at the last tile of the previous item)
                    wait_iss(&B->q_full[a], i & 1);
                    issue_qk(nkv == 1, g > 0 ? &B->s_free[a] : nullptr, (g - 1) & 1);
                    for (int j = 0; j < nkv; ++j, ++g) {
                        if (j + 1 < nkv) issue_qk(j + 2 == nkv, &B->s_free[a], g & 1);
                        if (j == 0) wait_iss(&B->o_free[a], (i & 1) ^ 1);
                        wait_iss(&B->v_full[vst], vph);              // the V tile landed long ago: not on the critical path
                        const uint64_t vdesc0 = umma_desc_sw128(skv_addr + vst * 2 * TILE_BYTES + TILE_BYTES, 16384, 1024);
                        wait_iss(&B->p_ready[a], g & 1);
                        K5_TRACE_ISSUER(4);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < KT / 16; ++k)
                            umma_ts(tO, tP + k * 8, vdesc0 + 128 * k, idesc_pv,
                                    (j != 0 || k != 0 || (PART && (p.part_mode & 2))) ? 1u : 0u);
                        umma_commit(&B->pv_done[a]);
                        if constexpr (PAIR) umma_commit_mc(&B->v_empty[vst], 0x3);
                        else umma_commit(&B->v_empty[vst]);
                        K5_TRACE_ISSUER(5);
                        if (++vst == KV_STAGES) {
                            vst = 0;
                            vph ^= 1;
                        }
                    }
                    ++i;                             // units this query tile took part in (phases of q_full / o_free)
                }
            }
        }
    } else {
        // ===================== softmax warpgroups =====================
        if constexpr (W16) reg_inc<104>();
        else reg_inc<216>();
        // general kernel: warps 0-3 / 4-7 own query tile 0 / 1, a thread owns a whole 128-column score row.
        // BOUNDED kernel: warps 0-7 / 8-15 own query tile 0 / 1; within a tile warps 0-3 take key columns 0-63 of every
        // KV tile and warps 4-7 columns 64-127 (two threads per query row: no row maximum means nothing to exchange per
        // tile), so each scheduler interleaves FOUR independent exponential streams instead of two.
        const int a = W16 ? warp >> 3 : warp >> 2;        // query tile of this warpgroup
        [[maybe_unused]] const int hf = W16 ? (warp >> 2) & 1 : 0;       // key column half
        const int wq = warp & 3;
        const int lane = threadIdx.x & 31;
        const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t tS = tmem_base + (a == 0 ? TM_S0 : TM_S1) + lane_off;
        const uint32_t tO = tmem_base + (a == 0 ? TM_O0 : TM_O1) + lane_off;
        const uint32_t tP = tmem_base + (a == 0 ? TM_P0 : TM_P1) + lane_off;
        const float sl2 = p.scale_log2;
        const uint64_t sl2x2 = pack_f32x2(sl2, sl2);
        uint32_t cnt = 0;
        [[maybe_unused]] uint32_t sf_ok = 0;             // BOUNDED: s_full of the next tile was already seen complete
        [[maybe_unused]] const uint32_t bar_s_full = smem_u32(&B->s_full[a]), bar_s_free = smem_u32(&B->s_free[a]),
                                        bar_p_ready = smem_u32(&B->p_ready[a]), bar_pv_done = smem_u32(&B->pv_done[a]);
        if (a == 1 && p.stagger > 0) {
            // The two warpgroups share the MUFU.  Started together they stay in lock-step (exp phases collide, the
            // MUFU idles while both load / reduce / store); started half a tile apart they interleave and stay so.
            const long long t0 = clock64();
            while (clock64() - t0 < p.stagger) {
            }
        }
        for (int u = 0; u < n_units; ++u) {
            int mask;
            const int item = unit_item(u, mask);
            if (!((mask >> a) & 1)) continue;             // half unit of the other query tile
            const int h = item / n_qpairs;
            const int row = (item % n_qpairs) * 2 * QT + a * QT + wq * 32 + lane;
            float l = 0.f;
            const int nkv = SPARSE ? p.item_count[item] : nkv_dense;
            const uint8_t* masks = SPARSE ? p.item_mask + static_cast<size_t>(item) * p.max_pairs : nullptr;
            const int qblk2 = (a * 2 + (wq >> 1)) * 2;     // bit position of this warp's 64-row query block
            int done = 0;                                  // KV tiles this query tile has taken part in (= j when dense)
            [[maybe_unused]] float l4[4] = {0.f, 0.f, 0.f, 0.f};      // the four partial row sums (part_mode 1)
            if constexpr (BOUNDED && !W16) {
                uint64_t sum_a = pack_f32x2(0.f, 0.f), sum_b = pack_f32x2(0.f, 0.f);   // row sum, carried over the item
                if constexpr (PART) {
                    if (p.part_mode & 2) {
                        // continue where the launch over the previous slabs stopped: its fp32 accumulators go back into TMEM
                        // (this thread read O of the previous item itself, so the columns are free), the row sums go on
                        uint32_t o0[64];
                        if (row < p.Sq) {
                            const uint4* src = reinterpret_cast<const uint4*>(p.part_o + (static_cast<size_t>(row) * p.heads + h) * HD);
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const uint4 v = src[c];          // (plain load: a middle launch writes these rows again at the end of the item)
                                o0[4 * c] = v.x; o0[4 * c + 1] = v.y; o0[4 * c + 2] = v.z; o0[4 * c + 3] = v.w;
                            }
                            const float4 lp = reinterpret_cast<const float4*>(p.part_l)[static_cast<size_t>(row) * p.heads + h];
                            sum_a = pack_f32x2(lp.x, lp.y);
                            sum_b = pack_f32x2(lp.z, lp.w);
                        } else {
#pragma unroll
                            for (int c = 0; c < 64; ++c) o0[c] = 0u;
                        }
                        tmem_st32(tO, o0);
                        tmem_st32(tO + 32, o0 + 32);
                        tmem_wait_st();
                        tc_fence_before();
                    }
                }
                for (int j = 0; j < nkv; ++j) {
                    bool actL = true, actR = true;
                    if constexpr (SPARSE) {
                        const uint32_t mb = masks[j];
                        if (((mb >> (4 * a)) & 0xFu) == 0u) continue;   // not selected by either block of this query tile
                        actL = (mb >> qblk2) & 1u;
                        actR = (mb >> (qblk2 + 1)) & 1u;
                    }
                    K5_TRACE(0);
                    if (!sf_ok) mbar_wait_a(bar_s_full, cnt & 1);
                    sf_ok = 0;
                    tc_fence_after();
                    if (SPARSE && !actL && !actR) {
                        // nothing selected for this warp's query block in this KV tile: P = 0, S is not needed
                        mbar_arrive_a(bar_s_free);
                        if (done > 0) {
                            mbar_wait_a(bar_pv_done, (cnt - 1) & 1);
                            tc_fence_after();
                        }
                        uint32_t z[32];
#pragma unroll
                        for (int c = 0; c < 32; ++c) z[c] = 0u;
                        tmem_st32(tP + 0, z);
                        tmem_st32(tP + 32, z);
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive_a(bar_p_ready);
                        ++cnt;
                        ++done;
                        continue;
                    }
                    // One tcgen05.ld.x32 moves 4 KB per warp at ~64 B/clk: a whole 128-column score row costs ~290 cycles
                    // before the first exponential can issue (tests/micro/pipe_rates.cu).  So only the first 32 columns
                    // are waited for; the other three loads stay in flight under the first 16 exponential pairs.
                    uint32_t s[128];
                    tmem_ld32(tS + 0, s);
                    tmem_wait_ld_regs(s);
                    K5_TRACE(1);
                    tmem_ld32(tS + 32, s + 32);
                    tmem_ld32(tS + 64, s + 64);
                    tmem_ld32(tS + 96, s + 96);
                    uint32_t pv_ok = done > 0 ? 0u : 1u;
                    const bool tail = !SPARSE && j == nkv - 1 && kv_rem < KT;
                    auto mask_cols = [&](int c0, int c1) {
                        if constexpr (SPARSE) {
                            if (!actL) {
#pragma unroll
                                for (int c = c0; c < (c1 < 64 ? c1 : 64); ++c) s[c] = 0xff800000u;
                            }
                            if (!actR) {
#pragma unroll
                                for (int c = (c0 > 64 ? c0 : 64); c < c1; ++c) s[c] = 0xff800000u;
                            }
                        } else if (tail) {
                            // a finite "minus infinity": exp2 gives 2^-126 (P rounds to ~1e-38, times the zero-filled V rows
                            // = 0; the row sum, >= 2^-60, does not see it), and the polynomial path needs no clamp
                            const uint32_t masked = __float_as_uint(-126.0f / sl2);
#pragma unroll
                            for (int c = c0; c < c1; ++c)
                                if (c >= kv_rem) s[c] = masked;
                        }
                    };
                    uint32_t dep = 0, p48 = 0;
                    auto pairs = [&](int q0, int q1) {
#pragma unroll
                        for (int q = q0; q < q1; ++q) {
                            float x0, x1, p0, p1;
                            unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(s[2 * q]), __uint_as_float(s[2 * q + 1])), sl2x2),
                                         x0, x1);
                            if ((q & 7) < NPOLY) {
                                exp2_poly2(x0, x1, p0, p1);
                            } else if (!SPARSE && ((K5_ATTN_POLY_MASK >> (q & 15)) & 1)) {
                                exp2_poly2_nc(x0, x1, p0, p1);
                            } else {
                                p0 = fast_exp2(x0);
                                p1 = fast_exp2(x1);
                            }
                            const uint64_t pr = pack_f32x2(p0, p1);
                            if (q & 1) sum_b = add_f32x2(sum_b, pr);
                            else sum_a = add_f32x2(sum_a, pr);
                            // P is packed in place: pair q overwrites s[q], an input of pair q / 2, so a pair may only run after
                            // pair q / 2.  Pair 48 runs early (see below) and parks its result until pair 24 is done.
                            if (q == 48) p48 = pack_bf16x2(p0, p1);
                            else s[q] = pack_bf16x2(p0, p1);
                            if (q == 16 || q == 32 || q == 48) dep |= __float_as_uint(p0);
                            // the next scores are usually there by now; the parity is made to depend on this pair's result
                            // (its sign bit, always 0) so that ptxas cannot hoist the probe to the head of the stream
                            if (q == PROBE_PAIR)
                                sf_ok = mbar_test_wait_a(bar_s_full, ((cnt + 1) & 1) + (__float_as_uint(p0) >> 31));
                            // PV_a of the previous tile (issued when the previous P was published, ~300-800 cycles ago by
                            // now) must have consumed that P before it is overwritten: probed here, pinned the same way
                            if (q == PV_PROBE_PAIR && done > 0)
                                pv_ok = mbar_test_wait_a(bar_pv_done, ((cnt - 1) & 1) + (__float_as_uint(p0) >> 31));
                        }
                    };
                    mask_cols(0, 32);
#if K5_ATTN_PINGPONG
                    if (!SPARSE && mask == 3) {
                        if (a == 1) named_bar_sync(1 + wq, 64);              // partner (tile 0) is past the middle of tile j
                        else if (j > 0) named_bar_sync(5 + wq, 64);          // partner (tile 1) is past the middle of tile j-1
                    }
#endif
                    pairs(0, 9);
                    tmem_wait_ld_regs(s + 32);
                    reg_fence32(s + 64);
                    reg_fence32(s + 96);
                    mask_cols(32, 128);
                    // ptxas realises tcgen05.wait::ld only through the scoreboards of the loaded registers and was seen to
                    // hoist the s_free arrive above it (SASS of this kernel, tools/sass_sched.py).  The arrive must not
                    // happen before the last load has read S, so its address is made to depend on one exponential from
                    // each of the three late chunks (sign bit of a positive number: always 0).  It comes as early as the
                    // loads allow (9 pairs ~ 200 cycles into the stream): the sooner S is handed back, the sooner
                    // QK(j+1) is issued and the likelier the s_full probe below finds the next scores in place.
                    pairs(16, 17);
                    pairs(32, 33);
                    pairs(48, 49);
                    tc_fence_before();
                    mbar_arrive_a(bar_s_free + (dep >> 31));   // the tensor pipe may overwrite S_a with the next scores
                    pairs(9, 16);
                    pairs(17, 32);
                    s[48] = p48;
                    if (K5_ATTN_HALF_AT > 33) pairs(33, K5_ATTN_HALF_AT < 48 ? K5_ATTN_HALF_AT : 48);
                    if (K5_ATTN_HALF_AT > 49) pairs(49, K5_ATTN_HALF_AT);
                    K5_TRACE(2);
#if K5_ATTN_PINGPONG
                    if (!SPARSE && mask == 3) {
                        // barrier id pinned behind pair 31 (sign bit of a positive number) so that the arrive stays in the middle
                        const uint32_t pin = s[31] >> 31;           // packed bf16 pair of positive numbers: 0
                        if (a == 0) named_bar_arrive(1 + wq + pin, 64);
                        else if (j + 1 < nkv) named_bar_arrive(5 + wq + pin, 64);
                    }
#endif
#if K5_ATTN_SPLITST
                    // the first half of P (64 key columns) goes to TMEM under the second half's exponentials
                    K5_TRACE_V(6, static_cast<long long>(pv_ok));
                    if (!pv_ok) mbar_wait_a(bar_pv_done, (cnt - 1) & 1);
                    pv_ok = 1u;
                    tc_fence_after();
                    tmem_st32(tP + 0, s);
                    K5_TRACE(7);
#endif
                    if (K5_ATTN_HALF_AT < 48) pairs(K5_ATTN_HALF_AT > 33 ? K5_ATTN_HALF_AT : 33, 48);
                    pairs(K5_ATTN_HALF_AT > 49 ? K5_ATTN_HALF_AT : 49, 64);
                    if (!pv_ok) mbar_wait_a(bar_pv_done, (cnt - 1) & 1);
                    tc_fence_after();
#if !K5_ATTN_SPLITST
                    tmem_st32(tP + 0, s);
#endif
                    tmem_st32(tP + 32, s + 32);
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive_a(bar_p_ready);
                    K5_TRACE(3);
                    ++cnt;
                    ++done;
                }
                float t0, t1;
                unpack_f32x2(add_f32x2(sum_a, sum_b), t0, t1);
                l = t0 + t1;
                if constexpr (PART) {
                    unpack_f32x2(sum_a, l4[0], l4[1]);
                    unpack_f32x2(sum_b, l4[2], l4[3]);
                }
            } else if constexpr (BOUNDED) {
                uint64_t sum_a = pack_f32x2(0.f, 0.f), sum_b = pack_f32x2(0.f, 0.f);   // half-row sum, carried over the item
                const uint32_t tSh = tS + 64 * hf;                  // this thread's 64 score columns of every KV tile
                const uint32_t tPh = tP + 32 * hf;                  // ... and their 32 packed bf16 P columns
                for (int j = 0; j < nkv; ++j) {
                    bool act = true;
                    if constexpr (SPARSE) {
                        const uint32_t mb = masks[j];
                        if (((mb >> (4 * a)) & 0xFu) == 0u) continue;   // not selected by either block of this query tile
                        act = (mb >> (qblk2 + hf)) & 1u;            // this warp's 64 x 64 block of the tile
                    }
                    K5_TRACE(0);
                    if (!sf_ok) mbar_wait_a(bar_s_full, cnt & 1);
                    sf_ok = 0;
                    tc_fence_after();
                    if (SPARSE && !act) {
                        // block not selected: P = 0 for it, S is not needed
                        mbar_arrive_a(bar_s_free);
                        if (done > 0) {
                            mbar_wait_a(bar_pv_done, (cnt - 1) & 1);
                            tc_fence_after();
                        }
                        uint32_t z[32];
#pragma unroll
                        for (int c = 0; c < 32; ++c) z[c] = 0u;
                        tmem_st32(tPh, z);
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive_a(bar_p_ready);
                        ++cnt;
                        ++done;
                        continue;
                    }
                    // One tcgen05.ld.x32 moves 4 KB per warp at ~64 B/clk (tests/micro/pipe_rates.cu): only the first 32
                    // columns are waited for, the second load stays in flight under the first exponentials.
                    uint32_t s[64];
                    tmem_ld32(tSh + 0, s);
                    tmem_wait_ld_regs(s);
                    K5_TRACE(1);
                    tmem_ld32(tSh + 32, s + 32);
                    uint32_t pv_ok = done > 0 ? 0u : 1u;
                    const bool tail = !SPARSE && j == nkv - 1 && kv_rem < KT;
                    auto mask_cols = [&](int c0, int c1) {
                        if (tail) {
#pragma unroll
                            for (int c = c0; c < c1; ++c)
                                if (64 * hf + c >= kv_rem) s[c] = 0xff800000u;      // -inf
                        }
                    };
                    uint32_t dep = 0;
                    auto pairs = [&](int q0, int q1) {
#pragma unroll
                        for (int q = q0; q < q1; ++q) {
                            float x0, x1, p0, p1;
                            unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(s[2 * q]), __uint_as_float(s[2 * q + 1])), sl2x2),
                                         x0, x1);
#ifdef K5_ATTN_NOEXP     // timing experiment only (wrong results): the softmax stream without its exponentials
                            p0 = x0 * 0.5f;
                            p1 = x1 * 0.5f;
#else
                            if ((q & 7) < NPOLY) {
                                exp2_poly2(x0, x1, p0, p1);
                            } else {
                                p0 = fast_exp2(x0);
                                p1 = fast_exp2(x1);
                            }
#endif
                            const uint64_t pr = pack_f32x2(p0, p1);
                            if (q & 1) sum_b = add_f32x2(sum_b, pr);
                            else sum_a = add_f32x2(sum_a, pr);
                            // P is packed in place: pair q overwrites s[q], an input of pair q / 2, so a pair may only run after
                            // pair q / 2 (the order below: 0..8, 16, 9..15, 17..31).
                            s[q] = pack_bf16x2(p0, p1);
                            if (q == 16) dep = __float_as_uint(p0);
                            // Probes are pinned behind a pair's result (its sign bit, always 0) so that ptxas cannot hoist them
                            // to the head of the stream.  s_full: the next scores, usually there by now.  pv_done: PV_a of the
                            // previous tile must have consumed the previous P before it is overwritten.
                            if (q == PROBE_PAIR16)
                                sf_ok = mbar_test_wait_a(bar_s_full, ((cnt + 1) & 1) + (__float_as_uint(p0) >> 31));
                            if (q == PV_PROBE_PAIR16 && done > 0)
                                pv_ok = mbar_test_wait_a(bar_pv_done, ((cnt - 1) & 1) + (__float_as_uint(p0) >> 31));
                        }
                    };
                    mask_cols(0, 32);
                    pairs(0, 9);
                    tmem_wait_ld_regs(s + 32);
                    mask_cols(32, 64);
                    // ptxas realises tcgen05.wait::ld only through the scoreboards of the loaded registers and was seen to
                    // hoist the s_free arrive above it (SASS of this kernel, tools/sass_sched.py).  The arrive must not
                    // happen before the last load has read S, so its address is made to depend on an exponential of the
                    // late chunk (sign bit of a positive number: always 0).  It comes as early as the loads allow: the
                    // sooner S is handed back, the sooner QK(j+1) is issued.
                    pairs(16, 17);
                    tc_fence_before();
                    mbar_arrive_a(bar_s_free + (dep >> 31));   // the tensor pipe may overwrite S_a with the next scores
                    pairs(9, 16);
                    pairs(17, 32);
                    K5_TRACE(2);
                    K5_TRACE_V(6, static_cast<long long>(pv_ok));
                    if (!pv_ok) mbar_wait_a(bar_pv_done, (cnt - 1) & 1);
                    tc_fence_after();
                    tmem_st32(tPh, s);
                    K5_TRACE(7);
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive_a(bar_p_ready);
                    K5_TRACE(3);
                    ++cnt;
                    ++done;
                }
                float t0, t1;
                unpack_f32x2(add_f32x2(sum_a, sum_b), t0, t1);
                l = t0 + t1;
            } else {
            float m_used = -INFINITY;
            bool started = false;
            for (int j = 0; j < nkv; ++j) {
                bool actL = true, actR = true;
                if constexpr (SPARSE) {
                    const uint32_t mb = masks[j];
                    if (((mb >> (4 * a)) & 0xFu) == 0u) continue;   // not selected by either block of this query tile
                    actL = (mb >> qblk2) & 1u;
                    actR = (mb >> (qblk2 + 1)) & 1u;
                }
                mbar_wait(&B->s_full[a], cnt & 1);
                tc_fence_after();
                if (SPARSE && !actL && !actR) {
                    // nothing selected for this warp's query block in this KV tile: P = 0, S is not needed
                    mbar_arrive(&B->s_free[a]);
                    if (done > 0) {
                        mbar_wait(&B->pv_done[a], (cnt - 1) & 1);
                        tc_fence_after();
                    }
                    uint32_t z[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) z[c] = 0u;
                    tmem_st32(tP + 0, z);
                    tmem_st32(tP + 32, z);
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(&B->p_ready[a]);
                    ++cnt;
                    ++done;
                    continue;
                }
                uint32_t s[128];
                tmem_ld32(tS + 0, s);
                tmem_ld32(tS + 32, s + 32);
                tmem_ld32(tS + 64, s + 64);
                tmem_ld32(tS + 96, s + 96);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&B->s_free[a]);              // the tensor pipe may overwrite S_a with the next scores
                if constexpr (SPARSE) {
                    if (!actL) {
#pragma unroll
                        for (int c = 0; c < 64; ++c) s[c] = 0xff800000u;
                    }
                    if (!actR) {
#pragma unroll
                        for (int c = 64; c < 128; ++c) s[c] = 0xff800000u;
                    }
                } else if (j == nkv - 1 && kv_rem < KT) {
#pragma unroll
                    for (int c = 0; c < 128; ++c)
                        if (c >= kv_rem) s[c] = 0xff800000u;      // -inf
                }
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int c = 0; c < 128; c += 8) {
                    mx0 = max3(mx0, __uint_as_float(s[c]), __uint_as_float(s[c + 1]));
                    mx1 = max3(mx1, __uint_as_float(s[c + 2]), __uint_as_float(s[c + 3]));
                    mx2 = max3(mx2, __uint_as_float(s[c + 4]), __uint_as_float(s[c + 5]));
                    mx3 = max3(mx3, __uint_as_float(s[c + 6]), __uint_as_float(s[c + 7]));
                }
                const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                float alpha = 1.0f;
                bool need = false;
                if (!started) {
                    m_used = mx;
                    started = true;
                } else if ((mx - m_used) * sl2 > RESCALE_THRESHOLD) {
                    alpha = fast_exp2((m_used - mx) * sl2);
                    m_used = mx;
                    need = true;
                }
                const float mneg = -m_used * sl2;
                const uint64_t mnegx2 = pack_f32x2(mneg, mneg);
                uint64_t sum_a = pack_f32x2(0.f, 0.f), sum_b = pack_f32x2(0.f, 0.f);
                // Software-pipelined over chunks of 8 pairs: while chunk c goes through the MUFU / polynomial, the
                // scaled arguments of chunk c+1 are formed (FFMA2) and the results of chunk c-1 are summed and packed,
                // so no instruction waits on a MUFU result that was issued less than ~16 instructions earlier.
#if K5_ATTN_PAD > 0
                float pad_chain = 1.0f;
#endif
                uint64_t xn[8];                          // x = s * scale*log2(e) - m * scale*log2(e) of the next chunk
                float pc[16], pp[16];                    // exp2 of the current / previous chunk
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    xn[q] = fma_f32x2_v(pack_f32x2(__uint_as_float(s[2 * q]), __uint_as_float(s[2 * q + 1])), sl2x2, mnegx2);
#pragma unroll
                for (int c = 0; c < 9; ++c) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (c < 8) {
                            float x0, x1;
                            unpack_f32x2(xn[q], x0, x1);
                            if (q < NPOLY) {
                                exp2_poly2(x0, x1, pc[2 * q], pc[2 * q + 1]);
                            } else {
                                pc[2 * q] = ex2_v(x0);
                                pc[2 * q + 1] = ex2_v(x1);
#if K5_ATTN_PAD > 0
                                // Spacers: a dependent FMA chain that does no useful work.  The denser the MUFU instructions in a
                                // warp's stream, the fuller the MIO queue they share with tcgen05.ld/st and the mbarrier
                                // traffic of the scheduler's other warps; removing instructions from this loop makes the
                                // kernel SLOWER, two spacers per pair make it 1.2 % faster isolated and 0.6 % in the step.
#pragma unroll
                                for (int z = 0; z < K5_ATTN_PAD; ++z)
                                    asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f0DA24260;" : "+f"(pad_chain));
#endif
                            }
                            if (c < 7) {
                                const int e = 16 * (c + 1) + 2 * q;
                                xn[q] = fma_f32x2_v(pack_f32x2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), sl2x2,
                                                    mnegx2);
                            }
                        }
                        if (c > 0) {
                            const uint64_t pr = pack_f32x2(pp[2 * q], pp[2 * q + 1]);
                            if (q & 1) sum_b = add_f32x2_v(sum_b, pr);
                            else sum_a = add_f32x2_v(sum_a, pr);
                            s[8 * (c - 1) + q] = pack_bf16x2_v(pp[2 * q], pp[2 * q + 1]);   // P packed in place
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) pp[q] = pc[q];
                }
                {
                    float t0, t1;
                    unpack_f32x2(add_f32x2(sum_a, sum_b), t0, t1);
                    l = l * alpha + (t0 + t1);
                }
                if (done > 0) {
                    // PV_a of the previous tile must have consumed the previous P (and produced the O this thread may rescale)
                    mbar_wait(&B->pv_done[a], (cnt - 1) & 1);
                    tc_fence_after();
                }
                tmem_st32(tP + 0, s);
                tmem_st32(tP + 32, s + 32);
                if (done > 0 && __any_sync(0xffffffffu, need)) {
                    uint32_t o[64];
                    tmem_ld32(tO, o);
                    tmem_ld32(tO + 32, o + 32);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 64; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
                    tmem_st32(tO, o);
                    tmem_st32(tO + 32, o + 32);
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&B->p_ready[a]);
                ++cnt;
                ++done;
            }
            }
            // ---- epilogue: O_a / l -> bf16 -> global
            if constexpr (W16) {
                // two threads per row: exchange the half-row sums through shared memory (once per item), then each
                // thread normalises and stores its 32 of the 64 output columns
                const int r = wq * 32 + lane;
                B->half_sum[a][hf][r] = l;
                named_bar_sync(1 + a, SM_THREADS);
                l += B->half_sum[a][hf ^ 1][r];
                uint32_t o[32];
                if (!SPARSE || done > 0) {
                    mbar_wait(&B->pv_done[a], (cnt - 1) & 1);
                    tc_fence_after();
                    tmem_ld32(tO + 32 * hf, o);
                    tmem_wait_ld();
                    tc_fence_before();
                    mbar_arrive(&B->o_free[a]);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) o[c] = 0u;
                    l = 0.f;
                }
                named_bar_sync(1 + a, SM_THREADS);               // half_sum may be overwritten by the next item
                if (row < p.Sq) {
                    const float inv = l > 0.f ? 1.0f / l : 0.f;
                    uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(row) * p.ldo + h * HD + 32 * hf);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 v;
                        v.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
                        v.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
                        v.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
                        v.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
                        dst[c] = v;
                    }
                }
                continue;
            }
            uint32_t o[64];
            if (!SPARSE || done > 0) {
                mbar_wait(&B->pv_done[a], (cnt - 1) & 1);
                tc_fence_after();
                tmem_ld32(tO, o);
                tmem_ld32(tO + 32, o + 32);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&B->o_free[a]);
            } else {
                // no KV tile at all for this query tile (empty block lists): zero rows, nothing was accumulated
#pragma unroll
                for (int c = 0; c < 64; ++c) o[c] = 0u;
                l = 0.f;
            }
            if constexpr (PART) {
                if (p.part_mode & 1) {
                    if (row < p.Sq) {          // unnormalised fp32 partials for the launch over the remaining slabs
                        uint4* dst = reinterpret_cast<uint4*>(p.part_o + (static_cast<size_t>(row) * p.heads + h) * HD);
#pragma unroll
                        for (int c = 0; c < 16; ++c) dst[c] = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                        reinterpret_cast<float4*>(p.part_l)[static_cast<size_t>(row) * p.heads + h] = make_float4(l4[0], l4[1], l4[2], l4[3]);
                    }
                    continue;
                }
            }
            if (row < p.Sq) {
                const float inv = l > 0.f ? 1.0f / l : 0.f;
                uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(row) * p.ldo + h * HD);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
                    v.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
                    v.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
                    v.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
                    dst[c] = v;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// Pre-pass for block-sparse attention: per (head, 256-row query item) the ascending list of 128-row KV tiles
// that contain at least one selected 64x64 block for any of the item's 4 query blocks, plus the 8-bit
// sub-block mask (bit qblk*2 + half).
__global__ void __launch_bounds__(256)
build_items_kernel(const int32_t* __restrict__ kv_count, const int32_t* __restrict__ kv_index, int nbq, int nbk,
                   int n_qpairs, int max_pairs, int32_t* __restrict__ item_count, int32_t* __restrict__ item_pairs,
                   uint8_t* __restrict__ item_mask) {
    __shared__ uint8_t act[4][2048];
    __shared__ int warp_counts[8];
    const int qp = blockIdx.x, h = blockIdx.y;
    const int item = h * n_qpairs + qp;
    for (int t = threadIdx.x; t < 4 * 2048; t += blockDim.x) (&act[0][0])[t] = 0;
    __syncthreads();
    for (int r = 0; r < 4; ++r) {
        const int qb = qp * 4 + r;
        if (qb >= nbq) break;
        const int n = kv_count[static_cast<size_t>(h) * nbq + qb];
        const int32_t* src = kv_index + (static_cast<size_t>(h) * nbq + qb) * nbk;
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const int b = src[t];
            if (b >= 0 && b < nbk) act[r][b] = 1;
        }
    }
    __syncthreads();
    const int npairs = (nbk + 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    for (int c0 = 0; c0 < npairs; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        uint32_t m = 0;
        if (c < npairs) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                m |= static_cast<uint32_t>(act[r][2 * c]) << (2 * r);
                if (2 * c + 1 < nbk) m |= static_cast<uint32_t>(act[r][2 * c + 1]) << (2 * r + 1);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, m != 0);
        if (lane == 0) warp_counts[warp] = __popc(bal);
        __syncthreads();
        int pre = 0, tot = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) pre += warp_counts[w];
            tot += warp_counts[w];
        }
        if (m != 0) {
            const int pos = base + pre + __popc(bal & ((1u << lane) - 1u));
            item_pairs[static_cast<size_t>(item) * max_pairs + pos] = c;
            item_mask[static_cast<size_t>(item) * max_pairs + pos] = static_cast<uint8_t>(m);
        }
        base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (base == 0) {   // cannot happen for valid NABLA lists; keep the pipeline protocol alive
            item_pairs[static_cast<size_t>(item) * max_pairs] = 0;
            item_mask[static_cast<size_t>(item) * max_pairs] = 0;
            base = 1;
        }
        item_count[item] = base;
    }
}

AttnSparseWs g_sparse_ws;

}  // namespace

int ensure_sparse_ws(AttnSparseWs& w, size_t items, size_t max_pairs) {
    if (w.items >= items && w.max_pairs >= max_pairs) return K5_OK;
    if (w.count) cudaFree(w.count);
    if (w.pairs) cudaFree(w.pairs);
    if (w.mask) cudaFree(w.mask);
    w.count = nullptr;
    w.pairs = nullptr;
    w.mask = nullptr;
    w.items = w.max_pairs = 0;
    K5_CHECK_CUDA(cudaMalloc(&w.count, items * sizeof(int32_t)));
    K5_CHECK_CUDA(cudaMalloc(&w.pairs, items * max_pairs * sizeof(int32_t)));
    K5_CHECK_CUDA(cudaMalloc(&w.mask, items * max_pairs));
    w.items = items;
    w.max_pairs = max_pairs;
    return K5_OK;
}

namespace {

constexpr int ATT_NPOLY_DEFAULT = 0;
constexpr float ATT_MAX_SCORE_BOUND = 60.f;   // log2 units
constexpr int ATT_STAGGER_DEFAULT = 0;
constexpr int ATT_PAIR_DEFAULT = 1;

template <bool SPARSE, bool BOUNDED>
void launch_kernel(int npoly, int grid, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                   const AttnParams& p, cudaStream_t st) {
    switch (npoly) {
        case 1: attention_fwd_kernel<SPARSE, 1, BOUNDED, SPARSE && BOUNDED><<<grid, SPARSE && BOUNDED ? ATT_THREADS_B : ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p); break;
        case 2: attention_fwd_kernel<SPARSE, 2, BOUNDED, SPARSE && BOUNDED><<<grid, SPARSE && BOUNDED ? ATT_THREADS_B : ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p); break;
        default: attention_fwd_kernel<SPARSE, 0, BOUNDED, SPARSE && BOUNDED><<<grid, SPARSE && BOUNDED ? ATT_THREADS_B : ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p); break;
    }
}

template <bool SPARSE, int NPOLY, bool BOUNDED>
int configure_one() {
    K5_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<SPARSE, NPOLY, BOUNDED, SPARSE && BOUNDED>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    return K5_OK;
}
template <bool SPARSE, bool BOUNDED>
int configure_set() {
    K5_TRY((configure_one<SPARSE, 0, BOUNDED>()));
    K5_TRY((configure_one<SPARSE, 1, BOUNDED>()));
    K5_TRY((configure_one<SPARSE, 2, BOUNDED>()));
    return K5_OK;
}
int configure_kernels() {
    K5_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false, 0, true, false, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    K5_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false, 0, true, false, true, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    K5_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false, 0, true, false, false, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    K5_TRY((configure_set<false, false>()));
    K5_TRY((configure_set<true, false>()));
    K5_TRY((configure_set<false, true>()));
    K5_TRY((configure_set<true, true>()));
    return K5_OK;
}

}  // namespace

int attention_fwd(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo, int Sq,
                  int Sk, int heads, float softmax_scale, const int32_t* kv_count, const int32_t* kv_index,
                  cudaStream_t st, AttnSparseWs* ws_in, float score_bound, const AttnSlabs* slabs, const AttnPartial* part) {
    K5_REQUIRE(Sq > 0 && Sk > 0 && heads > 0, "attention: empty problem");
    const bool sparse = kv_count != nullptr;
    K5_REQUIRE((kv_count == nullptr) == (kv_index == nullptr), "attention: kv_count and kv_index go together");
    K5_REQUIRE(!sparse || (Sq % 64 == 0 && Sk % 64 == 0 && Sk / 64 <= 2048),
               "attention: block-sparse mode needs Sq, Sk multiples of 64 and at most 2048 KV blocks");
    K5_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention: pitches must be x8 elements");
    K5_REQUIRE((reinterpret_cast<uintptr_t>(O) & 15) == 0, "attention: output must be 16-byte aligned");
    CUtensorMap tmQ, tmK, tmV;
    // rows of the K | V matrices: with a slab schedule that leaves out the own slab, Sk counts the walked rows only
    const int kv_rows = (slabs && slabs->n > 0) ? slabs->row0[slabs->n] : Sk;
    K5_TRY(make_tmap_2d_bf16(&tmQ, Q, Sq, static_cast<uint64_t>(heads) * HD, ldq, QT));
    K5_TRY(make_tmap_2d_bf16(&tmK, K, kv_rows, static_cast<uint64_t>(heads) * HD, ldk, KT));
    K5_TRY(make_tmap_2d_bf16(&tmV, V, kv_rows, static_cast<uint64_t>(heads) * HD, ldv, KT));
    static int npoly = -1, stagger = 0, split_tail = 1, use_bounded = 1;
    static std::mutex knob_mutex;
    static PerDevice<int> configured;
    {
        const int dev = current_device();
        std::lock_guard<std::mutex> lk(configured.m);
        if (!configured.set[dev]) {
            K5_TRY(configure_kernels());
            configured.set[dev] = true;
        }
    }
    std::lock_guard<std::mutex> knob_lock(knob_mutex);
    if (npoly < 0) {
        if (const char* sp = getenv("K5_ATTN_SPLIT_TAIL")) split_tail = atoi(sp) != 0;
        const char* sg = getenv("K5_ATTN_STAGGER");
        stagger = sg ? atoi(sg) : ATT_STAGGER_DEFAULT;
        // fraction of the exponentials evaluated on the FMA pipe (pairs out of every 8); tuning knob only
        const char* env = getenv("K5_ATTN_POLY");
        npoly = env ? atoi(env) : ATT_NPOLY_DEFAULT;
        if (npoly < 0 || npoly > 2) npoly = ATT_NPOLY_DEFAULT;
        if (const char* bd = getenv("K5_ATTN_BOUNDED")) use_bounded = atoi(bd) != 0;
    }
    AttnParams p;
    p.Sq = Sq;
    p.Sk = Sk;
    p.heads = heads;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.out = O;
    p.ldo = ldo;
    p.kv_count = kv_count;
    p.kv_index = kv_index;
    p.item_count = nullptr;
    p.item_pairs = nullptr;
    p.item_mask = nullptr;
    p.max_pairs = 0;
    p.stagger = stagger;
    p.split_tail = split_tail;
    p.part_o = nullptr;
    p.part_l = nullptr;
    p.part_mode = 0;
    p.slab_skip = 0;
    p.slab_flags = nullptr;
    p.slab_err = nullptr;
    p.slab_timeout_ns = 0;
    p.slab_epoch = 0;
    p.n_slabs = 0;
    p.slab_first = 0;
    for (int& t : p.slab_tile0) t = 0;
    if (slabs && slabs->n > 0) {     // flags == nullptr: the slab ORDER only (debug: single engine in a rank's order)
        K5_REQUIRE(!sparse, "attention: the overlapped gather is implemented for the dense kernel");
        // rows this launch walks: the slabs at positions [skip, skip + c) of the rotated order first, first + 1, ...
        bool walk_ok = slabs->skip == 0 && slabs->row0[slabs->n] == Sk;
        if (!walk_ok && slabs->first >= 0 && slabs->skip >= 0 && slabs->skip < slabs->n) {
            int walked = 0;
            for (int c = slabs->skip; c < slabs->n && !walk_ok; ++c) {
                const int sl = (slabs->first + c) % slabs->n;
                walked += slabs->row0[sl + 1] - slabs->row0[sl];
                walk_ok = walked == Sk;
            }
        }
        K5_REQUIRE(slabs->n >= 1 && slabs->n <= 8 && slabs->first >= -1 && slabs->first < slabs->n && Sk % KT == 0 &&
                       (slabs->first >= 0 || slabs->flags == nullptr) && slabs->row0[0] == 0 && walk_ok,
                   "attention: bad slab schedule");
        for (int i = 0; i <= slabs->n; ++i) {
            K5_REQUIRE(slabs->row0[i] % KT == 0 && (i == 0 || slabs->row0[i] > slabs->row0[i - 1]),
                       "attention: slab boundaries must be increasing multiples of 128 rows");
            K5_REQUIRE(slabs->first >= 0 || slabs->row0[i] % (2 * QT) == 0 || i == slabs->n,
                       "attention: the per-row debug order needs slab boundaries at multiples of 256 rows");
            p.slab_tile0[i] = slabs->row0[i] / KT;
        }
        K5_REQUIRE(slabs->flags == nullptr || slabs->err != nullptr, "attention: slab flags need an error word");
        p.slab_flags = slabs->flags;
        p.slab_err = slabs->err;
        p.slab_timeout_ns = slabs->timeout_ns;
        p.slab_epoch = slabs->epoch;
        p.n_slabs = slabs->n;
        p.slab_first = slabs->first;
        p.slab_skip = slabs->skip;
    }
    if (part && part->mode != 0) {
        K5_REQUIRE(!sparse && part->mode >= 1 && part->mode <= 3 && part->o && part->l &&
                       (reinterpret_cast<uintptr_t>(part->o) & 15) == 0 && (reinterpret_cast<uintptr_t>(part->l) & 15) == 0,
                   "attention: bad partial-sum buffers");
        p.part_o = part->o;
        p.part_l = part->l;
        p.part_mode = part->mode;
    }
    K5_REQUIRE(p.slab_skip == 0 || (p.part_mode & 2), "attention: skipping slabs needs the partials of the launches that consumed them");
    // fixed-offset softmax only under a proven bound that keeps exp2 and the fp32 row sums far from overflow
    const bool bounded = use_bounded && score_bound > 0.f && score_bound <= ATT_MAX_SCORE_BOUND;
    K5_REQUIRE(p.part_mode == 0 || bounded, "attention: partial sums are additive only under the fixed-offset softmax (score bound)");
    const int n_qpairs = (Sq + 2 * QT - 1) / (2 * QT);
    const int n_items = n_qpairs * heads;
    const int grid = n_items < sm_count() ? n_items : sm_count();
    p.item_count = nullptr;
    p.item_pairs = nullptr;
    p.item_mask = nullptr;
    p.max_pairs = 0;
    p.stagger = stagger;
    if (sparse) {
        const int nbq = Sq / 64, nbk = Sk / 64;
        const int max_pairs = (nbk + 1) / 2;
        AttnSparseWs& ws = ws_in ? *ws_in : g_sparse_ws;
        K5_TRY(ensure_sparse_ws(ws, n_items, max_pairs));
        build_items_kernel<<<dim3(n_qpairs, heads), 256, 0, st>>>(kv_count, kv_index, nbq, nbk, n_qpairs, max_pairs, ws.count,
                                                                  ws.pairs, ws.mask);
        K5_CHECK_CUDA(cudaGetLastError());
        p.item_count = ws.count;
        p.item_pairs = ws.pairs;
        p.item_mask = ws.mask;
        p.max_pairs = max_pairs;
        if (bounded) launch_kernel<true, true>(npoly, grid, tmQ, tmK, tmV, p, st);
        else launch_kernel<true, false>(npoly, grid, tmQ, tmK, tmV, p, st);
    } else {
        // CTA pairs sharing the K / V stream (see the kernel): adjacent items must belong to one head (even item count
        // per head) and the grid must consist of whole pairs.  K5_ATTN_PAIR=0 restores single CTAs (A/B).
        static int pair_env = -1;
        if (pair_env < 0) {
            const char* ev = getenv("K5_ATTN_PAIR");
            pair_env = ev ? (atoi(ev) != 0) : ATT_PAIR_DEFAULT;
        }
        const bool parted = p.part_mode != 0 || p.slab_skip != 0;      // one of several launches over key slabs (PART kernels)
        K5_REQUIRE(!parted || npoly == 0, "attention: launches split by key slab exist for K5_ATTN_POLY=0 only");
        if (bounded && npoly == 0 && pair_env && n_qpairs % 2 == 0 && grid % 2 == 0) {
            CUtensorMap tmK64, tmV64;
            K5_TRY(make_tmap_2d_bf16(&tmK64, K, kv_rows, static_cast<uint64_t>(heads) * HD, ldk, KT / 2));
            K5_TRY(make_tmap_2d_bf16(&tmV64, V, kv_rows, static_cast<uint64_t>(heads) * HD, ldv, KT / 2));
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(static_cast<unsigned>(sm_count() / 2 * 2));
            cfg.blockDim = dim3(ATT_THREADS);
            cfg.dynamicSmemBytes = ATT_SMEM;
            cfg.stream = st;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            static PerDevice<int> max_pairs;          // clusters that can be resident at once (pairs live inside a GPC)
            int resident = 0;
            {
                const int dev = current_device();
                std::lock_guard<std::mutex> lk(max_pairs.m);
                if (!max_pairs.set[dev]) {
                    int n_clusters = 0;
                    K5_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, attention_fwd_kernel<false, 0, true, false, true>, &cfg));
                    K5_REQUIRE(n_clusters > 0, "attention: no CTA pair fits this device");
                    max_pairs.v[dev] = n_clusters;
                    max_pairs.set[dev] = true;
                }
                resident = 2 * max_pairs.v[dev];
            }
            cfg.gridDim = dim3(static_cast<unsigned>(grid < resident ? grid : resident));
            if (parted) K5_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attention_fwd_kernel<false, 0, true, false, true, true>, tmQ, tmK64, tmV64, p));
            else K5_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attention_fwd_kernel<false, 0, true, false, true>, tmQ, tmK64, tmV64, p));
        } else if (bounded && parted) {
            attention_fwd_kernel<false, 0, true, false, false, true><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
        } else if (bounded) {
            launch_kernel<false, true>(npoly, grid, tmQ, tmK, tmV, p, st);
        } else {
            launch_kernel<false, false>(npoly, grid, tmQ, tmK, tmV, p, st);
        }
    }
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

int attention_debug_trace(long long* buf) {
#ifdef K5_ATTN_TRACE
    K5_CHECK_CUDA(cudaMemcpyToSymbol(g_attn_trace, &buf, sizeof(buf)));
    return K5_OK;
#else
    (void)buf;
    set_last_error("attention: built without -DK5_ATTN_TRACE");
    return K5_ERR_UNSUPPORTED;
#endif
}

AttnSparseWs::~AttnSparseWs() {
    if (count) cudaFree(count);
    if (pairs) cudaFree(pairs);
    if (mask) cudaFree(mask);
}

}  // namespace k5
