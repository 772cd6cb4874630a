// Attention forward, head_dim 64: software-pipelined softmax over 64-row KV tiles (v4).
//
// Same contract as attention.cu (flash_attn_func at kandinsky/models/nn.py:201,254,336; flex_attention with the
// NABLA block lists at nn.py:257-280).  The ncu source pages of v2 / v3 (profiles/r1_attention_v2_stalls.md,
// profiles/r1_attention_v3_v4.md) show the MUFU (the unit that bounds head_dim-64 attention: 16 exponentials per clock
// per SM against 512 tensor cycles per 1024 MUFU cycles) idle ~20 % of the time because every softmax warp spends
// as long OUTSIDE its exponential loop (mbarrier round trips, tcgen05.ld + wait, row max, argument scaling, P
// store) as the MUFU needs for the loop itself, and the warps of one scheduler behave like independent customers
// of one server.  Here the per-tile work of a warp is restructured so that nothing but the P hand-over is left
// outside the exponential stream: while the exponentials of KV tile j are issued, the SAME thread
//   * probes / waits the barrier of S(j+1) and issues its tcgen05.ld     (after the 1st quarter of the exponentials),
//   * collects it, releases the S buffer to the tensor pipe              (after the 2nd quarter),
//   * reduces the row max of S(j+1) and takes the lazy-rescale decision  (interleaved with the 3rd quarter),
//   * forms the scaled arguments of tile j+1 in place                    (interleaved with the 4th quarter),
// so the non-MUFU instructions sit in the issue slots the MUFU leaves free (a warp can issue one MUFU per 8 cycles).
// Two register tiles (current / next) ping-pong; the loop is unrolled by two so no register copies are needed.
//   TMEM  : S_a (2 x 64 columns, double-buffered) | O_a (64) | P_a (2 x 32, double-buffered) per query tile a = 512
//   warps : 4 * NQ softmax warps (thread = one query row), then TMA producer (+ TMEM allocation), then one
//           tcgen05.mma issuer warp per query tile
//   smem  : NQ Q tiles (16 KB each) + a ring of K | V tiles (8 KB + 8 KB)
#include <cstdlib>

#include "attention.h"
#include "common.h"
#include "ptx.cuh"

namespace k5 {

namespace {

constexpr int QT = 128;            // query rows per tile
constexpr int KT = 64;             // kv rows per tile (= one NABLA block)
constexpr int HD = 64;
constexpr int Q_BYTES = QT * HD * 2;
constexpr int KV_BYTES = KT * HD * 2;
constexpr int STAGES = 8;
constexpr float RESCALE_THRESHOLD = 8.0f;

template <int NQ>
struct Cfg {
    static constexpr int THREADS = NQ * 128 + 128;
    static constexpr int SMEM = NQ * Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 1024;
    static constexpr int TMA_WARP = 4 * NQ;
    static constexpr int ISSUER0 = 4 * NQ + 1;
    // TMEM columns: S_a(buf) = a * 128 + buf * 64, O_a = NQ * 128 + a * 64, P_a(buf) = NQ * 192 + a * 64 + buf * 32
    static constexpr uint32_t TM_S = 0, TM_O = NQ * 128, TM_P = NQ * 192;
    static_assert(NQ * 256 <= 512, "TMEM columns");
    // register pool = THREADS x (65536 / THREADS, rounded down to 8): softmax warps take what the others release
    static constexpr int REG_OTHER = 56;
    static constexpr int REG_SOFTMAX = 224;
    // mbarrier block (byte offsets from its shared-memory base)
    static constexpr uint32_t Q_FULL = 0, Q_EMPTY = Q_FULL + 8 * NQ;
    static constexpr uint32_t K_FULL = Q_EMPTY + 8 * NQ, K_EMPTY = K_FULL + 8 * STAGES;
    static constexpr uint32_t V_FULL = K_EMPTY + 8 * STAGES, V_EMPTY = V_FULL + 8 * STAGES;
    static constexpr uint32_t S_FULL = V_EMPTY + 8 * STAGES;       // [a][buf]
    static constexpr uint32_t S_FREE = S_FULL + 16 * NQ;           // [a][buf]
    static constexpr uint32_t P_READY = S_FREE + 16 * NQ;          // [a][buf]
    static constexpr uint32_t PV_DONE = P_READY + 16 * NQ;         // [a][buf]
    static constexpr uint32_t O_FREE = PV_DONE + 16 * NQ;          // [a]
    static constexpr uint32_t TMEM_SLOT = O_FREE + 8 * NQ;
    static_assert(TMEM_SLOT + 4 <= 1024, "barrier block");
};

template <bool SPARSE, int NQ>
__global__ void __launch_bounds__(Cfg<NQ>::THREADS, 1)
attention4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, AttnParams p) {
    using C = Cfg<NQ>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;     // shared::cta byte addresses from here on
    const uint32_t sQ = smem;
    const uint32_t sKV = smem + NQ * Q_BYTES;            // stage s: K at s * 2 * KV_BYTES, V right behind it
    const uint32_t bars = smem + NQ * Q_BYTES + STAGES * 2 * KV_BYTES;

    const int warp = threadIdx.x >> 5;
    const int n_qg = (p.Sq + NQ * QT - 1) / (NQ * QT);
    const int n_items = n_qg * p.heads;
    const int nkv_dense = (p.Sk + KT - 1) / KT;
    const int kv_rem = p.Sk - (nkv_dense - 1) * KT;      // valid kv rows in the last tile (1..64)

    if (warp == C::TMA_WARP && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == C::ISSUER0 && elect_one()) {
        for (int a = 0; a < NQ; ++a) {
            mbar_init_a(bars + C::Q_FULL + 8 * a, 1);
            mbar_init_a(bars + C::Q_EMPTY + 8 * a, 1);
            for (int b = 0; b < 2; ++b) {
                mbar_init_a(bars + C::S_FULL + 16 * a + 8 * b, 1);
                mbar_init_a(bars + C::S_FREE + 16 * a + 8 * b, 128);
                mbar_init_a(bars + C::P_READY + 16 * a + 8 * b, 128);
                mbar_init_a(bars + C::PV_DONE + 16 * a + 8 * b, 1);
            }
            mbar_init_a(bars + C::O_FREE + 8 * a, 128);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init_a(bars + C::K_FULL + 8 * s, 1);
            mbar_init_a(bars + C::K_EMPTY + 8 * s, NQ);  // every issuer commits on a stage before it is refilled
            mbar_init_a(bars + C::V_FULL + 8 * s, 1);
            mbar_init_a(bars + C::V_EMPTY + 8 * s, NQ);
        }
        fence_barrier_init();
    }
    if (warp == C::TMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(bars + C::TMEM_SLOT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bars + C::TMEM_SLOT));

    if (warp >= 4 * NQ) {
        reg_dec<C::REG_OTHER>();
        if (warp == C::TMA_WARP) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                int st = 0;
                uint32_t ph = 0;
                int i = 0;
                for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
                    const int h = item / n_qg;
                    const int q0 = (item % n_qg) * NQ * QT;
                    for (int a = 0; a < NQ; ++a) {
                        mbar_wait_parked_a(bars + C::Q_EMPTY + 8 * a, (i & 1) ^ 1);
                        mbar_expect_tx_a(bars + C::Q_FULL + 8 * a, Q_BYTES);
                        tma_load_2d_a(sQ + a * Q_BYTES, &tmQ, bars + C::Q_FULL + 8 * a, h * HD, q0 + a * QT);
                    }
                    const int nkv = SPARSE ? p.item_count[item] : nkv_dense;
                    const int32_t* blocks = SPARSE ? p.item_pairs + static_cast<size_t>(item) * p.max_pairs : nullptr;
                    for (int j = 0; j < nkv; ++j) {
                        const int kv0 = (SPARSE ? blocks[j] : j) * KT;
                        const uint32_t sk = sKV + st * 2 * KV_BYTES;
                        mbar_wait_parked_a(bars + C::K_EMPTY + 8 * st, ph ^ 1);
                        mbar_expect_tx_a(bars + C::K_FULL + 8 * st, KV_BYTES);
                        tma_load_2d_a(sk, &tmK, bars + C::K_FULL + 8 * st, h * HD, kv0);
                        mbar_wait_parked_a(bars + C::V_EMPTY + 8 * st, ph ^ 1);
                        mbar_expect_tx_a(bars + C::V_FULL + 8 * st, KV_BYTES);
                        tma_load_2d_a(sk + KV_BYTES, &tmV, bars + C::V_FULL + 8 * st, h * HD, kv0);
                        if (++st == STAGES) {
                            st = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        } else if (warp < C::ISSUER0 + NQ) {
            // ===================== MMA issuers: one warp per query tile =====================
            if (elect_one()) {
                const int a = warp - C::ISSUER0;
                constexpr uint32_t idesc_qk = umma_idesc_bf16(QT, KT, 0, 0);
                constexpr uint32_t idesc_pv = umma_idesc_bf16(QT, HD, 0, 1);
                const uint32_t tS = tmem_base + C::TM_S + a * 128;
                const uint32_t tO = tmem_base + C::TM_O + a * 64;
                const uint32_t tP = tmem_base + C::TM_P + a * 64;
                const uint32_t qa = sQ + a * Q_BYTES;
                int kst = 0, vst = 0;
                uint32_t kph = 0, vph = 0;
                uint32_t g = 0;                     // KV tiles handled so far (over all items)
                int i = 0;

                // S_a(gt) goes to buffer gt & 1, free once the softmax warps have pulled S_a(gt - 2)
                auto issue_qk = [&](uint32_t gt, bool last_of_item) {
                    const uint32_t b = gt & 1;
                    if (gt >= 2) mbar_wait_parked_a(bars + C::S_FREE + 16 * a + 8 * b, ((gt >> 1) - 1) & 1);
                    mbar_wait_parked_a(bars + C::K_FULL + 8 * kst, kph);
                    tc_fence_after();
                    const uint32_t ka = sKV + kst * 2 * KV_BYTES;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_ss(tS + b * 64, umma_desc_sw128(qa + k * 32, 0, 1024), umma_desc_sw128(ka + k * 32, 0, 1024),
                                idesc_qk, k != 0 ? 1u : 0u);
                    umma_commit_a(bars + C::S_FULL + 16 * a + 8 * b);
                    umma_commit_a(bars + C::K_EMPTY + 8 * kst);
                    if (last_of_item) umma_commit_a(bars + C::Q_EMPTY + 8 * a);
                    if (++kst == STAGES) {
                        kst = 0;
                        kph ^= 1;
                    }
                };
                for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
                    const int nkv = SPARSE ? p.item_count[item] : nkv_dense;
                    mbar_wait_parked_a(bars + C::Q_FULL + 8 * a, i & 1);
                    issue_qk(g, nkv == 1);
                    if (nkv > 1) issue_qk(g + 1, nkv == 2);
                    for (int j = 0; j < nkv; ++j, ++g) {
                        // the scores run two tiles ahead of the exponentials: S(j+2) is issued as soon as S(j) has
                        // been pulled, which the softmax warps do in the middle of tile j-1
                        if (j + 2 < nkv) issue_qk(g + 2, j + 3 == nkv);
                        if (j == 0) mbar_wait_parked_a(bars + C::O_FREE + 8 * a, (i & 1) ^ 1);
                        mbar_wait_parked_a(bars + C::P_READY + 16 * a + 8 * (g & 1), (g >> 1) & 1);
                        mbar_wait_parked_a(bars + C::V_FULL + 8 * vst, vph);
                        tc_fence_after();
                        const uint32_t va = sKV + vst * 2 * KV_BYTES + KV_BYTES;
#pragma unroll
                        for (int k = 0; k < KT / 16; ++k)
                            umma_ts(tO, tP + (g & 1) * 32 + k * 8, umma_desc_sw128(va + k * 2048, KV_BYTES, 1024), idesc_pv,
                                    (j != 0 || k != 0) ? 1u : 0u);
                        umma_commit_a(bars + C::PV_DONE + 16 * a + 8 * (g & 1));
                        umma_commit_a(bars + C::V_EMPTY + 8 * vst);
                        if (++vst == STAGES) {
                            vst = 0;
                            vph ^= 1;
                        }
                    }
                }
            }
        }
    } else {
        // ===================== softmax warpgroups =====================
        reg_inc<C::REG_SOFTMAX>();
        const int a = warp >> 2;                          // query tile of this warpgroup
        const int wq = warp & 3;
        const int lane = threadIdx.x & 31;
        const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t tS0 = tmem_base + C::TM_S + a * 128 + lane_off;   // + (tile & 1) * 64
        const uint32_t tO = tmem_base + C::TM_O + a * 64 + lane_off;
        const uint32_t tP0 = tmem_base + C::TM_P + a * 64 + lane_off;    // + (tile & 1) * 32
        const uint32_t b_sfull = bars + C::S_FULL + 16 * a, b_sfree = bars + C::S_FREE + 16 * a;
        const uint32_t b_pready = bars + C::P_READY + 16 * a, b_pvdone = bars + C::PV_DONE + 16 * a;
        const float sl2 = p.scale_log2;
        const uint64_t sl2x2 = pack_f32x2(sl2, sl2);
        const int qblk = a * 2 + (wq >> 1);               // this warp's 64-row query block inside the item
        uint32_t cnt = 0;                                 // KV tiles handled so far (over all items)

        // per-item state: running stabiliser (raw score units), row sum, and the rescale of the CURRENT tile
        // (decided when the tile was collected)
        float m_used = 0.f, l = 0.f, alpha_c = 1.f;
        bool started = false, need_c = false;
        int j = 0, nkv = 0;
        const uint8_t* masks = nullptr;

        auto row_max = [&](const uint32_t (&t)[KT]) -> float {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < KT; c += 8) {
                m0 = max3(m0, __uint_as_float(t[c]), __uint_as_float(t[c + 1]));
                m1 = max3(m1, __uint_as_float(t[c + 2]), __uint_as_float(t[c + 3]));
                m2 = max3(m2, __uint_as_float(t[c + 4]), __uint_as_float(t[c + 5]));
                m3 = max3(m3, __uint_as_float(t[c + 6]), __uint_as_float(t[c + 7]));
            }
            return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        };
        auto mask_tail = [&](uint32_t (&t)[KT]) {
#pragma unroll
            for (int c = 0; c < KT; ++c)
                if (c >= kv_rem) t[c] = 0xff800000u;      // -inf
        };
        // Lazy rescale: the stabiliser only moves when the tile max outgrows it by more than 2^8; returns -m * scale.
        auto decide = [&](float mx, float& alpha, bool& need) -> float {
            const bool grow = started && (mx - m_used) * sl2 > RESCALE_THRESHOLD;
            const float al = fast_exp2((m_used - mx) * sl2);
            alpha = grow ? al : 1.0f;
            need = grow;
            m_used = (!started || grow) ? mx : m_used;
            started = true;
            return -m_used * sl2;
        };
        auto scale_pair = [&](uint32_t (&t)[KT], int e, uint64_t mnegx2) {
            float x0, x1;
            unpack_f32x2(fma_f32x2_v(pack_f32x2(__uint_as_float(t[e]), __uint_as_float(t[e + 1])), sl2x2, mnegx2), x0, x1);
            t[e] = __float_as_uint(x0);
            t[e + 1] = __float_as_uint(x1);
        };
        // Collect tile `tile_cnt` without overlap: used for the first tile of an item, for a ragged last tile and
        // behind tiles this warp skips in block-sparse mode.  Leaves the scaled arguments in t, (alpha_c, need_c) set.
        auto load_blocking = [&](uint32_t (&t)[KT], uint32_t tile_cnt, bool act, bool last) {
            const uint32_t b = tile_cnt & 1;
            const uint32_t tS = tS0 + b * 64;
            mbar_wait_a(b_sfull + 8 * b, (tile_cnt >> 1) & 1);
            tc_fence_after();
            if (SPARSE && !act) {
                mbar_arrive_a(b_sfree + 8 * b);
                alpha_c = 1.0f;
                need_c = false;
                return;
            }
            tmem_ld32(tS + 0, t);
            tmem_ld32(tS + 32, t + 32);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive_a(b_sfree + 8 * b);
            if (!SPARSE && last && kv_rem < KT) mask_tail(t);
            const float mneg = decide(row_max(t), alpha_c, need_c);
            const uint64_t mnegx2 = pack_f32x2(mneg, mneg);
#pragma unroll
            for (int e = 0; e < KT; e += 2) scale_pair(t, e, mnegx2);
        };
        // Hand P (packed bf16 in t[0..31]) of tile cnt to the tensor pipe through P buffer cnt & 1, which PV(cnt-2)
        // read last (pv2_ok: already seen complete); rescale O first when the stabiliser moved (needs PV(cnt-1)).
        // A warp may therefore run one tile ahead of the other warps of its query tile, which is why p_ready (like
        // s_full / s_free / pv_done) is one barrier PER BUFFER: arrivals for tile cnt+1 must not count for tile cnt.
        auto publish = [&](const uint32_t (&t)[KT], bool pv2_ok) {
            const uint32_t b = cnt & 1;
            if (cnt >= 2 && !pv2_ok) mbar_wait_a(b_pvdone + 8 * b, ((cnt >> 1) - 1) & 1);
            tc_fence_after();
            tmem_st32(tP0 + b * 32, t);
            if (j > 0 && __any_sync(0xffffffffu, need_c)) {
                mbar_wait_a(b_pvdone + 8 * (b ^ 1), ((cnt - 1) >> 1) & 1);
                tc_fence_after();
                uint32_t o[64];
                tmem_ld32(tO, o);
                tmem_ld32(tO + 32, o + 32);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 64; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha_c);
                tmem_st32(tO, o);
                tmem_st32(tO + 32, o + 32);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_a(b_pready + 8 * b);
        };
        // The exponentials of one quarter (8 pairs) of the current tile.
        auto ex_quarter = [&](const uint32_t (&cur)[KT], int c, float (&dst)[16], int q) {
            dst[2 * q] = ex2_v(__uint_as_float(cur[16 * c + 2 * q]));
            dst[2 * q + 1] = ex2_v(__uint_as_float(cur[16 * c + 2 * q + 1]));
        };
        // One active tile whose successor is collected in the same instruction stream.  Requires: a next tile
        // exists, this warp's query block selects it, and it needs no tail masking.  Apart from the (normally
        // satisfied) barrier checks the body is straight-line code, which is what lets the scheduler interleave.
        auto phase_pipe = [&](uint32_t (&cur)[KT], uint32_t (&nxt)[KT]) {
            uint64_t sum_a = pack_f32x2(0.f, 0.f), sum_b = pack_f32x2(0.f, 0.f);
            float pc[16] = {}, pp[16] = {};
            auto drain = [&](int c, int q) {              // results of quarter c-1: row sum + bf16 pack in place
#ifndef K5_V4_NOSUM
                const uint64_t pr = pack_f32x2(pp[2 * q], pp[2 * q + 1]);
                if (q & 1) sum_b = add_f32x2_v(sum_b, pr);
                else sum_a = add_f32x2_v(sum_a, pr);
#else
                if (c == 1 && q == 0) sum_a = pack_f32x2(pp[0], pp[1]);   // timing experiment only
#endif
                cur[8 * (c - 1) + q] = pack_bf16x2_v(pp[2 * q], pp[2 * q + 1]);
            };
            const uint32_t nb = (cnt + 1) & 1, nph = ((cnt + 1) >> 1) & 1;
            const bool s_ok = mbar_test_wait_a(b_sfull + 8 * nb, nph);     // issued two tiles ago: normally complete
            // ---- quarter 0 (its MUFU queue covers the barrier round trip)
#pragma unroll
            for (int q = 0; q < 8; ++q) ex_quarter(cur, 0, pp, q);
            if (!s_ok) mbar_wait_a(b_sfull + 8 * nb, nph);
            tc_fence_after();
            tmem_ld32(tS0 + nb * 64, nxt);
            tmem_ld32(tS0 + nb * 64 + 32, nxt + 32);
            // ---- quarter 1
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ex_quarter(cur, 1, pc, q);
                drain(1, q);
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) pp[q] = pc[q];
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive_a(b_sfree + 8 * nb);              // this S buffer may take the scores of tile j+3
            const bool pv2_ok = mbar_test_wait_a(b_pvdone + 8 * (cnt & 1), ((cnt >> 1) - 1) & 1);   // unused if cnt < 2
            // ---- quarter 2, with the row max of the next tile
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ex_quarter(cur, 2, pc, q);
                drain(2, q);
                m0 = max3_v(m0, __uint_as_float(nxt[8 * q]), __uint_as_float(nxt[8 * q + 1]));
                m1 = max3_v(m1, __uint_as_float(nxt[8 * q + 2]), __uint_as_float(nxt[8 * q + 3]));
                m2 = max3_v(m2, __uint_as_float(nxt[8 * q + 4]), __uint_as_float(nxt[8 * q + 5]));
                m3 = max3_v(m3, __uint_as_float(nxt[8 * q + 6]), __uint_as_float(nxt[8 * q + 7]));
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) pp[q] = pc[q];
            float alpha_n;
            bool need_n;
            const float mneg = decide(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), alpha_n, need_n);
            const uint64_t mnegx2 = pack_f32x2(mneg, mneg);
            // ---- quarter 3, with the scaled arguments of the next tile formed in place
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ex_quarter(cur, 3, pc, q);
                drain(3, q);
#pragma unroll
                for (int e = 0; e < 8; e += 2) scale_pair(nxt, 8 * q + e, mnegx2);
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) pp[q] = pc[q];
#pragma unroll
            for (int q = 0; q < 8; ++q) drain(4, q);
            {
                float t0, t1;
                unpack_f32x2(add_f32x2(sum_a, sum_b), t0, t1);
                l = l * alpha_c + (t0 + t1);
            }
            publish(cur, pv2_ok);
            alpha_c = alpha_n;
            need_c = need_n;
        };
        // One active tile without look-ahead (last tile of an item, or the next one is skipped / needs tail masking).
        auto phase_plain = [&](uint32_t (&cur)[KT]) {
            uint64_t sum_a = pack_f32x2(0.f, 0.f), sum_b = pack_f32x2(0.f, 0.f);
            float pc[16] = {}, pp[16] = {};
#pragma unroll
            for (int c = 0; c <= 4; ++c) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (c < 4) ex_quarter(cur, c, pc, q);
                    if (c > 0) {
                        const uint64_t pr = pack_f32x2(pp[2 * q], pp[2 * q + 1]);
                        if (q & 1) sum_b = add_f32x2_v(sum_b, pr);
                        else sum_a = add_f32x2_v(sum_a, pr);
                        cur[8 * (c - 1) + q] = pack_bf16x2_v(pp[2 * q], pp[2 * q + 1]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) pp[q] = pc[q];
            }
            {
                float t0, t1;
                unpack_f32x2(add_f32x2(sum_a, sum_b), t0, t1);
                l = l * alpha_c + (t0 + t1);
            }
            publish(cur, false);
        };

        if (a == 1 && p.stagger > 0) {
            const long long t0 = clock64();               // tuning knob: start query tile 1 behind query tile 0
            while (clock64() - t0 < p.stagger) {
            }
        }
        uint32_t ta[KT], tb[KT];
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int h = item / n_qg;
            const int row = (item % n_qg) * NQ * QT + a * QT + wq * 32 + lane;
            m_used = -INFINITY;
            l = 0.f;
            started = false;
            nkv = SPARSE ? p.item_count[item] : nkv_dense;
            if constexpr (SPARSE) masks = p.item_mask + static_cast<size_t>(item) * p.max_pairs;
            const bool tail = !SPARSE && kv_rem < KT;
            j = 0;
            load_blocking(ta, cnt, SPARSE ? ((masks[0] >> qblk) & 1u) : true, nkv == 1);
            if constexpr (!SPARSE) {
                // dense main loop: every tile but the last one (two when the last is ragged) runs pipelined
                const int jlim = nkv - (tail ? 2 : 1);    // tiles [0, jlim) have a successor that needs no masking
                while (j + 2 <= jlim) {
                    phase_pipe(ta, tb);
                    ++j, ++cnt;
                    phase_pipe(tb, ta);
                    ++j, ++cnt;
                }
            }
            bool flip = false;                            // false: the current tile lives in ta
            while (j < nkv) {
                bool act = true, act_next = j + 1 < nkv;
                if constexpr (SPARSE) {
                    act = (masks[j] >> qblk) & 1u;
                    act_next = act_next && ((masks[j + 1] >> qblk) & 1u);
                }
                bool loaded = false;                      // next tile already collected by this phase
                if (SPARSE && !act) {
                    // this KV block is not selected for the warp's query block: P = 0
                    uint32_t z[KT];
#pragma unroll
                    for (int c = 0; c < 32; ++c) z[c] = 0u;
                    need_c = false;
                    publish(z, false);
                } else if (act_next && !(tail && j + 2 == nkv)) {
                    if (!flip) phase_pipe(ta, tb);
                    else phase_pipe(tb, ta);
                    flip = !flip;
                    loaded = true;
                } else {
                    if (!flip) phase_plain(ta);
                    else phase_plain(tb);
                }
                ++j, ++cnt;
                if (!loaded && j < nkv) {
                    flip = false;
                    load_blocking(ta, cnt, SPARSE ? ((masks[j] >> qblk) & 1u) : true, j + 1 == nkv);
                }
            }
            // ---- epilogue: O_a / l -> bf16 -> global
            mbar_wait_a(b_pvdone + 8 * ((cnt - 1) & 1), ((cnt - 1) >> 1) & 1);
            tc_fence_after();
            uint32_t o[64];
            tmem_ld32(tO, o);
            tmem_ld32(tO + 32, o + 32);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive_a(bars + C::O_FREE + 8 * a);
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
    if (warp == C::TMA_WARP) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// Pre-pass for block-sparse attention: per (head, NQ*128-row query item) the ascending list of 64-row KV blocks
// selected for at least one of the item's 2*NQ query blocks, plus the mask (bit r = selected for query block r).
template <int NQ>
__global__ void __launch_bounds__(256)
build_items4_kernel(const int32_t* __restrict__ kv_count, const int32_t* __restrict__ kv_index, int nbq, int nbk, int n_qg,
                    int max_pairs, int32_t* __restrict__ item_count, int32_t* __restrict__ item_pairs,
                    uint8_t* __restrict__ item_mask) {
    constexpr int R = 2 * NQ;
    __shared__ uint8_t act[R][2048];
    __shared__ int warp_counts[8];
    const int qg = blockIdx.x, h = blockIdx.y;
    const int item = h * n_qg + qg;
    for (int t = threadIdx.x; t < R * 2048; t += blockDim.x) (&act[0][0])[t] = 0;
    __syncthreads();
    for (int r = 0; r < R; ++r) {
        const int qb = qg * R + r;
        if (qb >= nbq) break;
        const int n = kv_count[static_cast<size_t>(h) * nbq + qb];
        const int32_t* src = kv_index + (static_cast<size_t>(h) * nbq + qb) * nbk;
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            const int b = src[t];
            if (b >= 0 && b < nbk) act[r][b] = 1;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    for (int c0 = 0; c0 < nbk; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        uint32_t m = 0;
        if (c < nbk) {
#pragma unroll
            for (int r = 0; r < R; ++r) m |= static_cast<uint32_t>(act[r][c]) << r;
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

template <bool SPARSE, int NQ>
int launch4(int grid, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p, cudaStream_t st) {
    static PerDevice<int> pd;
    {
        const int dev = current_device();
        std::lock_guard<std::mutex> lk(pd.m);
        if (!pd.set[dev]) {
            K5_CHECK_CUDA(cudaFuncSetAttribute(attention4_kernel<SPARSE, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               Cfg<NQ>::SMEM));
            pd.set[dev] = true;
        }
    }
    attention4_kernel<SPARSE, NQ><<<grid, Cfg<NQ>::THREADS, Cfg<NQ>::SMEM, st>>>(tmQ, tmK, tmV, p);
    return K5_OK;
}

template <int NQ>
int run4(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, AttnParams p, cudaStream_t st, AttnSparseWs& ws) {
    CUtensorMap tmQ, tmK, tmV;
    K5_TRY(make_tmap_2d_bf16(&tmQ, Q, p.Sq, static_cast<uint64_t>(p.heads) * HD, ldq, QT));
    K5_TRY(make_tmap_2d_bf16(&tmK, K, p.Sk, static_cast<uint64_t>(p.heads) * HD, ldk, KT));
    K5_TRY(make_tmap_2d_bf16(&tmV, V, p.Sk, static_cast<uint64_t>(p.heads) * HD, ldv, KT));
    const int n_qg = (p.Sq + NQ * QT - 1) / (NQ * QT);
    const int n_items = n_qg * p.heads;
    const int grid = n_items < sm_count() ? n_items : sm_count();
    if (p.kv_count != nullptr) {
        const int nbq = p.Sq / 64, nbk = p.Sk / 64;
        K5_TRY(ensure_sparse_ws(ws, n_items, nbk));
        build_items4_kernel<NQ><<<dim3(n_qg, p.heads), 256, 0, st>>>(p.kv_count, p.kv_index, nbq, nbk, n_qg, nbk, ws.count,
                                                                     ws.pairs, ws.mask);
        K5_CHECK_CUDA(cudaGetLastError());
        p.item_count = ws.count;
        p.item_pairs = ws.pairs;
        p.item_mask = ws.mask;
        p.max_pairs = nbk;
        K5_TRY((launch4<true, NQ>(grid, tmQ, tmK, tmV, p, st)));
    } else {
        K5_TRY((launch4<false, NQ>(grid, tmQ, tmK, tmV, p, st)));
    }
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

}  // namespace

int attention_fwd_v4(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, AttnParams p, cudaStream_t st,
                     AttnSparseWs& ws) {
    return run4<2>(Q, ldq, K, ldk, V, ldv, p, st, ws);   // 3 query tiles do not fit TMEM with double-buffered S and P
}

}  // namespace k5
