// bf16 GEMM on the 5th-gen tensor cores:  C[M,N] = A[M,K] . W[N,K]^T  (+ fused epilogue).
//
// This one kernel replaces every nn.Linear on the DiT hot path of the reference
// (kandinsky/models/nn.py:181-184,206,235-237,284,317-319,341,354-361,376-382; SURVEY.md K4-K6,K13):
//   * operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a multi-stage shared
//     memory ring, one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1),
//   * the fp32 accumulator lives in TMEM, double-buffered (2 x BN columns) so that the epilogue of
//     tile i overlaps the main loop of tile i+1,
//   * the kernel is persistent (one CTA per SM, static tile schedule, N fastest so that the A rows
//     of a wave are shared through L2 and W stays L2 resident),
//   * the epilogue warps read TMEM with tcgen05.ld (thread = one output row) and apply the op that
//     follows the Linear in the reference, so that no extra HBM round trip is needed:
//       EPI_STORE : bf16(acc + bias)
//       EPI_GELU  : bf16(gelu_erf(bf16(acc)))                               nn.py:356
//       EPI_GATE  : bf16(x + gate * bf16(acc + bias))                       nn.py:30-33 (apply_gate_sum)
//       EPI_HEADS : bf16(acc + bias) -> per-head RMSNorm (fp32) -> bf16 -> RoPE (fp32) -> bf16
//                                                                           nn.py:246-250, 35-40
// Rounding points follow SURVEY.md Appendix A.
#include <cstdlib>

#include "common.h"
#include "gemm.h"
#include "ptx.cuh"

namespace k5 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
// 0 = single CTA, 1 = CTA pair + W multicast, 3 = CTA pair + 2-SM MMA (K5_GEMM_CLUSTER overrides; profiles/r1_gemm_cluster.md)
constexpr int GEMM_CLUSTER_DEFAULT = 3;
constexpr int GEMM_THREADS = 384;     // warps 0-2: TMA / MMA / TMEM alloc; warps 4-11: epilogue, two per TMEM lane quarter

// CL: 1 = single CTA; 2 = CTA pair, W tile multicast, cta_group::1 MMAs; 3 = CTA pair, ONE cta_group::2 MMA per pair
// (each CTA keeps only its half of the W tile).
template <int BN, int CL = 1>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (CL == 3 ? BN / 2 : BN) * BK * 2;      // bytes of W held per CTA and stage
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = CL == 3 ? ((BN == 256) ? 6 : 8) : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int XPOSE_BYTES = 8 * 32 * 128;   // per epilogue warp: 32 rows x 128 B, for the peer scatter
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + XPOSE_BYTES;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Static tile schedule.  CL = 1: tile = blockIdx.x + i * gridDim.x, N fastest.  CL = 2 (thread-block cluster of two
// CTAs): the pair works on the two M tiles of one 256-row "super tile" and the same N tile, so both need the same W
// tile: each CTA fetches one half of it and TMA multicasts it into both shared memories - 32 KB instead of 48 KB
// of L2 -> SM traffic per CTA and k-block at BN = 256.  (At 128 x 256 x 64 per 512 tensor cycles the single-CTA kernel
// asks L2 for ~13 TB/s over 148 SMs, which is what the L2 can deliver; profiles/r1_gemm_cluster.md.)
template <int CL>
struct TileIter {
    int i, step, count, n_tiles_n, crank;
    __device__ TileIter(int n_tiles_m, int n_tiles_n_, int crank_) : n_tiles_n(n_tiles_n_), crank(crank_) {
        i = blockIdx.x / CL;
        step = gridDim.x / CL;
        count = ((n_tiles_m + CL - 1) / CL) * n_tiles_n;
    }
    __device__ bool valid() const { return i < count; }
    __device__ void next() { i += step; }
    __device__ int m0() const { return ((i / n_tiles_n) * CL + crank) * BM; }
    __device__ int n0(int BN) const { return (i % n_tiles_n) * BN; }
};

template <int BN, int EPI, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 GemmEpilogue e) {
    using Cfg = GemmCfg<BN, CL>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NC = CL == 1 ? 1 : 2;          // CTAs per cluster
    constexpr bool MMA2 = CL == 3;               // 2-SM MMA
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint8_t* xpose = smem + STAGES * Cfg::STAGE_BYTES + 256;

    const int warp = threadIdx.x >> 5;
    const int n_tiles_n = N / BN;
    const int n_tiles_m = (M + BM - 1) / BM;
    const int nkb = (K + BK - 1) / BK;
    const int crank = NC > 1 ? static_cast<int>(cluster_ctarank()) : 0;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL == 2 ? 2 : 1);   // CL = 2: the peer's multicast also lands here, both MMA warps free it
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], MMA2 ? 512 : 256);   // 2-SM MMA: the leader's barrier collects both CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (MMA2) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
        else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (NC > 1) cluster_sync_all();   // the peer's barriers exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (TileIter<NC> t(n_tiles_m, n_tiles_n, crank); t.valid(); t.next()) {
                const int m0 = t.m0();
                const int n0 = t.n0(BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_parked(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    if constexpr (MMA2) {
                        // both CTAs load their A tile and their half of the W tile into their own shared memory; the
                        // bytes of both are counted on the leader's barrier, which only the leader arms
                        if (crank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                        tma_load_2d_2sm(sa, &tmA, &full[stage], kb * BK, m0);
                        tma_load_2d_2sm(sa + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, n0 + crank * (BN / 2));
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
                    if constexpr (CL == 2) {
                        // tmB's box is BN / 2 rows here: this CTA's slice of the W tile goes to every CTA of the pair
                        tma_load_2d_mc(sa + Cfg::A_BYTES + crank * (Cfg::B_BYTES / 2), &tmB, &full[stage], kb * BK,
                                       n0 + crank * (BN / 2), static_cast<uint16_t>(3));
                    } else {
                        tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, n0);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (2-SM MMA: the leader CTA issues for the pair) =====================
        if ((!MMA2 || crank == 0) && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(MMA2 ? 2 * BM : BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (TileIter<NC> t(n_tiles_m, n_tiles_n, crank); t.valid(); t.next(), ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_parked(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_parked(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = umma_desc_sw128(sa + k * 32, 0, 1024);
                        const uint64_t bd = umma_desc_sw128(sb + k * 32, 0, 1024);
                        if constexpr (MMA2) umma_ss_2sm(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // smem slot reusable once these MMAs have read it (in both CTAs when the pair shares the stage)
                    if constexpr (MMA2) umma_commit_2sm_mc(&empty[stage], static_cast<uint16_t>(3));
                    else if constexpr (CL == 2) umma_commit_mc(&empty[stage], static_cast<uint16_t>(3));
                    else umma_commit(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                // accumulator complete (2-SM MMA: each CTA's epilogue waits on its own copy of the barrier)
                if constexpr (MMA2) umma_commit_2sm_mc(&tfull[acc], static_cast<uint16_t>(3));
                else umma_commit(&tfull[acc]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps, thread = output row x half of the tile's columns) =====================
        // A warp may only touch TMEM lanes 32 * (warp % 4) .. +31, so warps 4-7 and 8-11 both cover the 128 rows; the
        // first set takes the lower half of the BN columns, the second the upper half.
        const int wq = warp & 3;
        const int chalf = (warp - 4) >> 2;
        const int lane = threadIdx.x & 31;
        int it = 0;
        for (TileIter<NC> t(n_tiles_m, n_tiles_n, crank); t.valid(); t.next(), ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = t.m0();
            const int n0 = t.n0(BN);
            const int row = m0 + wq * 32 + lane;
            const bool row_ok = row < M;
            // RoPE angles depend on the row only: fetch this row's 32 (cos, sin) pairs once per tile, BEFORE waiting for
            // the accumulator, so that the L2 round trip hides behind the main loop instead of stalling every head.
            float4 rope_row[16];
            bool has_rope = false;
            if constexpr (EPI == EPI_HEADS) {
                has_rope = row_ok && n0 < e.rope_cols;
                if (has_rope) {
                    const float4* rp = reinterpret_cast<const float4*>(e.rope + static_cast<size_t>(row) * 32);
#pragma unroll
                    for (int i = 0; i < 16; ++i) rope_row[i] = __ldg(rp + i);
                }
            }
            // Gated residual (HBM, not L2 resident): fetched before the wait too, in the TRANSPOSED pattern the stores
            // use below (8 lanes cover one 128-byte line of a row; lane = (row 4i + lane / 8, 16-byte chunk lane % 8)).
            constexpr int BPW_ = (BN / 64 + 1) / 2;          // 64-column blocks per warp set
            uint4 resid_t[EPI == EPI_GATE ? 8 * BPW_ : 1];
            if constexpr (EPI == EPI_GATE) {
#pragma unroll
                for (int c = 0; c < BPW_; ++c) {
                    const int cc = chalf * BPW_ + c;
                    if (cc < BN / 64) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = m0 + wq * 32 + 4 * i + (lane >> 3);
                            if (rr < M)
                                resid_t[8 * c + i] = *reinterpret_cast<const uint4*>(
                                    e.resid + static_cast<size_t>(rr) * e.ldr + n0 + cc * 64 + (lane & 7) * 8);
                        }
                    }
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(wq * 32) << 16);

            if constexpr (EPI == EPI_HEADS) {
                // one head (64 columns) at a time
                constexpr int HPW = (BN / 64 + 1) / 2;          // heads per warp set
                for (int c = chalf * HPW; c < BN / 64 && c < (chalf + 1) * HPW; ++c) {
                    uint32_t raw[64];
                    tmem_ld32(t_row + c * 64, raw);
                    tmem_ld32(t_row + c * 64 + 32, raw + 32);
                    tmem_wait_ld();
                    const int col0 = n0 + c * 64;
                    float x[64];
                    float ss = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        // bias / norm weights / rope are read 16 bytes at a time (all pointers are 16-byte aligned:
                        // col0 is a multiple of 64, the rope row is 256 bytes)
                        const float4 b = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + col0) + i)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                        x[4 * i + 0] = __uint_as_float(raw[4 * i + 0]) + b.x;
                        x[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + b.y;
                        x[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + b.z;
                        x[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + b.w;
                    }
                    // Rounding points of the reference (bf16 Linear output; bf16 after RMSNorm; bf16 after RoPE) are kept,
                    // but a value that is only rounded to be packed right away is rounded once, by the pack itself.
                    // norm_period > 0: the N axis is a stack of groups of norm_period columns (the cross-attention K | V
                    // projections of all visual blocks in one GEMM), each with its own 64-float norm weight
                    const int pc0 = e.norm_period > 0 ? col0 % e.norm_period : col0;
                    const int grp64 = e.norm_period > 0 ? (col0 / e.norm_period) * 64 : 0;
                    const bool normed = pc0 < e.norm_cols;
                    if (normed) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            bf16_round2(x[2 * i], x[2 * i + 1]);
                            ss = fmaf(x[2 * i], x[2 * i], ss);
                            ss = fmaf(x[2 * i + 1], x[2 * i + 1], ss);
                        }
                        const float4* w4 = reinterpret_cast<const float4*>(((pc0 < e.norm_split) ? e.norm_w0 : e.norm_w1) + grp64);
                        const float inv = rsqrtf(ss * (1.0f / 64.0f) + 1.1920928955078125e-07f);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 w = __ldg(w4 + i);
                            x[4 * i + 0] = x[4 * i + 0] * inv * w.x;
                            x[4 * i + 1] = x[4 * i + 1] * inv * w.y;
                            x[4 * i + 2] = x[4 * i + 2] * inv * w.z;
                            x[4 * i + 3] = x[4 * i + 3] * inv * w.w;
                        }
                        if (pc0 < e.rope_cols && has_rope) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float4 cs = rope_row[i];          // (cos, sin) of pairs 2i and 2i + 1
                                bf16_round2(x[4 * i], x[4 * i + 1]);
                                bf16_round2(x[4 * i + 2], x[4 * i + 3]);
                                const float a0 = x[4 * i], b0 = x[4 * i + 1], a1 = x[4 * i + 2], b1 = x[4 * i + 3];
                                // reference: (rope * x_).sum(-1): products rounded separately, then added
                                x[4 * i + 0] = __fadd_rn(__fmul_rn(cs.x, a0), __fmul_rn(-cs.y, b0));
                                x[4 * i + 1] = __fadd_rn(__fmul_rn(cs.y, a0), __fmul_rn(cs.x, b0));
                                x[4 * i + 2] = __fadd_rn(__fmul_rn(cs.z, a1), __fmul_rn(-cs.w, b1));
                                x[4 * i + 3] = __fadd_rn(__fmul_rn(cs.w, a1), __fmul_rn(cs.z, b1));
                            }
                        }
                    }
                    // Transpose the warp's 32 rows x 128 B through shared memory (XOR-swizzled 16-byte chunks, conflict
                    // free) so that 8 lanes cover one 128-byte line and every store instruction writes 4 full lines
                    // (a row-per-thread store touches 32 different lines per instruction).
                    uint4* xw = reinterpret_cast<uint4*>(xpose + (warp - 4) * 4096);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint4 v;
                        v.x = pack_bf16x2(x[8 * i + 0], x[8 * i + 1]);
                        v.y = pack_bf16x2(x[8 * i + 2], x[8 * i + 3]);
                        v.z = pack_bf16x2(x[8 * i + 4], x[8 * i + 5]);
                        v.w = pack_bf16x2(x[8 * i + 6], x[8 * i + 7]);
                        xw[lane * 8 + (i ^ (lane & 7))] = v;
                    }
                    __syncwarp();
                    const int chunk = lane & 7;
                    uint4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + (lane >> 3);
                        v[i] = xw[r * 8 + (chunk ^ (r & 7))];
                    }
                    __syncwarp();
                    if (e.peers.n > 0 && col0 >= e.peers.col0) {
                        // K | V columns of a temporal shard: all-gather fused into the epilogue - the lines go to every
                        // rank's K|V buffer (this rank's own included) over NVLink peer memory
                        const size_t colo = static_cast<size_t>(col0 - e.peers.col0) + chunk * 8;
                        for (int pr = 0; pr < e.peers.n; ++pr) {
                            bf16* base = e.peers.dst[pr];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int rr = m0 + wq * 32 + 4 * i + (lane >> 3);
                                if (rr < M)
                                    *reinterpret_cast<uint4*>(base + (static_cast<size_t>(e.peers.row0) + rr) * e.peers.ld + colo) = v[i];
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = m0 + wq * 32 + 4 * i + (lane >> 3);
                            if (rr < M)
                                *reinterpret_cast<uint4*>(e.out + static_cast<size_t>(rr) * e.ldo + col0 + chunk * 8) = v[i];
                        }
                    }
                }
            } else {
                // 64 columns at a time.  The accumulator arrives with thread = row; a row-per-thread store would touch 32
                // different 128-byte lines per instruction, so the bf16 values are transposed through the warp's
                // shared-memory patch (XOR-swizzled 16-byte chunks, conflict free) and leave as full lines: 8 lanes
                // per row, 4 rows per instruction.  The residual of the gated epilogue comes in the same pattern.
                constexpr int BPW = (BN / 64 + 1) / 2;          // 64-column blocks per warp set
                uint4* xw = reinterpret_cast<uint4*>(xpose + (warp - 4) * 4096);
                const int chunk = lane & 7;
#pragma unroll
                for (int cl = 0; cl < BPW; ++cl) {
                    const int c = chalf * BPW + cl;
                    if (c >= BN / 64) break;
                    uint32_t raw[64];
                    tmem_ld32(t_row + c * 64, raw);
                    tmem_ld32(t_row + c * 64 + 32, raw + 32);
                    tmem_wait_ld();
                    const int col0 = n0 + c * 64;
                    if constexpr (EPI == EPI_F32) {
                        // unrounded scores for a softmax (VAE mid-block attention): thread = row, 256 contiguous bytes
                        if (row_ok) {
                            float4* dst = reinterpret_cast<float4*>(e.out_f32 + static_cast<size_t>(row) * e.ldo + col0);
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + i);
                                dst[i] = make_float4(__uint_as_float(raw[4 * i]) + b.x, __uint_as_float(raw[4 * i + 1]) + b.y,
                                                     __uint_as_float(raw[4 * i + 2]) + b.z, __uint_as_float(raw[4 * i + 3]) + b.w);
                            }
                        }
                    } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float y[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(raw[8 * i + j]);
                        if (e.bias) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + 2 * i);
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + 2 * i + 1);
                            y[0] += b0.x; y[1] += b0.y; y[2] += b0.z; y[3] += b0.w;
                            y[4] += b1.x; y[5] += b1.y; y[6] += b1.z; y[7] += b1.w;
                        }
                        if constexpr (EPI == EPI_GELU) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                bf16_round2(y[2 * j], y[2 * j + 1]);
                                y[2 * j] = gelu_erf(y[2 * j]);
                                y[2 * j + 1] = gelu_erf(y[2 * j + 1]);
                            }
                        }
                        uint4 v;                                  // bf16(acc + bias) / bf16(gelu(bf16(acc + bias)))
                        v.x = pack_bf16x2(y[0], y[1]);
                        v.y = pack_bf16x2(y[2], y[3]);
                        v.z = pack_bf16x2(y[4], y[5]);
                        v.w = pack_bf16x2(y[6], y[7]);
                        xw[lane * 8 + (i ^ (lane & 7))] = v;
                    }
                    __syncwarp();
                    uint4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + (lane >> 3);
                        v[i] = xw[r * 8 + (chunk ^ (r & 7))];
                    }
                    __syncwarp();
                    float gg[8];
                    if constexpr (EPI == EPI_GATE) {
                        const float4 g0 = __ldg(reinterpret_cast<const float4*>(e.gate + col0 + chunk * 8));
                        const float4 g1 = __ldg(reinterpret_cast<const float4*>(e.gate + col0 + chunk * 8) + 1);
                        gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w;
                        gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = m0 + wq * 32 + 4 * i + (lane >> 3);
                        if (rr >= M) continue;
                        uint4 o = v[i];
                        if constexpr (EPI == EPI_GATE) {
                            // x + gate * bf16(acc + bias), products and sums rounded like the reference's fp32 ops (nn.py:30-33)
                            const uint4 r = resid_t[8 * cl + i];
                            const uint32_t rr4[4] = {r.x, r.y, r.z, r.w};
                            const uint32_t yy4[4] = {o.x, o.y, o.z, o.w};
                            uint32_t oo[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float lo = __fadd_rn(bf16_lo(rr4[j]), __fmul_rn(gg[2 * j], bf16_lo(yy4[j])));
                                const float hi = __fadd_rn(bf16_hi(rr4[j]), __fmul_rn(gg[2 * j + 1], bf16_hi(yy4[j])));
                                oo[j] = pack_bf16x2(lo, hi);
                            }
                            o = make_uint4(oo[0], oo[1], oo[2], oo[3]);
                        }
                        *reinterpret_cast<uint4*>(e.out + static_cast<size_t>(rr) * e.ldo + col0 + chunk * 8) = o;
                    }
                    }   // EPI != EPI_F32
                }
            }
            tc_fence_before();
            if constexpr (MMA2) mbar_arrive_leader(&tempty[acc]);
            else mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (NC > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast to it or free its stages
    if (warp == 2) {
        tc_fence_after();
        if constexpr (MMA2) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
        else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int BN, int EPI, int CL>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const GemmEpilogue& e, cudaStream_t st) {
    using Cfg = GemmCfg<BN, CL>;
    constexpr int NC = CL == 1 ? 1 : 2;
    static PerDevice<int> pd;          // CTAs that can be resident at once (whole clusters only when NC > 1), per device
    auto kern = gemm_bf16_kernel<BN, EPI, CL>;
    const int dev = current_device();
    std::lock_guard<std::mutex> lk(pd.m);
    int& max_ctas = pd.v[dev];
    if (!pd.set[dev]) {
        K5_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        max_ctas = sm_count();
        if (NC > 1) {
            // clusters live inside one GPC; a GPC with an odd SM count leaves one SM out
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(static_cast<unsigned>(sm_count() / NC * NC));
            q.blockDim = dim3(GEMM_THREADS);
            q.dynamicSmemBytes = Cfg::SMEM_BYTES;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = NC;
            qa[0].val.clusterDim.y = 1;
            qa[0].val.clusterDim.z = 1;
            q.attrs = qa;
            q.numAttrs = 1;
            int n_clusters = 0;
            K5_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, kern, &q));
            K5_REQUIRE(n_clusters > 0, "GEMM: no thread-block cluster fits this device");
            max_ctas = n_clusters * NC;
        }
        pd.set[dev] = true;
    }
    const int units = (((M + BM - 1) / BM + NC - 1) / NC) * (N / BN) * NC;
    const int grid = units < max_ctas ? units : max_ctas;
    if (NC == 1) {
        kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, M, N, K, e);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(grid));
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = NC;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        K5_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, M, N, K, e));
    }
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

template <int BN, int CL>
int launch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const GemmEpilogue& e,
               cudaStream_t st) {
    switch (epi) {
        case EPI_STORE: return launch<BN, EPI_STORE, CL>(tmA, tmB, M, N, K, e, st);
        case EPI_GELU: return launch<BN, EPI_GELU, CL>(tmA, tmB, M, N, K, e, st);
        case EPI_GATE: return launch<BN, EPI_GATE, CL>(tmA, tmB, M, N, K, e, st);
        case EPI_HEADS: return launch<BN, EPI_HEADS, CL>(tmA, tmB, M, N, K, e, st);
        case EPI_F32: return launch<BN, EPI_F32, CL>(tmA, tmB, M, N, K, e, st);
    }
    set_last_error("unknown GEMM epilogue");
    return K5_ERR_INVALID;
}

}  // namespace

int gemm_bf16(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, int epi, const GemmEpilogue& e,
              cudaStream_t st) {
    K5_REQUIRE(M > 0 && N > 0 && K > 0, "GEMM: empty problem");
    K5_REQUIRE(N % 64 == 0, "GEMM: N must be a multiple of 64");
    K5_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "GEMM: K / pitches must be multiples of 8 elements");
    K5_REQUIRE((epi == EPI_F32 ? static_cast<const void*>(e.out_f32) : static_cast<const void*>(e.out)) != nullptr && e.ldo % 8 == 0,
               "GEMM: output pitch must be a multiple of 8 elements");
    K5_REQUIRE(epi != EPI_F32 || (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0, "GEMM: fp32 output must be 16-byte aligned");
    if (epi == EPI_GATE) K5_REQUIRE(e.resid && e.gate && e.ldr % 8 == 0, "GEMM: gate epilogue needs resid/gate");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    K5_REQUIRE(al16(e.bias) && al16(e.gate) && al16(e.norm_w0) && al16(e.norm_w1) && al16(e.rope),
               "GEMM: bias / gate / norm weights / rope must be 16-byte aligned");
    if (epi == EPI_HEADS) {
        K5_REQUIRE(e.norm_cols % 64 == 0 && e.norm_split % 64 == 0 && e.rope_cols % 64 == 0, "GEMM: head split must be x64");
        K5_REQUIRE(e.norm_cols == 0 || (e.norm_w0 && e.norm_w1), "GEMM: head epilogue needs norm weights");
        K5_REQUIRE(e.rope_cols == 0 || e.rope, "GEMM: head epilogue needs a rope table");
        K5_REQUIRE(e.norm_period >= 0 && e.norm_period % 64 == 0 && (e.norm_period == 0 || e.peers.n == 0),
                   "GEMM: the norm period must be x64 and excludes the peer scatter");
        K5_REQUIRE(e.peers.n >= 0 && e.peers.n <= MAX_PEERS, "GEMM: at most 8 scatter destinations");
        K5_REQUIRE(e.peers.n == 0 || (e.peers.col0 % 64 == 0 && e.peers.ld % 8 == 0), "GEMM: scatter split must be x64");
    } else {
        K5_REQUIRE(e.peers.n == 0, "GEMM: the peer scatter belongs to the head epilogue");
    }
    const int BN = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
    // CTA pairs pay as soon as there are two M tiles to pair (K5_GEMM_CLUSTER = 0 / 1 / 3 overrides: tuning)
    static int cluster_env = -1;
    if (cluster_env < 0) {
        const char* ev = getenv("K5_GEMM_CLUSTER");
        cluster_env = ev ? (atoi(ev) == 3 ? 3 : (atoi(ev) != 0 ? 1 : 0)) : 2;
    }
    const bool pair = cluster_env == 2 ? (GEMM_CLUSTER_DEFAULT != 0 && M > BM) : (cluster_env >= 1);
    CUtensorMap tmA, tmB;
    K5_TRY(make_tmap_2d_bf16(&tmA, A, M, K, lda, BM));
    K5_TRY(make_tmap_2d_bf16(&tmB, W, N, K, ldw, pair ? BN / 2 : BN));
    if (pair && (cluster_env == 3 || (cluster_env == 2 && GEMM_CLUSTER_DEFAULT == 3))) {   // 2-SM MMA
        if (BN == 256) return launch_epi<256, 3>(epi, tmA, tmB, M, N, K, e, st);
        if (BN == 128) return launch_epi<128, 3>(epi, tmA, tmB, M, N, K, e, st);
        return launch_epi<64, 3>(epi, tmA, tmB, M, N, K, e, st);
    }
    if (pair) {
        if (BN == 256) return launch_epi<256, 2>(epi, tmA, tmB, M, N, K, e, st);
        if (BN == 128) return launch_epi<128, 2>(epi, tmA, tmB, M, N, K, e, st);
        return launch_epi<64, 2>(epi, tmA, tmB, M, N, K, e, st);
    }
    if (BN == 256) return launch_epi<256, 1>(epi, tmA, tmB, M, N, K, e, st);
    if (BN == 128) return launch_epi<128, 1>(epi, tmA, tmB, M, N, K, e, st);
    return launch_epi<64, 1>(epi, tmA, tmB, M, N, K, e, st);
}

}  // namespace k5
