// Causal 3x3x3 convolution of the HunyuanVideo VAE decoder as an implicit GEMM on the 5th-gen tensor cores.
//
// Replaces HunyuanVideoCausalConv3d (kandinsky/models/vae.py:125-163: F.pad replicate + nn.Conv3d) for every 3x3x3
// convolution of the decoder (conv_in, the two convs of each of the 14 resnets, the 3 up-sampler convs, conv_out:
// 98 % of the VAE FLOPs).  Activations are channels-last (NDHWC) bf16.  The GEMM view is
//   M = T*H*W output positions, N = Cout, K = 27 taps x Cin,
// and no im2col matrix ever exists: the producer warp walks the K dimension tap by tap and, for each 64-channel slice,
// issues ONE 4-D TMA box load {64 ch, bw, bh, 1 frame} at the tap-shifted coordinate of the padded input; the box lands
// in shared memory as 128 rows x 128 B with the 128-byte swizzle, i.e. exactly the K-major A operand tcgen05.mma
// wants.  The M tile is therefore a bh x bw patch of one frame (bw * bh = 128) rather than 128 consecutive rows.
// Everything else is the pipeline of gemm.cu: multi-stage TMA ring, one elected MMA thread, fp32 accumulators
// double-buffered in TMEM, 4 epilogue warps (thread = output position) that add the bias, round to bf16, add the
// residual (bf16 + bf16 like the reference's tensor add) and store channels-last.
#include <cstdlib>

#include "common.h"
#include "conv3d.h"
#include "ptx.cuh"

namespace k5 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int CONV_CLUSTER_DEFAULT = 1;   // VAE decode -5.4 % (profiles/r1_gemm_cluster.md)

// MT = M tiles (128 positions each) that share one weight tile: with 128 output channels a 128 x 128 tile pulls
// 128 B / clk / SM of operands out of L2 (measured 594 TFLOP/s); two position tiles per weight tile cut the
// weight traffic and the TMA requests per FLOP to the level of the 128 x 256 tile.
// CL: 1 = single CTA; 3 = CTA pair driven by ONE cta_group::2 MMA (M = 256: each CTA supplies its 128 positions and
// keeps only its half of the weight tile), as in gemm.cu.
template <int BN, int MT, int CL = 1>
struct ConvCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (CL == 3 ? BN / 2 : BN) * BK * 2;    // bytes of the weight tile held per CTA and stage
    static constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
    static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) < 8 ? (200 * 1024 / STAGE_BYTES) : 8;
    static constexpr int TMEM_COLS = 2 * MT * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(TMEM_COLS <= 512, "accumulators must fit TMEM");
};
constexpr int CONV_THREADS = 384;      // warps 0-2: TMA / MMA / TMEM alloc; warps 4-7, 8-11: epilogue of M tile 0, 1

struct ConvParams {
    int T, H, W, Cin, Cout;
    int bw, bh, bt;             // M tile = bt frames x bh rows x bw columns (bw * bh * bt = 128)
    int tiles_w, tiles_h;       // W / bw, H / bh; ceil(T / bt) tiles along time (frames >= T are masked)
    const float* bias;
    const bf16* resid;
    int ldr;
    bf16* out;
    int ldo;
};

// 2-SM form: lands in this CTA's shared memory, completes on the leader CTA's mbarrier (ptx.cuh)
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & K5_PEER_BIT_MASK), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}

template <int BN, int MT, int CL>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ConvParams p) {
    using Cfg = ConvCfg<BN, MT, CL>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NC = CL == 1 ? 1 : 2;          // CTAs per cluster
    constexpr bool MMA2 = CL == 3;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int n_tiles_n = (p.Cout + BN - 1) / BN;
    const int tiles_per_frame = p.tiles_w * p.tiles_h;
    const int n_tiles_m = ((p.T + p.bt - 1) / p.bt) * tiles_per_frame;
    const int n_groups_m = (n_tiles_m + MT - 1) / MT;         // groups of MT consecutive position tiles
    // work unit of a cluster = NC consecutive groups x one weight tile; CTA `crank` takes group unit_m * NC + crank
    const int num_units = ((n_groups_m + NC - 1) / NC) * n_tiles_n;
    const int crank = NC > 1 ? static_cast<int>(cluster_ctarank()) : 0;
    const int unit0 = blockIdx.x / NC, unit_step = gridDim.x / NC;
    const int cblocks = p.Cin / BK;
    const int nkb = 27 * cblocks;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 128 * MT * NC);      // 2-SM MMA: the leader's barrier collects both CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (MMA2) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
        else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (NC > 1) cluster_sync_all();    // the peer's barriers exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: implicit im2col =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = unit0; unit < num_units; unit += unit_step) {
                const int mg = (unit / n_tiles_n) * NC + crank;
                const int n0 = (unit % n_tiles_n) * BN;
                int t[MT], h0[MT], w0[MT];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const int mt = mg * MT + m;                   // may run past the volume: TMA zero-fills, epilogue skips
                    const int r = mt % tiles_per_frame;
                    t[m] = (mt / tiles_per_frame) * p.bt;
                    h0[m] = (r / p.tiles_w) * p.bh;
                    w0[m] = (r % p.tiles_w) * p.bw;
                }
                int kb = 0;
                for (int tap = 0; tap < 27; ++tap) {
                    const int kt = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                    for (int cb = 0; cb < cblocks; ++cb, ++kb) {
                        mbar_wait_parked(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        // padded coordinates: output (t, h, w), tap (kt, kh, kw) reads xpad[t + kt, h + kh, w + kw]
                        if constexpr (MMA2) {
                            // the bytes of both CTAs are counted on the leader's barrier, which only the leader arms
                            if (crank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
#pragma unroll
                            for (int m = 0; m < MT; ++m)
                                tma_load_4d_2sm(sa + m * Cfg::A_BYTES, &tmX, &full[stage], cb * BK, w0[m] + kw, h0[m] + kh,
                                                t[m] + kt);
                            tma_load_2d_2sm(sa + MT * Cfg::A_BYTES, &tmW, &full[stage], kb * BK, n0 + crank * (BN / 2));
                        } else {
                            mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
#pragma unroll
                            for (int m = 0; m < MT; ++m)
                                tma_load_4d(sa + m * Cfg::A_BYTES, &tmX, &full[stage], cb * BK, w0[m] + kw, h0[m] + kh, t[m] + kt);
                            tma_load_2d(sa + MT * Cfg::A_BYTES, &tmW, &full[stage], kb * BK, n0);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
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
            for (int unit = unit0; unit < num_units; unit += unit_step, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_parked(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * MT * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_parked(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + MT * Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t bd = umma_desc_sw128(sb + k * 32, 0, 1024);
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            const uint64_t ad = umma_desc_sw128(sa + m * Cfg::A_BYTES + k * 32, 0, 1024);
                            if constexpr (MMA2) umma_ss_2sm(d_tmem + m * BN, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                            else umma_ss(d_tmem + m * BN, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    if constexpr (MMA2) umma_commit_2sm_mc(&empty[stage], static_cast<uint16_t>(3));
                    else umma_commit(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if constexpr (MMA2) umma_commit_2sm_mc(&tfull[acc], static_cast<uint16_t>(3));
                else umma_commit(&tfull[acc]);
            }
        }
    } else if (warp >= 4 && warp < 4 + 4 * MT) {
        // ===================== epilogue (4 warps per M tile, thread = output position) =====================
        const int wq = warp & 3;
        const int msel = (warp - 4) >> 2;                         // which of the MT position tiles
        const int lane = threadIdx.x & 31;
        int it = 0;
        for (int unit = unit0; unit < num_units; unit += unit_step, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int mt = ((unit / n_tiles_n) * NC + crank) * MT + msel;
            const int n0 = (unit % n_tiles_n) * BN;
            const int r = mt % tiles_per_frame;
            const int row = wq * 32 + lane;                       // row of the tile = (tl * bh + hl) * bw + wl
            const int t = (mt / tiles_per_frame) * p.bt + row / (p.bw * p.bh);
            const int h = (r / p.tiles_w) * p.bh + (row / p.bw) % p.bh;
            const int w = (r % p.tiles_w) * p.bw + row % p.bw;
            const bool row_ok = t < p.T && mt < n_tiles_m;
            const size_t pos = (static_cast<size_t>(row_ok ? t : 0) * p.H + h) * p.W + w;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (acc * MT + msel) * BN + (static_cast<uint32_t>(wq * 32) << 16);
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n0 + c * 32;
                if (col0 >= p.Cout) break;                        // warp-uniform
                uint32_t raw[32];
                tmem_ld32(t_row + c * 32, raw);
                tmem_wait_ld();
                bf16* dst = p.out + pos * p.ldo + col0;
                const bf16* res = p.resid ? p.resid + pos * p.ldr + col0 : nullptr;
                if (!row_ok) {
                    // frame past the end of the volume (time-masked tile): nothing to store
                } else if (col0 + 32 <= p.Cout && (p.ldo & 7) == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float y[8];
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + 2 * i);
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + 2 * i + 1);
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(raw[8 * i + j]) + bb[j];
                        if (res) {
                            // conv output is a bf16 tensor before the residual add (vae.py:274): round, then add
                            const uint4 rv = *reinterpret_cast<const uint4*>(res + 8 * i);
                            const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                bf16_round2(y[2 * j], y[2 * j + 1]);
                                y[2 * j] = __fadd_rn(y[2 * j], bf16_lo(rr[j]));
                                y[2 * j + 1] = __fadd_rn(y[2 * j + 1], bf16_hi(rr[j]));
                            }
                        }
                        uint4 o;
                        o.x = pack_bf16x2(y[0], y[1]);
                        o.y = pack_bf16x2(y[2], y[3]);
                        o.z = pack_bf16x2(y[4], y[5]);
                        o.w = pack_bf16x2(y[6], y[7]);
                        *reinterpret_cast<uint4*>(dst + 8 * i) = o;
                    }
                } else {
                    // ragged channel count (conv_out: 3 channels): scalar stores
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (col0 + j < p.Cout) {
                            float y = bf16_round(__uint_as_float(raw[j]) + __ldg(p.bias + col0 + j));
                            if (res) y = __fadd_rn(y, __bfloat162float(res[j]));
                            dst[j] = __float2bfloat16_rn(y);
                        }
                    }
                }
            }
            tc_fence_before();
            if constexpr (MMA2) mbar_arrive_leader(&tempty[acc]);
            else mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (NC > 1) cluster_sync_all();    // no CTA leaves while the leader may still free its stages
    if (warp == 2) {
        tc_fence_after();
        if constexpr (MMA2) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
        else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_4d(CUtensorMap* out, const void* base, uint64_t C, uint64_t Wp, uint64_t Hp, uint64_t Tp, uint32_t bw,
                 uint32_t bh, uint32_t bt) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            set_last_error("cuTensorMapEncodeTiled entry point not available");
            return K5_ERR_CUDA;
        }
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    K5_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "conv3d: input must be 16-byte aligned");
    cuuint64_t dims[4] = {C, Wp, Hp, Tp};
    cuuint64_t strides[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
    cuuint32_t box[4] = {64, bw, bh, bt};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled (4-D) failed with CUresult " + std::to_string(static_cast<int>(r)));
        return K5_ERR_CUDA;
    }
    return K5_OK;
}

template <int BN, int MT, int CL>
int launch_conv(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvParams& p, cudaStream_t st) {
    using Cfg = ConvCfg<BN, MT, CL>;
    constexpr int NC = CL == 1 ? 1 : 2;
    static PerDevice<int> pd;           // CTAs that can be resident at once (whole clusters only when NC > 1), per device
    auto kern = conv3d_kernel<BN, MT, CL>;
    const int dev = current_device();
    std::lock_guard<std::mutex> lk(pd.m);
    int& max_ctas = pd.v[dev];
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(CONV_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (!pd.set[dev]) {
        K5_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        max_ctas = sm_count();
        if (NC > 1) {
            cfg.gridDim = dim3(static_cast<unsigned>(sm_count() / NC * NC));
            int n_clusters = 0;
            K5_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg));
            K5_REQUIRE(n_clusters > 0, "conv3d: no thread-block cluster fits this device");
            max_ctas = n_clusters * NC;
        }
        pd.set[dev] = true;
    }
    const int mtiles = ((p.T + p.bt - 1) / p.bt) * p.tiles_w * p.tiles_h;
    const int groups = (mtiles + MT - 1) / MT;
    const int units = ((groups + NC - 1) / NC) * ((p.Cout + BN - 1) / BN) * NC;
    cfg.gridDim = dim3(static_cast<unsigned>(units < max_ctas ? units : max_ctas));
    K5_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmX, tmW, p));
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

}  // namespace

int conv3d_causal(const bf16* xpad, int T, int H, int W, int Cin, const bf16* w, int Cout, int Cout_pad, const float* bias,
                  const bf16* resid, int ldr, bf16* out, int ldo, cudaStream_t st) {
    K5_REQUIRE(T > 0 && H > 0 && W > 0, "conv3d: empty volume");
    K5_REQUIRE(Cin % 64 == 0 && Cout_pad % 64 == 0 && Cout > 0 && Cout <= Cout_pad, "conv3d: channels must be padded to x64");
    K5_REQUIRE(bias != nullptr && out != nullptr && xpad != nullptr && w != nullptr, "conv3d: null tensor");
    // 128-position M tile = bt frames x bh rows x bw columns, powers of two dividing W and H (time is masked instead)
    int bw = 128;
    while (bw > 1 && W % bw != 0) bw >>= 1;
    int bh = 128 / bw;
    while (bh > 1 && H % bh != 0) bh >>= 1;
    const int bt = 128 / (bw * bh);
    K5_REQUIRE(bh <= 256 && bt <= 256, "conv3d: H x W has too few power-of-two factors for a 128-position tile");
    ConvParams p;
    p.T = T;
    p.H = H;
    p.W = W;
    p.Cin = Cin;
    p.Cout = Cout;
    p.bw = bw;
    p.bh = bh;
    p.bt = bt;
    p.tiles_w = W / bw;
    p.tiles_h = H / bh;
    p.bias = bias;
    p.resid = resid;
    p.ldr = ldr;
    p.out = out;
    p.ldo = ldo;
    CUtensorMap tmX, tmW;
    K5_TRY(make_tmap_4d(&tmX, xpad, Cin, W + 2, H + 2, T + 2, bw, bh, bt));
    const int BN = (Cout_pad % 256 == 0) ? 256 : (Cout_pad % 128 == 0 ? 128 : 64);
    // CTA pairs under one 2-SM MMA (K5_CONV_CLUSTER=0 restores the single-CTA kernel: tuning / A-B)
    static int pair = -1;
    if (pair < 0) {
        const char* ev = getenv("K5_CONV_CLUSTER");
        pair = ev ? (atoi(ev) != 0) : CONV_CLUSTER_DEFAULT;
    }
    K5_TRY(make_tmap_2d_bf16(&tmW, w, Cout_pad, static_cast<uint64_t>(27) * Cin, static_cast<uint64_t>(27) * Cin,
                             pair ? BN / 2 : BN));
    if (pair) {
        if (BN == 256) return launch_conv<256, 1, 3>(tmX, tmW, p, st);
        if (BN == 128) return launch_conv<128, 2, 3>(tmX, tmW, p, st);
        return launch_conv<64, 2, 3>(tmX, tmW, p, st);
    }
    if (BN == 256) return launch_conv<256, 1, 1>(tmX, tmW, p, st);
    if (BN == 128) return launch_conv<128, 2, 1>(tmX, tmW, p, st);
    return launch_conv<64, 2, 1>(tmX, tmW, p, st);
}

}  // namespace k5
