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
#include "common.h"
#include "conv3d.h"
#include "ptx.cuh"

namespace k5 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;

// MT = M tiles (128 positions each) that share one weight tile: with 128 output channels a 128 x 128 tile pulls
// 128 B / clk / SM of operands out of L2 (measured 594 TFLOP/s); two position tiles per weight tile cut the
// weight traffic and the TMA requests per FLOP to the level of the 128 x 256 tile.
template <int BN, int MT>
struct ConvCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
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

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}

template <int BN, int MT>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ConvParams p) {
    using Cfg = ConvCfg<BN, MT>;
    constexpr int STAGES = Cfg::STAGES;
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
    const int num_tiles = n_groups_m * n_tiles_n;
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
            mbar_init(&tempty[s], 128 * MT);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: implicit im2col =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mg = tile / n_tiles_n;
                const int n0 = (tile % n_tiles_n) * BN;
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
                        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        // padded coordinates: output (t, h, w), tap (kt, kh, kw) reads xpad[t + kt, h + kh, w + kw]
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                            tma_load_4d(sa + m * Cfg::A_BYTES, &tmX, &full[stage], cb * BK, w0[m] + kw, h0[m] + kh, t[m] + kt);
                        tma_load_2d(sa + MT * Cfg::A_BYTES, &tmW, &full[stage], kb * BK, n0);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
                        for (int m = 0; m < MT; ++m)
                            umma_ss(d_tmem + m * BN, umma_desc_sw128(sa + m * Cfg::A_BYTES + k * 32, 0, 1024), bd, idesc,
                                    (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else if (warp >= 4 && warp < 4 + 4 * MT) {
        // ===================== epilogue (4 warps per M tile, thread = output position) =====================
        const int wq = warp & 3;
        const int msel = (warp - 4) >> 2;                         // which of the MT position tiles
        const int lane = threadIdx.x & 31;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int mt = (tile / n_tiles_n) * MT + msel;
            const int n0 = (tile % n_tiles_n) * BN;
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
            mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
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

template <int BN, int MT>
int launch_conv(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvParams& p, cudaStream_t st) {
    using Cfg = ConvCfg<BN, MT>;
    static bool configured = false;
    auto kern = conv3d_kernel<BN, MT>;
    if (!configured) {
        K5_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int mtiles = ((p.T + p.bt - 1) / p.bt) * p.tiles_w * p.tiles_h;
    const int tiles = ((mtiles + MT - 1) / MT) * ((p.Cout + BN - 1) / BN);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, CONV_THREADS, Cfg::SMEM_BYTES, st>>>(tmX, tmW, p);
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
    K5_TRY(make_tmap_2d_bf16(&tmW, w, Cout_pad, static_cast<uint64_t>(27) * Cin, static_cast<uint64_t>(27) * Cin, BN));
    if (BN == 256) return launch_conv<256, 1>(tmX, tmW, p, st);
    if (BN == 128) return launch_conv<128, 2>(tmX, tmW, p, st);
    return launch_conv<64, 2>(tmX, tmW, p, st);
}

}  // namespace k5
