#pragma once
#include "common.h"

namespace k5 {

// HBM-bound passes of the VAE decoder (kandinsky/models/vae.py); channels-last bf16 activations [P, C].

// Per-channel sum and sum of squares over P positions, in double, as one row of partials per thread block:
// sums[b][0..C) and sums[b][C..2C) for b < *nblocks <= GN_MAX_BLOCKS (no atomics: bit-reproducible).
constexpr int GN_MAX_BLOCKS = 1184;          // 8 blocks on each of 148 SMs
int gn_channel_sums(const bf16* x, size_t P, int C, double* sums, int* nblocks, cudaStream_t st);
// GroupNorm(32 groups, eps) statistics from the block partials (fixed summation order): mean_rstd[2g] = mean, [2g + 1] = rstd.
int gn_finalize(const double* sums, int nblocks, size_t P, int C, int groups, float eps, float* mean_rstd, cudaStream_t st);

// The gather that feeds every 3x3x3 convolution: builds the replicate-padded (2 frames in front, 1 pixel around),
// optionally nearest-up-sampled, optionally GroupNorm + SiLU'd copy of x.
//   x    : [Ts, Hs, Ws, C]            source volume
//   out  : [T + 2, H + 2, W + 2, C]   with (T, H, W) = up-sampled extent: H = Hs * fh, W = Ws * fw,
//          T = Ts if ft == 1 else 1 + (Ts - 1) * 2  (first frame is never repeated in time: vae.py:187-205)
//   mean_rstd / gamma / beta : GroupNorm(32) statistics and affine (null = plain copy); silu: apply SiLU after it
int pad_gather(const bf16* x, int Ts, int Hs, int Ws, int C, int ft, int fh, int fw, const float* mean_rstd,
               const float* gamma, const float* beta, int groups, bool silu, bf16* out, cudaStream_t st);
// GroupNorm without padding (the attention's group_norm): out[P, C] = bf16(GN(x))
int gn_apply(const bf16* x, size_t P, int C, const float* mean_rstd, const float* gamma, const float* beta, int groups,
             bool silu, bf16* out, cudaStream_t st);

// post_quant_conv (1x1x1, 16 -> 16; vae.py:874) on the fp32 latent z [C=16, T, H, W] (the reference's NCTHW layout),
// written as the padded channels-last input of conv_in: out [T + 2, H + 2, W + 2, 64], channels 16..63 zero.
int post_quant_pad(const float* z, int Cz, int T, int H, int W, int t0, int Tz, const float* w, const float* b, bf16* out,
                   cudaStream_t st);

// Masked softmax of the mid-block attention scores (vae.py:110-122, 343-359): row r of frame f = r / hw keeps columns
// [0, (f + 1) * hw); p <- bf16(softmax(s * scale)) over them, s fp32 (unrounded GEMM accumulators), all math fp32.
int softmax_frame_causal(const float* s, int N, int lds, bf16* p, int ldp, int hw, float scale, int row0, int rows,
                         cudaStream_t st);

// Temporal tile assembly (vae.py:1144-1204, 928-936).  cur / prev: decoded tiles, channels-last [F, H, W, 3] bf16.
// Writes `count` frames starting at local frame `src0` of cur into frames [dst0, dst0 + count) of the NCTHW bf16 output
// [3, Fout, H, W]; the first `blend` of them are blended with prev's frames [prev0, prev0 + blend):
//   out = bf16(bf16(prev * (1 - x / blend)) + bf16(cur * (x / blend))).
int emit_frames(const bf16* cur, int src0, const bf16* prev, int prev0, int blend, int count, int H, int W, int Fout,
                int dst0, bf16* out, cudaStream_t st);

// out[c, r] = in[r, c]  (bf16; in [R, C] with pitch ldi, out [C, R] with pitch ldo)
int transpose_bf16(const bf16* in, int R, int C, int ldi, bf16* out, int ldo, cudaStream_t st);

// Conv weight [Cout, Cin, taps] (any of f32 / bf16 / f16, the checkpoint layout with the 3x3x3 taps innermost) ->
// bf16 [Cout_pad, taps, Cin_pad] (tap-major, channel-minor; padding zero-filled by the caller's memset).
int repack_conv_weight(const void* src, int dtype, int Cout, int Cin, int taps, int Cin_pad, bf16* dst, cudaStream_t st);

}  // namespace k5
