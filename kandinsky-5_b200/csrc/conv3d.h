#pragma once
#include "common.h"

namespace k5 {

// Causal 3x3x3 convolution (HunyuanVideoCausalConv3d, kandinsky/models/vae.py:125-163) as an implicit GEMM on
// tcgen05: out[t,h,w,:] = sum over the 27 taps of  W_tap[Cout,Cin] . xpad[t+kt, h+kh, w+kw, :]  (+ bias, + residual).
//   xpad : bf16 NDHWC [T+2, H+2, W+2, Cin]  the input ALREADY padded (replicate; 2 frames in front, none behind), as
//          written by pad_gather (vae_ops.cu); Cin a multiple of 64
//   w    : bf16 [Cout_pad, 27 * Cin]  (tap-major, channel-minor), Cout_pad a multiple of 64, zero rows beyond Cout
//   out  : bf16 NDHWC [T, H, W, ldo]; only the first `Cout` channels of a position are written
//   resid: optional bf16 [T*H*W, ldr]: out = bf16(bf16(acc + bias) + resid)   (vae.py:274, bf16 tensor add)
// The 128-position M tile is a (bt frames x bh rows x bw columns) patch with power-of-two factors of W and H (frames
// past T are masked), so every decoder resolution (64x96 ... 512x768) and the 8x8 ... 64x64 test volumes qualify.
int conv3d_causal(const bf16* xpad, int T, int H, int W, int Cin, const bf16* w, int Cout, int Cout_pad, const float* bias,
                  const bf16* resid, int ldr, bf16* out, int ldo, cudaStream_t st);

}  // namespace k5
