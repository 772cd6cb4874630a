#pragma once
#include "common.h"
#include <cuda_fp16.h>

namespace k5 {

// dtype codes used across the C ABI: 0 = float32, 1 = bfloat16, 2 = float16
int ln_rows(const bf16* x, int ldx, bf16* out, int ldo, int S, int D, const float* mul, const float* add, bool plus_one,
            float eps, cudaStream_t st);
int gemv_f32(const float* W, const float* bias, const float* x, float* out, int N, int K, bool silu_in, bool silu_out,
             cudaStream_t st);
int time_features(const float* freqs, float t, float* out, int half, cudaStream_t st);
int pooled_embed(const bf16* W, const float* bias, const float* lnw, const float* lnb, const bf16* pooled, int Kin,
                 int time_dim, float* time_embed, float eps, cudaStream_t st);
int patchify(const float* x, int Cx, int C, int T, int Hp, int Wp, bool fractal, bf16* A, int KP, cudaStream_t st);
int unpatchify(const bf16* y, int ldy, int T, int Hp, int Wp, bool fractal, int Cout, bf16* out, cudaStream_t st);
int rope3d_table(const float* at, const float* ah, const float* aw, int nt, int nh, int nw, const int* pt, const int* ph,
                 const int* pw, const float sf[3], int T, int Hp, int Wp, bool fractal, float2* table, cudaStream_t st);
int rope1d_table(const float* args, int np, const int* pos, int L, float2* table, cudaStream_t st);
int cfg_combine(const bf16* vc, const bf16* vu, float w, bf16* out, size_t n, cudaStream_t st);
int euler_step(float* img, const bf16* v, float dt, size_t n, cudaStream_t st);
int bf16_addsub(const bf16* a, const bf16* b, bf16* out, size_t n, bool subtract, cudaStream_t st);
int convert_to_bf16(const void* src, int src_dtype, bf16* dst, int rows, int cols, int ld_dst, cudaStream_t st);
int convert_to_f32(const void* src, int src_dtype, float* dst, size_t n, bool round_bf16, cudaStream_t st);

}  // namespace k5
