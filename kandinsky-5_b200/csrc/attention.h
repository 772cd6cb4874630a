#pragma once
#include "common.h"

namespace k5 {

struct AttnParams {
    int Sq, Sk, heads;
    float scale_log2;      // softmax scale * log2(e)
    bf16* out;             // [Sq, heads*64], pitch ldo
    int ldo;
    // optional block-sparse KV lists over 64x64 blocks (NABLA); null = dense
    const int32_t* kv_count;   // [heads, Sq/64]
    const int32_t* kv_index;   // [heads, Sq/64, Sk/64], first kv_count entries valid, ascending
};

// O[Sq, heads*64] = softmax(Q K^T * softmax_scale) V per head, head_dim 64, non-causal.
// Q/K/V/O are row-major token matrices whose head h occupies columns [h*64, h*64+64).
int attention_fwd(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo, int Sq,
                  int Sk, int heads, float softmax_scale, const int32_t* kv_count, const int32_t* kv_index,
                  cudaStream_t st);

}  // namespace k5
