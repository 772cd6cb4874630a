#pragma once
#include "common.h"

namespace k5 {

enum : int { EPI_STORE = 0, EPI_GELU = 1, EPI_GATE = 2, EPI_HEADS = 3, EPI_F32 = 4 };

// Fused all-gather of the temporal shard (EPI_HEADS only): output columns >= col0 (the K | V part of the fused
// QKV projection) are not written to `out` but to row (row0 + m), column (n - col0) of every destination in
// dst[0..n) -- the [S, 2D] K|V buffers of all ranks of the node, peers mapped through CUDA IPC, own buffer
// included.  The stores are plain st.global over NVLink, staged through shared memory so that every warp
// instruction writes whole 128-byte lines.
constexpr int MAX_PEERS = 8;
struct PeerScatter {
    bf16* dst[MAX_PEERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    int ld = 0;
    int col0 = 0;
    long long row0 = 0;
};

struct GemmEpilogue {
    bf16* out = nullptr;            // [M, N] row-major, pitch ldo
    float* out_f32 = nullptr;       // EPI_F32: the fp32 accumulators (+ bias) are stored unrounded, pitch ldo (elements)
    int ldo = 0;
    const float* bias = nullptr;    // [N] (bf16-representable values held in fp32) or null
    // EPI_GATE: out = bf16(resid + gate * bf16(acc + bias)); out may alias resid
    const bf16* resid = nullptr;
    int ldr = 0;
    const float* gate = nullptr;    // [N] fp32
    // EPI_HEADS (head_dim 64): columns [0, norm_cols) get per-head RMSNorm with weight norm_w0 for
    // columns < norm_split and norm_w1 otherwise; columns [0, rope_cols) then get RoPE.
    const float* norm_w0 = nullptr;
    const float* norm_w1 = nullptr;
    // norm_period > 0: N is a stack of groups of norm_period columns; the three column limits then apply to the
    // column WITHIN its group, and group g uses norm_w0 + 64 g / norm_w1 + 64 g (weights stacked [groups, 64]).
    int norm_split = 0;
    int norm_cols = 0;
    int rope_cols = 0;
    int norm_period = 0;
    const float2* rope = nullptr;   // [M, 32] (cos, sin) per row and rotation pair
    PeerScatter peers;              // n == 0: everything goes to `out`
};

// C[M,N] = A[M,K] . W[N,K]^T with a fused epilogue; A, W bf16 row-major (K contiguous).
int gemm_bf16(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, int epi, const GemmEpilogue& e,
              cudaStream_t st);

}  // namespace k5
