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
    // derived per-item lists (built by a pre-pass from kv_count / kv_index): item = head * n_qpairs + qpair
    const int32_t* item_count; // [items]            number of 128-row KV tiles the item visits
    const int32_t* item_pairs; // [items, max_pairs] KV tile ids (ascending)
    const uint8_t* item_mask;  // [items, max_pairs] bit (qblk*2 + half): 64x64 sub-block selected
    int max_pairs;
    // Temporal shard with an overlapped all-gather (engine.cu): K | V rows arrive slab by slab (one slab per source rank)
    // while the kernel runs.  The TMA producer walks the slabs starting at the rank's own one and, before the first
    // tile of a foreign slab, waits until slab_flags[slab] >= slab_epoch (written by the source rank after its copy).
    const uint32_t* slab_flags;   // null = everything is in place (no waiting)
    volatile uint32_t* slab_err;  // mapped host word: set to K5_DIST_ERR_SLAB when a slab does not arrive in time
    unsigned long long slab_timeout_ns;

    uint32_t slab_epoch;
    int n_slabs, slab_first;      // n_slabs == 0: natural tile order
    int slab_tile0[9];            // first 128-row KV tile of each slab; slab_tile0[n_slabs] = number of KV tiles
    // Launch split by key slab (dense fixed-offset kernel; partial sums are additive because no row maximum exists):
    // part_mode bit 0: the launch ends by writing the UNNORMALISED fp32 accumulators O [Sq, heads * 64] and the four fp32
    // partial row sums [Sq, heads, 4] instead of the bf16 output; bit 1: every row starts from those partials (1 = first
    // launch, 3 = a middle launch, 2 = the last launch, which normalises).
    float* part_o;
    float* part_l;
    int part_mode;
    int slab_skip;         // slab walk: leave out the first slab_skip slabs of the rotated order (earlier launches consumed them)
    int stagger;           // cycles query tile 1 starts behind query tile 0 (keeps the two exp phases apart)
    int split_tail;        // split the items of a partial last round into their two query tiles (attention.cu)
};

// Scratch of the block-sparse pre-pass (per-item KV tile lists).  Owned by the caller so that several engines of one
// process can run concurrently on different streams; nullptr = a process-wide instance (single-stream callers only).
struct AttnSparseWs {
    int32_t* count = nullptr;
    int32_t* pairs = nullptr;
    uint8_t* mask = nullptr;
    size_t items = 0, max_pairs = 0;
    ~AttnSparseWs();
};

// O[Sq, heads*64] = softmax(Q K^T * softmax_scale) V per head, head_dim 64, non-causal.
// Arrival schedule of the K | V rows for the overlapped all-gather of the temporal shard (see AttnParams).
constexpr uint32_t K5_DIST_ERR_BARRIER = 1, K5_DIST_ERR_SLAB = 2;
struct AttnSlabs {
    const uint32_t* flags = nullptr;
    uint32_t* err = nullptr;      // see AttnParams::slab_err (required with flags)
    unsigned long long timeout_ns = 600ull * 1000000000ull;
    uint32_t epoch = 0;
    int n = 0, first = 0;         // first = -1 (debug, flags == nullptr): every query row starts at the slab holding it
    int row0[9] = {};             // first row of each slab (multiples of 128); row0[n] = rows of the K | V matrices
    int skip = 0;                 // leave out the first `skip` slabs of the order first, first + 1, ... (consumed by earlier launches
                                  // of the same attention; needs `part` mode 2 or 3); Sk of the call = the rows of the slabs walked
};
// Split of one attention over two launches (see AttnParams::part_mode); o: fp32 [Sq, heads * 64], l: fp32 [Sq, heads, 4]
struct AttnPartial {
    float* o = nullptr;
    float* l = nullptr;
    int mode = 0;                 // bit 0: write partials (no bf16 output); bit 1: start from partials; 3 = a middle launch
};

// Q/K/V/O are row-major token matrices whose head h occupies columns [h*64, h*64+64).
// score_bound: an upper bound on |q . k| * softmax_scale * log2(e) over all query / key pairs that the caller can
// PROVE (0 = none known).  With a bound <= 60 the kernel runs the fixed-offset softmax (no running row max, see
// attention.cu); without one it keeps the running max with lazy rescaling.  Both compute the same softmax.
int attention_fwd(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo, int Sq,
                  int Sk, int heads, float softmax_scale, const int32_t* kv_count, const int32_t* kv_index,
                  cudaStream_t st, AttnSparseWs* ws = nullptr, float score_bound = 0.f, const AttnSlabs* slabs = nullptr,
                  const AttnPartial* part = nullptr);

// Debug builds (-DK5_ATTN_TRACE) only: device buffer [2][512][4] of clock64 stamps written by CTA 0 (attention.cu).
int attention_debug_trace(long long* buf);

// Grows the pre-pass scratch to at least items x max_pairs entries.
int ensure_sparse_ws(AttnSparseWs& w, size_t items, size_t max_pairs);


}  // namespace k5
