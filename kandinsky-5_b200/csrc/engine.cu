// The DiT denoising engine behind the C ABI (include/k5.h): owns repacked weights and workspace, runs
// DiffusionTransformer3D.forward (kandinsky/models/dit.py:155-181) and the flow-matching sampler
// (kandinsky/generation_utils.py:39-129) as a fixed sequence of the kernels in gemm.cu / attention.cu /
// rowops.cu / nabla.cu on one stream.  No host synchronisation inside forward / sample.
#include <unistd.h>

#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/k5.h"
#include "attention.h"
#include "common.h"
#include "gemm.h"
#include "nabla.h"
#include "rowops.h"

namespace k5 {

static int64_t g_launches = 0;
void count_launch(int n) { g_launches += n; }
int64_t launch_count(bool reset) {
    int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

namespace {

constexpr float LN_EPS = 1e-5f;

// ---- temporal shard: cross-GPU barrier on flags in peer memory --------------------------------------------
// Thread p publishes this rank's epoch into slot `rank` of rank p's flag array (release, system scope: the K/V
// rows the preceding GEMM epilogue stored into p's buffer are visible before the flag), then waits until rank
// p's epoch has arrived in the local array.  A wait that outlives the time-out (K5_DIST_TIMEOUT_S, default 600 s: ranks
// may be skewed by first-call allocations, rank-0-only I/O, text encoders) does not trap - a trap poisons the CUDA
// context of every rank for good - but records a code in `err` (a word of mapped HOST memory the host reads before
// every forward, engine_forward) and gives up; once the word is set every later wait returns at once, so the queue
// drains and the next call reports the failure.
struct PeerFlags {
    uint32_t* f[MAX_PEERS];
};
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void dist_barrier_kernel(PeerFlags peers, uint32_t* mine, int rank, int world, uint32_t epoch,
                                    volatile uint32_t* err, unsigned long long timeout_ns) {
    const int p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.f[p] + rank), "r"(epoch) : "memory");
    const unsigned long long t0 = global_ns();
    for (uint32_t spins = 0;; ++spins) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + p) : "memory");
        if (static_cast<int32_t>(v - epoch) >= 0) break;
        if ((spins & 1023u) == 1023u) {
            if (*err != 0u) break;
            if (global_ns() - t0 > timeout_ns) {
                *err = K5_DIST_ERR_BARRIER;
                __threadfence_system();
                break;
            }
        }
    }
}

// Overlapped all-gather (cross-process shards): flag slots inside each rank's 1 KB flag area.
constexpr int FLAG_READY = 16;      // + buf * 8 + source rank: "my slab of this epoch is in your K|V buffer `buf`"
constexpr int FLAG_DONE = 32;       // + source rank: "my attention of this epoch has finished reading"

struct Lin {
    bf16* W = nullptr;     // [out, ld] bf16
    float* b = nullptr;    // [out] fp32 (bf16-rounded) or null
    int out = 0, in = 0, ld = 0;
};
struct AttnW {
    Lin qkv;               // self: [3D, D];  cross: q [D, D]
    Lin kv;                // cross only: [2D, D]
    Lin o;
    float *qn = nullptr, *kn = nullptr;   // [64]
    // Proven bound on |q . k| / 8 * log2(e) after the per-head RMSNorm (nn.py:246-250): |q| <= 8 max|w_q|,
    // |k| <= 8 max|w_k| (RoPE is a rotation); 2 % slack for the bf16 roundings.  Lets the attention kernel drop the
    // running row max (attention.h).  Set at finalize.
    float score_bound = 0.f;
};
struct Block {
    size_t mod_off = 0;    // row offset into the concatenated modulation output
    AttnW self, cross;
    Lin ff_in, ff_out;
};

template <typename T>
int dalloc(T** p, size_t n) {
    K5_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
    return K5_OK;
}

}  // namespace

struct Engine {
    k5_config c{};
    int D = 0, F = 0, Td = 0, heads = 0, Cin = 0, KP = 0;
    size_t mod_rows = 0;
    std::vector<void*> allocs;
    std::set<std::string> expected, loaded;

    // weights
    float *modW = nullptr, *modB = nullptr;
    Lin time_in_dummy;
    float *time_in_W = nullptr, *time_in_b = nullptr, *time_out_W = nullptr, *time_out_b = nullptr;
    Lin text_in, vis_in, out_lin, pooled_in;
    float *text_ln_w = nullptr, *text_ln_b = nullptr, *pooled_ln_w = nullptr, *pooled_ln_b = nullptr;
    std::vector<Block> tblocks, vblocks;
    Lin ckv_all;                    // [num_visual_blocks * 2D, D]: every block's cross-attention [Wk | Wv], stacked
    float* ckn_all = nullptr;       // [num_visual_blocks, 64] key-norm weights of the cross-attentions
    size_t out_mod_off = 0;
    float *freqs = nullptr, *args_text = nullptr, *args_ax[3] = {nullptr, nullptr, nullptr};

    // workspace
    bf16 *x = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *hid = nullptr, *patchA = nullptr, *y64 = nullptr;
    bf16 *te = nullptr, *ten = nullptr, *tqkv = nullptr, *tatt = nullptr, *thid = nullptr, *tproj = nullptr, *ckv = nullptr;
    float *tfeat = nullptr, *t1 = nullptr, *tembed = nullptr, *modOut = nullptr;
    float2 *rope_v = nullptr, *rope_t = nullptr, *rope_t_arange = nullptr;
    int *pos_dev = nullptr;         // [T + Hp + Wp] then text positions at +4096
    int *pos_pinned = nullptr;      // pinned staging of caller-given text positions (lazily allocated, 4096 ints)
    cudaEvent_t pos_ev = nullptr;   // recorded behind the copy that reads pos_pinned
    bf16 *v_c = nullptr, *v_u = nullptr;
    // MagCache (magcache_utils.py): the embedded input of the visual stack and one cached block-stack residual per
    // CFG branch; allocated on first use
    bf16 *mag_x0 = nullptr, *mag_res[2] = {nullptr, nullptr};
    bool mag_valid[2] = {false, false};
    // NABLA
    uint8_t* sta = nullptr;
    int32_t *kv_count = nullptr, *kv_index = nullptr;
    float* nabla_ws = nullptr;
    float* density_acc = nullptr;
    AttnSparseWs sparse_ws;          // per-engine scratch of the block-sparse pre-pass
    int sta_key[6] = {0, 0, 0, 0, 0, 0};
    bool last_sparse = false;
    cudaStream_t last_stream = nullptr;

    // optional CUDA-event timing of the dominant kernel (visual self-attention) for bench.py's roofline
    bool timing = false;
    std::vector<cudaEvent_t> ev;     // pairs (start, stop)
    size_t ev_used = 0;

    // staging for load_tensor
    void* stage = nullptr;
    size_t stage_bytes = 0;

    // temporal shard over the GPUs of one node (k5_dist_*): this rank owns frames [f0, f0 + Tl) = tokens
    // [tok0, tok0 + Sl); K | V of all tokens are all-gathered per visual block into kv_all[buf]
    struct Dist {
        bool on = false;
        int rank = 0, world = 1;
        void* block = nullptr;            // one allocation: flags | kv_all[0] | kv_all[1]  (one IPC handle)
        size_t block_bytes = 0;
        uint32_t* flags = nullptr;
        bf16* kv[2] = {nullptr, nullptr};
        uint32_t* peer_flags[MAX_PEERS] = {};
        bf16* peer_kv[2][MAX_PEERS] = {};
        std::vector<void*> opened;
        uint32_t epoch = 0;
        int buf = 0;
        // overlapped all-gather (ranks in different processes): the projection writes the local slab only, copy
        // engines push it to the peers on `copy_st` while the attention kernel already runs on the local slab
        bool overlap = false;
        // The pushes of a block go out in the order their consumers need them (rank r - 1 reads slab r first), round-robin
        // over a few copy streams.  Measured on 8 ranks (44 MB slabs): ONE stream moves ~120 GB/s, so slab k lands at
        // k x 0.36 ms against a need at k x 0.30 ms and the attention waits 0.45 ms per block; ONE STREAM PER PEER reaches
        // ~350 GB/s in total but every slab lands at the same late moment (0.9 ms, wait 0.59 ms); three streams keep the
        // order and triple the rate (K5_DIST_COPY_STREAMS = 1 .. 7).
        int n_copy_st = 3;
        cudaStream_t copy_st[MAX_PEERS] = {};
        cudaEvent_t ev_kv = nullptr;
        cudaEvent_t ev_copied[2][MAX_PEERS] = {};        // the pushes out of kv[buf] on copy stream i have read their source
        bool copied_valid[2] = {false, false};
        // time-out reporting of the cross-GPU waits (dist_barrier_kernel, the slab wait of the attention producer)
        // overlapped all-gather: fp32 partials handed from the attention launch over the local slab to the one over the
        // foreign slabs (attention.h: AttnPartial)
        float *part_o = nullptr, *part_l = nullptr;
        uint32_t* err_host = nullptr;     // cudaHostAlloc(mapped): 0 = fine, else K5_DIST_ERR_*
        uint32_t* err_dev = nullptr;      // the same word as the device sees it
        unsigned long long timeout_ns = 600ull * 1000000000ull;
    } dist;
    int f0 = 0, Tl = 0, tok0 = 0, Sl = 0;
    // test hook (K5_DEBUG_KV_ORDER=<world>, read at set_grid): a single engine walks the KV tiles of the dense visual
    // attention, for every query row, in the slab order the rank owning that row in a `world`-rank shard with the
    // overlapped all-gather uses - so the whole latent can be compared BIT FOR BIT (tests/gpu_shard_ranks.py)
    int dbg_world = 0;

    // grid
    int T = 0, Hp = 0, Wp = 0, S = 0, fractal = 0;
    bool grid_set = false;
    bool finalized = false;

    template <typename T_>
    int alloc(T_** p, size_t n) {
        K5_TRY(dalloc(p, n));
        allocs.push_back(*p);
        return K5_OK;
    }
    int alloc_lin(Lin& l, int out, int in, bool bias, int ld = 0) {
        l.out = out;
        l.in = in;
        l.ld = ld ? ld : in;
        K5_TRY(alloc(&l.W, static_cast<size_t>(out) * l.ld));
        K5_CHECK_CUDA(cudaMemset(l.W, 0, static_cast<size_t>(out) * l.ld * sizeof(bf16)));
        if (bias) K5_TRY(alloc(&l.b, out));
        return K5_OK;
    }
    ~Engine() {
        if (dist.err_host) cudaFreeHost(dist.err_host);
        if (pos_pinned) cudaFreeHost(pos_pinned);
        if (pos_ev) cudaEventDestroy(pos_ev);
        for (cudaStream_t cs : dist.copy_st)
            if (cs) cudaStreamDestroy(cs);
        if (dist.ev_kv) cudaEventDestroy(dist.ev_kv);
        for (auto& row : dist.ev_copied)
            for (cudaEvent_t ev : row)
                if (ev) cudaEventDestroy(ev);
        for (void* p : dist.opened) cudaIpcCloseMemHandle(p);
        if (dist.block) cudaFree(dist.block);
        for (auto& v : ev) cudaEventDestroy(v);
        for (void* p : allocs) cudaFree(p);
        if (stage) cudaFree(stage);
    }
};

namespace {

void expect_lin(Engine* e, const std::string& name, bool bias = true) {
    e->expected.insert(name + ".weight");
    if (bias) e->expected.insert(name + ".bias");
}
void expect_attn(Engine* e, const std::string& p) {
    for (const char* n : {"to_query", "to_key", "to_value", "out_layer"}) expect_lin(e, p + n);
    e->expected.insert(p + "query_norm.weight");
    e->expected.insert(p + "key_norm.weight");
}

// cross_slot >= 0: cross-attention of visual block `cross_slot`; its K | V projection and key-norm weight are slices of
// the engine-wide stacks (Engine::ckv_all), so that one GEMM projects the text for all blocks at once
int alloc_attn(Engine* e, AttnW& a, int cross_slot = -1) {
    const int D = e->D;
    if (cross_slot < 0) {
        K5_TRY(e->alloc_lin(a.qkv, 3 * D, D, true));
        K5_TRY(e->alloc(&a.kn, 64));
    } else {
        K5_TRY(e->alloc_lin(a.qkv, D, D, true));
        a.kv.out = 2 * D;
        a.kv.in = a.kv.ld = D;
        a.kv.W = e->ckv_all.W + static_cast<size_t>(cross_slot) * 2 * D * D;
        a.kv.b = e->ckv_all.b + static_cast<size_t>(cross_slot) * 2 * D;
        a.kn = e->ckn_all + static_cast<size_t>(cross_slot) * 64;
    }
    K5_TRY(e->alloc_lin(a.o, D, D, true));
    K5_TRY(e->alloc(&a.qn, 64));
    return K5_OK;
}

int engine_init(Engine* e) {
    const k5_config& c = e->c;
    K5_REQUIRE(c.patch_size[0] == 1 && c.patch_size[1] == 2 && c.patch_size[2] == 2, "only patch_size (1,2,2) is supported");
    K5_REQUIRE(c.axes_dims[0] + c.axes_dims[1] + c.axes_dims[2] == 64, "head_dim (sum of axes_dims) must be 64");
    K5_REQUIRE(c.model_dim % 256 == 0 && c.model_dim <= 2048, "model_dim must be a multiple of 256 and <= 2048");
    K5_REQUIRE(c.ff_dim % 64 == 0 && c.time_dim % 128 == 0 && c.time_dim <= 1024, "ff_dim x64, time_dim x128 <= 1024");
    K5_REQUIRE(c.in_text_dim % 8 == 0 && c.in_text_dim2 % 8 == 0, "text dims must be multiples of 8");
    K5_REQUIRE(c.max_tokens > 0 && c.max_text_tokens > 0 && c.max_text_tokens <= 1024, "bad workspace bounds");
    K5_REQUIRE(4 * c.out_visual_dim == 64, "out_visual_dim must be 16 (patch 1x2x2 -> 64 output features)");
    e->D = c.model_dim;
    e->F = c.ff_dim;
    e->Td = c.time_dim;
    e->heads = c.model_dim / 64;
    e->Cin = c.visual_cond ? 2 * c.in_visual_dim + 1 : c.in_visual_dim;
    e->KP = ((4 * e->Cin + 63) / 64) * 64;
    const int D = e->D, F = e->F, Td = e->Td;
    const size_t S = c.max_tokens, L = c.max_text_tokens;

    e->mod_rows = static_cast<size_t>(c.num_text_blocks) * 6 * D + static_cast<size_t>(c.num_visual_blocks) * 9 * D + 2 * D;
    K5_TRY(e->alloc(&e->modW, e->mod_rows * Td));
    K5_TRY(e->alloc(&e->modB, e->mod_rows));
    K5_TRY(e->alloc(&e->modOut, e->mod_rows));
    K5_TRY(e->alloc(&e->time_in_W, static_cast<size_t>(Td) * D));
    K5_TRY(e->alloc(&e->time_in_b, Td));
    K5_TRY(e->alloc(&e->time_out_W, static_cast<size_t>(Td) * Td));
    K5_TRY(e->alloc(&e->time_out_b, Td));
    K5_TRY(e->alloc_lin(e->text_in, D, c.in_text_dim, true));
    K5_TRY(e->alloc_lin(e->pooled_in, Td, c.in_text_dim2, true));
    K5_TRY(e->alloc_lin(e->vis_in, D, 4 * e->Cin, true, e->KP));
    K5_TRY(e->alloc_lin(e->out_lin, 4 * c.out_visual_dim, D, true));
    K5_TRY(e->alloc(&e->text_ln_w, D));
    K5_TRY(e->alloc(&e->text_ln_b, D));
    K5_TRY(e->alloc(&e->pooled_ln_w, Td));
    K5_TRY(e->alloc(&e->pooled_ln_b, Td));
    K5_TRY(e->alloc(&e->freqs, D / 2));
    K5_TRY(e->alloc(&e->args_text, 1024 * 32));
    for (int i = 0; i < 3; ++i) K5_TRY(e->alloc(&e->args_ax[i], 128 * (c.axes_dims[i] / 2)));

    expect_lin(e, "time_embeddings.in_layer");
    expect_lin(e, "time_embeddings.out_layer");
    expect_lin(e, "text_embeddings.in_layer");
    e->expected.insert("text_embeddings.norm.weight");
    e->expected.insert("text_embeddings.norm.bias");
    expect_lin(e, "pooled_text_embeddings.in_layer");
    e->expected.insert("pooled_text_embeddings.norm.weight");
    e->expected.insert("pooled_text_embeddings.norm.bias");
    expect_lin(e, "visual_embeddings.in_layer");
    expect_lin(e, "out_layer.modulation.out_layer");
    expect_lin(e, "out_layer.out_layer");
    for (const char* n : {"time_embeddings.freqs", "text_rope_embeddings.args", "visual_rope_embeddings.args_0",
                          "visual_rope_embeddings.args_1", "visual_rope_embeddings.args_2"})
        e->expected.insert(n);

    size_t off = 0;
    e->tblocks.resize(c.num_text_blocks);
    for (int i = 0; i < c.num_text_blocks; ++i) {
        Block& b = e->tblocks[i];
        b.mod_off = off;
        off += 6 * D;
        K5_TRY(alloc_attn(e, b.self));
        K5_TRY(e->alloc_lin(b.ff_in, F, D, false));
        K5_TRY(e->alloc_lin(b.ff_out, D, F, false));
        const std::string p = "text_transformer_blocks." + std::to_string(i) + ".";
        expect_lin(e, p + "text_modulation.out_layer");
        expect_attn(e, p + "self_attention.");
        expect_lin(e, p + "feed_forward.in_layer", false);
        expect_lin(e, p + "feed_forward.out_layer", false);
    }
    e->vblocks.resize(c.num_visual_blocks);
    K5_TRY(e->alloc_lin(e->ckv_all, c.num_visual_blocks * 2 * D, D, true));
    K5_TRY(e->alloc(&e->ckn_all, static_cast<size_t>(c.num_visual_blocks) * 64));
    for (int i = 0; i < c.num_visual_blocks; ++i) {
        Block& b = e->vblocks[i];
        b.mod_off = off;
        off += 9 * D;
        K5_TRY(alloc_attn(e, b.self));
        K5_TRY(alloc_attn(e, b.cross, i));
        K5_TRY(e->alloc_lin(b.ff_in, F, D, false));
        K5_TRY(e->alloc_lin(b.ff_out, D, F, false));
        const std::string p = "visual_transformer_blocks." + std::to_string(i) + ".";
        expect_lin(e, p + "visual_modulation.out_layer");
        expect_attn(e, p + "self_attention.");
        expect_attn(e, p + "cross_attention.");
        expect_lin(e, p + "feed_forward.in_layer", false);
        expect_lin(e, p + "feed_forward.out_layer", false);
    }
    e->out_mod_off = off;

    // workspace
    K5_TRY(e->alloc(&e->x, S * D));
    K5_TRY(e->alloc(&e->xn, S * D));
    K5_TRY(e->alloc(&e->qkv, S * 3 * D));
    K5_TRY(e->alloc(&e->att, S * D));
    K5_TRY(e->alloc(&e->hid, S * F));
    K5_TRY(e->alloc(&e->patchA, S * e->KP));
    K5_TRY(e->alloc(&e->y64, S * 64));
    K5_TRY(e->alloc(&e->te, L * D));
    K5_TRY(e->alloc(&e->ten, L * D));
    K5_TRY(e->alloc(&e->tqkv, L * 3 * D));
    K5_TRY(e->alloc(&e->tatt, L * D));
    K5_TRY(e->alloc(&e->thid, L * F));
    K5_TRY(e->alloc(&e->tproj, L * D));
    K5_TRY(e->alloc(&e->ckv, L * 2 * D * c.num_visual_blocks));     // cross-attention K | V of ALL visual blocks
    K5_TRY(e->alloc(&e->tfeat, D));
    K5_TRY(e->alloc(&e->t1, Td));
    K5_TRY(e->alloc(&e->tembed, Td));
    K5_TRY(e->alloc(&e->rope_v, S * 32));
    K5_TRY(e->alloc(&e->rope_t, L * 32));
    K5_TRY(e->alloc(&e->rope_t_arange, L * 32));
    K5_TRY(e->alloc(&e->pos_dev, 8192));
    K5_TRY(e->alloc(&e->v_c, S * 64));
    K5_TRY(e->alloc(&e->v_u, S * 64));
    K5_TRY(e->alloc(&e->density_acc, 2));
    return K5_OK;
}


struct Dest {
    enum Kind { NONE, BF16_MAT, F32_ROUND, F32 } kind = NONE;
    void* ptr = nullptr;
    int rows = 0, cols = 0, ld = 0;
};

// key -> destination inside the engine's repacked storage
Dest resolve(Engine* e, const std::string& key) {
    Dest d;
    const int D = e->D, Td = e->Td;
    auto mat = [&](Lin& l, int row0, int rows) {
        d.kind = Dest::BF16_MAT;
        d.ptr = l.W + static_cast<size_t>(row0) * l.ld;
        d.rows = rows;
        d.cols = l.in;
        d.ld = l.ld;
    };
    auto vec_round = [&](float* p, int n) {
        d.kind = Dest::F32_ROUND;
        d.ptr = p;
        d.rows = 1;
        d.cols = n;
    };
    auto f32 = [&](float* p, int rows, int cols) {
        d.kind = Dest::F32;
        d.ptr = p;
        d.rows = rows;
        d.cols = cols;
    };
    auto lin = [&](Lin& l, const std::string& suffix, int row0 = 0, int rows = -1) {
        if (rows < 0) rows = l.out;
        if (suffix == "weight") mat(l, row0, rows);
        else if (suffix == "bias" && l.b) vec_round(l.b + row0, rows);
    };
    auto attn = [&](AttnW& a, bool cross, const std::string& rest) {
        const std::string suf = rest.substr(rest.rfind('.') + 1);
        if (rest.rfind("to_query.", 0) == 0) lin(a.qkv, suf, 0, D);
        else if (rest.rfind("to_key.", 0) == 0) cross ? lin(a.kv, suf, 0, D) : lin(a.qkv, suf, D, D);
        else if (rest.rfind("to_value.", 0) == 0) cross ? lin(a.kv, suf, D, D) : lin(a.qkv, suf, 2 * D, D);
        else if (rest.rfind("out_layer.", 0) == 0) lin(a.o, suf);
        else if (rest == "query_norm.weight") f32(a.qn, 1, 64);
        else if (rest == "key_norm.weight") f32(a.kn, 1, 64);
    };
    auto modl = [&](size_t off, int n, const std::string& suf) {
        if (suf == "weight") f32(e->modW + off * Td, n, Td);
        else if (suf == "bias") f32(e->modB + off, 1, n);
    };
    const std::string suf = key.substr(key.rfind('.') + 1);
    if (key == "time_embeddings.freqs") f32(e->freqs, 1, D / 2);
    else if (key == "text_rope_embeddings.args") f32(e->args_text, 1024, 32);
    else if (key.rfind("visual_rope_embeddings.args_", 0) == 0) {
        const int i = key.back() - '0';
        if (i >= 0 && i < 3) f32(e->args_ax[i], 128, e->c.axes_dims[i] / 2);
    } else if (key == "time_embeddings.in_layer.weight") f32(e->time_in_W, Td, D);
    else if (key == "time_embeddings.in_layer.bias") f32(e->time_in_b, 1, Td);
    else if (key == "time_embeddings.out_layer.weight") f32(e->time_out_W, Td, Td);
    else if (key == "time_embeddings.out_layer.bias") f32(e->time_out_b, 1, Td);
    else if (key.rfind("text_embeddings.in_layer.", 0) == 0) lin(e->text_in, suf);
    else if (key == "text_embeddings.norm.weight") f32(e->text_ln_w, 1, D);
    else if (key == "text_embeddings.norm.bias") f32(e->text_ln_b, 1, D);
    else if (key.rfind("pooled_text_embeddings.in_layer.", 0) == 0) lin(e->pooled_in, suf);
    else if (key == "pooled_text_embeddings.norm.weight") f32(e->pooled_ln_w, 1, Td);
    else if (key == "pooled_text_embeddings.norm.bias") f32(e->pooled_ln_b, 1, Td);
    else if (key.rfind("visual_embeddings.in_layer.", 0) == 0) lin(e->vis_in, suf);
    else if (key.rfind("out_layer.modulation.out_layer.", 0) == 0) modl(e->out_mod_off, 2 * D, suf);
    else if (key.rfind("out_layer.out_layer.", 0) == 0) lin(e->out_lin, suf);
    else {
        const bool text = key.rfind("text_transformer_blocks.", 0) == 0;
        const bool vis = key.rfind("visual_transformer_blocks.", 0) == 0;
        if (text || vis) {
            const size_t p0 = key.find('.') + 1;
            const size_t p1 = key.find('.', p0);
            const int idx = atoi(key.substr(p0, p1 - p0).c_str());
            std::vector<Block>& blocks = text ? e->tblocks : e->vblocks;
            if (idx >= 0 && idx < static_cast<int>(blocks.size())) {
                Block& b = blocks[idx];
                const std::string rest = key.substr(p1 + 1);
                if (rest.rfind("text_modulation.out_layer.", 0) == 0 && text) modl(b.mod_off, 6 * D, suf);
                else if (rest.rfind("visual_modulation.out_layer.", 0) == 0 && vis) modl(b.mod_off, 9 * D, suf);
                else if (rest.rfind("self_attention.", 0) == 0) attn(b.self, false, rest.substr(15));
                else if (rest.rfind("cross_attention.", 0) == 0 && vis) attn(b.cross, true, rest.substr(16));
                else if (rest == "feed_forward.in_layer.weight") mat(b.ff_in, 0, b.ff_in.out);
                else if (rest == "feed_forward.out_layer.weight") mat(b.ff_out, 0, b.ff_out.out);
            }
        }
    }
    return d;
}

int dtype_size(int dt) { return dt == 0 ? 4 : 2; }

}  // namespace

int engine_load_tensor(Engine* e, const char* key_c, const void* data, int dtype, const int64_t* shape, int ndim) {
    K5_REQUIRE(key_c && data && shape, "load_tensor: null argument");
    K5_REQUIRE(dtype >= 0 && dtype <= 2, "load_tensor: dtype must be 0 (f32), 1 (bf16) or 2 (f16)");
    const std::string key(key_c);
    if (!e->expected.count(key)) {
        set_last_error("load_tensor: unexpected key '" + key + "'");
        return K5_ERR_INVALID;
    }
    Dest d = resolve(e, key);
    if (d.kind == Dest::NONE) {
        set_last_error("load_tensor: no destination for key '" + key + "'");
        return K5_ERR_INVALID;
    }
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= static_cast<size_t>(shape[i]);
    const size_t want = static_cast<size_t>(d.rows) * d.cols;
    bool shape_ok = (n == want);
    if (shape_ok && ndim == 2) shape_ok = (shape[0] == d.rows && shape[1] == d.cols);
    if (!shape_ok) {
        set_last_error("load_tensor: shape mismatch for '" + key + "': expected [" + std::to_string(d.rows) + ", " +
                       std::to_string(d.cols) + "], got " + std::to_string(n) + " elements");
        return K5_ERR_INVALID;
    }
    const size_t bytes = n * dtype_size(dtype);
    if (bytes > e->stage_bytes) {
        if (e->stage) cudaFree(e->stage);
        e->stage = nullptr;
        e->stage_bytes = 0;
        K5_CHECK_CUDA(cudaMalloc(&e->stage, bytes));
        e->stage_bytes = bytes;
    }
    K5_CHECK_CUDA(cudaMemcpy(e->stage, data, bytes, cudaMemcpyDefault));
    if (d.kind == Dest::BF16_MAT) K5_TRY(convert_to_bf16(e->stage, dtype, static_cast<bf16*>(d.ptr), d.rows, d.cols, d.ld, 0));
    else K5_TRY(convert_to_f32(e->stage, dtype, static_cast<float*>(d.ptr), n, d.kind == Dest::F32_ROUND, 0));
    K5_CHECK_CUDA(cudaStreamSynchronize(0));
    e->loaded.insert(key);
    return K5_OK;
}

int engine_finalize(Engine* e) {
    std::string missing;
    int n = 0;
    for (const std::string& k : e->expected)
        if (!e->loaded.count(k)) {
            if (n++ < 8) missing += (missing.empty() ? "" : ", ") + k;
        }
    if (n) {
        set_last_error("finalize: " + std::to_string(n) + " tensors missing: " + missing + (n > 8 ? ", ..." : ""));
        return K5_ERR_STATE;
    }
    if (e->stage) {
        cudaFree(e->stage);
        e->stage = nullptr;
        e->stage_bytes = 0;
    }
    {
        const int L = e->c.max_text_tokens;
        std::vector<int> pos(L);
        for (int i = 0; i < L; ++i) pos[i] = i;
        K5_CHECK_CUDA(cudaMemcpy(e->pos_dev + 4096, pos.data(), L * sizeof(int), cudaMemcpyHostToDevice));
        K5_TRY(rope1d_table(e->args_text, 32, e->pos_dev + 4096, L, e->rope_t_arange, 0));
        K5_CHECK_CUDA(cudaStreamSynchronize(0));
    }
    {
        auto bound = [&](AttnW& a) -> int {
            float q[64], k[64];
            K5_CHECK_CUDA(cudaMemcpy(q, a.qn, sizeof(q), cudaMemcpyDeviceToHost));
            K5_CHECK_CUDA(cudaMemcpy(k, a.kn, sizeof(k), cudaMemcpyDeviceToHost));
            float mq = 0.f, mk = 0.f;
            for (int i = 0; i < 64; ++i) {
                mq = fmaxf(mq, fabsf(q[i]));
                mk = fmaxf(mk, fabsf(k[i]));
            }
            a.score_bound = 8.f * mq * 8.f * mk * 0.125f * 1.4426950408889634f * 1.02f;
            if (!(a.score_bound > 0.f) || !std::isfinite(a.score_bound)) a.score_bound = 0.f;   // unknown -> running max
            return K5_OK;
        };
        for (Block& b : e->tblocks) K5_TRY(bound(b.self));
        for (Block& b : e->vblocks) {
            K5_TRY(bound(b.self));
            K5_TRY(bound(b.cross));
        }
    }
    e->finalized = true;
    return K5_OK;
}

int engine_set_grid(Engine* e, int T, int H, int W, const int32_t* pt, const int32_t* ph, const int32_t* pw,
                    const float sf[3], int fractal) {
    K5_REQUIRE(e->finalized, "set_grid: call k5_engine_finalize first (the RoPE buffers come with the weights)");
    K5_REQUIRE(T > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "set_grid: latent H, W must be even and positive");
    const int Hp = H / 2, Wp = W / 2;
    const size_t S = static_cast<size_t>(T) * Hp * Wp;
    K5_REQUIRE(S <= static_cast<size_t>(e->c.max_tokens), "set_grid: token count exceeds max_tokens of the engine");
    K5_REQUIRE(T <= 128 && Hp <= 128 && Wp <= 128, "set_grid: RoPE tables hold 128 positions per axis");
    K5_REQUIRE(!fractal || (Hp % 8 == 0 && Wp % 8 == 0), "set_grid: fractal order needs H/2, W/2 divisible by 8");
    K5_REQUIRE(sf && sf[0] != 0.f && sf[1] != 0.f && sf[2] != 0.f, "set_grid: scale_factor must be non-zero");
    std::vector<int> pos(T + Hp + Wp);
    for (int i = 0; i < T; ++i) pos[i] = pt ? pt[i] : i;
    for (int i = 0; i < Hp; ++i) pos[T + i] = ph ? ph[i] : i;
    for (int i = 0; i < Wp; ++i) pos[T + Hp + i] = pw ? pw[i] : i;
    for (int v : pos) K5_REQUIRE(v >= 0 && v < 128, "set_grid: RoPE position out of range [0,128)");
    K5_CHECK_CUDA(cudaMemcpy(e->pos_dev, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice));
    const k5_config& c = e->c;
    K5_TRY(rope3d_table(e->args_ax[0], e->args_ax[1], e->args_ax[2], c.axes_dims[0] / 2, c.axes_dims[1] / 2,
                        c.axes_dims[2] / 2, e->pos_dev, e->pos_dev + T, e->pos_dev + T + Hp, sf, T, Hp, Wp, fractal != 0,
                        e->rope_v, 0));
    K5_CHECK_CUDA(cudaStreamSynchronize(0));
    e->T = T;
    e->Hp = Hp;
    e->Wp = Wp;
    e->S = static_cast<int>(S);
    e->fractal = fractal ? 1 : 0;
    e->f0 = 0;
    e->Tl = T;
    if (e->dist.on) {
        K5_REQUIRE(T >= e->dist.world, "set_grid: the temporal shard needs at least one frame per rank");
        const int base = T / e->dist.world, rem = T % e->dist.world, r = e->dist.rank;
        e->f0 = r * base + (r < rem ? r : rem);
        e->Tl = base + (r < rem ? 1 : 0);
    }
    e->tok0 = e->f0 * Hp * Wp;
    e->Sl = e->Tl * Hp * Wp;
    e->dbg_world = 0;
    if (const char* ko = getenv("K5_DEBUG_KV_ORDER")) {
        const int w = atoi(ko);
        if (!e->dist.on && w > 1 && w <= MAX_PEERS && T >= w) e->dbg_world = w;
    }
    e->grid_set = true;
    return K5_OK;
}

// ---- temporal shard set-up (include/k5.h: k5_dist_export / k5_dist_init) ----------------------------------
namespace {
struct DistHandle {                 // K5_DIST_HANDLE_BYTES = 192
    uint64_t base;                  // raw device pointer (valid inside the exporting process)
    uint64_t bytes;
    int64_t pid;
    int32_t device;
    int32_t pad[9];
    cudaIpcMemHandle_t ipc;         // 64 bytes
    uint8_t pad2[64];
};
static_assert(sizeof(DistHandle) == 192, "handle layout is part of the C ABI");
constexpr size_t DIST_FLAG_BYTES = 1024;
}  // namespace

int engine_dist_export(Engine* e, void* out) {
    K5_REQUIRE(out, "dist_export: null handle buffer");
    if (!e->dist.block) {
        const size_t kv_bytes = static_cast<size_t>(e->c.max_tokens) * 2 * e->D * sizeof(bf16);
        e->dist.block_bytes = DIST_FLAG_BYTES + 2 * kv_bytes;
        K5_CHECK_CUDA(cudaMalloc(&e->dist.block, e->dist.block_bytes));
        K5_CHECK_CUDA(cudaMemset(e->dist.block, 0, DIST_FLAG_BYTES));
        K5_CHECK_CUDA(cudaDeviceSynchronize());
        e->dist.flags = static_cast<uint32_t*>(e->dist.block);
        e->dist.kv[0] = reinterpret_cast<bf16*>(static_cast<uint8_t*>(e->dist.block) + DIST_FLAG_BYTES);
        e->dist.kv[1] = reinterpret_cast<bf16*>(static_cast<uint8_t*>(e->dist.block) + DIST_FLAG_BYTES + kv_bytes);
    }
    DistHandle h;
    memset(&h, 0, sizeof(h));
    h.base = reinterpret_cast<uint64_t>(e->dist.block);
    h.bytes = e->dist.block_bytes;
    h.pid = static_cast<int64_t>(getpid());
    int dev = 0;
    K5_CHECK_CUDA(cudaGetDevice(&dev));
    h.device = dev;
    K5_CHECK_CUDA(cudaIpcGetMemHandle(&h.ipc, e->dist.block));
    memcpy(out, &h, sizeof(h));
    return K5_OK;
}

int engine_dist_init(Engine* e, int rank, int world, const void* handles) {
    K5_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, "dist_init: rank / world out of range (1..8)");
    K5_REQUIRE(handles, "dist_init: null handles");
    K5_REQUIRE(e->dist.block, "dist_init: call k5_dist_export first");
    K5_REQUIRE(!e->dist.on, "dist_init: already initialised");
    const DistHandle* hs = static_cast<const DistHandle*>(handles);
    const size_t kv_bytes = static_cast<size_t>(e->c.max_tokens) * 2 * e->D * sizeof(bf16);
    for (int p = 0; p < world; ++p) {
        K5_REQUIRE(hs[p].bytes == e->dist.block_bytes, "dist_init: peers were created with a different max_tokens / model_dim");
        uint8_t* base = nullptr;
        if (p == rank) {
            base = static_cast<uint8_t*>(e->dist.block);
        } else if (hs[p].pid == static_cast<int64_t>(getpid())) {
            base = reinterpret_cast<uint8_t*>(hs[p].base);      // engines of one process (tests): plain pointers
        } else {
            void* ptr = nullptr;
            K5_CHECK_CUDA(cudaIpcOpenMemHandle(&ptr, hs[p].ipc, cudaIpcMemLazyEnablePeerAccess));
            e->dist.opened.push_back(ptr);
            base = static_cast<uint8_t*>(ptr);
        }
        e->dist.peer_flags[p] = reinterpret_cast<uint32_t*>(base);
        e->dist.peer_kv[0][p] = reinterpret_cast<bf16*>(base + DIST_FLAG_BYTES);
        e->dist.peer_kv[1][p] = reinterpret_cast<bf16*>(base + DIST_FLAG_BYTES + kv_bytes);
    }
    e->dist.rank = rank;
    e->dist.world = world;
    e->dist.on = world > 1;
    if (!e->dist.err_host) {
        K5_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&e->dist.err_host), sizeof(uint32_t), cudaHostAllocMapped));
        *e->dist.err_host = 0;
        K5_CHECK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->dist.err_dev), e->dist.err_host, 0));
    }
    if (const char* to = getenv("K5_DIST_TIMEOUT_S")) {
        const double sec = atof(to);
        if (sec > 0.0) e->dist.timeout_ns = static_cast<unsigned long long>(sec * 1e9);
    }
    // Peers of the same process (the single-GPU tests drive several engines on one device) keep the scatter + barrier
    // form: an attention kernel that waits for a slab inside its producer would occupy the SMs its peer needs.
    bool cross = world > 1;
    for (int p = 0; p < world; ++p)
        if (p != rank && hs[p].pid == static_cast<int64_t>(getpid())) cross = false;
    // Two forms of the all-gather.  Scatter + barrier: the QKV epilogue stores the K | V columns into every rank's buffer,
    // one flag barrier, then one attention launch over all of S.  Overlapped: copy engines push the slab, attention runs
    // as two launches (local slab first, fp32 partials, then the foreign slabs as they arrive); bit-identical to a single
    // engine walking the slabs in the owners' order.  Measured (ms per step at the 5 s size): 2 GPUs
    // 375.4 scatter / 374.8 overlapped / 379.1 overlapped + split; 8 GPUs 111.0 scatter / 109.5 - 113 overlapped /
    // 107.3 overlapped + split; 4 GPUs 198.8 scatter / 196.1 overlapped + split - all with the attention loop of that day,
    // which was 2.6 % slower than the tuned one in BOTH forms (profiles/r2_attention_part_template.md).  Since then the plain
    // dense kernel (scatter form) has the tuned loop back while the split launches (PART kernels) keep the slower one, which
    // moves the scatter numbers to ~194.8 (4 GPUs) and ~109.0 (8 GPUs): the overlapped form is the default from 8 ranks on,
    // scatter + barrier below; K5_DIST_OVERLAP=0 / 1 forces either.
    const char* ov = getenv("K5_DIST_OVERLAP");
    cross = cross && (ov != nullptr ? atoi(ov) != 0 : world >= 8);
    e->dist.overlap = cross;
    if (cross && !e->dist.ev_kv) {
        K5_CHECK_CUDA(cudaEventCreateWithFlags(&e->dist.ev_kv, cudaEventDisableTiming));
        K5_TRY(e->alloc(&e->dist.part_o, static_cast<size_t>(e->c.max_tokens) * e->D));
        K5_TRY(e->alloc(&e->dist.part_l, static_cast<size_t>(e->c.max_tokens) * e->heads * 4));
        if (const char* ns = getenv("K5_DIST_COPY_STREAMS")) e->dist.n_copy_st = atoi(ns);
        e->dist.n_copy_st = e->dist.n_copy_st < 1 ? 1 : (e->dist.n_copy_st > MAX_PEERS - 1 ? MAX_PEERS - 1 : e->dist.n_copy_st);
        for (int i = 0; i < e->dist.n_copy_st; ++i) {
            K5_CHECK_CUDA(cudaStreamCreateWithFlags(&e->dist.copy_st[i], cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) K5_CHECK_CUDA(cudaEventCreateWithFlags(&e->dist.ev_copied[b][i], cudaEventDisableTiming));
        }
    }
    e->dist.epoch = 0;
    e->dist.buf = 0;
    e->grid_set = false;            // the local slab depends on (rank, world)
    return K5_OK;
}

int engine_dist_barrier(Engine* e, cudaStream_t st) {
    if (!e->dist.on) return K5_OK;
    PeerFlags pf;
    for (int p = 0; p < MAX_PEERS; ++p) pf.f[p] = e->dist.peer_flags[p];
    ++e->dist.epoch;
    count_launch(1);
    dist_barrier_kernel<<<1, 32, 0, st>>>(pf, e->dist.flags, e->dist.rank, e->dist.world, e->dist.epoch, e->dist.err_dev,
                                          e->dist.timeout_ns);
    K5_CHECK_CUDA(cudaGetLastError());
    return K5_OK;
}

void engine_dist_info(Engine* e, int* f0, int* frames) {
    if (f0) *f0 = e->f0;
    if (frames) *frames = e->Tl;
}

int engine_dist_mode(Engine* e) { return !e->dist.on ? 0 : (e->dist.overlap ? 2 : 1); }

namespace {

int lin_gemm(const bf16* A, int lda, const Lin& l, int M, int epi, GemmEpilogue& ep, cudaStream_t st) {
    ep.bias = l.b;
    count_launch(1);
    return gemm_bf16(A, lda, l.W, l.ld, M, l.out, l.ld, epi, ep, st);
}

int ensure_nabla(Engine* e, const k5_sparse* sp) {
    const int nb = e->S / 64;
    const int Tb = e->T, Hb = e->Hp / 8, Wb = e->Wp / 8;
    if (!e->kv_count) {
        const size_t nbmax = e->c.max_tokens / 64;
        K5_TRY(e->alloc(&e->kv_count, static_cast<size_t>(e->heads) * nbmax));
        K5_TRY(e->alloc(&e->kv_index, static_cast<size_t>(e->heads) * nbmax * nbmax));
        K5_TRY(e->alloc(&e->nabla_ws, nabla_workspace_floats(static_cast<int>(nbmax) * 64, e->heads)));
        K5_TRY(e->alloc(&e->sta, nbmax * nbmax));
    }
    const int key[6] = {Tb, Hb, Wb, sp->wT, sp->wH, sp->wW};
    bool same = true;
    for (int i = 0; i < 6; ++i) same = same && key[i] == e->sta_key[i];
    if (!same) {
        K5_REQUIRE(Tb * Hb * Wb == nb, "NABLA: block grid does not match the token count");
        count_launch(1);
        K5_TRY(sta_mask(Tb, Hb, Wb, sp->wT, sp->wH, sp->wW, e->sta, 0));
        K5_CHECK_CUDA(cudaStreamSynchronize(0));
        for (int i = 0; i < 6; ++i) e->sta_key[i] = key[i];
    }
    return K5_OK;
}

// One attention sub-layer output projection + gated residual: x = bf16(x + gate * (att . Wo^T + b))
int out_proj_gate(Engine* e, const bf16* att, const Lin& o, bf16* x, const float* gate, int M, cudaStream_t st) {
    GemmEpilogue ep;
    ep.out = x;
    ep.ldo = e->D;
    ep.resid = x;
    ep.ldr = e->D;
    ep.gate = gate;
    return lin_gemm(att, e->D, o, M, EPI_GATE, ep, st);
}

int feed_forward(Engine* e, const Block& b, bf16* x, bf16* xn, bf16* hid, const float* mod, int M, cudaStream_t st) {
    const int D = e->D, F = e->F;
    count_launch(1);
    K5_TRY(ln_rows(x, D, xn, D, M, D, mod + D, mod, true, LN_EPS, st));
    GemmEpilogue g1;
    g1.out = hid;
    g1.ldo = F;
    K5_TRY(lin_gemm(xn, D, b.ff_in, M, EPI_GELU, g1, st));
    GemmEpilogue g2;
    g2.out = x;
    g2.ldo = D;
    g2.resid = x;
    g2.ldr = D;
    g2.gate = mod + 2 * D;
    return lin_gemm(hid, F, b.ff_out, M, EPI_GATE, g2, st);
}

int self_attention(Engine* e, const Block& b, bf16* x, bf16* xn, bf16* qkv, bf16* att, const float* mod, int M,
                   const float2* rope, const k5_sparse* sp, bool visual, cudaStream_t st) {
    const int D = e->D;
    const bool shard = visual && e->dist.on;
    count_launch(1);
    K5_TRY(ln_rows(x, D, xn, D, M, D, mod + D, mod, true, LN_EPS, st));
    GemmEpilogue g;
    g.out = qkv;
    g.ldo = 3 * D;
    g.norm_w0 = b.self.qn;
    g.norm_w1 = b.self.kn;
    g.norm_split = D;
    g.norm_cols = 2 * D;
    g.rope_cols = 2 * D;
    g.rope = rope;
    const bf16 *kp = qkv + D, *vp = qkv + 2 * D;
    int ldkv = 3 * D, Sk = M;
    const bool overlap = shard && e->dist.overlap && !sp && e->S % 128 == 0 && (e->Hp * e->Wp) % 128 == 0;
    AttnSlabs slabs;
    int buf = 0;
    if (shard) {
        buf = e->dist.buf;
        e->dist.buf ^= 1;          // the other buffer may still be read by a slower rank's previous block
        if (overlap) ++e->dist.epoch;   // (the barrier form advances the epoch in engine_dist_barrier)
        g.peers.ld = 2 * D;
        g.peers.col0 = D;
        g.peers.row0 = e->tok0;
        if (overlap) {
            g.peers.n = 1;         // the projection writes this rank's slab of its OWN buffer only
            g.peers.dst[0] = e->dist.kv[buf];
            // ... which the copy engines may still be reading for the pushes of two blocks ago
            if (e->dist.copied_valid[buf])
                for (int i = 0; i < e->dist.n_copy_st; ++i) K5_CHECK_CUDA(cudaStreamWaitEvent(st, e->dist.ev_copied[buf][i], 0));
        } else {
            // all-gather fused into the projection: the K | V columns of this rank's rows go straight from the GEMM
            // epilogue into every rank's [S, 2D] buffer over NVLink; one flag barrier, then attention over all of S
            g.peers.n = e->dist.world;
            for (int p = 0; p < e->dist.world; ++p) g.peers.dst[p] = e->dist.peer_kv[buf][p];
        }
        kp = e->dist.kv[buf];
        vp = e->dist.kv[buf] + D;
        ldkv = 2 * D;
        Sk = e->S;
    }
    K5_TRY(lin_gemm(xn, D, b.self.qkv, M, EPI_HEADS, g, st));
    if (overlap) {
        // Overlapped all-gather: copy engines push the slab to the peers (nearest consumer first: rank r-1 reads slab r
        // right after its own) and flag each arrival, while the attention kernel below starts on the local slab and
        // waits per slab inside its TMA producer.  A peer's buffer may only be overwritten once that peer has finished
        // the attention that read it two blocks ago (FLAG_DONE), which the copy stream - not the compute stream - waits for.
        Engine::Dist& d = e->dist;
        const int W_ = d.world, r = d.rank;
        K5_CHECK_CUDA(cudaEventRecord(d.ev_kv, st));
        const size_t off = static_cast<size_t>(e->tok0) * 2 * D, bytes = static_cast<size_t>(e->Sl) * 2 * D * sizeof(bf16);
        for (int i = 0; i < d.n_copy_st; ++i) K5_CHECK_CUDA(cudaStreamWaitEvent(d.copy_st[i], d.ev_kv, 0));
        for (int k = 1; k < W_; ++k) {
            const int p = (r - k + W_) % W_;          // rank r - 1 reads slab r right after its own: it is served first
            cudaStream_t cs = d.copy_st[(k - 1) % d.n_copy_st];
            // a wrap-safe ">= epoch - 2" on the destination's DONE slot (stream memory operations: no SM needed, so they
            // make progress while the persistent attention kernel holds every SM - a signalling KERNEL on a second stream may not)
            if (d.epoch > 2) K5_TRY(stream_wait_geq_u32(cs, d.flags + FLAG_DONE + p, d.epoch - 2));
            K5_CHECK_CUDA(cudaMemcpyAsync(d.peer_kv[buf][p] + off, d.kv[buf] + off, bytes, cudaMemcpyDeviceToDevice, cs));
            K5_TRY(stream_write_u32(cs, d.peer_flags[p] + FLAG_READY + buf * 8 + r, d.epoch));
        }
        for (int i = 0; i < d.n_copy_st; ++i) K5_CHECK_CUDA(cudaEventRecord(d.ev_copied[buf][i], d.copy_st[i]));
        d.copied_valid[buf] = true;
        slabs.flags = d.flags + FLAG_READY + buf * 8;
        slabs.err = d.err_dev;
        slabs.timeout_ns = d.timeout_ns;
        slabs.epoch = d.epoch;
        slabs.n = W_;
        slabs.first = r;
        const int base = e->T / W_, rem = e->T % W_;
        for (int q = 0; q <= W_; ++q) slabs.row0[q] = (q * base + (q < rem ? q : rem)) * e->Hp * e->Wp;
    } else if (shard) {
        K5_TRY(engine_dist_barrier(e, st));
    } else if (visual && e->dbg_world > 1 && !sp && (e->Hp * e->Wp) % 256 == 0) {
        slabs.n = e->dbg_world;
        slabs.first = -1;
        const int base = e->T / slabs.n, rem = e->T % slabs.n;
        for (int q = 0; q <= slabs.n; ++q) slabs.row0[q] = (q * base + (q < rem ? q : rem)) * e->Hp * e->Wp;
    }
    const int32_t *cnt = nullptr, *idx = nullptr;
    if (sp) {
        count_launch(nabla_select_launches());
        // local query blocks against ALL key blocks (on a shard: the gathered K); STA rows of this rank's blocks
        K5_TRY(nabla_select(qkv, 3 * D, M, kp, ldkv, Sk, e->heads, sp->P, sp->add_sta ? e->sta : nullptr, e->tok0 / 64,
                            e->kv_count, e->kv_index, e->nabla_ws, e->density_acc, st));
        cnt = e->kv_count;
        idx = e->kv_index;
    }
    count_launch(1);
    const bool timed = e->timing && visual && e->ev_used + 2 <= e->ev.size();
    if (timed) K5_CHECK_CUDA(cudaEventRecord(e->ev[e->ev_used], st));
    const bool split = overlap && e->dist.world > 1 && b.self.score_bound > 0.f && b.self.score_bound <= 60.f &&
                       !(getenv("K5_DIST_SPLIT") && atoi(getenv("K5_DIST_SPLIT")) == 0);
    if (split) {
        // The kernel runs heads outermost and every item sweeps ALL key tiles, so one launch over the whole K | V needs every
        // foreign slab within its FIRST item (0.55 ms on 8 ranks) - the transfer would not hide.  Two launches instead:
        // the local slab for all items first (its partial sums of O and l are additive under the fixed-offset softmax and
        // travel as fp32), then the foreign slabs in arrival order.  The accumulation order per row is that of one launch
        // starting at the own slab, so the result is bit-identical to it (tests/gpu_shard_ranks.py).
        // K5_DIST_SPLIT_GROUPS=2 walks the foreign slabs in TWO launches (the nearest (W - 1) / 2 first, a middle launch that
        // reads and writes the partials): with all ranks pushing at once a copy engine delivers ~370 GB/s
        // (tests/gpu_p2p_bandwidth.py), one 44 MB slab per 0.12 ms on 8 ranks, while an item consumes one per 0.06 ms, so the
        // last launch would start with its slabs landed.  Bit-identical on 8 GPUs as well, but measured no faster (108.7
        // against 107.3 ms per step, profiles/r2_shard_final_8gpu.log): late slabs are not what the split form still loses.
        AttnPartial part;
        part.o = e->dist.part_o;
        part.l = e->dist.part_l;
        part.mode = 1;
        const size_t own = static_cast<size_t>(e->tok0) * ldkv;
        const int W_ = e->dist.world, r = e->dist.rank;
        int groups = 1;
        if (const char* gs = getenv("K5_DIST_SPLIT_GROUPS")) groups = atoi(gs) >= 2 && W_ >= 3 ? 2 : 1;
        count_launch(groups);
        K5_TRY(attention_fwd(qkv, 3 * D, kp + own, ldkv, vp + own, ldkv, att, D, M, e->Sl, e->heads, 0.125f, nullptr, nullptr, st,
                             &e->sparse_ws, b.self.score_bound, nullptr, &part));
        const int near = groups == 2 ? (W_ - 1) / 2 : W_ - 1;     // foreign slabs of the first foreign launch
        int rows_near = 0;
        for (int c = 1; c <= near; ++c) {
            const int sl = (r + c) % W_;
            rows_near += slabs.row0[sl + 1] - slabs.row0[sl];
        }
        if (groups == 2) {
            part.mode = 3;
            slabs.skip = 1;
            K5_TRY(attention_fwd(qkv, 3 * D, kp, ldkv, vp, ldkv, att, D, M, rows_near, e->heads, 0.125f, nullptr, nullptr, st,
                                 &e->sparse_ws, b.self.score_bound, &slabs, &part));
            slabs.skip = 1 + near;
            rows_near = Sk - e->Sl - rows_near;
        } else {
            slabs.skip = 1;
        }
        part.mode = 2;
        K5_TRY(attention_fwd(qkv, 3 * D, kp, ldkv, vp, ldkv, att, D, M, rows_near, e->heads, 0.125f, nullptr, nullptr, st,
                             &e->sparse_ws, b.self.score_bound, &slabs, &part));
    } else {
        K5_TRY(attention_fwd(qkv, 3 * D, kp, ldkv, vp, ldkv, att, D, M, Sk, e->heads, 0.125f, cnt, idx, st, &e->sparse_ws,
                             b.self.score_bound, slabs.n > 0 ? &slabs : nullptr));
    }
    if (timed) {
        K5_CHECK_CUDA(cudaEventRecord(e->ev[e->ev_used + 1], st));
        e->ev_used += 2;
    }
    if (overlap) {     // this rank no longer reads K|V buffer `buf` of this epoch: peers may push the block after next
        for (int p = 0; p < e->dist.world; ++p)
            if (p != e->dist.rank)
                K5_TRY(stream_write_u32(st, e->dist.peer_flags[p] + FLAG_DONE + e->dist.rank, e->dist.epoch));
    }
    return out_proj_gate(e, att, b.self.o, x, mod + 2 * D, M, st);
}

// Cross-attention K | V of every visual block in ONE GEMM (dit.py:170-171 hands the same text_embed to all blocks;
// nn.py:317-319,343-349: k = RMSNorm(to_key(text)), v = to_value(text), no RoPE): [L, D] x [NB * 2D, D]^T with the head
// epilogue in stacked-group mode (period 2D, first D columns of a group normed with that block's key_norm weight).
// Replaces 32 launches of 28 tiles each (19 % of the SMs) by one launch of 2 x 448 tiles.
int cross_kv_all(Engine* e, const bf16* text, int L, cudaStream_t st) {
    const int D = e->D;
    GemmEpilogue g;
    g.out = e->ckv;
    g.ldo = 2 * D * static_cast<int>(e->vblocks.size());
    g.norm_w0 = e->ckn_all;
    g.norm_w1 = e->ckn_all;
    g.norm_split = D;
    g.norm_cols = D;
    g.norm_period = 2 * D;
    return lin_gemm(text, D, e->ckv_all, L, EPI_HEADS, g, st);
}

int cross_attention(Engine* e, const Block& b, int L, const float* mod, cudaStream_t st) {
    const int D = e->D, M = e->Sl;
    count_launch(1);
    K5_TRY(ln_rows(e->x, D, e->xn, D, M, D, mod + D, mod, true, LN_EPS, st));
    GemmEpilogue gq;                 // q = RMSNorm(to_query(x)); no RoPE (nn.py:343-349)
    gq.out = e->qkv;
    gq.ldo = D;
    gq.norm_w0 = b.cross.qn;
    gq.norm_w1 = b.cross.qn;
    gq.norm_split = D;
    gq.norm_cols = D;
    K5_TRY(lin_gemm(e->xn, D, b.cross.qkv, M, EPI_HEADS, gq, st));
    // [k | v] of this block = columns [slot * 2D, (slot + 1) * 2D) of the stacked projection (cross_kv_all)
    const int ldc = 2 * D * static_cast<int>(e->vblocks.size());
    const bf16* kv = e->ckv + (&b - e->vblocks.data()) * 2 * D;
    count_launch(1);
    K5_TRY(attention_fwd(e->qkv, D, kv, ldc, kv + D, ldc, e->att, D, M, L, e->heads, 0.125f, nullptr, nullptr, st,
                         nullptr, b.cross.score_bound));
    return out_proj_gate(e, e->att, b.cross.o, e->x, mod + 2 * D, M, st);
}

}  // namespace

// mag_slot < 0: plain forward.  mag_slot 0 / 1 (MagCache, conditional / unconditional branch): mag_skip == 0 runs the
// visual blocks and stores their residual (output - embedded input) in the slot, mag_skip != 0 replaces the 32 visual
// blocks by "embedded input + cached residual" (magcache_utils.py:64-88).
int engine_forward(Engine* e, const float* x, int Cx, const bf16* text, int L, const int32_t* text_pos, const bf16* pooled,
                   float time, const k5_sparse* sp, bf16* out, cudaStream_t st, int mag_slot = -1, int mag_skip = 0) {
    K5_REQUIRE(e->finalized, "forward: call k5_engine_finalize first");
    K5_REQUIRE(e->grid_set, "forward: call k5_engine_set_grid first");
    if (e->dist.on && e->dist.err_host && *static_cast<volatile uint32_t*>(e->dist.err_host) != 0u) {
        set_last_error("temporal shard: a peer rank did not arrive within the time-out (K5_DIST_TIMEOUT_S) at " +
                       std::string(*e->dist.err_host == K5_DIST_ERR_BARRIER ? "the K|V barrier" : "a K|V slab") +
                       "; every result since then is invalid and the ranks are out of step - re-create the engines");
        return K5_ERR_CUDA;
    }
    K5_REQUIRE(x && text && pooled && out, "forward: null tensor");
    K5_REQUIRE(Cx == e->Cin || Cx == e->c.in_visual_dim, "forward: x must have model-input or latent channel count");
    K5_REQUIRE(L > 0 && L <= e->c.max_text_tokens, "forward: text length out of range");
    K5_REQUIRE(mag_slot >= -1 && mag_slot <= 1, "forward: MagCache slot must be 0 (conditional) or 1 (unconditional)");
    if (mag_slot >= 0 && !e->mag_x0) {
        const size_t n = static_cast<size_t>(e->c.max_tokens) * e->D;
        K5_TRY(e->alloc(&e->mag_x0, n));
        K5_TRY(e->alloc(&e->mag_res[0], n));
        K5_TRY(e->alloc(&e->mag_res[1], n));
    }
    K5_REQUIRE(mag_slot < 0 || !mag_skip || e->mag_valid[mag_slot], "forward: MagCache skip before any residual was stored");
    K5_REQUIRE(!sp || e->fractal, "forward: NABLA needs the fractal token order (set_grid fractal=1)");
    K5_REQUIRE(!sp || (e->S % 64 == 0 && e->Sl % 64 == 0 && e->tok0 % 64 == 0),
               "forward: NABLA needs token counts (per frame, too, on a shard) divisible by 64");
    const int D = e->D, Td = e->Td;
    const int S = e->Sl;                                   // rows this rank owns (all of them without a shard)
    const size_t frame_in = static_cast<size_t>(e->Hp) * 2 * e->Wp * 2;   // latent pixels per frame
    if (sp) K5_TRY(ensure_nabla(e, sp));
    e->last_sparse = sp != nullptr;
    e->last_stream = st;
    if (sp) K5_CHECK_CUDA(cudaMemsetAsync(e->density_acc, 0, 2 * sizeof(float), st));

    // --- before_text_transformer_blocks (dit.py:130-137)
    count_launch(5);
    K5_TRY(time_features(e->freqs, time, e->tfeat, D / 2, st));
    K5_TRY(gemv_f32(e->time_in_W, e->time_in_b, e->tfeat, e->t1, Td, D, false, true, st));
    K5_TRY(gemv_f32(e->time_out_W, e->time_out_b, e->t1, e->tembed, Td, Td, false, false, st));
    K5_TRY(pooled_embed(e->pooled_in.W, e->pooled_in.b, e->pooled_ln_w, e->pooled_ln_b, pooled, e->c.in_text_dim2, Td,
                        e->tembed, LN_EPS, st));
    // every Modulation layer of the forward in one GEMV (nn.py:161-164): Linear(SiLU(time_embed))
    K5_TRY(gemv_f32(e->modW, e->modB, e->tembed, e->modOut, static_cast<int>(e->mod_rows), Td, true, false, st));

    {   // text_embeddings (nn.py:70-72)
        GemmEpilogue g;
        g.out = e->tproj;
        g.ldo = D;
        K5_TRY(lin_gemm(text, e->c.in_text_dim, e->text_in, L, EPI_STORE, g, st));
        count_launch(1);
        K5_TRY(ln_rows(e->tproj, D, e->te, D, L, D, e->text_ln_w, e->text_ln_b, false, LN_EPS, st));
    }
    {   // visual_embeddings (nn.py:81-96), rows written directly in engine token order (own frames only)
        count_launch(1);
        K5_TRY(patchify(x + static_cast<size_t>(e->f0) * frame_in * Cx, Cx, e->Cin, e->Tl, e->Hp, e->Wp, e->fractal != 0,
                        e->patchA, e->KP, st));
        GemmEpilogue g;
        g.out = e->x;
        g.ldo = D;
        K5_TRY(lin_gemm(e->patchA, e->KP, e->vis_in, S, EPI_STORE, g, st));
    }
    // text RoPE (nn.py:110-116): arange positions use the table built at finalize (a prefix of it)
    const float2* rope_t = e->rope_t_arange;
    if (text_pos) {
        K5_REQUIRE(L <= 4096, "forward: more than 4096 text positions");
        for (int i = 0; i < L; ++i)
            K5_REQUIRE(text_pos[i] >= 0 && text_pos[i] < 1024, "forward: text RoPE position out of range [0,1024)");
        // Pinned staging owned by the engine: the copy is asynchronous and the stream is NOT drained (a pageable source
        // would make cudaMemcpyAsync synchronise it).  The only wait is for the previous forward's copy of this buffer.
        if (!e->pos_pinned) {
            K5_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&e->pos_pinned), 4096 * sizeof(int), cudaHostAllocDefault));
            K5_CHECK_CUDA(cudaEventCreateWithFlags(&e->pos_ev, cudaEventDisableTiming));
        } else {
            K5_CHECK_CUDA(cudaEventSynchronize(e->pos_ev));
        }
        std::memcpy(e->pos_pinned, text_pos, static_cast<size_t>(L) * sizeof(int));
        K5_CHECK_CUDA(cudaMemcpyAsync(e->pos_dev + 4096, e->pos_pinned, L * sizeof(int), cudaMemcpyHostToDevice, st));
        K5_CHECK_CUDA(cudaEventRecord(e->pos_ev, st));
        count_launch(1);
        K5_TRY(rope1d_table(e->args_text, 32, e->pos_dev + 4096, L, e->rope_t, st));
        rope_t = e->rope_t;
    }
    // --- text transformer blocks (dit.py:33-44): replicated on every rank of a shard
    for (const Block& b : e->tblocks) {
        const float* mod = e->modOut + b.mod_off;
        K5_TRY(self_attention(e, b, e->te, e->ten, e->tqkv, e->tatt, mod, L, rope_t, nullptr, false, st));
        K5_TRY(feed_forward(e, b, e->te, e->ten, e->thid, mod + 3 * D, L, st));
    }
    // --- visual transformer blocks (dit.py:61-79)
    const float2* rope_v = e->rope_v + static_cast<size_t>(e->tok0) * 32;
    const size_t nx = static_cast<size_t>(S) * D;
    if (mag_slot >= 0 && mag_skip) {
        count_launch(1);
        K5_TRY(bf16_addsub(e->x, e->mag_res[mag_slot], e->x, nx, false, st));
    } else {
        if (mag_slot >= 0) K5_CHECK_CUDA(cudaMemcpyAsync(e->mag_x0, e->x, nx * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
        if (!e->vblocks.empty()) K5_TRY(cross_kv_all(e, e->te, L, st));
        for (const Block& b : e->vblocks) {
            const float* mod = e->modOut + b.mod_off;
            K5_TRY(self_attention(e, b, e->x, e->xn, e->qkv, e->att, mod, S, rope_v, sp, true, st));
            K5_TRY(cross_attention(e, b, L, mod + 3 * D, st));
            K5_TRY(feed_forward(e, b, e->x, e->xn, e->hid, mod + 6 * D, S, st));
        }
        if (mag_slot >= 0) {
            count_launch(1);
            K5_TRY(bf16_addsub(e->x, e->mag_x0, e->mag_res[mag_slot], nx, true, st));
            e->mag_valid[mag_slot] = true;
        }
    }
    // --- after_blocks / OutLayer (dit.py:150-153, nn.py:374-400): own frames of `out`
    {
        const float* mod = e->modOut + e->out_mod_off;     // (shift, scale)
        count_launch(2);
        K5_TRY(ln_rows(e->x, D, e->xn, D, S, D, mod + D, mod, true, LN_EPS, st));
        GemmEpilogue g;
        g.out = e->y64;
        g.ldo = 64;
        K5_TRY(lin_gemm(e->xn, D, e->out_lin, S, EPI_STORE, g, st));
        K5_TRY(unpatchify(e->y64, 64, e->Tl, e->Hp, e->Wp, e->fractal != 0, e->c.out_visual_dim,
                          out + static_cast<size_t>(e->f0) * frame_in * e->c.out_visual_dim, st));
    }
    return K5_OK;
}

// skip_schedule (MagCache, optional): [num_steps * 2] bytes, entry 2 i + slot != 0 = forward `slot` (0 conditional,
// 1 unconditional) of step i replaces its visual blocks by the cached residual (magcache_utils.py:64-88).  The decisions
// depend on the calibrated magnitude ratios only, never on data, so the host computes the whole schedule up front.
int engine_sample(Engine* e, float* img, int num_steps, float w, float sched, const bf16* text, int L, const bf16* pooled,
                  const bf16* ntext, int Ln, const bf16* npooled, const k5_sparse* sp, cudaStream_t st,
                  const uint8_t* skip_schedule) {
    K5_REQUIRE(e->grid_set, "sample: call k5_engine_set_grid first");
    K5_REQUIRE(num_steps > 0 && img, "sample: bad arguments");
    const bool cfg = fabsf(w - 1.0f) > 1e-6f;
    K5_REQUIRE(!cfg || (ntext && npooled && Ln > 0), "sample: guidance needs the null-text embeddings");
    // T*H*W*16 latent elements; on a temporal shard only this rank's frames are integrated (the caller gathers
    // the slabs once, after the last step: a forward never reads other ranks' latent frames)
    const size_t n = static_cast<size_t>(e->Sl) * 64;
    const size_t off = static_cast<size_t>(e->tok0) * 64;
    // timesteps (generation_utils.py:102-103): linspace(1, 0, N+1), t <- s t / (1 + (s - 1) t), all fp32 like torch
    std::vector<float> ts(num_steps + 1);
    for (int i = 0; i <= num_steps; ++i) {
        // torch.linspace: start + step * i for the first half, end - step * (N - i) for the second
        const float step = (0.0f - 1.0f) / static_cast<float>(num_steps);
        const float t = (i < (num_steps + 1) / 2) ? 1.0f + step * static_cast<float>(i)
                                                  : 0.0f - step * static_cast<float>(num_steps - i);
        ts[i] = (sched * t) / (1.0f + (sched - 1.0f) * t);
    }
    for (int i = 0; i < num_steps; ++i) {
        const float t = ts[i], dt = ts[i + 1] - ts[i];
        const bool mag = skip_schedule != nullptr;
        K5_TRY(engine_forward(e, img, e->c.in_visual_dim, text, L, nullptr, pooled, t * 1000.0f, sp, e->v_c, st, mag ? 0 : -1,
                              mag ? skip_schedule[2 * i] : 0));
        const bf16* v = e->v_c + off;
        if (cfg) {
            K5_TRY(engine_forward(e, img, e->c.in_visual_dim, ntext, Ln, nullptr, npooled, t * 1000.0f, sp, e->v_u, st,
                                  mag ? 1 : -1, mag ? skip_schedule[2 * i + 1] : 0));
            count_launch(1);
            K5_TRY(cfg_combine(e->v_c + off, e->v_u + off, w, e->v_c + off, n, st));
        }
        count_launch(1);
        K5_TRY(euler_step(img + off, v, dt, n, st));
    }
    return K5_OK;
}

float engine_density(Engine* e) {
    if (!e->last_sparse) return 1.0f;
    float h[2] = {0.f, 0.f};
    cudaStreamSynchronize(e->last_stream);
    cudaMemcpy(h, e->density_acc, sizeof(h), cudaMemcpyDeviceToHost);
    return h[1] > 0.f ? h[0] / h[1] : 1.0f;
}

int engine_timing(Engine* e, int enable, double* total_ms, int64_t* launches) {
    // read back what was recorded so far (synchronises), then switch recording on / off
    double tot = 0.0;
    int64_t n = 0;
    if (e->ev_used) {
        K5_CHECK_CUDA(cudaEventSynchronize(e->ev[e->ev_used - 1]));
        for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
            float ms = 0.f;
            K5_CHECK_CUDA(cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
            tot += ms;
            ++n;
        }
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    e->ev_used = 0;
    e->timing = enable != 0;
    if (e->timing && e->ev.empty()) {
        e->ev.resize(2 * 4096);
        for (auto& v : e->ev) K5_CHECK_CUDA(cudaEventCreate(&v));
    }
    return K5_OK;
}

Engine* engine_new(const k5_config* cfg, int* rc) {
    Engine* e = new Engine();
    e->c = *cfg;
    *rc = engine_init(e);
    if (*rc != K5_OK) {
        delete e;
        return nullptr;
    }
    return e;
}
void engine_delete(Engine* e) { delete e; }

}  // namespace k5
