// extern "C" surface of libk5 (include/k5.h): plain pointers and sizes only.
#include "../../include/k5.h"
#include "attention.h"
#include "common.h"
#include "gemm.h"
#include "nabla.h"
#include "conv3d.h"
#include "rowops.h"
#include "vae_ops.h"

namespace k5 {
struct Vae;
Vae* vae_new(const k5_vae_config* cfg, int* rc);
void vae_delete(Vae* e);
int vae_load_tensor(Vae* e, const char* key, const void* data, int dtype, const int64_t* shape, int ndim);
int vae_finalize(Vae* e);
int vae_decode(Vae* e, const float* z, int T, int H, int W, int tile_frames, int stride_frames, bf16* out, cudaStream_t st);
struct Engine;
Engine* engine_new(const k5_config* cfg, int* rc);
void engine_delete(Engine* e);
int engine_load_tensor(Engine* e, const char* key, const void* data, int dtype, const int64_t* shape, int ndim);
int engine_finalize(Engine* e);
int engine_set_grid(Engine* e, int T, int H, int W, const int32_t* pt, const int32_t* ph, const int32_t* pw,
                    const float sf[3], int fractal);
int engine_forward(Engine* e, const float* x, int Cx, const bf16* text, int L, const int32_t* text_pos, const bf16* pooled,
                   float time, const k5_sparse* sp, bf16* out, cudaStream_t st, int mag_slot, int mag_skip);
int engine_sample(Engine* e, float* img, int num_steps, float w, float sched, const bf16* text, int L, const bf16* pooled,
                  const bf16* ntext, int Ln, const bf16* npooled, const k5_sparse* sp, cudaStream_t st,
                  const uint8_t* skip_schedule = nullptr);
float engine_density(Engine* e);
int engine_timing(Engine* e, int enable, double* total_ms, int64_t* launches);
int engine_dist_export(Engine* e, void* out);
int engine_dist_init(Engine* e, int rank, int world, const void* handles);
int engine_dist_barrier(Engine* e, cudaStream_t st);
void engine_dist_info(Engine* e, int* f0, int* frames);
int engine_dist_mode(Engine* e);
int64_t launch_count(bool reset);
void count_launch(int n);
}  // namespace k5

using namespace k5;

#define K5_NEED(p)                                         \
    do {                                                   \
        if (!(p)) {                                        \
            set_last_error("null argument: " #p);          \
            return K5_ERR_INVALID;                         \
        }                                                  \
    } while (0)

extern "C" {

const char* k5_last_error(void) { return get_last_error(); }
int k5_version(void) { return 100; }

int k5_engine_create(const k5_config* cfg, k5_engine** out) {
    K5_NEED(cfg);
    K5_NEED(out);
    int rc = 0;
    Engine* e = engine_new(cfg, &rc);
    *out = reinterpret_cast<k5_engine*>(e);
    return rc;
}
void k5_engine_destroy(k5_engine* e) {
    if (e) engine_delete(reinterpret_cast<Engine*>(e));
}
int k5_engine_load_tensor(k5_engine* e, const char* key, const void* data, int dtype, const int64_t* shape, int ndim) {
    K5_NEED(e);
    return engine_load_tensor(reinterpret_cast<Engine*>(e), key, data, dtype, shape, ndim);
}
int k5_engine_finalize(k5_engine* e) {
    K5_NEED(e);
    return engine_finalize(reinterpret_cast<Engine*>(e));
}
int k5_engine_set_grid(k5_engine* e, int T, int H, int W, const int32_t* pt, const int32_t* ph, const int32_t* pw,
                       const float scale_factor[3], int fractal) {
    K5_NEED(e);
    return engine_set_grid(reinterpret_cast<Engine*>(e), T, H, W, pt, ph, pw, scale_factor, fractal);
}
int k5_dit_forward(k5_engine* e, const float* x, int Cx, const void* text, int L, const int32_t* text_pos_host,
                   const void* pooled, float time, const k5_sparse* sparse, void* out, void* stream) {
    K5_NEED(e);
    return engine_forward(reinterpret_cast<Engine*>(e), x, Cx, static_cast<const bf16*>(text), L, text_pos_host,
                          static_cast<const bf16*>(pooled), time, sparse, static_cast<bf16*>(out),
                          static_cast<cudaStream_t>(stream), -1, 0);
}
int k5_dit_forward_magcache(k5_engine* e, const float* x, int Cx, const void* text, int L, const int32_t* text_pos_host,
                            const void* pooled, float time, const k5_sparse* sparse, void* out, int slot, int skip,
                            void* stream) {
    K5_NEED(e);
    if (slot < 0 || slot > 1) {
        set_last_error("MagCache slot must be 0 or 1");
        return K5_ERR_INVALID;
    }
    return engine_forward(reinterpret_cast<Engine*>(e), x, Cx, static_cast<const bf16*>(text), L, text_pos_host,
                          static_cast<const bf16*>(pooled), time, sparse, static_cast<bf16*>(out),
                          static_cast<cudaStream_t>(stream), slot, skip);
}
int k5_sample(k5_engine* e, float* img, int num_steps, float guidance_weight, float scheduler_scale, const void* text, int L,
              const void* pooled, const void* null_text, int Ln, const void* null_pooled, const k5_sparse* sparse,
              void* stream) {
    K5_NEED(e);
    return engine_sample(reinterpret_cast<Engine*>(e), img, num_steps, guidance_weight, scheduler_scale,
                         static_cast<const bf16*>(text), L, static_cast<const bf16*>(pooled),
                         static_cast<const bf16*>(null_text), Ln, static_cast<const bf16*>(null_pooled), sparse,
                         static_cast<cudaStream_t>(stream));
}
int k5_sample_magcache(k5_engine* e, float* img, int num_steps, float guidance_weight, float scheduler_scale, const void* text,
                       int L, const void* pooled, const void* null_text, int Ln, const void* null_pooled, const k5_sparse* sparse,
                       const uint8_t* skip_schedule, void* stream) {
    K5_NEED(e);
    K5_NEED(skip_schedule);
    return engine_sample(reinterpret_cast<Engine*>(e), img, num_steps, guidance_weight, scheduler_scale,
                         static_cast<const bf16*>(text), L, static_cast<const bf16*>(pooled),
                         static_cast<const bf16*>(null_text), Ln, static_cast<const bf16*>(null_pooled), sparse,
                         static_cast<cudaStream_t>(stream), skip_schedule);
}
int k5_engine_attention_timing(k5_engine* e, int enable, double* total_ms, int64_t* launches) {
    K5_NEED(e);
    return engine_timing(reinterpret_cast<Engine*>(e), enable, total_ms, launches);
}
int k5_dist_export(k5_engine* e, void* handle_out) {
    K5_NEED(e);
    return engine_dist_export(reinterpret_cast<Engine*>(e), handle_out);
}
int k5_dist_init(k5_engine* e, int rank, int world, const void* handles) {
    K5_NEED(e);
    return engine_dist_init(reinterpret_cast<Engine*>(e), rank, world, handles);
}
int k5_dist_barrier(k5_engine* e, void* stream) {
    K5_NEED(e);
    return engine_dist_barrier(reinterpret_cast<Engine*>(e), static_cast<cudaStream_t>(stream));
}
int k5_dist_local_frames(k5_engine* e, int* first_frame, int* num_frames) {
    K5_NEED(e);
    engine_dist_info(reinterpret_cast<Engine*>(e), first_frame, num_frames);
    return K5_OK;
}
int k5_dist_mode(k5_engine* e) { return e ? engine_dist_mode(reinterpret_cast<Engine*>(e)) : 0; }
int64_t k5_launch_count(int reset) { return launch_count(reset != 0); }
float k5_last_sparse_density(k5_engine* e) { return e ? engine_density(reinterpret_cast<Engine*>(e)) : 1.0f; }

int k5_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int epilogue, void* out, int ldo,
                 const float* bias, const void* resid, int ldr, const float* gate, const float* norm_w0,
                 const float* norm_w1, int norm_split, int norm_cols, int rope_cols, const void* rope, void* stream) {
    K5_NEED(A);
    K5_NEED(W);
    K5_NEED(out);
    GemmEpilogue e;
    e.out = static_cast<bf16*>(out);
    e.out_f32 = static_cast<float*>(out);        // K5_EPI_F32: `out` is an fp32 matrix
    e.ldo = ldo;
    e.bias = bias;
    e.resid = static_cast<const bf16*>(resid);
    e.ldr = ldr;
    e.gate = gate;
    e.norm_w0 = norm_w0;
    e.norm_w1 = norm_w1;
    e.norm_split = norm_split;
    e.norm_cols = norm_cols;
    e.rope_cols = rope_cols;
    e.rope = static_cast<const float2*>(rope);
    count_launch(1);
    return gemm_bf16(static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, M, N, K, epilogue, e,
                     static_cast<cudaStream_t>(stream));
}

int k5_attention(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int Sq, int Sk,
                 int heads, float scale, const int32_t* kv_count, const int32_t* kv_index, void* stream) {
    K5_NEED(Q);
    K5_NEED(K);
    K5_NEED(V);
    K5_NEED(O);
    count_launch(1);
    return attention_fwd(static_cast<const bf16*>(Q), ldq, static_cast<const bf16*>(K), ldk, static_cast<const bf16*>(V),
                         ldv, static_cast<bf16*>(O), ldo, Sq, Sk, heads, scale, kv_count, kv_index,
                         static_cast<cudaStream_t>(stream));
}

int k5_attention_bounded(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int Sq,
                         int Sk, int heads, float scale, const int32_t* kv_count, const int32_t* kv_index,
                         float score_bound_log2, void* stream) {
    K5_NEED(Q);
    K5_NEED(K);
    K5_NEED(V);
    K5_NEED(O);
    count_launch(1);
    return attention_fwd(static_cast<const bf16*>(Q), ldq, static_cast<const bf16*>(K), ldk, static_cast<const bf16*>(V),
                         ldv, static_cast<bf16*>(O), ldo, Sq, Sk, heads, scale, kv_count, kv_index,
                         static_cast<cudaStream_t>(stream), nullptr, score_bound_log2);
}

int k5_attention_bounded_split(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int Sq,
                               int Sk, int heads, float scale, float score_bound_log2, int split_row, int split_row2,
                               float* workspace, void* stream) {
    K5_NEED(Q);
    K5_NEED(K);
    K5_NEED(V);
    K5_NEED(O);
    K5_NEED(workspace);
    if (split_row <= 0 || split_row >= Sk || split_row % 128 != 0 || Sk % 128 != 0 ||
        (split_row2 != 0 && (split_row2 <= split_row || split_row2 >= Sk || split_row2 % 128 != 0))) {
        set_last_error("attention split: the split rows and Sk must be multiples of 128 with 0 < split_row < split_row2 < Sk");
        return K5_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bf16 *q = static_cast<const bf16*>(Q), *k = static_cast<const bf16*>(K), *v = static_cast<const bf16*>(V);
    AttnPartial part;
    part.o = workspace;
    part.l = workspace + static_cast<size_t>(Sq) * heads * 64;
    part.mode = 1;                   // first slab: plain launch over rows [0, split_row) that leaves its partials
    const int launches = split_row2 != 0 ? 3 : 2;
    count_launch(launches);
    K5_TRY(attention_fwd(q, ldq, k, ldk, v, ldv, static_cast<bf16*>(O), ldo, Sq, split_row, heads, scale, nullptr, nullptr, st,
                         nullptr, score_bound_log2, nullptr, &part));
    AttnSlabs slabs;                 // the slabs already consumed are skipped; no arrival flags
    slabs.n = launches;
    slabs.first = 0;
    slabs.row0[0] = 0;
    slabs.row0[1] = split_row;
    slabs.row0[2] = split_row2 != 0 ? split_row2 : Sk;
    slabs.row0[3] = Sk;
    if (split_row2 != 0) {
        part.mode = 3;               // middle launch: starts from the partials and leaves them again
        slabs.skip = 1;
        K5_TRY(attention_fwd(q, ldq, k, ldk, v, ldv, static_cast<bf16*>(O), ldo, Sq, split_row2 - split_row, heads, scale, nullptr,
                             nullptr, st, nullptr, score_bound_log2, &slabs, &part));
    }
    part.mode = 2;
    slabs.skip = launches - 1;
    return attention_fwd(q, ldq, k, ldk, v, ldv, static_cast<bf16*>(O), ldo, Sq, Sk - slabs.row0[launches - 1], heads, scale,
                         nullptr, nullptr, st, nullptr, score_bound_log2, &slabs, &part);
}

int k5_debug_attn_trace(void* buf) { return attention_debug_trace(static_cast<long long*>(buf)); }

int k5_ln_rows(const void* x, int ldx, void* out, int ldo, int S, int D, const float* mul, const float* add, int plus_one,
               float eps, void* stream) {
    K5_NEED(x);
    K5_NEED(out);
    K5_NEED(mul);
    K5_NEED(add);
    count_launch(1);
    return ln_rows(static_cast<const bf16*>(x), ldx, static_cast<bf16*>(out), ldo, S, D, mul, add, plus_one != 0, eps,
                   static_cast<cudaStream_t>(stream));
}

int k5_nabla_select(const void* q, int ldq, const void* k, int ldk, int S, int heads, float P, const uint8_t* sta,
                    int32_t* kv_count, int32_t* kv_index, float* workspace, void* stream) {
    K5_NEED(q);
    K5_NEED(k);
    K5_NEED(kv_count);
    K5_NEED(kv_index);
    K5_NEED(workspace);
    count_launch(nabla_select_launches());
    return nabla_select(static_cast<const bf16*>(q), ldq, S, static_cast<const bf16*>(k), ldk, S, heads, P, sta, 0, kv_count,
                        kv_index, workspace, nullptr, static_cast<cudaStream_t>(stream));
}
int k5_vae_create(const k5_vae_config* cfg, k5_vae** out) {
    K5_NEED(cfg);
    K5_NEED(out);
    int rc = 0;
    Vae* v = vae_new(cfg, &rc);
    *out = reinterpret_cast<k5_vae*>(v);
    return rc;
}
void k5_vae_destroy(k5_vae* v) {
    if (v) vae_delete(reinterpret_cast<Vae*>(v));
}
int k5_vae_load_tensor(k5_vae* v, const char* key, const void* data, int dtype, const int64_t* shape, int ndim) {
    K5_NEED(v);
    return vae_load_tensor(reinterpret_cast<Vae*>(v), key, data, dtype, shape, ndim);
}
int k5_vae_finalize(k5_vae* v) {
    K5_NEED(v);
    return vae_finalize(reinterpret_cast<Vae*>(v));
}
int k5_vae_decode(k5_vae* v, const float* z, int T, int H, int W, int tile_frames, int stride_frames, void* out,
                  void* stream) {
    K5_NEED(v);
    return vae_decode(reinterpret_cast<Vae*>(v), z, T, H, W, tile_frames, stride_frames, static_cast<bf16*>(out),
                      static_cast<cudaStream_t>(stream));
}

int k5_conv3d_causal(const void* x, int T, int H, int W, int Cin, const void* w, int Cout, const float* bias,
                     const void* resid, void* out, void* workspace, void* stream) {
    K5_NEED(x);
    K5_NEED(w);
    K5_NEED(bias);
    K5_NEED(out);
    K5_NEED(workspace);
    if (Cin % 64 != 0 || Cout <= 0 || (Cout % 64 != 0 && Cout > 64)) {
        set_last_error("conv3d: Cin must be a multiple of 64, Cout a multiple of 64 or below 64 (conv_out)");
        return K5_ERR_INVALID;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int Cout_pad = (Cout + 63) / 64 * 64;          // weight rows are padded with zeros (conv_out: 3 -> 64)
    bf16* pad = static_cast<bf16*>(workspace);
    bf16* wr = pad + static_cast<size_t>(T + 2) * (H + 2) * (W + 2) * Cin;
    count_launch(3);
    K5_TRY(pad_gather(static_cast<const bf16*>(x), T, H, W, Cin, 1, 1, 1, nullptr, nullptr, nullptr, 32, false, pad, st));
    if (Cout_pad != Cout) K5_CHECK_CUDA(cudaMemsetAsync(wr, 0, static_cast<size_t>(Cout_pad) * 27 * Cin * sizeof(bf16), st));
    K5_TRY(repack_conv_weight(w, 1, Cout, Cin, 27, Cin, wr, st));
    return conv3d_causal(pad, T, H, W, Cin, wr, Cout, Cout_pad, bias, static_cast<const bf16*>(resid), Cout,
                         static_cast<bf16*>(out), Cout, st);
}

int k5_sta_mask(int T, int Hb, int Wb, int wT, int wH, int wW, uint8_t* out, void* stream) {
    K5_NEED(out);
    count_launch(1);
    return sta_mask(T, Hb, Wb, wT, wH, wW, out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
