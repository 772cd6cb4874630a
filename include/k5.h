/* libk5 — C ABI of the B200-native Kandinsky-5 DiT denoising engine.
 *
 * The reference (ai-forever/Kandinsky-5) is pure Python and has no FFI of its own; these entry points are
 * what a binding for its hot path replaces (all citations relative to the reference tree):
 *
 *   k5_engine_create / k5_engine_load_tensor / k5_engine_finalize
 *        <- get_dit(conf.model.dit_params) + dit.load_state_dict(load_file(ckpt), assign=True)
 *           kandinsky/utils.py:105,115-116 ; kandinsky/models/dit.py:82-127,184-186
 *   k5_engine_set_grid
 *        <- RoPE3D.forward + fractal_flatten set-up, kandinsky/models/dit.py:140-147, nn.py:132-150
 *   k5_dit_forward
 *        <- DiffusionTransformer3D.forward, kandinsky/models/dit.py:155-181
 *   k5_sample
 *        <- generate() / get_velocity(), kandinsky/generation_utils.py:39-129
 *   k5_gemm_bf16, k5_attention, k5_ln_rows, k5_nabla_*      (operator level; used by the parity tests)
 *        <- nn.Linear call sites nn.py:181-184,206,235-237,284,317-319,341,354-361,376-382;
 *           flash_attn_func nn.py:201,254,336; apply_scale_shift_norm nn.py:25-28;
 *           nablaT_v2 models/utils.py:136-163 + flex_attention nn.py:257-280
 *
 * Conventions
 *   - every data pointer is CALLER-OWNED DEVICE memory unless the parameter name ends in _host;
 *     the engine never frees caller memory and owns its weights / workspace;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); calls are asynchronous with
 *     respect to the host except create / load_tensor / finalize / set_grid;
 *   - return value 0 = success; non-zero = error code, message via k5_last_error() (thread local).
 *     The Python binding maps K5_ERR_INVALID to ValueError and the rest to RuntimeError, mirroring the
 *     reference's exceptions;
 *   - one engine per GPU per process, not re-entrant (the reference drives one model per rank from a
 *     single Python thread, README.md:271-276);
 *   - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16.
 */
#ifndef K5_H
#define K5_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define K5_OK 0
#define K5_ERR_INVALID 1
#define K5_ERR_CUDA 2
#define K5_ERR_STATE 3
#define K5_ERR_UNSUPPORTED 4

typedef struct k5_engine k5_engine;

/* Mirrors conf.model.dit_params (configs/config_5s_sft.yaml:11-29) plus workspace bounds. */
typedef struct k5_config {
    int32_t in_visual_dim;     /* latent channels (16) */
    int32_t out_visual_dim;    /* 16 */
    int32_t time_dim;          /* 512 */
    int32_t patch_size[3];     /* (1,2,2) — the only patching the engine supports */
    int32_t model_dim;         /* 1792 */
    int32_t ff_dim;            /* 7168 */
    int32_t num_text_blocks;   /* 2 */
    int32_t num_visual_blocks; /* 32 */
    int32_t axes_dims[3];      /* (16,24,24): must sum to head_dim 64 */
    int32_t visual_cond;       /* 1: input has 2*in_visual_dim+1 channels */
    int32_t in_text_dim;       /* 3584 */
    int32_t in_text_dim2;      /* 768 */
    int32_t max_tokens;        /* workspace bound on visual tokens S (47616 for 5 s, 93696 for 10 s) */
    int32_t max_text_tokens;   /* workspace bound on text tokens L (<= 1024) */
} k5_config;

/* NABLA sparse attention parameters (generation_utils.py:10-36; models/utils.py:108-163). */
typedef struct k5_sparse {
    float P;                   /* cumulative-probability threshold (conf.model.attention.P) */
    int32_t wT, wH, wW;        /* STA window over the 64-token block grid */
    int32_t add_sta;           /* OR the STA mask into the adaptive mask (the reference always does) */
} k5_sparse;

const char* k5_last_error(void);
int k5_version(void);

int k5_engine_create(const k5_config* cfg, k5_engine** out);
void k5_engine_destroy(k5_engine* e);

/* One call per state-dict entry, keys exactly as in the reference checkpoint, plus the
 * reference's non-persistent buffers: "time_embeddings.freqs", "text_rope_embeddings.args",
 * "visual_rope_embeddings.args_{0,1,2}".  `data` may be host or device memory. The engine converts and
 * repacks (fused QKV, bf16 GEMM operands, fp32 norm / modulation / time-MLP tensors) into its own storage. */
int k5_engine_load_tensor(k5_engine* e, const char* key, const void* data, int dtype, const int64_t* shape, int ndim);
/* Fails with K5_ERR_STATE (message lists them) if any tensor of the contract was not loaded. */
int k5_engine_finalize(k5_engine* e);

/* Latent grid [T, H, W] (H, W in latent pixels, i.e. before 2x2 patching), RoPE positions (host arrays of
 * length T, H/2, W/2; NULL = arange), RoPE scale_factor (conf.metrics.scale_factor) and token order
 * (fractal = 1 is the NABLA order, needs H/2 and W/2 divisible by 8). */
int k5_engine_set_grid(k5_engine* e, int T, int H, int W, const int32_t* pos_t_host, const int32_t* pos_h_host,
                       const int32_t* pos_w_host, const float scale_factor[3], int fractal);

/* One DiT forward.  x: float32 [T,H,W,Cx] with Cx = in channels of the model (33) or Cx = in_visual_dim
 * (then the zero visual-cond / mask channels are implied); text: bf16 [L, in_text_dim]; pooled: bf16
 * [in_text_dim2]; time = t * 1000; text_pos_host: L positions or NULL = arange; sparse: NULL = dense
 * attention; out: bf16 [T,H,W,out_visual_dim]. */
int k5_dit_forward(k5_engine* e, const float* x, int Cx, const void* text, int L, const int32_t* text_pos_host,
                   const void* pooled, float time, const k5_sparse* sparse, void* out, void* stream);

/* MagCache variant of k5_dit_forward (kandinsky/magcache_utils.py:41-101, enabled by get_T2V_pipeline(magcache=True),
 * kandinsky/utils.py:107-113).  slot: 0 = conditional, 1 = unconditional branch (the reference's cnt % 2).  skip == 0:
 * run the visual blocks and cache their residual (output minus embedded input, bf16) in the slot; skip != 0: replace
 * the visual blocks by "embedded input + cached residual".  The skip decision (accumulated magnitude-ratio error,
 * magcache_utils.py:64-80) is host logic and lives in the Python mirror. */
int k5_dit_forward_magcache(k5_engine* e, const float* x, int Cx, const void* text, int L, const int32_t* text_pos_host,
                            const void* pooled, float time, const k5_sparse* sparse, void* out, int slot, int skip,
                            void* stream);

/* The whole flow-matching Euler loop on the device.  img: float32 [T,H,W,in_visual_dim], noise in / latent
 * out (updated in place).  null_text / null_pooled may be NULL when |guidance_weight - 1| <= 1e-6. */
int k5_sample(k5_engine* e, float* img, int num_steps, float guidance_weight, float scheduler_scale, const void* text,
              int L, const void* pooled, const void* null_text, int Ln, const void* null_pooled, const k5_sparse* sparse,
              void* stream);

/* k5_sample with MagCache (kandinsky/magcache_utils.py:41-101 inside the loop of generation_utils.py:105-128).
 * skip_schedule: host array of num_steps * 2 bytes; entry [2 i + slot] != 0 makes forward `slot` (0 = conditional,
 * 1 = unconditional) of step i reuse the cached residual of its slot instead of running the visual blocks.  The
 * decisions follow from the calibrated magnitude ratios alone (magcache_utils.py:64-80), so the caller computes them up
 * front (the Python mirror's MagCacheState) and no host round trip is left inside the loop. */
int k5_sample_magcache(k5_engine* e, float* img, int num_steps, float guidance_weight, float scheduler_scale,
                       const void* text, int L, const void* pooled, const void* null_text, int Ln, const void* null_pooled,
                       const k5_sparse* sparse, const uint8_t* skip_schedule, void* stream);

/* ---- temporal shard over the GPUs of one node -------------------------------------------
 * The reference scales with a DTensor tensor-parallel plan (kandinsky/models/parallelize.py:11-102, entered from
 * kandinsky/utils.py:40-87 when WORLD_SIZE > 1); these entry points replace it: one engine per rank, rank r owns a
 * contiguous slab of latent frames, every per-token op is local, and the K | V rows of each visual block are
 * all-gathered by the QKV projection itself (stores over NVLink into every rank's buffer) followed by one flag
 * barrier.  Protocol: every rank calls k5_dist_export, the K5_DIST_HANDLE_BYTES blobs are exchanged by the host
 * (torch.distributed.all_gather_object in the Python mirror), every rank calls k5_dist_init with all blobs in rank
 * order, then k5_engine_set_grid.  All ranks must issue the same sequence of forwards.  In a shard, k5_dit_forward
 * writes (and k5_sample integrates) only this rank's frames [first_frame, first_frame + num_frames) of out / img. */
#define K5_DIST_HANDLE_BYTES 192
int k5_dist_export(k5_engine* e, void* handle_out);
int k5_dist_init(k5_engine* e, int rank, int world, const void* handles);
int k5_dist_barrier(k5_engine* e, void* stream);
int k5_dist_local_frames(k5_engine* e, int* first_frame, int* num_frames);
/* Form of the K | V all-gather this engine uses for dense attention: 0 = not a shard, 1 = scatter from the QKV epilogue +
 * flag barrier, 2 = overlapped (copy-engine pushes, per-slab arrival flags, attention split into local / foreign
 * slabs).  Chosen in k5_dist_init: 2 from 8 ranks on, K5_DIST_OVERLAP=0|1 forces either (ranks in one process: 1). */
int k5_dist_mode(k5_engine* e);

/* Number of kernels launched by this library since the counter was last reset (bench.py's gpu_launches). */
int64_t k5_launch_count(int reset);

/* CUDA-event timing of the dominant kernel (visual self-attention), on the stream it is launched on.
 * Returns the accumulated duration and launch count recorded since the previous call (synchronises on the last
 * recorded event), then enables / disables recording for subsequent forwards.  Used by bench.py's roofline. */
int k5_engine_attention_timing(k5_engine* e, int enable, double* total_ms, int64_t* launches);

/* Realised NABLA block density of the last sparse forward (selected / total 64x64 blocks, averaged over
 * blocks and heads); 1.0 if the last forward was dense.  Synchronises the stream it was produced on. */
float k5_last_sparse_density(k5_engine* e);

/* ---- VAE decode (config 5) -----------------------------------------------------------------------------
 *   k5_vae_create / k5_vae_load_tensor / k5_vae_finalize
 *        <- build_vae(conf.model.vae): AutoencoderKLHunyuanVideo.from_pretrained(..., subfolder="vae",
 *           torch_dtype=float16), kandinsky/models/vae.py:1276-1282; keys = the diffusers checkpoint's decoder.* and
 *           post_quant_conv.* tensors, any of f32 / bf16 / f16
 *   k5_vae_decode
 *        <- vae.decode(z).sample, kandinsky/generation_utils.py:220-221 -> vae.py:880-906 (decode), 847-877 (_decode),
 *           1144-1204 (_temporal_tiled_decode), 928-936 (blend_t), 682-696 (HunyuanVideoDecoder3D.forward)
 * z: float32 [latent_channels, T, H, W] (the reference's NCTHW latent, batch 1), already divided by
 * scaling_factor; out: bf16 [3, 4 (T - 1) + 1, 8 H, 8 W].  tile_frames / stride_frames: temporal tiling in SAMPLE
 * frames as chosen by get_dec_optimal_tiling (vae.py:1246-1273; (17, 8) for 121 and 241 frames); tile_frames <= 0
 * decodes in one piece.  Spatial tiling (only for sqrt(H W) > 900 pixels) is not part of the T2V path. */
typedef struct k5_vae k5_vae;
typedef struct k5_vae_config {
    int32_t block_out_channels[4]; /* (128, 256, 512, 512) */
    int32_t latent_channels;       /* 16 */
    int32_t out_channels;          /* 3 */
    int32_t max_tile_frames;       /* workspace bound: latent frames decoded at once (5 for the (17, 8) tiling) */
    int32_t max_height;            /* workspace bound: latent height (64) */
    int32_t max_width;             /* workspace bound: latent width (96) */
} k5_vae_config;
int k5_vae_create(const k5_vae_config* cfg, k5_vae** out);
void k5_vae_destroy(k5_vae* v);
int k5_vae_load_tensor(k5_vae* v, const char* key, const void* data, int dtype, const int64_t* shape, int ndim);
int k5_vae_finalize(k5_vae* v);
int k5_vae_decode(k5_vae* v, const float* z, int T, int H, int W, int tile_frames, int stride_frames, void* out,
                  void* stream);

/* ---- operator level ---------------------------------------------------------------------------------- */
#define K5_EPI_STORE 0
#define K5_EPI_GELU 1
#define K5_EPI_GATE 2
#define K5_EPI_HEADS 3
#define K5_EPI_F32 4            /* out is float [M, N]: fp32 accumulators (+ bias), unrounded (scores for a softmax) */

/* out[M,N] = epilogue(A[M,K] . W[N,K]^T); A, W, out, resid bf16; bias, gate, norm weights float32;
 * rope: float2 [M,32] (cos,sin).  See csrc/gemm.h for the epilogue semantics. */
int k5_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int epilogue, void* out, int ldo,
                 const float* bias, const void* resid, int ldr, const float* gate, const float* norm_w0,
                 const float* norm_w1, int norm_split, int norm_cols, int rope_cols, const void* rope, void* stream);

/* O = softmax(Q K^T * scale) V, head_dim 64, non-causal; head h = columns [64h, 64h+64) of each matrix.
 * kv_count / kv_index: optional block-sparse lists over 64x64 blocks: int32 [heads, Sq/64] and
 * [heads, Sq/64, Sk/64] (first kv_count entries valid), NULL = dense. */
int k5_attention(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int Sq, int Sk,
                 int heads, float scale, const int32_t* kv_count, const int32_t* kv_index, void* stream);

/* Same, for callers that can PROVE |q . k| * scale * log2(e) <= score_bound_log2 for every query / key pair (the DiT
 * can: q and k are RMS-normalised per head, nn.py:246-250, so the bound follows from the two norm weight vectors).
 * With a bound <= 60 the kernel drops the running row maximum (softmax is shift invariant); a bound of 0 or above 60
 * selects the general kernel of k5_attention.  A WRONG bound can overflow the exponentials. */
int k5_attention_bounded(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int Sq,
                         int Sk, int heads, float scale, const int32_t* kv_count, const int32_t* kv_index,
                         float score_bound_log2, void* stream);

/* The same attention split over TWO or THREE launches by key rows - [0, split_row) first, whose unnormalised fp32
 * accumulators and row sums travel through `workspace` (float [Sq * heads * 68]), then [split_row, split_row2) (when
 * split_row2 != 0: a middle launch that starts from the partials and leaves them again), then the rest - the form the
 * temporal shard uses to start on its local K | V slab while the foreign slabs are still arriving (csrc/engine.cu).
 * Partial sums are additive because the fixed-offset softmax has no row maximum: the result is BIT-IDENTICAL to
 * k5_attention_bounded.  Needs a score bound in (0, 60]; the split rows and Sk multiples of 128.  (Parity-test entry point.) */
int k5_attention_bounded_split(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int Sq,
                               int Sk, int heads, float scale, float score_bound_log2, int split_row, int split_row2,
                               float* workspace, void* stream);

/* Debug builds of the library (-DK5_ATTN_TRACE) only: device buffer of 2*512*4 int64 clock stamps that CTA 0 of the
 * attention kernel fills (tools/attn_trace.py); K5_ERR_UNSUPPORTED otherwise. */
int k5_debug_attn_trace(void* buf);

/* out = bf16(LN(x) * (mul + plus_one) + add), fp32 statistics, eps; x, out bf16 [S,D]; mul, add float32 [D]. */
int k5_ln_rows(const void* x, int ldx, void* out, int ldo, int S, int D, const float* mul, const float* add, int plus_one,
               float eps, void* stream);

/* NABLA block selection (nablaT_v2): q, k bf16 [S, heads*64] (post-norm, post-RoPE, fractal order) ->
 * kv_count int32 [heads, S/64], kv_index int32 [heads, S/64, S/64] (ascending block ids).
 * sta: uint8 [S/64, S/64] or NULL.  workspace: float32 [heads * (S/64)^2 + 2 * S/64 * heads * 64]. */
int k5_nabla_select(const void* q, int ldq, const void* k, int ldk, int S, int heads, float P, const uint8_t* sta,
                    int32_t* kv_count, int32_t* kv_index, float* workspace, void* stream);
/* STA block mask (fast_sta_nabla): uint8 [T*Hb*Wb, T*Hb*Wb], row-major (t,h,w) block order. */
int k5_sta_mask(int T, int Hb, int Wb, int wT, int wH, int wW, uint8_t* out, void* stream);

/* Causal 3x3x3 convolution on channels-last bf16 (HunyuanVideoCausalConv3d, vae.py:125-163).  x: [T, H, W, Cin];
 * w: the checkpoint layout [Cout, Cin, 3, 3, 3] in bf16; bias float32 [Cout]; resid (optional) / out: [T, H, W, Cout].
 * Cin a multiple of 64; Cout a multiple of 64, or below 64 (conv_out's 3 channels; weight rows are zero-padded to 64).
 * workspace: bf16 [(T + 2) (H + 2) (W + 2) Cin + 27 ceil64(Cout) Cin].  (Parity-test entry
 * point: pads, repacks and convolves; the engine keeps repacked weights and fuses GroupNorm + SiLU into the pad.) */
int k5_conv3d_causal(const void* x, int T, int H, int W, int Cin, const void* w, int Cout, const float* bias,
                     const void* resid, void* out, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* K5_H */
