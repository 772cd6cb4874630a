// HunyuanVideo causal-conv3d VAE decoder engine behind the C ABI (include/k5.h: k5_vae_*).
//
// Replaces AutoencoderKLHunyuanVideo.decode -> _decode -> _temporal_tiled_decode -> HunyuanVideoDecoder3D.forward
// (kandinsky/models/vae.py:880-906, 847-877, 1144-1204, 682-696) for the latent a T2V sample produces
// (generation_utils.py:210-222).  Activations are channels-last bf16 in engine-owned HBM buffers; every 3x3x3
// convolution is the implicit-GEMM tcgen05 kernel of conv3d.cu fed by the fused GroupNorm + SiLU + up-sample +
// replicate-pad gather of vae_ops.cu; 1x1x1 shortcuts, the attention projections and the two attention
// contractions run on the GEMM kernel of gemm.cu.  The mid-block attention (1 head, head_dim = channels, frame-causal
// mask) is three passes: S = Q K^T per frame against the visible keys only, in-place masked softmax, O = P V.
// Rounding points follow SURVEY.md Appendix A (conv / linear results bf16, GroupNorm / SiLU fp32, residual adds bf16).
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/k5.h"
#include "common.h"
#include "conv3d.h"
#include "gemm.h"
#include "rowops.h"
#include "vae_ops.h"

namespace k5 {

void count_launch(int n);

namespace {

constexpr int GROUPS = 32;
constexpr float GN_EPS = 1e-6f;

int pad64(int c) { return (c + 63) / 64 * 64; }

struct Conv {
    bf16* w = nullptr;       // [cout_pad, taps, cin_pad]
    float* b = nullptr;      // [cout_pad], bf16-rounded values
    int cin = 0, cout = 0, cin_pad = 0, cout_pad = 0, taps = 27;
};
struct Norm {
    float *g = nullptr, *b = nullptr;
    int c = 0;
};
struct Resnet {
    Norm n1, n2;
    Conv c1, c2, sc;
    bool has_sc = false;
};
struct Linear {
    bf16* w = nullptr;       // [out, in]
    float* b = nullptr;
    int n = 0;
};

}  // namespace

struct Vae {
    k5_vae_config c{};
    int width[4] = {0, 0, 0, 0};            // block_out_channels (128, 256, 512, 512)
    int top = 0;
    std::vector<void*> allocs;
    std::set<std::string> expected, loaded;
    std::map<std::string, Conv*> convs;
    std::map<std::string, Norm*> norms;
    std::map<std::string, Linear*> lins;

    Conv conv_in, conv_out, ups[3];
    Norm norm_out, attn_gn;
    Resnet mid[2], up[4][3];
    Linear q, k, v, o;
    float *pq_w = nullptr, *pq_b = nullptr;          // post_quant_conv, bf16-rounded fp32

    // workspace
    bf16* buf[3] = {nullptr, nullptr, nullptr};       // activations [P, C]
    bf16* pad = nullptr;                              // padded conv input
    float* scores = nullptr;                          // attention: S [N, N] fp32
    bf16 *probs = nullptr, *vt = nullptr;             // P [N, N] bf16, V^T [C, N]
    bf16* tile[2] = {nullptr, nullptr};               // decoded tiles [F, 8H, 8W, 3]
    double* sums = nullptr;
    float *stats = nullptr, *ones = nullptr;
    void* stage = nullptr;
    size_t stage_bytes = 0;
    bool finalized = false;

    template <typename T_>
    int alloc(T_** p, size_t n) {
        K5_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T_)));
        allocs.push_back(*p);
        return K5_OK;
    }
    ~Vae() {
        for (void* p : allocs) cudaFree(p);
        if (stage) cudaFree(stage);
    }
};

namespace {

int make_conv(Vae* e, Conv& c, const std::string& name, int cout, int cin, int taps) {
    c.cin = cin;
    c.cout = cout;
    c.taps = taps;
    c.cin_pad = pad64(cin);
    c.cout_pad = pad64(cout);
    const size_t n = static_cast<size_t>(c.cout_pad) * taps * c.cin_pad;
    K5_TRY(e->alloc(&c.w, n));
    K5_CHECK_CUDA(cudaMemset(c.w, 0, n * sizeof(bf16)));
    K5_TRY(e->alloc(&c.b, c.cout_pad));
    K5_CHECK_CUDA(cudaMemset(c.b, 0, c.cout_pad * sizeof(float)));
    e->convs[name] = &c;
    e->expected.insert(name + ".weight");
    e->expected.insert(name + ".bias");
    return K5_OK;
}
int make_norm(Vae* e, Norm& n, const std::string& name, int c) {
    n.c = c;
    K5_TRY(e->alloc(&n.g, c));
    K5_TRY(e->alloc(&n.b, c));
    e->norms[name] = &n;
    e->expected.insert(name + ".weight");
    e->expected.insert(name + ".bias");
    return K5_OK;
}
int make_resnet(Vae* e, Resnet& r, const std::string& p, int cin, int cout) {
    K5_TRY(make_norm(e, r.n1, p + "norm1", cin));
    K5_TRY(make_conv(e, r.c1, p + "conv1.conv", cout, cin, 27));
    K5_TRY(make_norm(e, r.n2, p + "norm2", cout));
    K5_TRY(make_conv(e, r.c2, p + "conv2.conv", cout, cout, 27));
    r.has_sc = cin != cout;
    if (r.has_sc) K5_TRY(make_conv(e, r.sc, p + "conv_shortcut.conv", cout, cin, 1));
    return K5_OK;
}
int make_linear(Vae* e, Linear& l, const std::string& name, int n) {
    l.n = n;
    K5_TRY(e->alloc(&l.w, static_cast<size_t>(n) * n));
    K5_TRY(e->alloc(&l.b, n));
    e->lins[name] = &l;
    e->expected.insert(name + ".weight");
    e->expected.insert(name + ".bias");
    return K5_OK;
}

// up-sampling factors of up block i (vae.py:644-659, time_compression 4, spatial_compression 8)
void up_factor(int i, int& ft, int& fs) {
    fs = i < 3 ? 2 : 1;
    ft = (i == 1 || i == 2) ? 2 : 1;
}

int vae_init(Vae* e) {
    const k5_vae_config& c = e->c;
    for (int i = 0; i < 4; ++i) {
        e->width[i] = c.block_out_channels[i];
        K5_REQUIRE(e->width[i] >= 64 && e->width[i] % 64 == 0 && 256 % (e->width[i] / 8) == 0 && e->width[i] <= 2048,
                   "vae: block_out_channels must be 64 / 128 / 256 / 512 / 1024 / 2048");
    }
    K5_REQUIRE(c.latent_channels > 0 && c.latent_channels <= 16 && c.out_channels > 0 && c.out_channels <= 64,
               "vae: latent_channels <= 16, out_channels <= 64");
    K5_REQUIRE(c.out_channels == 3, "vae: the tile assembly writes RGB video (out_channels must be 3)");
    K5_REQUIRE(c.max_tile_frames > 0 && c.max_height > 0 && c.max_width > 0, "vae: bad workspace bounds");
    K5_REQUIRE((c.max_height * c.max_width) % 8 == 0, "vae: latent H * W must be a multiple of 8");
    e->top = e->width[3];
    const int top = e->top;
    K5_TRY(e->alloc(&e->pq_w, 16 * 16));
    K5_TRY(e->alloc(&e->pq_b, 16));
    e->expected.insert("post_quant_conv.weight");
    e->expected.insert("post_quant_conv.bias");
    K5_TRY(make_conv(e, e->conv_in, "decoder.conv_in.conv", top, c.latent_channels, 27));
    K5_TRY(make_resnet(e, e->mid[0], "decoder.mid_block.resnets.0.", top, top));
    K5_TRY(make_resnet(e, e->mid[1], "decoder.mid_block.resnets.1.", top, top));
    const std::string a = "decoder.mid_block.attentions.0.";
    K5_TRY(make_norm(e, e->attn_gn, a + "group_norm", top));
    K5_TRY(make_linear(e, e->q, a + "to_q", top));
    K5_TRY(make_linear(e, e->k, a + "to_k", top));
    K5_TRY(make_linear(e, e->v, a + "to_v", top));
    K5_TRY(make_linear(e, e->o, a + "to_out.0", top));
    int prev = top;
    for (int i = 0; i < 4; ++i) {
        const int co = e->width[3 - i];
        for (int j = 0; j < 3; ++j)
            K5_TRY(make_resnet(e, e->up[i][j], "decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j) + ".",
                               j == 0 ? prev : co, co));
        if (i < 3) K5_TRY(make_conv(e, e->ups[i], "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv.conv", co, co, 27));
        prev = co;
    }
    K5_TRY(make_norm(e, e->norm_out, "decoder.conv_norm_out", e->width[0]));
    K5_TRY(make_conv(e, e->conv_out, "decoder.conv_out.conv", c.out_channels, e->width[0], 27));

    // workspace: walk the decoder once to find the largest activation / padded volume
    size_t max_act = 0, max_pad = 0;
    {
        size_t T = c.max_tile_frames, H = c.max_height, W = c.max_width;
        auto see = [&](size_t t, size_t h, size_t w, size_t ch) {
            max_act = std::max(max_act, t * h * w * ch);
            max_pad = std::max(max_pad, (t + 2) * (h + 2) * (w + 2) * ch);
        };
        see(T, H, W, std::max(64, 4 * top));          // the attention keeps q | k | v | o in one activation buffer
        int ch = top;
        for (int i = 0; i < 4; ++i) {
            const int co = e->width[3 - i];
            see(T, H, W, std::max(ch, co));
            ch = co;
            int ft, fs;
            up_factor(i, ft, fs);
            if (i < 3) {
                T = ft == 2 ? 1 + (T - 1) * 2 : T;
                H *= fs;
                W *= fs;
                see(T, H, W, ch);
            }
        }
        const size_t N = static_cast<size_t>(c.max_tile_frames) * c.max_height * c.max_width;
        for (int i = 0; i < 3; ++i) K5_TRY(e->alloc(&e->buf[i], max_act));
        K5_TRY(e->alloc(&e->pad, max_pad));
        K5_TRY(e->alloc(&e->scores, N * N));
        K5_TRY(e->alloc(&e->probs, N * N));
        K5_TRY(e->alloc(&e->vt, N * top));
        for (int i = 0; i < 2; ++i) K5_TRY(e->alloc(&e->tile[i], T * H * W * 3));
    }
    K5_TRY(e->alloc(&e->sums, static_cast<size_t>(GN_MAX_BLOCKS) * 2 *
                                  std::max(std::max(e->width[0], e->width[1]), std::max(e->width[2], e->width[3]))));   // block partials
    K5_TRY(e->alloc(&e->stats, 2 * GROUPS));
    K5_TRY(e->alloc(&e->ones, 2048));
    {
        std::vector<float> one(2048, 1.0f);
        K5_CHECK_CUDA(cudaMemcpy(e->ones, one.data(), one.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return K5_OK;
}

int dsize(int dt) { return dt == 0 ? 4 : 2; }

}  // namespace

int vae_load_tensor(Vae* e, const char* key_c, const void* data, int dtype, const int64_t* shape, int ndim) {
    K5_REQUIRE(key_c && data && shape, "vae_load_tensor: null argument");
    K5_REQUIRE(dtype >= 0 && dtype <= 2, "vae_load_tensor: dtype must be 0 (f32), 1 (bf16) or 2 (f16)");
    const std::string key(key_c);
    if (!e->expected.count(key)) {
        set_last_error("vae_load_tensor: unexpected key '" + key + "'");
        return K5_ERR_INVALID;
    }
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= static_cast<size_t>(shape[i]);
    const size_t bytes = n * dsize(dtype);
    if (bytes > e->stage_bytes) {
        if (e->stage) cudaFree(e->stage);
        e->stage = nullptr;
        e->stage_bytes = 0;
        K5_CHECK_CUDA(cudaMalloc(&e->stage, bytes));
        e->stage_bytes = bytes;
    }
    K5_CHECK_CUDA(cudaMemcpy(e->stage, data, bytes, cudaMemcpyDefault));
    const size_t dot = key.rfind('.');
    const std::string base = key.substr(0, dot), suf = key.substr(dot + 1);
    auto bad_shape = [&](const std::string& want) {
        set_last_error("vae_load_tensor: shape mismatch for '" + key + "': expected " + want);
        return K5_ERR_INVALID;
    };
    if (base == "post_quant_conv") {
        const int L = e->c.latent_channels;
        if (suf == "weight") {
            if (n != static_cast<size_t>(L) * L) return bad_shape("[L, L, 1, 1, 1]");
            K5_TRY(convert_to_f32(e->stage, dtype, e->pq_w, n, true, 0));
        } else {
            if (n != static_cast<size_t>(L)) return bad_shape("[L]");
            K5_TRY(convert_to_f32(e->stage, dtype, e->pq_b, n, true, 0));
        }
    } else if (e->convs.count(base)) {
        Conv& c = *e->convs[base];
        if (suf == "weight") {
            if (n != static_cast<size_t>(c.cout) * c.cin * c.taps || (ndim == 5 && (shape[0] != c.cout || shape[1] != c.cin)))
                return bad_shape("[" + std::to_string(c.cout) + ", " + std::to_string(c.cin) + ", k, k, k]");
            K5_TRY(repack_conv_weight(e->stage, dtype, c.cout, c.cin, c.taps, c.cin_pad, c.w, 0));
        } else {
            if (n != static_cast<size_t>(c.cout)) return bad_shape("[" + std::to_string(c.cout) + "]");
            K5_TRY(convert_to_f32(e->stage, dtype, c.b, n, true, 0));
        }
    } else if (e->norms.count(base)) {
        Norm& nn = *e->norms[base];
        if (n != static_cast<size_t>(nn.c)) return bad_shape("[" + std::to_string(nn.c) + "]");
        K5_TRY(convert_to_f32(e->stage, dtype, suf == "weight" ? nn.g : nn.b, n, false, 0));
    } else if (e->lins.count(base)) {
        Linear& l = *e->lins[base];
        if (suf == "weight") {
            if (n != static_cast<size_t>(l.n) * l.n) return bad_shape("[C, C]");
            K5_TRY(convert_to_bf16(e->stage, dtype, l.w, l.n, l.n, l.n, 0));
        } else {
            if (n != static_cast<size_t>(l.n)) return bad_shape("[C]");
            K5_TRY(convert_to_f32(e->stage, dtype, l.b, n, true, 0));
        }
    } else {
        set_last_error("vae_load_tensor: no destination for key '" + key + "'");
        return K5_ERR_INVALID;
    }
    K5_CHECK_CUDA(cudaStreamSynchronize(0));
    e->loaded.insert(key);
    return K5_OK;
}

int vae_finalize(Vae* e) {
    std::string missing;
    int n = 0;
    for (const std::string& k : e->expected)
        if (!e->loaded.count(k) && n++ < 8) missing += (missing.empty() ? "" : ", ") + k;
    if (n) {
        set_last_error("vae_finalize: " + std::to_string(n) + " tensors missing: " + missing + (n > 8 ? ", ..." : ""));
        return K5_ERR_STATE;
    }
    if (e->stage) {
        cudaFree(e->stage);
        e->stage = nullptr;
        e->stage_bytes = 0;
    }
    e->finalized = true;
    return K5_OK;
}

namespace {

int group_stats(Vae* e, const bf16* x, size_t P, int C, cudaStream_t st) {
    count_launch(2);
    int nblocks = 0;
    K5_TRY(gn_channel_sums(x, P, C, e->sums, &nblocks, st));
    return gn_finalize(e->sums, nblocks, P, C, GROUPS, GN_EPS, e->stats, st);
}

// GroupNorm + SiLU + pad, then the 3x3x3 convolution
int norm_conv(Vae* e, const bf16* x, int T, int H, int W, const Norm& n, const Conv& c, const bf16* resid, bf16* out,
              cudaStream_t st) {
    K5_TRY(group_stats(e, x, static_cast<size_t>(T) * H * W, n.c, st));
    count_launch(2);
    K5_TRY(pad_gather(x, T, H, W, n.c, 1, 1, 1, e->stats, n.g, n.b, GROUPS, true, e->pad, st));
    return conv3d_causal(e->pad, T, H, W, c.cin_pad, c.w, c.cout, c.cout_pad, c.b, resid, c.cout, out, c.cout, st);
}

// HunyuanVideoResnetBlockCausal3D (vae.py:254-275).  x = buf[xi]; returns the index of the buffer holding the result.
int resnet(Vae* e, const Resnet& r, int xi, int T, int H, int W, int* out_idx, cudaStream_t st) {
    bf16* x = e->buf[xi];
    bf16* h = e->buf[(xi + 1) % 3];
    K5_TRY(norm_conv(e, x, T, H, W, r.n1, r.c1, nullptr, h, st));
    if (!r.has_sc) {
        // out = conv2(...) + x, written over x (each thread reads exactly the residual elements it overwrites)
        K5_TRY(norm_conv(e, h, T, H, W, r.n2, r.c2, x, x, st));
        *out_idx = xi;
        return K5_OK;
    }
    bf16* s = e->buf[(xi + 2) % 3];
    GemmEpilogue g;
    g.out = s;
    g.ldo = r.sc.cout;
    g.bias = r.sc.b;
    count_launch(1);
    K5_TRY(gemm_bf16(x, r.sc.cin, r.sc.w, r.sc.cin_pad, T * H * W, r.sc.cout, r.sc.cin, EPI_STORE, g, st));
    K5_TRY(norm_conv(e, h, T, H, W, r.n2, r.c2, s, s, st));
    *out_idx = (xi + 2) % 3;
    return K5_OK;
}

// vae.py:343-359 + diffusers Attention (1 head, head_dim = C, frame-causal additive mask); in place on buf[xi]
int mid_attention(Vae* e, int xi, int T, int H, int W, cudaStream_t st) {
    const int C = e->top, hw = H * W, N = T * hw;
    bf16* x = e->buf[xi];
    bf16* xn = e->buf[(xi + 1) % 3];                  // GroupNorm(x)          [N, C]
    bf16* qkv = e->buf[(xi + 2) % 3];                 // q | k | v | o         4 x [N, C]
    K5_REQUIRE(hw % 64 == 0, "vae attention: latent H * W must be a multiple of 64");
    bf16 *q = qkv, *k = qkv + static_cast<size_t>(N) * C, *v = k + static_cast<size_t>(N) * C, *o = v + static_cast<size_t>(N) * C;
    K5_TRY(group_stats(e, x, N, C, st));
    count_launch(1);
    K5_TRY(gn_apply(x, N, C, e->stats, e->attn_gn.g, e->attn_gn.b, GROUPS, false, xn, st));
    const Linear* lin[3] = {&e->q, &e->k, &e->v};
    bf16* dst[3] = {q, k, v};
    for (int i = 0; i < 3; ++i) {
        GemmEpilogue g;
        g.out = dst[i];
        g.ldo = C;
        g.bias = lin[i]->b;
        count_launch(1);
        K5_TRY(gemm_bf16(xn, C, lin[i]->w, C, N, C, C, EPI_STORE, g, st));
    }
    count_launch(1);
    K5_TRY(transpose_bf16(v, N, C, C, e->vt, N, st));
    const float scale = 1.0f / sqrtf(static_cast<float>(C));
    for (int f = 0; f < T; ++f) {
        const int keys = (f + 1) * hw;                // frames <= f are visible
        GemmEpilogue gs;
        gs.out_f32 = e->scores + static_cast<size_t>(f) * hw * N;
        gs.ldo = N;
        count_launch(3);
        K5_TRY(gemm_bf16(q + static_cast<size_t>(f) * hw * C, C, k, C, hw, keys, C, EPI_F32, gs, st));
        K5_TRY(softmax_frame_causal(e->scores, N, N, e->probs, N, hw, scale, f * hw, hw, st));
        GemmEpilogue go;
        go.out = o + static_cast<size_t>(f) * hw * C;
        go.ldo = C;
        K5_TRY(gemm_bf16(e->probs + static_cast<size_t>(f) * hw * N, N, e->vt, N, hw, C, keys, EPI_STORE, go, st));
    }
    GemmEpilogue g;                                    // x = bf16(bf16(to_out(o) + b) + x)
    g.out = x;
    g.ldo = C;
    g.bias = e->o.b;
    g.resid = x;
    g.ldr = C;
    g.gate = e->ones;
    count_launch(1);
    return gemm_bf16(o, C, e->o.w, C, N, C, C, EPI_GATE, g, st);
}

// post_quant_conv + HunyuanVideoDecoder3D.forward on latent frames [t0, t0 + T) -> out [F, 8H, 8W, 3]
int decode_tile(Vae* e, const float* z, int Tz, int t0, int T, int H, int W, bf16* out, cudaStream_t st) {
    count_launch(2);
    K5_TRY(post_quant_pad(z, e->c.latent_channels, T, H, W, t0, Tz, e->pq_w, e->pq_b, e->pad, st));
    int xi = 0;
    K5_TRY(conv3d_causal(e->pad, T, H, W, e->conv_in.cin_pad, e->conv_in.w, e->conv_in.cout, e->conv_in.cout_pad, e->conv_in.b,
                         nullptr, 0, e->buf[xi], e->conv_in.cout, st));
    K5_TRY(resnet(e, e->mid[0], xi, T, H, W, &xi, st));
    K5_TRY(mid_attention(e, xi, T, H, W, st));
    K5_TRY(resnet(e, e->mid[1], xi, T, H, W, &xi, st));
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 3; ++j) K5_TRY(resnet(e, e->up[i][j], xi, T, H, W, &xi, st));
        if (i < 3) {
            int ft, fs;
            up_factor(i, ft, fs);
            const Conv& c = e->ups[i];
            count_launch(2);
            K5_TRY(pad_gather(e->buf[xi], T, H, W, c.cin, ft, fs, fs, nullptr, nullptr, nullptr, GROUPS, false, e->pad, st));
            T = ft == 2 ? 1 + (T - 1) * 2 : T;
            H *= fs;
            W *= fs;
            const int yi = (xi + 1) % 3;
            K5_TRY(conv3d_causal(e->pad, T, H, W, c.cin_pad, c.w, c.cout, c.cout_pad, c.b, nullptr, 0, e->buf[yi], c.cout, st));
            xi = yi;
        }
    }
    return norm_conv(e, e->buf[xi], T, H, W, e->norm_out, e->conv_out, nullptr, out, st);
}

}  // namespace

// vae.py:1144-1204 (_temporal_tiled_decode) with the tiling apply_tiling() installed: tile_frames / stride_frames are
// the SAMPLE-frame tile size and stride chosen by get_dec_optimal_tiling ((17, 8) for the 5 s and 10 s videos);
// tile_frames <= 0 decodes the latent in one piece.  z: fp32 [C, T, H, W]; out: bf16 [3, 4 (T - 1) + 1, 8H, 8W].
int vae_decode(Vae* e, const float* z, int T, int H, int W, int tile_frames, int stride_frames, bf16* out, cudaStream_t st) {
    K5_REQUIRE(e->finalized, "vae_decode: call k5_vae_finalize first");
    K5_REQUIRE(z && out && T > 0 && H > 0 && W > 0, "vae_decode: bad arguments");
    // every workspace size is symmetric in (H, W) - activations T*H*W*C, padded volume (T+2)(H+2)(W+2)*C, scores (T*H*W)^2 -
    // so a portrait latent (96 x 64 for 768 x 512 pixels, t2v_pipeline.py:122-125) fits the landscape workspace
    K5_REQUIRE((H <= e->c.max_height && W <= e->c.max_width) || (H <= e->c.max_width && W <= e->c.max_height),
               "vae_decode: latent larger than the engine's workspace");
    const int F = (T - 1) * 4 + 1, HH = 8 * H, WW = 8 * W;
    const int lat_min = tile_frames > 0 ? (tile_frames - 1) / 4 : 0;
    if (tile_frames <= 0 || T <= lat_min + 1) {
        K5_REQUIRE(T <= e->c.max_tile_frames, "vae_decode: too many latent frames for an un-tiled decode (workspace)");
        K5_TRY(decode_tile(e, z, T, 0, T, H, W, e->tile[0], st));
        count_launch(1);
        return emit_frames(e->tile[0], 0, nullptr, 0, 0, F, HH, WW, F, 0, out, st);
    }
    const int min_frames = tile_frames - 1;                     // tile_sample_min_num_frames
    K5_REQUIRE(min_frames % 4 == 0 && stride_frames % 4 == 0 && stride_frames > 0 && stride_frames <= min_frames,
               "vae_decode: tile / stride must be 4k+1 / 4k sample frames");
    const int lat_stride = stride_frames / 4, blend = min_frames - stride_frames;
    K5_REQUIRE(lat_min + 1 <= e->c.max_tile_frames, "vae_decode: tile larger than the engine's workspace");
    // a tile's own blended frames [0, blend) must not reach the frames [kept - blend, kept) its successor reads
    K5_REQUIRE(2 * blend <= min_frames, "vae_decode: blend regions of consecutive tiles overlap (unsupported tiling)");
    int dst = 0, prev_total = 0;
    for (int i = 0, n = 0; i < T - lat_min + 1; i += lat_stride, ++n) {
        bf16* cur = e->tile[n & 1];
        const bf16* prev = e->tile[(n & 1) ^ 1];
        const int Tt = std::min(lat_min + 1, T - i);
        K5_TRY(decode_tile(e, z, T, i, Tt, H, W, cur, st));
        const int frames = (Tt - 1) * 4 + 1;
        const bool last = i + lat_stride >= T - lat_min + 1;
        count_launch(1);
        if (n == 0) {
            const int cnt = std::min(stride_frames + 1, F - dst);
            K5_TRY(emit_frames(cur, 0, nullptr, 0, 0, cnt, HH, WW, F, dst, out, st));
            dst += cnt;
        } else {
            // tile n (first frame dropped): frames 0..blend-1 blended with the last `blend` frames of tile n-1
            const int have = frames - 1;
            int cnt = std::min(last ? min_frames : stride_frames, have);
            cnt = std::min(cnt, F - dst);
            K5_TRY(emit_frames(cur, 1, prev, prev_total - blend, std::min(blend, cnt), cnt, HH, WW, F, dst, out, st));
            dst += cnt;
        }
        prev_total = frames;
    }
    K5_REQUIRE(dst == F, "vae_decode: tile schedule did not cover the video (unsupported tiling)");
    return K5_OK;
}

Vae* vae_new(const k5_vae_config* cfg, int* rc) {
    Vae* e = new Vae();
    e->c = *cfg;
    *rc = vae_init(e);
    if (*rc != K5_OK) {
        delete e;
        return nullptr;
    }
    return e;
}
void vae_delete(Vae* e) { delete e; }

}  // namespace k5
