"""VAE decoder on a real B200 through the C ABI: the implicit-GEMM causal convolution against torch's conv3d, and the
whole decoder (k5_vae_decode via the AutoencoderKLHunyuanVideo mirror) against the vectors minted by the reference's
own vae.py and against the CPU oracle.  Tolerances are calibrated against the same graph in fp32 (the decoder is ~60
bf16 rounding points deep; see tests/test_vae_oracle.py)."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as VO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _conv(x, w, bias, resid=None):
    from kandinsky._lib import check, lib, ptr, stream_ptr

    T, H, W, Cin = x.shape
    Cout = w.shape[0]
    out = torch.empty(T, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    ws = torch.empty((T + 2) * (H + 2) * (W + 2) * Cin + 27 * ((Cout + 63) // 64 * 64) * Cin, device="cuda", dtype=torch.bfloat16)
    check(lib().k5_conv3d_causal(ptr(x), T, H, W, Cin, ptr(w), Cout, ptr(bias), ptr(resid), ptr(out), ptr(ws), stream_ptr()))
    return out


@pytest.mark.parametrize("T,H,W,Cin,Cout,res", [(3, 8, 8, 64, 64, False), (5, 16, 16, 128, 64, True), (2, 4, 32, 64, 128, False),
                                                (3, 64, 96, 128, 256, True), (1, 32, 48, 256, 64, False), (5, 8, 24, 64, 512, True)])
def test_conv3d_causal_matches_torch(T, H, W, Cin, Cout, res):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(T, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) * (27 * Cin) ** -0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device="cuda", generator=g).to(torch.bfloat16).float()
    r = torch.randn(T, H, W, Cout, device="cuda", generator=g).to(torch.bfloat16) if res else None
    out = _conv(x, w, b, r)
    xn = x.permute(3, 0, 1, 2)[None].float()                                     # NCTHW
    xp = F.pad(xn, (1, 1, 1, 1, 2, 0), mode="replicate")                        # vae.py:138-161
    ref = (F.conv3d(xp, w.float()) + b.view(1, -1, 1, 1, 1)).to(torch.bfloat16)
    ref = ref[0].permute(1, 2, 3, 0)
    if res:
        ref = (ref.float() + r.float()).to(torch.bfloat16)
    assert rel_l2(out, ref) < 3e-3
    assert float((out.float() - ref.float()).abs().max()) < 0.06


def _build(widths, max_latent):
    from kandinsky.models.vae import AutoencoderKLHunyuanVideo

    sd = VO.synthetic_state_dict(widths, seed=0)
    vae = AutoencoderKLHunyuanVideo(block_out_channels=widths, max_latent=max_latent)
    vae.load_state_dict(sd)
    return vae.to("cuda"), sd


@pytest.mark.parametrize("name", ["vae_full_width_3x8x8", "vae_tiled_9x8x8"])
def test_decoder_matches_reference_golden(name):
    rec = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    z = rec["z"]
    vae, sd = _build(rec["widths"], (5, 8, 8))
    tile = None
    if rec["tiling"] is not None:
        (_, ft, ht, wt), (fs, hs, ws) = rec["tiling"]
        vae.apply_tiling((1, ft, ht, wt), (fs, hs, ws))
        out = vae._decode(z.cuda()).sample
        tile = (ft, fs)
    else:
        out = vae.decode(z.cuda()).sample
    assert out.shape == rec["out"].shape and out.dtype == torch.bfloat16
    VO.ROUNDING = False
    try:
        gold = VO.decode(sd, z, tile)
    finally:
        VO.ROUNDING = True
    err, ref_noise, own_noise = rel_l2(out, rec["out"]), rel_l2(rec["out"], gold), rel_l2(out, gold)
    print(f"{name}: engine-vs-reference {err:.2e}  engine-vs-fp32 {own_noise:.2e}  reference-vs-fp32 {ref_noise:.2e}")
    assert err < 1.25 * ref_noise and own_noise < 1.25 * ref_noise
    u8, u8_ref = VO.to_uint8(out.cpu()), VO.to_uint8(rec["out"])               # generation_utils.py:222
    assert float((u8.float() - u8_ref.float()).abs().mean()) < 2.0


def test_portrait_latent_decodes_in_the_landscape_workspace():
    """768 x 512 pixels = latent 96 x 64 (t2v_pipeline.py:122-125) must decode in an engine sized for 64 x 96: here 16 x 8
    in a workspace declared as 8 x 16, against the CPU oracle (bf16 rounding points on / off = the noise floor)."""
    widths = (64, 64, 128, 128)
    vae, sd = _build(widths, (5, 8, 16))
    z = torch.randn(1, 16, 3, 16, 8, generator=torch.Generator().manual_seed(3))
    out = vae.decode(z.cuda()).sample
    assert out.shape == (1, 3, 9, 128, 64)
    ref = VO.decode(sd, z, None)
    VO.ROUNDING = False
    try:
        gold = VO.decode(sd, z, None)
    finally:
        VO.ROUNDING = True
    own, oracle = rel_l2(out, gold), rel_l2(ref, gold)
    print(f"portrait: engine-vs-fp32 {own:.2e}  oracle-vs-fp32 {oracle:.2e}  engine-vs-oracle {rel_l2(out, ref):.2e}")
    assert own < 1.25 * oracle


def test_decode_is_deterministic_and_tiles_agree_with_oracle_schedule():
    """Same latent twice -> identical video; the tiled schedule produces exactly 4 (T - 1) + 1 frames."""
    rec = torch.load(os.path.join(GOLD, "vae_tiled_9x8x8.pt"), weights_only=False)
    vae, _ = _build(rec["widths"], (5, 8, 8))
    (_, ft, ht, wt), (fs, hs, ws) = rec["tiling"]
    vae.apply_tiling((1, ft, ht, wt), (fs, hs, ws))
    a = vae._decode(rec["z"].cuda()).sample
    b = vae._decode(rec["z"].cuda()).sample
    assert a.shape[2] == 4 * (rec["z"].shape[2] - 1) + 1
    assert torch.equal(a, b)            # GroupNorm statistics are block partials summed in a fixed order: no atomics


def test_vae_rejects_bad_arguments():
    vae, _ = _build((64, 64, 128, 128), (5, 8, 8))
    with pytest.raises(ValueError):
        vae._decode(torch.zeros(1, 16, 3, 16, 16).cuda())                        # larger than the engine's workspace
    with pytest.raises(ValueError):
        vae.decode(torch.zeros(1, 8, 3, 8, 8).cuda())                            # wrong latent channel count
    vae.apply_tiling((1, 9, 64, 64), (9, 64, 64))
    with pytest.raises(NotImplementedError):
        vae._decode(torch.zeros(1, 16, 3, 16, 16).cuda())                        # would need spatial tiles
    from kandinsky.models.vae import AutoencoderKLHunyuanVideo

    v2 = AutoencoderKLHunyuanVideo(block_out_channels=(64, 64, 128, 128), max_latent=(5, 8, 8))
    with pytest.raises(RuntimeError):
        v2.load_state_dict({})                                                   # strict: every decoder tensor is required


# ---------------------------------------------------------------------------------------------------------------------
# Real shapes of the 5 s decode: one temporal tile is latent 5 x 64 x 96 -> 17 x 512 x 768 pixels.
# The goldens above are 8 x 8 latents; these compare the kernels that only exist at full size (conv3d<128, MT = 2> at
# 512 x 768, the 3-channel conv_out, the up-sampling gathers, hd = 512 attention over 30 720 tokens) with torch on the
# same GPU, fp32 without TF32.

@pytest.fixture
def no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("T,H,W,Cin,Cout,res", [(17, 512, 768, 128, 128, True),      # up_blocks.3 resnets (35 % of the FLOPs)
                                                (17, 512, 768, 128, 3, False),       # conv_out
                                                (9, 256, 384, 256, 256, True),       # up_blocks.2
                                                (9, 256, 384, 256, 128, False),      # 256 -> 128 (first resnet of up_blocks.3)
                                                (5, 128, 192, 512, 512, True),       # up_blocks.1
                                                (5, 64, 96, 16, 512, False)])        # conv_in
def test_conv3d_real_shapes_match_torch(no_tf32, T, H, W, Cin, Cout, res):
    g = torch.Generator(device="cuda").manual_seed(11)
    Cp = max(Cin, 64)                                                            # the engine pads input channels to 64
    x = torch.zeros(T, H, W, Cp, device="cuda", dtype=torch.bfloat16)
    x[..., :Cin] = torch.randn(T, H, W, Cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.zeros(Cout, Cp, 3, 3, 3, device="cuda", dtype=torch.bfloat16)
    w[:, :Cin] = (torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) * (27 * Cin) ** -0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device="cuda", generator=g).to(torch.bfloat16).float()
    r = torch.randn(T, H, W, Cout, device="cuda", generator=g).to(torch.bfloat16) if res else None
    out = _conv(x, w, b, r)
    torch.cuda.synchronize()
    # torch reference frame by frame (a 17-frame fp32 volume plus cuDNN workspace is needlessly large)
    xn = x[..., :Cin].permute(3, 0, 1, 2)[None]
    worst, num, den = 0.0, 0.0, 0.0
    for t in range(T):
        lo = max(t - 2, 0)
        xt = xn[:, :, lo:t + 1].float()
        if t < 2:
            xt = torch.cat([xt[:, :, :1]] * (2 - t) + [xt], dim=2)              # replicate pad in front (vae.py:138-161)
        xp = F.pad(xt, (1, 1, 1, 1, 0, 0), mode="replicate")
        ref = (F.conv3d(xp, w[:, :Cin].float()) + b.view(1, -1, 1, 1, 1)).to(torch.bfloat16)[0, :, 0].permute(1, 2, 0)
        if res:
            ref = (ref.float() + r[t].float()).to(torch.bfloat16)
        d = out[t].float() - ref.float()
        worst = max(worst, float(d.abs().max()))
        num += float(d.pow(2).sum())
        den += float(ref.float().pow(2).sum())
    err = (num / den) ** 0.5
    print(f"conv3d {T}x{H}x{W} {Cin}->{Cout}: rel-L2 {err:.2e} max {worst:.3f}")
    assert err < 3e-3 and worst < 0.07


def test_full_size_tile_matches_the_oracle_on_the_gpu(no_tf32):
    """One whole temporal tile at the 5 s size with the real widths (128, 256, 512, 512): k5_vae_decode against the
    oracle graph run by torch on the same GPU, with and without its bf16 rounding points (the noise floor)."""
    widths = (128, 256, 512, 512)
    vae, sd = _build(widths, (5, 64, 96))
    z = torch.randn(1, 16, 5, 64, 96, generator=torch.Generator().manual_seed(5))
    vae.apply_tiling((1, 17, 512, 768), (16, 512, 768))                          # one tile: no temporal tiling
    out = vae._decode(z.cuda()).sample
    again = vae._decode(z.cuda()).sample
    assert out.shape == (1, 3, 17, 512, 768) and torch.equal(out, again)
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = VO.decode(sdc, z.cuda(), None).cpu()
        torch.cuda.empty_cache()
        VO.ROUNDING = False
        try:
            gold = VO.decode(sdc, z.cuda(), None).cpu()
        finally:
            VO.ROUNDING = True
    torch.cuda.empty_cache()
    own, oracle, direct = rel_l2(out, gold), rel_l2(ref, gold), rel_l2(out, ref)
    print(f"full-size tile: engine-vs-fp32 {own:.2e}  oracle(bf16 points)-vs-fp32 {oracle:.2e}  engine-vs-oracle {direct:.2e}")
    assert own < 1.25 * oracle
    u8, u8_ref = VO.to_uint8(out.cpu()), VO.to_uint8(ref)
    assert float((u8.float() - u8_ref.float()).abs().mean()) < 1.0


def test_mid_block_attention_scores_stay_fp32(no_tf32):
    """v6 at its real size (30 720 tokens, hd = 512) with LARGE logits: q = k makes the diagonal score ~ |q|^2 / sqrt(512),
    where a bf16-rounded score would move the softmax by percents.  Checked through the operator pieces the engine uses
    (K5_EPI_F32 GEMM) against torch SDPA semantics in fp32 on a row sample."""
    from kandinsky import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    N, C = 6144, 512
    q = (torch.randn(N, C, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    k = (q.float() + 0.3 * torch.randn(N, C, device="cuda", generator=g)).to(torch.bfloat16)
    s = ops.linear(q, k, None, epilogue="f32")
    ref = q.float() @ k.float().t()
    assert float(ref.abs().max()) * C ** -0.5 > 30.0                              # logits far beyond bf16's exact range
    assert float((s - ref).abs().max()) < 8e-6 * float(ref.abs().max())          # fp32 summation order only (a bf16 ulp there: 4)
    p, pr = torch.softmax(s * C ** -0.5, -1), torch.softmax(ref * C ** -0.5, -1)
    assert float((p - pr).abs().max()) < 1e-3
    rounded = torch.softmax(ref.to(torch.bfloat16).float() * C ** -0.5, -1)        # what bf16 scores would have given
    assert float((rounded - pr).abs().max()) > 20 * float((p - pr).abs().max())
