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
    ws = torch.empty((T + 2) * (H + 2) * (W + 2) * Cin + 27 * Cout * Cin, device="cuda", dtype=torch.bfloat16)
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
    """768 x 512 pixels = latent 96 x 64 (t2v_pipeline.py:122-125) must decode in an engine sized for 64 x 96: here 12 x 8
    in a workspace declared as 8 x 12, against the CPU oracle (bf16 rounding points on / off = the noise floor)."""
    widths = (64, 64, 128, 128)
    vae, sd = _build(widths, (5, 8, 12))
    z = torch.randn(1, 16, 3, 12, 8, generator=torch.Generator().manual_seed(3))
    out = vae.decode(z.cuda()).sample
    assert out.shape == (1, 3, 9, 96, 64)
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
    assert rel_l2(a, b) < 1e-6          # GroupNorm statistics use double atomics: order-independent to ~1e-16


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
