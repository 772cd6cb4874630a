#!/usr/bin/env python
"""Times the VAE decode of one 5 s video latent [1,16,31,64,96] -> [1,3,121,512,768] (BASELINE.json configs[4], the
decode leg) through the reference-facing mirror (vae.decode(z).sample -> k5_vae_decode), random-init weights.
Algorithmic work: 118.84 TFLOP per 5-latent-frame tile x 14 tiles.  Not a pytest file."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))
sys.path.insert(0, ROOT)
from kandinsky.models.vae import AutoencoderKLHunyuanVideo, decoder_state_dict_shapes  # noqa: E402


def main():
    T = int(os.environ.get("K5_VAE_T", 31))
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    sd = {}
    shapes = decoder_state_dict_shapes()
    for k, shp in shapes.items():
        if "norm" in k:
            t = 1.0 + 0.1 * torch.randn(shp, device=dev, generator=g) if k.endswith("weight") else 0.05 * torch.randn(shp, device=dev, generator=g)
        else:
            w = shp if k.endswith("weight") else shapes[k[:-4] + "weight"]
            fan = 1
            for d in w[1:]:
                fan *= d
            t = (torch.rand(shp, device=dev, generator=g) * 2 - 1) / fan ** 0.5
        sd[k] = t.half()
    vae = AutoencoderKLHunyuanVideo(max_latent=(5, 64, 96))
    vae.load_state_dict(sd)
    vae.to(dev)
    z = torch.randn(1, 16, T, 64, 96, device=dev, generator=g)
    out = vae.decode(z).sample
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.float()).all())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = vae.decode(z).sample
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    ntiles = len(range(0, T - 4 + 1, 2)) if T > 5 else 1
    tf = 118.84 * ntiles
    print(f"vae decode T={T}: {ms:.1f} ms for {ntiles} tiles = {ms / ntiles:.1f} ms/tile = {tf / ms * 1e3:.0f} TFLOP/s (algorithmic "
          f"{tf:.0f} TFLOP), output {tuple(out.shape)} {out.dtype}, mean |x| {float(out.float().abs().mean()):.3f}", flush=True)


if __name__ == "__main__":
    main()
