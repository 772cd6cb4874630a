#!/usr/bin/env python
"""Block-sparse (NABLA) attention timing at the 10 s size (S = 93 696 tokens = 1 464 blocks of 64, 28 heads) for block
selections of different density: the STA window alone (wT = 11, wH = wW = 3 -> 4.8 %) and the window OR'ed with a random
selection.  Prints the time per launch and the rate over the SELECTED blocks (K5_VARIANT_BOUND=1: the DiT's bounded kernel).
Not a pytest file."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))
from kandinsky import ops  # noqa: E402


def main():
    T, Hb, Wb, heads, D = 61, 4, 6, 28, 1792
    nb = T * Hb * Wb
    S = nb * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(S, 3 * D, device="cuda", generator=g).bfloat16()
    bound = None
    if os.environ.get("K5_VARIANT_BOUND") == "1":       # RMS-normalised q / k + the proven score bound -> fixed-offset kernel
        for c in (0, D):
            x = qkv[:, c:c + D].float().reshape(S, heads, 64)
            qkv[:, c:c + D] = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).reshape(S, D).bfloat16()
        bound = 8 * 8 * 0.125 * 1.4426950408889634 * 1.02
    o = torch.empty(S, D, device="cuda", dtype=torch.bfloat16)
    sta = ops.sta_mask(T, Hb, Wb, 11, 3, 3).bool()
    for dens in [float(x) for x in os.environ.get("K5_DENS", "0,0.1,0.3").split(",")]:
        sel = sta[None].expand(heads, nb, nb).clone()
        if dens > 0:
            sel |= torch.rand(heads, nb, nb, device="cuda", generator=g) < dens
        cnt = sel.sum(-1).to(torch.int32).contiguous()
        idx = torch.argsort((~sel).to(torch.int8), dim=-1, stable=True).to(torch.int32).contiguous()
        rho = float(cnt.sum()) / (heads * nb * nb)
        del sel

        def run():
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], heads, kv_count=cnt, kv_index=idx, out=o, score_bound=bound)

        for _ in range(2):
            run()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(4):
            run()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 4
        fl = 4.0 * S * S * D * rho
        print(f"bounded={bound is not None} density {rho:.3f}: {ms:.2f} ms = {fl / ms / 1e9:.0f} TFLOP/s of the selected blocks", flush=True)


if __name__ == "__main__":
    main()
