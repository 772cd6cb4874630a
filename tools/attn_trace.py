#!/usr/bin/env python
"""Per-tile timeline of the softmax warps of CTA 0 (debug build of the library with -DK5_ATTN_TRACE, selected through
K5_LIB_PATH): clock64 stamps at tile top (0), first exponential (1), middle (2), P published (3), first half of P
stored (7), the pv_done probe result (6), and from the MMA issuer of each query tile: p_ready seen (4), PV issued (5)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))
from kandinsky import ops  # noqa: E402
from kandinsky._lib import check, lib, ptr  # noqa: E402

S, heads, D = 47616, 28, 1792
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(S, 3 * D, device="cuda", generator=g).bfloat16()
for c in (0, D):
    x = qkv[:, c:c + D].float().reshape(S, heads, 64)
    qkv[:, c:c + D] = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).reshape(S, D).bfloat16()
o = torch.empty(S, D, device="cuda", dtype=torch.bfloat16)
buf = torch.zeros(2, 4, 512, 8, device="cuda", dtype=torch.int64)
bound = 8 * 8 * 0.125 * 1.4426950408889634 * 1.02
for _ in range(2):
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], heads, out=o, score_bound=bound)
check(lib().k5_debug_attn_trace(ptr(buf)))
ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], heads, out=o, score_bound=bound)
torch.cuda.synchronize()
t = buf.cpu().double()[:, :, 50:370]          # [tile a, warp, kv tile, stamp]
t0 = t[0, 0, 50, 0]
print("kv tile | per warp (a.w): first exponential, P published   (cycles relative to warp 0.0)")
for j in range(50, 58):
    print(j, " ".join(f"{a}.{w}:{t[a, w, j, 1] - t0:6.0f}/{t[a, w, j, 3] - t0:6.0f}" for a in (0, 1) for w in range(4)))
for a in (0, 1):
    per = (t[a, :, 1:, 0] - t[a, :, :-1, 0]).mean(1)
    print(f"query tile {a}: period per warp", " ".join(f"{x:.0f}" for x in per))
    for name, x, y in (("top->exp", 0, 1), ("exp->mid", 1, 2), ("mid->P half stored", 2, 7), ("half stored->pub", 7, 3)):
        print(f"   {name}: " + " ".join(f"{(t[a, w, :, y] - t[a, w, :, x]).mean():.0f}" for w in range(4)))
    print("   pub->next top: " + " ".join(f"{(t[a, w, 1:, 0] - t[a, w, :-1, 3]).mean():.0f}" for w in range(4)))
    pub = t[a, :, :, 3]
    last = pub.max(0).values
    print("   publish time behind the first warp of the tile: " + " ".join(f"{(pub[w] - pub.min(0).values).mean():.0f}" for w in range(4)))
    print(f"   last warp published -> issuer saw p_ready {(t[a, 0, :, 4] - last).mean():.0f} -> PV issued {(t[a, 0, :, 5] - last).mean():.0f}")
    print("   pv_done probe hit rate: " + " ".join(f"{t[a, w, 1:, 6].mean():.2f}" for w in range(4)))
lag = t[1, 0, :, 1] - t[0, 0, :, 1]
print("lag of warp 1.0's first exponential behind warp 0.0's: mean %.0f std %.0f" % (lag.mean(), lag.std()))
