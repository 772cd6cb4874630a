#!/usr/bin/env python
"""Times the dense attention of the 5 s size (S = 47 616, 28 heads) as ONE launch (the tuned kernel) and as two launches split
by key rows (the PART instantiation, whose loop ptxas schedules differently - profiles/r2_attention_part_template.md), for the
library named by K5_LIB_PATH.  Not a pytest file; used to A/B the position of the s_full probe in the PART kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))
from kandinsky import ops  # noqa: E402


def main():
    S, heads = 47616, 28
    g = torch.Generator(device="cuda").manual_seed(0)

    def rms(x):
        x = x.float().view(S, heads, 64)
        return (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).view(S, heads * 64).to(torch.bfloat16)

    q = rms(torch.randn(S, heads * 64, device="cuda", generator=g))
    k = rms(torch.randn(S, heads * 64, device="cuda", generator=g))
    v = torch.randn(S, heads * 64, device="cuda", generator=g).to(torch.bfloat16)
    bound = 64 * 0.125 * 1.4426950408889634 * 1.02
    out = torch.empty_like(q)

    def timed(fn, reps=6):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    one = timed(lambda: ops.attention(q, k, v, heads, score_bound=bound, out=out))
    ref = out.clone()
    two = timed(lambda: ops.attention_split(q, k, v, heads, bound, 23808, out=out))
    same = bool(torch.equal(out, ref))
    print(f"{os.environ.get('K5_LIB_PATH') or 'libk5.so'}: one launch {one:.3f} ms, split in two {two:.3f} ms "
          f"(+{100 * (two / one - 1):.1f} %), bit-identical {same}", flush=True)


if __name__ == "__main__":
    main()
