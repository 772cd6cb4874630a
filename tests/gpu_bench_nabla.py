#!/usr/bin/env python
"""Times the NABLA block selection (k5_nabla_select: pooling + one map row per thread block) at the 10 s size
(S = 93 696 tokens = 1 464 blocks, 28 heads) with the STA window OR'ed in.  Not a pytest file."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))
from kandinsky import ops  # noqa: E402
from kandinsky._lib import lib  # noqa: E402
from kandinsky.ops import check, ptr, stream_ptr  # noqa: E402


def main():
    T, Hb, Wb, heads, D = 61, 4, 6, 28, 1792
    nb = T * Hb * Wb
    S = nb * 64
    P = float(os.environ.get("K5_NABLA_P", 0.9))
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(S, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(S, D, device="cuda", generator=g).bfloat16()
    sta = ops.sta_mask(T, Hb, Wb, 11, 3, 3)
    cnt = torch.empty(heads, nb, device="cuda", dtype=torch.int32)
    idx = torch.zeros(heads, nb, nb, device="cuda", dtype=torch.int32)
    ws = torch.empty(heads * nb * nb + 2 * nb * heads * 64, device="cuda", dtype=torch.float32)

    def run():
        check(lib().k5_nabla_select(ptr(q), q.stride(0), ptr(k), k.stride(0), S, heads, P, ptr(sta), ptr(cnt), ptr(idx),
                                    ptr(ws), stream_ptr()))

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        run()
    e.record()
    torch.cuda.synchronize()
    print(f"nabla_select S={S} heads={heads} P={P}: {s.elapsed_time(e) / 5:.3f} ms per layer, density "
          f"{float(cnt.sum()) / (heads * nb * nb):.3f}", flush=True)


if __name__ == "__main__":
    main()
