#!/usr/bin/env python
"""Micro-benchmarks of the two tensor-core kernels at the 5 s problem sizes, next to the
library kernels the reference would use on this box (cuBLAS via torch.matmul, FlashAttention-2).  CUDA-event timing,
inputs larger than L2 or rotated, 3 warm-ups.  Prints one line per case; not a pytest file."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))
from kandinsky import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    S = int(os.environ.get("K5_BENCH_S", 47616))
    D, F = 1792, 7168
    dev = "cuda"
    torch.manual_seed(0)
    only = os.environ.get("K5_BENCH_ONLY", "")
    gemms = [(S, 3 * D, D, "heads"), (S, D, D, "gate"), (S, F, D, "gelu"), (S, D, F, "gate"), (S, D, D, "store")]
    for (M, N, K, epi) in ([] if only in ("attn", "ln") else gemms):
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw = {}
        if epi == "gate":
            kw = dict(resid=out, gate=torch.randn(N, device=dev), bias=torch.randn(N, device=dev))
        if epi == "heads":
            kw = dict(bias=torch.randn(N, device=dev), norm_w0=torch.ones(64, device=dev), norm_w1=torch.ones(64, device=dev),
                      norm_split=D, norm_cols=2 * D, rope_cols=2 * D, rope=torch.randn(M, 32, 2, device=dev))
        ms = timeit(lambda: ops.linear(a, w, epilogue=epi, out=out, **kw))
        ms_t = timeit(lambda: torch.matmul(a, w.t()))
        fl = 2.0 * M * N * K
        print(f"gemm {epi:6s} M={M} N={N} K={K}: k5 {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s | torch.matmul {ms_t:.3f} ms = {fl / ms_t / 1e9:.0f} TFLOP/s", flush=True)
        del a, w, out
    if only == "gemm":
        return
    heads = 28
    if only == "ln":
        x = torch.randn(S, D, device=dev).bfloat16()
        sc, sh = torch.randn(D, device=dev), torch.randn(D, device=dev)
        y = torch.empty_like(x)
        ms = timeit(lambda: ops.ln_rows(x, sc, sh, out=y), iters=50)
        print(f"ln_modulate S={S}: {ms:.4f} ms = {2 * S * D * 2 / ms / 1e6:.0f} GB/s", flush=True)
        return
    qkv = torch.randn(S, 3 * D, device=dev).bfloat16()
    o = torch.empty(S, D, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], heads, out=o), iters=5, warm=2)
    fl = 4.0 * S * S * D
    line = f"attn S={S} h={heads} poly={os.environ.get('K5_ATTN_POLY', 'default')}: k5 {ms:.2f} ms = {fl / ms / 1e9:.0f} TFLOP/s"
    if only == "attn":
        print(line, flush=True)
        return
    try:
        from flash_attn import flash_attn_func

        q4 = qkv[:, :D].reshape(1, S, heads, 64)
        k4 = qkv[:, D:2 * D].reshape(1, S, heads, 64)
        v4 = qkv[:, 2 * D:].reshape(1, S, heads, 64)
        ms_f = timeit(lambda: flash_attn_func(q4, k4, v4), iters=3, warm=1)
        line += f" | flash_attn2 {ms_f:.2f} ms = {fl / ms_f / 1e9:.0f} TFLOP/s"
    except Exception as ex:  # noqa: BLE001
        line += f" | flash_attn unavailable ({type(ex).__name__})"
    print(line, flush=True)
    L = 256
    kv = torch.randn(L, 2 * D, device=dev).bfloat16()
    ms = timeit(lambda: ops.attention(qkv[:, :D], kv[:, :D], kv[:, D:], heads, out=o), iters=5, warm=2)
    print(f"cross-attn S={S} L={L}: k5 {ms:.3f} ms = {4.0 * S * L * D / ms / 1e9:.0f} TFLOP/s", flush=True)
    x = torch.randn(S, D, device=dev).bfloat16()
    sc, sh = torch.randn(D, device=dev), torch.randn(D, device=dev)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.ln_rows(x, sc, sh, out=y))
    print(f"ln_modulate S={S}: {ms:.3f} ms = {2 * S * D * 2 / ms / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
