"""Scratch: dump what k5_nabla_select writes for the case that tripped the scatter assertion."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200")); sys.path.insert(0, ROOT)
from kandinsky import ops
from oracle import dit_oracle as O
S, heads, P = 1024, 4, 0.6
g = torch.Generator(device="cuda").manual_seed(50)
q = torch.randn(S, heads * 64, device="cuda", generator=g).bfloat16()
k = torch.randn(S, heads * 64, device="cuda", generator=g).bfloat16()
nb = S // 64
sta_cpu = O.sta_mask(nb // 4, 2, 2, 3, 3, 3)
sta = sta_cpu.to(torch.uint8).cuda()
cnt, idx = ops.nabla_select(q, k, heads, P, sta)
torch.cuda.synchronize()
print("cnt", cnt.min().item(), cnt.max().item(), cnt.dtype, cnt.shape)
print("idx", idx.min().item(), idx.max().item(), idx.dtype, idx.shape)
print(cnt[0]); print(idx[0, :4])
