#!/usr/bin/env python
"""Cross-process check of the temporal shard over real NVLink peer memory (not a pytest file: launch with
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/gpu_shard_ranks.py`).
Every rank runs the sharded forward and, on its own GPU, the single-engine forward of the same inputs; its frames
must be BIT-IDENTICAL.  With the scatter + barrier form of the all-gather (K5_DIST_OVERLAP=0) both walk the KV tiles
in natural order.  With the overlapped all-gather (K5_DIST_OVERLAP=1, across processes) a rank walks the slabs starting at its
own, so the single engine is told to walk them, for every query row, in the order of the rank owning that row
(K5_DEBUG_KV_ORDER, csrc/engine.cu): any difference left would be a slab read before it arrived.  Also runs the device sampler with CFG and gathers the latent
through generate()."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kandinsky-5_b200"))
sys.path.insert(0, ROOT)
from kandinsky.generation_utils import generate  # noqa: E402
from kandinsky.models.dit import DiffusionTransformer3D  # noqa: E402
from kandinsky.models.parallelize import frame_partition, parallelize_dit  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(O.LITE_CFG, num_visual_blocks=3)
    T, H, W, L, Ln = 2 * world + 1, 32, 64, 40, 17       # 512 tokens per frame: slab boundaries at multiples of 256 rows
    S = T * (H // 2) * (W // 2)
    sd = O.synthetic_state_dict(cfg, seed=0)
    models = []
    for _ in range(2):
        m = DiffusionTransformer3D(**cfg, max_tokens=S, max_text_tokens=64)
        m.load_state_dict(sd, assign=True)
        models.append(m.to(dev))
    full, shard = models
    parallelize_dit(shard)
    g = torch.Generator().manual_seed(1)
    img = torch.randn(T, H, W, 16, generator=g).to(dev)
    text = torch.randn(L, 3584, generator=g).to(torch.bfloat16).to(dev)
    pooled = torch.randn(1, 768, generator=g).to(torch.bfloat16).to(dev)
    ntext = torch.randn(Ln, 3584, generator=g).to(torch.bfloat16).to(dev)
    npooled = torch.randn(1, 768, generator=g).to(torch.bfloat16).to(dev)
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    ov = os.environ.get("K5_DIST_OVERLAP", "")
    overlapped = ov not in ("", "0") if ov != "" else world >= 8       # the engine's default (csrc/engine.cu, engine_dist_init)
    ref_nat = None
    if overlapped:
        ref_nat = full(img, text, pooled, 700.0, pos, torch.arange(L), scale_factor=(1.0, 2.0, 2.0)).clone()
        os.environ["K5_DEBUG_KV_ORDER"] = str(world)      # read by the single engine at set_grid
        full._grid_key = None
    ref = full(img, text, pooled, 700.0, pos, torch.arange(L), scale_factor=(1.0, 2.0, 2.0))

    def close(x, y):
        err = float((x.float() - y.float()).norm() / y.float().norm())
        return bool(torch.equal(x, y)), err

    ok, errs = True, []
    for it in range(3):
        out = shard(img, text, pooled, 700.0, pos, torch.arange(L), scale_factor=(1.0, 2.0, 2.0))
        torch.cuda.synchronize()
        f0, n = frame_partition(T, world)[rank]
        same, err = close(out[f0:f0 + n], ref[f0:f0 + n])
        errs.append(err)
        ok = ok and same and shard.local_frames() == (f0, n)
    conf = {"metrics": {"scale_factor": (1.0, 2.0, 2.0)}, "model": {"dit_params": dict(cfg), "attention": {"type": "flash"}}}
    te, nte = {"text_embeds": text, "pooled_embed": pooled}, {"text_embeds": ntext, "pooled_embed": npooled}
    a = generate(full, dev, (T, H, W, 16), 3, te, nte, pos, torch.arange(L), torch.arange(Ln), 5.0, 5.0, conf, noise=img)
    b = generate(shard, dev, (T, H, W, 16), 3, te, nte, pos, torch.arange(L), torch.arange(Ln), 5.0, 5.0, conf, noise=img)
    torch.cuda.synchronize()
    f0, n = frame_partition(T, world)[rank]
    same, err = close(b[f0:f0 + n], a[f0:f0 + n])
    ok = ok and same
    if os.environ.get("K5_SHARD_VERBOSE"):
        f0, n = frame_partition(T, world)[rank]
        extra = "" if ref_nat is None else f" order effect (rotated vs natural single engine) {close(ref[f0:f0 + n], ref_nat[f0:f0 + n])[1]:.2e}"
        print(f"  rank {rank}: forward errs {['%.2e' % e for e in errs]} sampler {err:.2e}{extra}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    worst = torch.tensor([max(errs), err], device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"shard x{world} ({'overlapped all-gather' if overlapped else 'scatter + barrier'}): forward + sampler "
              f"bit-identical to the single-engine run on every rank: {bool(flag.item())} (worst rel-L2 "
              f"{worst[0].item():.2e} forward / {worst[1].item():.2e} sampler)", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
