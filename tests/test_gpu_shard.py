"""Temporal shard on ONE GPU: `world` engines live in this process, each on its own CUDA stream,
and exchange K | V through each other's buffers exactly as ranks on different GPUs do (k5_dist_init maps a peer
of the same process by plain pointer instead of CUDA IPC).  Every rank's frames must be BIT-IDENTICAL to the
single-engine forward: rows are independent in every kernel and the KV tiles are visited in the same order.
The cross-process / NVLink leg of the same code is exercised by bench.py --gpus N and tests/gpu_shard_ranks.py."""
import ctypes
import os

import pytest
import torch

from oracle import dit_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _models(cfg, max_tokens, world):
    from kandinsky.models.dit import DiffusionTransformer3D

    sd = O.synthetic_state_dict(cfg, seed=0)
    out = []
    for _ in range(world + 1):
        m = DiffusionTransformer3D(**cfg, max_tokens=max_tokens, max_text_tokens=256)
        m.load_state_dict(sd, assign=True)
        out.append(m.to("cuda"))
    full, ranks = out[0], out[1:]
    handles = [m.dist_export() for m in ranks]
    for r, m in enumerate(ranks):
        m.dist_init(r, world, handles)
    return full, ranks


def _inputs(rec):
    g = torch.Generator().manual_seed(rec.get("input_seed", 1))
    T, H, W, L = rec["T"], rec["H"], rec["W"], rec["L"]
    img = torch.randn(T, H, W, 16, generator=g)
    text = torch.randn(L, 3584, generator=g).to(torch.bfloat16)
    pooled = torch.randn(1, 768, generator=g).to(torch.bfloat16)
    return img.cuda(), text.cuda(), pooled.cuda()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_forward_is_bit_identical_to_single_engine(world):
    from kandinsky.models.parallelize import frame_partition

    rec = torch.load(os.path.join(GOLD, "tiny_flash_3x16x16.pt"), weights_only=False)
    cfg = rec["cfg"]
    T, H, W, L = rec["T"], rec["H"], rec["W"], rec["L"]
    full, ranks = _models(cfg, T * (H // 2) * (W // 2), world)
    img, text, pooled = _inputs(rec)
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    ref = full(img, text, pooled, 700.0, pos, torch.arange(L), scale_factor=rec["scale_factor"])
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in ranks]
    outs = []
    for _ in range(2):                                    # twice: the K|V double buffer and the epochs roll over
        outs = []
        for m, s in zip(ranks, streams):
            with torch.cuda.stream(s):
                outs.append(m(img, text, pooled, 700.0, pos, torch.arange(L), scale_factor=rec["scale_factor"]))
        torch.cuda.synchronize()
    parts = frame_partition(T, world)
    for r, (m, o) in enumerate(zip(ranks, outs)):
        assert m.local_frames() == parts[r]
        f0, n = parts[r]
        assert torch.equal(o[f0:f0 + n], ref[f0:f0 + n]), f"rank {r} frames differ from the single-engine forward"


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nabla_forward_is_bit_identical_to_single_engine(world):
    """NABLA on a shard: block selection of the rank's query blocks against ALL key blocks (gathered K), STA rows of its
    own blocks, block-sparse attention over the gathered K|V - same lists, same result as one engine."""
    from kandinsky.models.parallelize import frame_partition

    rec = torch.load(os.path.join(GOLD, "tiny_nabla_4x16x16.pt"), weights_only=False)
    cfg = rec["cfg"]
    T, H, W, L = rec["T"], rec["H"], rec["W"], rec["L"]
    full, ranks = _models(cfg, T * (H // 2) * (W // 2), world)
    img, text, pooled = _inputs(rec)
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    nb = rec["nabla"]
    sparse = {"to_fractal": True, "P": nb["P"], "wT": nb["wT"], "wH": nb["wH"], "wW": nb["wW"], "add_sta": True}
    ref = full(img, text, pooled, 500.0, pos, torch.arange(L), scale_factor=rec["scale_factor"], sparse_params=sparse)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in ranks]
    outs = []
    for m, s in zip(ranks, streams):
        with torch.cuda.stream(s):
            outs.append(m(img, text, pooled, 500.0, pos, torch.arange(L), scale_factor=rec["scale_factor"],
                          sparse_params=sparse))
    torch.cuda.synchronize()
    dens = []
    for r, (o, (f0, n)) in enumerate(zip(outs, frame_partition(T, world))):
        assert torch.equal(o[f0:f0 + n], ref[f0:f0 + n]), f"rank {r} frames differ from the single-engine NABLA forward"
        dens.append(ranks[r].last_sparse_density())
    assert abs(sum(dens) / len(dens) - full.last_sparse_density()) < 1e-3      # equal slabs: densities average


def test_sharded_sampler_with_cfg_matches_single_engine():
    """k5_sample on a 2-way shard (CFG: two forwards per step): each rank integrates its own frames only."""
    from kandinsky._lib import check, lib, ptr
    from kandinsky.models.parallelize import frame_partition

    rec = torch.load(os.path.join(GOLD, "tiny_sampler_cfg.pt"), weights_only=False)
    cfg = rec["cfg"]
    T, H, W, L, Ln = rec["T"], rec["H"], rec["W"], rec["L"], rec["Ln"]
    world = 2 if T >= 2 else 1
    if world == 1:
        pytest.skip("golden sampler case has a single frame")
    full, ranks = _models(cfg, T * (H // 2) * (W // 2), world)
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(T, H, W, 16, generator=g).cuda()
    text = torch.randn(L, 3584, generator=g).to(torch.bfloat16).cuda()
    pooled = torch.randn(768, generator=g).to(torch.bfloat16).cuda()
    ntext = torch.randn(Ln, 3584, generator=g).to(torch.bfloat16).cuda()
    npooled = torch.randn(768, generator=g).to(torch.bfloat16).cuda()
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]

    def run(m, img, stream):
        m.set_grid((T, H, W), pos, rec["scale_factor"], False)
        with torch.cuda.stream(stream):
            check(lib().k5_sample(m._engine, ptr(img), rec["steps"], float(rec["guidance_weight"]),
                                  float(rec["scheduler_scale"]), ptr(text), L, ptr(pooled), ptr(ntext), Ln, ptr(npooled),
                                  None, ctypes.c_void_p(stream.cuda_stream)))

    ref = noise.clone()
    run(full, ref, torch.cuda.current_stream())
    torch.cuda.synchronize()
    imgs = [noise.clone() for _ in ranks]
    streams = [torch.cuda.Stream() for _ in ranks]
    torch.cuda.synchronize()
    for m, im, s in zip(ranks, imgs, streams):
        run(m, im, s)
    torch.cuda.synchronize()
    for r, (f0, n) in enumerate(frame_partition(T, world)):
        assert torch.equal(imgs[r][f0:f0 + n], ref[f0:f0 + n])
        other = [i for i in range(T) if not f0 <= i < f0 + n]
        assert torch.equal(imgs[r][other], noise[other])          # frames of other ranks are left untouched


def test_sharded_magcache_sampler_matches_single_engine():
    """k5_sample_magcache on a 2-way shard: the residual caches hold each rank's own rows, the skip schedule is the
    same host array on every rank, and the frames stay bit-identical to the single-engine run (which differs from the
    plain sampler, i.e. the skips really happened)."""
    from kandinsky._lib import check, lib, ptr
    from kandinsky.models.parallelize import frame_partition

    rec = torch.load(os.path.join(GOLD, "tiny_sampler_cfg.pt"), weights_only=False)
    cfg = rec["cfg"]
    T, H, W, L, Ln = rec["T"], rec["H"], rec["W"], rec["L"], rec["Ln"]
    if T < 2:
        pytest.skip("golden sampler case has a single frame")
    steps = max(int(rec["steps"]), 4)
    full, ranks = _models(cfg, T * (H // 2) * (W // 2), 2)
    g = torch.Generator().manual_seed(1)
    noise = torch.randn(T, H, W, 16, generator=g).cuda()
    text = torch.randn(L, 3584, generator=g).to(torch.bfloat16).cuda()
    pooled = torch.randn(768, generator=g).to(torch.bfloat16).cuda()
    ntext = torch.randn(Ln, 3584, generator=g).to(torch.bfloat16).cuda()
    npooled = torch.randn(768, generator=g).to(torch.bfloat16).cuda()
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    sched = (ctypes.c_uint8 * (2 * steps))(*[1 if (i // 2) % 2 == 1 else 0 for i in range(2 * steps)])   # every other step

    def run(m, img, stream, schedule):
        m.set_grid((T, H, W), pos, rec["scale_factor"], False)
        args = (m._engine, ptr(img), steps, float(rec["guidance_weight"]), float(rec["scheduler_scale"]), ptr(text), L,
                ptr(pooled), ptr(ntext), Ln, ptr(npooled), None)
        with torch.cuda.stream(stream):
            if schedule is None:
                check(lib().k5_sample(*args, ctypes.c_void_p(stream.cuda_stream)))
            else:
                check(lib().k5_sample_magcache(*args, schedule, ctypes.c_void_p(stream.cuda_stream)))

    ref, plain = noise.clone(), noise.clone()
    run(full, ref, torch.cuda.current_stream(), sched)
    run(full, plain, torch.cuda.current_stream(), None)
    torch.cuda.synchronize()
    assert not torch.equal(ref, plain)
    imgs = [noise.clone() for _ in ranks]
    streams = [torch.cuda.Stream() for _ in ranks]
    torch.cuda.synchronize()
    for m, im, st in zip(ranks, imgs, streams):
        run(m, im, st, sched)
    torch.cuda.synchronize()
    for r, (f0, n) in enumerate(frame_partition(T, 2)):
        assert torch.equal(imgs[r][f0:f0 + n], ref[f0:f0 + n])
    bad = (ctypes.c_uint8 * (2 * steps))(*([1] * (2 * steps)))           # a skip before anything was cached
    fresh, _ = _models(cfg, T * (H // 2) * (W // 2), 1)
    with pytest.raises(ValueError):
        run(fresh, noise.clone(), torch.cuda.current_stream(), bad)


def test_dist_init_rejects_bad_arguments():
    rec = torch.load(os.path.join(GOLD, "cfg1_block_1x8x8.pt"), weights_only=False)
    from kandinsky.models.dit import DiffusionTransformer3D

    m = DiffusionTransformer3D(**rec["cfg"], max_tokens=64, max_text_tokens=64)
    m.load_state_dict(O.synthetic_state_dict(rec["cfg"], seed=0), assign=True)
    m.to("cuda")
    h = m.dist_export()
    with pytest.raises(ValueError):
        m.dist_init(0, 2, [h])                               # one handle per rank
    with pytest.raises(ValueError):
        m.dist_init(3, 2, [h, h])                            # rank out of range
    m.dist_init(0, 2, [h, h])
    with pytest.raises(ValueError):                          # one frame, two ranks
        m.set_grid((1, 16, 16), [torch.arange(1), torch.arange(8), torch.arange(8)], (1.0, 2.0, 2.0), False)
