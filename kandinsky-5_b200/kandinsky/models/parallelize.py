"""Multi-GPU set-up of the DiT.  The reference's `kandinsky/models/parallelize.py:11-102` applies a DTensor
tensor-parallel plan (3 all-reduces of [S, 1792] per block); here the same entry point sets up the engine's
temporal shard instead: rank r owns a contiguous slab of latent frames, all per-token work is
local, and K | V are all-gathered inside the QKV projection kernel over NVLink peer memory.

`frame_partition` is the host-side statement of the split the engine uses (csrc/engine.cu engine_set_grid)."""
import torch


def frame_partition(num_frames, world):
    """[(first_frame, count)] per rank: as even as possible, the first `num_frames % world` ranks get one more."""
    if world < 1 or num_frames < world:
        raise ValueError("the temporal shard needs at least one frame per rank")
    base, rem = divmod(num_frames, world)
    out, f = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((f, n))
        f += n
    return out


def parallelize_dit(dit, tp_mesh=None, group=None):
    """Same call shape as the reference's parallelize_dit(dit, tp_mesh) (utils.py:80-87).  `tp_mesh` is accepted
    for compatibility (its process group is used when given).  Exchanges the IPC handles of the per-rank K|V
    buffers through torch.distributed and arms the engine's shard."""
    import torch.distributed as dist

    if not dist.is_initialized():
        return dit
    if group is None and tp_mesh is not None and hasattr(tp_mesh, "get_group"):
        group = tp_mesh.get_group()
    world = dist.get_world_size(group)
    if world == 1:
        return dit
    rank = dist.get_rank(group)
    handles = [None] * world
    dist.all_gather_object(handles, dit.dist_export(), group=group)
    dit.dist_init(rank, world, handles)
    dist.barrier(group=group)            # nobody starts a forward before every rank has mapped every buffer
    dit._dist_group = group
    return dit


def gather_frames(dit, latent, group=None):
    """After sampling on a temporal shard every rank holds only its own frames of `latent` [T,H,W,C] up to date:
    exchange the slabs (one broadcast per rank, uneven slabs allowed) so that every rank returns the full latent,
    like the reference's replicated output."""
    import torch.distributed as dist

    world = getattr(dit, "dist_world", 1)
    if world == 1:
        return latent
    group = group if group is not None else getattr(dit, "_dist_group", None)
    for r, (f0, n) in enumerate(frame_partition(latent.shape[0], world)):
        slab = latent[f0:f0 + n].contiguous()
        dist.broadcast(slab, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        latent[f0:f0 + n] = slab
    return latent
