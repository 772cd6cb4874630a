#!/usr/bin/env python
"""Copy-engine peer-to-peer rates over NVLink on one node (not a pytest file): what the overlapped K|V all-gather of the
temporal shard (csrc/engine.cu) can count on.  One process, all visible GPUs: GPU 0 pushes a slab-sized buffer (44 MB =
6 144 rows x 3 584 bf16, the slab of an 8-rank shard at the 5 s size) to one peer on one stream, then to all peers at
once on one stream per peer - the two forms the engine used (round 2: one stream, then one per peer).
Then ALL GPUs push their slab to all peers at once (the real traffic pattern of a block: every rank r serves r-1, r-2, ...
round-robin over a few copy streams), every copy queued behind one gate event so that host enqueue time stays outside.
`nvidia-smi nvlink -gt d` reports N/A on these boxes, so timing the copies is the NVLink evidence there is."""
import torch


def main():
    n = torch.cuda.device_count()
    if n < 2:
        print("p2p: fewer than 2 GPUs visible")
        return
    nbytes = 6144 * 3584 * 2
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}") for i in range(1, n)]
    for i in range(1, n):
        assert torch.cuda.can_device_access_peer(0, i), f"no peer access 0 -> {i}"
    torch.cuda.set_device(0)
    streams = [torch.cuda.Stream(device="cuda:0") for _ in dst]

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def one_peer():
        with torch.cuda.stream(streams[0]):
            dst[0].copy_(src, non_blocking=True)

    def all_serial():
        with torch.cuda.stream(streams[0]):
            for d in dst:
                d.copy_(src, non_blocking=True)

    def all_parallel():
        for s, d in zip(streams, dst):
            with torch.cuda.stream(s):
                d.copy_(src, non_blocking=True)

    ms = timed(one_peer)
    print(f"p2p 0 -> 1, one stream: {nbytes / 1e6:.0f} MB in {ms:.3f} ms = {nbytes / ms / 1e6:.0f} GB/s")
    if n > 2:
        ms = timed(all_serial)
        print(f"p2p 0 -> {n - 1} peers, ONE stream: {ms:.3f} ms for {(n - 1) * nbytes / 1e6:.0f} MB = {(n - 1) * nbytes / ms / 1e6:.0f} GB/s")
        ms = timed(all_parallel)
        print(f"p2p 0 -> {n - 1} peers, one stream PER PEER: {ms:.3f} ms for {(n - 1) * nbytes / 1e6:.0f} MB = {(n - 1) * nbytes / ms / 1e6:.0f} GB/s")
    if n > 2:
        all_to_all(n, nbytes)


def all_to_all(n, nbytes):
    """Every GPU pushes one slab to every peer, nearest consumer first, over `ns` copy streams per GPU; all copies wait
    for one gate event, each GPU times its own pushes (gate passed -> last copy done); reported: the slowest GPU."""
    src = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{r}") for r in range(n)]
    dst = [[torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{p}") if p != r else None for p in range(n)] for r in range(n)]
    for ns in (1, 2, 4, n - 1):
        streams = [[torch.cuda.Stream(device=f"cuda:{r}") for _ in range(ns)] for r in range(n)]
        worst = []
        for rep in range(6):
            torch.cuda.set_device(0)
            gate = torch.cuda.Event()
            torch.cuda._sleep(int(3e7))                      # ~15 ms: everything below is queued before the gate opens
            gate.record()
            t0, t1 = [], []
            for r in range(n):
                torch.cuda.set_device(r)
                a = torch.cuda.Event(enable_timing=True)
                for st in streams[r]:
                    st.wait_event(gate)
                a.record(streams[r][0])
                for k in range(1, n):
                    p = (r - k) % n
                    with torch.cuda.stream(streams[r][(k - 1) % ns]):
                        dst[r][p].copy_(src[r], non_blocking=True)
                for st in streams[r][1:]:
                    streams[r][0].wait_stream(st)
                b = torch.cuda.Event(enable_timing=True)
                b.record(streams[r][0])
                t0.append(a)
                t1.append(b)
            for r in range(n):
                torch.cuda.synchronize(r)
            if rep >= 2:
                worst.append(max(a.elapsed_time(b) for a, b in zip(t0, t1)))
        ms = sum(worst) / len(worst)
        print(f"p2p ALL {n} GPUs -> all peers at once, {ns} copy stream(s) per GPU: slowest GPU {ms:.3f} ms for "
              f"{(n - 1) * nbytes / 1e6:.0f} MB out (and in) = {(n - 1) * nbytes / ms / 1e6:.0f} GB/s per GPU and direction")
    torch.cuda.set_device(0)


if __name__ == "__main__":
    main()
