#!/usr/bin/env python
"""Copy-engine peer-to-peer rates over NVLink on one node (not a pytest file): what the overlapped K|V all-gather of the
temporal shard (csrc/engine.cu) can count on.  One process, all visible GPUs: GPU 0 pushes a slab-sized buffer (44 MB =
6 144 rows x 3 584 bf16, the slab of an 8-rank shard at the 5 s size) to one peer on one stream, then to all peers at
once on one stream per peer - the two forms the engine used (round 2: one stream, then one per peer).
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


if __name__ == "__main__":
    main()
