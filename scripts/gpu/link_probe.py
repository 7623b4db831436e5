#!/usr/bin/env python
"""Raw host-link ceiling: plain pinned-memory copies, N ranks at once (one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/gpu/link_probe.py --out gpurun_out/link_probe_N.json

Every rank allocates pinned host buffers and times, between barriers, (a) H2D alone, (b) D2H alone, (c) both directions
at once on two streams.  Reported: per-rank GB/s (min/max over ranks) and the aggregate over all ranks -- the number the
end-to-end figures of bench.py are bounded by, since plaac_score()'s timed region is made of exactly these copies.
No kernel of the library runs here; this is a platform measurement.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mib << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(7)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(h2d, d2h):
        for it in range(args.reps + 1):
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        return n * args.reps / dt / 1e9  # GB/s per direction per rank (max-over-ranks time through the barrier)

    res = {}
    for tag, (a, b) in (("h2d_alone", (1, 0)), ("d2h_alone", (0, 1)), ("bidirectional", (1, 1))):
        gbs = run(a, b)
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        if world > 1:
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            lo, hi = float(lo), float(hi)
        else:
            lo = hi = gbs
        res[tag] = {"per_rank_gbs_per_direction": lo, "aggregate_gbs_per_direction": lo * world,
                    "directions": a + b, "per_rank_max": hi}
    if rank == 0:
        line = {"ranks": world, "mib_per_copy": args.mib, "reps": args.reps, "results": res,
                "note": "pinned host memory (torch pin_memory), cudaMemcpyAsync on one stream per direction; wall clock "
                        "between barriers, so every rate is bounded by the slowest rank"}
        txt = json.dumps(line)
        print(txt, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "w") as f:
                f.write(txt + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
