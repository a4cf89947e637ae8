"""All exchange modes of the row-partitioned SpMV in ONE launch (N GPUs are charged N-fold: one
matrix generation, one rendezvous and one partition calibration for every mode instead of one
per mode).  Same workload, partition and timing rules as bench.py (BASELINE config 5, CUDA
events, barrier on both sides, max over ranks); the safest modes run first and every result is
printed as soon as it exists, so a trap in a never-run mode loses only what follows it.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
      --master-port 29611 tools/scale_modes.py --steps 20 \\
      --modes "push fused nccl mcast-push mcast mcast-chunked chunked mcast-stream stream"

Prints one JSON line per (mode, barrier); `speedup_vs` divides --n1-ms (the measured 1-GPU
step, default round 1's 4.18 ms) by the step time.  Not a bench arm: bench.py stays the contract.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    # the modes whose put kernel waits on the SpMV (they can trap) go last
    ap.add_argument("--modes", default="push fused nccl mcast-push mcast mcast-chunked chunked "
                                       "mcast-stream stream")
    ap.add_argument("--barriers", default="nccl symm", help="tried for the mcast modes")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--npr", type=int, default=100)
    ap.add_argument("--n1-ms", type=float, default=4.18)
    ap.add_argument("--recuts", type=int, default=3)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import sprs_b200 as sp
    from sprs_b200 import generate as G
    from sprs_b200.dist import (ChunkedPushAllGatherSpMV, FusedAllGatherSpMV, McastAllGatherSpMV,
                                PushAllGatherSpMV, RowPartitionedSpMV, StreamAllGatherSpMV,
                                fit_row_cost, nnz_balanced_bounds, rebalance_bounds)
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = sp.Context.default(local)
    t0 = time.time()
    full = G.make_matrix(ctx, "rmat", args.n, args.npr, 0x5EED0005)
    n, nnz = args.n, full.nnz
    x = G.normal_vector(ctx, n)

    def say(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(flag):
        t = torch.tensor([1.0 if flag else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def make(mode, barrier, blk, bnds):
        if mode == "nccl":
            yb = torch.zeros(n, device=dev, dtype=torch.float64)
            return RowPartitionedSpMV(bnds, rank, world, yb, lambda xv, ys: G.spmv(ctx, blk, xv, ys),
                                      dist=dist)
        if mode.startswith("mcast"):
            return McastAllGatherSpMV(ctx, blk.mirror, bnds, rank, world, n, dist, dev,
                                      mode=mode.partition("-")[2] or "fused", barrier=barrier)
        cls = {"push": PushAllGatherSpMV, "fused": FusedAllGatherSpMV,
               "stream": StreamAllGatherSpMV, "chunked": ChunkedPushAllGatherSpMV}[mode]
        return cls(ctx, blk.mirror, bnds, rank, world, n, dist, dev)

    def timed(op, steps, split=False):
        """(ms per step, compute-only ms per step), max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        if not split:
            e0.record()
            for _ in range(steps):
                op.step(x)
            e1.record()
            torch.cuda.synchronize()
            dist.barrier()
            return max_over_ranks(e0.elapsed_time(e1) / steps), None
        tsum = 0.0
        for _ in range(steps):
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            op.compute(x)
            e1.record()
            op.exchange()
            torch.cuda.synchronize()
            tsum += e0.elapsed_time(e1)
        return None, tsum / steps

    # ---- partition: row cost fitted from the plain kernel, then equal-time re-cuts measured
    #      with the push operator (validated in round 1); every mode then runs on the same cut
    bounds = nnz_balanced_bounds(full.indptr, world)
    blk = full.slice_rows(bounds[rank], bounds[rank + 1])
    yt = torch.empty(max(blk.rows, 1), device=dev, dtype=torch.float64)
    for _ in range(2):
        G.spmv(ctx, blk, x, yt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        G.spmv(ctx, blk, x, yt)
    e1.record()
    torch.cuda.synchronize()
    mine = torch.tensor([blk.nnz, blk.rows, e0.elapsed_time(e1) / 3e3], device=dev,
                        dtype=torch.float64)
    allm = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allm, mine)
    row_cost = fit_row_cost([m.tolist() for m in allm])
    del blk, yt
    bounds = nnz_balanced_bounds(full.indptr, world, row_cost=row_cost)
    best = None
    for cut in range(args.recuts + 1):
        blk = full.slice_rows(bounds[rank], bounds[rank + 1])
        op = make("push", "nccl", blk, bounds)
        for _ in range(3):
            op.step(x)
        _, comp = timed(op, 4, split=True)
        tm = torch.tensor([comp], device=dev, dtype=torch.float64)
        allt = [torch.empty_like(tm) for _ in range(world)]
        dist.all_gather(allt, tm)
        times = [float(t.item()) for t in allt]
        op.close()
        del op, blk
        torch.cuda.empty_cache()
        if best is None or max(times) < best[0]:
            best = (max(times), list(bounds))
        say({"partition_round": cut, "row_cost": row_cost, "compute_ms_max": max(times),
             "compute_ms_mean": sum(times) / world})
        if cut == args.recuts or max(times) <= 1.015 * sum(times) / world:
            break
        nb = rebalance_bounds(full.indptr, bounds, times, row_cost=row_cost)
        if nb == bounds:
            break
        bounds = nb
    bounds = best[1]
    blk = full.slice_rows(bounds[rank], bounds[rank + 1])
    ref = torch.empty(n, device=dev, dtype=torch.float64)
    G.spmv(ctx, full, x, ref)
    scale = float(ref.abs().max().item())
    del full
    torch.cuda.empty_cache()
    say({"setup_seconds": round(time.time() - t0, 1), "nnz": nnz, "world": world, "bounds": bounds})

    for mode in args.modes.split():
        for barrier in (args.barriers.split() if mode.startswith("mcast") else ["nccl"]):
            op, err = None, None
            try:
                op = make(mode, barrier, blk, bounds)
            except Exception as e:  # e.g. no multicast support: every rank must agree to skip
                err = repr(e)
            if not all_ok(op is not None):
                say({"mode": mode, "barrier": barrier, "skipped": err or "failed on another rank"})
                if op is not None and hasattr(op, "close"):
                    op.close()
                continue
            for _ in range(args.warmup):
                op.step(x)
            torch.cuda.synchronize()
            good = bool(((op.y - ref).abs() <= 1e-9 * scale).all().item())
            ms, _ = timed(op, args.steps)
            _, comp = timed(op, 5, split=True)
            comp = max_over_ranks(comp)
            say({"mode": mode, "barrier": barrier, "ms_per_step": ms, "compute_ms": comp,
                 "gflops": 2.0 * nnz / ms / 1e6, "speedup_vs_n1": args.n1_ms / ms,
                 "correct": all_ok(good)})
            if hasattr(op, "close"):
                op.close()
            del op
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
