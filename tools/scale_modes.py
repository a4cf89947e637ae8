"""All exchange modes of the row-partitioned SpMV in ONE multi-GPU launch (8 GPUs are charged
8-fold): the matrix, the rendezvous and the measured partition are shared, every mode is checked
against the single-GPU product of the same matrix (computed on every rank) and timed with CUDA
events, max over ranks.  One JSON line per mode on rank 0.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/scale_modes.py \
      [--n 10000000 --npr 100 --steps 20 --warmup 5 --modes "nccl push fused push+mc fused+mc"]

Modes: nccl = NCCL all_gather after the kernel; push / fused = the library's communicator over
CUDA IPC peer mappings; "+mc" = the same through the NVSwitch multicast address of y.  fused =
the SpMV kernel stages a tile's rows in shared memory and sends them with one TMA bulk store per
target; "+direct" = a plain store per finished row instead (SPRS_B200_SPMV_PEER_STORES=direct,
the round's first form).  The multicast modes run last (they are the ones that can fail hard on
an unsupported box).
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
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--npr", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--modes", default="nccl push fused+direct fused push+mc fused+mc+direct fused+mc")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import sprs_b200 as sp
    from sprs_b200 import generate as G
    from sprs_b200.dist import (Comm, CommHostSpMV, CommSpMV, RowPartitionedSpMV, fit_row_cost,
                                nnz_balanced_bounds, rebalance_bounds)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = sp.Context.default(local)
    n = args.n

    def say(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    def allmax(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    t0 = time.time()
    full = G.make_matrix(ctx, "rmat", n, args.npr, 0x5EED0005)
    x = G.normal_vector(ctx, n)
    y_ref = torch.empty(n, device=dev, dtype=torch.float64)
    G.spmv(ctx, full, x, y_ref)
    # single-GPU time of the same matrix on this box (the denominator of the speed-up)
    for _ in range(3):
        G.spmv(ctx, full, x, y_ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        G.spmv(ctx, full, x, y_ref)
    e1.record()
    torch.cuda.synchronize()
    (n1_ms,) = allmax([e0.elapsed_time(e1) / 10])
    scale = float(y_ref.abs().max().item()) + 1e-300
    ids = [Comm.unique_id(ctx) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = Comm(ctx, ids[0], rank, world)
    say({"setup_seconds": round(time.time() - t0, 1), "n1_ms_this_box": n1_ms, "world": world,
         "multicast_supported": comm.multicast, "nnz": full.nnz})

    # ---- partition: nnz balance, fitted row cost, then measured equal-time re-cuts (plain SpMV)
    bounds = nnz_balanced_bounds(full.indptr, world)
    row_cost = 0.0

    def block_time(b):
        a = full.slice_rows(b[rank], b[rank + 1])
        yt = torch.empty(max(b[rank + 1] - b[rank], 1), device=dev, dtype=torch.float64)
        for _ in range(2):
            G.spmv(ctx, a, x, yt)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(4):
            G.spmv(ctx, a, x, yt)
        c1.record()
        torch.cuda.synchronize()
        return a, c0.elapsed_time(c1) / 4

    a, ms = block_time(bounds)
    allm = comm.allgather_f64([a.nnz, bounds[rank + 1] - bounds[rank], ms / 1e3])
    row_cost = fit_row_cost([m.tolist() for m in allm])
    bounds = nnz_balanced_bounds(full.indptr, world, row_cost=row_cost)
    for rnd in range(5):  # `a` always belongs to `bounds` when the loop is left
        del a
        torch.cuda.empty_cache()
        a, ms = block_time(bounds)
        times = [float(v[0]) for v in comm.allgather_f64([ms])]
        say({"partition_round": rnd, "row_cost": round(row_cost, 2), "block_ms": [round(t, 4) for t in times]})
        if max(times) <= 1.02 * (sum(times) / world) or rnd == 4:
            break
        nb = rebalance_bounds(full.indptr, bounds, times, row_cost=row_cost)
        if nb == bounds:
            break
        bounds = nb
    r0, r1 = bounds[rank], bounds[rank + 1]

    def run_mode(mode):
        parts = mode.split("+")
        name, mc = parts[0], ("mc" if "mc" in parts[1:] else "")
        os.environ["SPRS_B200_SPMV_PEER_STORES"] = "direct" if "direct" in parts[1:] else "tma"
        if name == "nccl":
            yb = torch.zeros(n, device=dev, dtype=torch.float64)
            op = RowPartitionedSpMV(bounds, rank, world, yb, lambda xv, ys: G.spmv(ctx, a, xv, ys), dist=dist)
        else:
            op = CommSpMV(comm, a.mirror, bounds, n, dev, exchange=name, multicast=(mc == "mc"))
            if mc == "mc" and not op.multicast:
                op.close()
                return {"mode": mode, "skipped": "no multicast binding on this box"}
        op.y.fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        op.step(x)
        torch.cuda.synchronize()
        comm.check()
        ok = bool(((op.y - y_ref).abs() <= 1e-9 * scale).all().item())
        (bad,) = allmax([0.0 if ok else 1.0])
        for _ in range(args.warmup):
            op.step(x)
        torch.cuda.synchronize()
        dist.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
                torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            evs[i][0].record()
            op.compute(x)
            evs[i][1].record()
            op.exchange()
            evs[i][2].record()
        s1.record()
        torch.cuda.synchronize()
        comm.check()
        kern = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
        coll = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
        total, kmax, cmax = allmax([s0.elapsed_time(s1) / args.steps, kern, coll])
        per_rank = [round(float(v[0]), 4) for v in comm.allgather_f64([kern])]
        if hasattr(op, "close"):
            op.close()
        return {"mode": mode, "correct": bad == 0.0, "ms_per_step": total, "compute_ms_max": kmax,
                "barrier_or_collective_ms_max": cmax, "compute_ms_per_rank": per_rank,
                "speedup_vs_n1": n1_ms / total, "efficiency": n1_ms / total / world}

    for mode in args.modes.split():
        try:
            res = run_mode(mode)
        except Exception as e:  # one mode failing must not hide the others' numbers
            res = {"mode": mode, "error": repr(e)[:300]}
        say(res)

    os.environ["SPRS_B200_SPMV_PEER_STORES"] = "tma"
    # ---- host-vector form (e2e): every rank moves only its own slices.  "same cut": x and y
    # sliced like the SpMV's row blocks; "pcie cut": x in equal slices, row blocks re-balanced
    # with 8 bytes of PCIe per y row on top of the SpMV cost (what bench.py's e2e does)
    for mc, pcie_cut in ((False, False), (True, False), (True, True)):
        try:
            if pcie_cut:
                c_pcie = 8.0 / 50e9 * (full.nnz / world / (n1_ms / world * 1e-3))
                eb = nnz_balanced_bounds(full.indptr, world, row_cost=row_cost + c_pcie)
                xb = [n * g // world for g in range(world + 1)]
                del a
                torch.cuda.empty_cache()
                a = full.slice_rows(eb[rank], eb[rank + 1])
                bounds, (r0, r1) = eb, (eb[rank], eb[rank + 1])
            else:
                xb = bounds
            hop = CommHostSpMV(comm, a.mirror, bounds, n, multicast=mc, x_bounds=xb)
            hx = torch.empty(max(xb[rank + 1] - xb[rank], 1), dtype=torch.float64).pin_memory()
            hx[:xb[rank + 1] - xb[rank]].copy_(x[xb[rank]:xb[rank + 1]])
            hy = torch.empty(max(r1 - r0, 1), dtype=torch.float64).pin_memory()
            for _ in range(2):
                hop.step(hx.data_ptr(), hy.data_ptr())
            dist.barrier()
            t1 = time.perf_counter()
            for _ in range(5):
                hop.step(hx.data_ptr(), hy.data_ptr())
            torch.cuda.synchronize()
            (ms,) = allmax([(time.perf_counter() - t1) * 1e3 / 5])
            ok = bool(((hy[:r1 - r0].to(dev) - y_ref[r0:r1]).abs() <= 1e-9 * scale).all().item())
            (bad,) = allmax([0.0 if ok else 1.0])
            say({"e2e_host_slices": "multicast" if mc and hop.x.multicast_ptr else "ipc",
                 "cut": "pcie" if pcie_cut else "same", "rows_per_rank_max": max(
                     bounds[g + 1] - bounds[g] for g in range(world)),
                 "ms_per_step": ms, "correct": bad == 0.0})
            hop.close()
        except Exception as e:
            say({"e2e_host_slices": "multicast" if mc else "ipc", "error": repr(e)[:300]})
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
