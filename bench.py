#!/usr/bin/env python
"""bench.py -- headline benchmark of the sprs product path on B200.

Metric (BASELINE.json): CSR SpMV f64 GFLOP/s and achieved fraction of the HBM roofline
(2*nnz flops over 12*nnz + 8*n bytes), at 1/2/4/8 B200, beside the sprs CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one `y = A x` over the whole matrix (all ranks together).  The default
workload is BASELINE config 5 -- the configuration the metric is quoted on: 10M x 10M
R-MAT, ~100 nnz/row (~1e9 nnz, 12 GB, fits one GPU) -- at every N (strong scaling:
contiguous nnz-balanced row blocks per rank, x replicated, all-gather of y).
Other workloads: spmv_rand_1m (config 2), spmm_rand_1m_k64 (config 3),
spgemm_rmat_500k (config 4).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_other  # noqa: E402  (needs ROOT on sys.path)

WORKLOADS = {
    # name: (kind, n, nnz_per_row, generator)
    "spmv_rmat_10m": ("spmv", 10_000_000, 100, "rmat"),
    "spmv_rand_1m": ("spmv", 1_000_000, 32, "rand"),
    "spmm_rand_1m_k64": ("spmm", 1_000_000, 32, "rand"),
    "spgemm_rmat_500k": ("spgemm", 500_000, 16, "rmat"),
    # small variants for quick checks
    "spmv_rmat_1m": ("spmv", 1_000_000, 100, "rmat"),
}
SEEDS = {"spmv_rmat_10m": 0x5EED0005, "spmv_rand_1m": 0x5EED0002, "spmm_rand_1m_k64": 0x5EED0002,
         "spgemm_rmat_500k": 0x5EED0004, "spmv_rmat_1m": 0x5EED0005}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (NVML, every 2 ms;
    same fields as the profiling recipe's nvidia-smi clocks line)."""

    def __init__(self, device):
        self.device, self.samples, self.stop_flag, self.t = device, [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device
            if vis:
                try:
                    idx = int(vis.split(",")[device])
                except ValueError:
                    idx = device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.time(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self, t0, t1):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + self.err]}
        self.stop_flag = True
        self.t.join()
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        bits = 0
        for s in inside:
            bits |= s[2]
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        return {"sm_mhz": statistics.median([s[1] for s in inside]) if inside else None,
                "sm_max_mhz": self.max, "reasons": sorted(k for k, v in names.items() if bits & v),
                "power_w_max": max([s[3] for s in inside]) if inside else None,
                "samples": len(inside)}


def sample_rows_to_host(a, target_nnz, nblocks=8):
    """A bounded sample of the SAME matrix for the CPU baseline: `nblocks` contiguous row
    blocks spread over the matrix, ~target_nnz non-zeros in total, as one host CSR."""
    import torch
    ip = a.indptr.to(torch.int64)
    n = a.rows
    per = max(1, target_nnz // nblocks)
    parts_ip, parts_ind, parts_dat, total, rows = [np.zeros(1, np.int64)], [], [], 0, 0
    for b in range(nblocks):
        r0 = (n * b) // nblocks
        s = int(ip[r0])
        r1 = int(torch.searchsorted(ip, torch.tensor([s + per], device=ip.device))[0])
        r1 = max(r0 + 1, min(r1, (n * (b + 1)) // nblocks))
        e = int(ip[r1])
        parts_ip.append((ip[r0 + 1:r1 + 1] - s + total).cpu().numpy())
        parts_ind.append(a.indices[s:e].cpu().numpy().view(np.uint32))
        parts_dat.append(a.data[s:e].cpu().numpy())
        total += e - s
        rows += r1 - r0
    return (np.concatenate(parts_ip).astype(np.uint32), np.concatenate(parts_ind),
            np.concatenate(parts_dat), rows)


def cpu_spmv_baseline(a, x_t, budget_nnz):
    """Faithful CPU restatement (oracle port), 1 thread -- sprs SpMV is single-threaded
     -- on a bounded sample of the same matrix; plus the all-cores row-chunked
    extension, labelled as not in the reference."""
    from oracle import oracle as O
    hip, hind, hdat, rows = sample_rows_to_host(a, budget_nnz)
    hx = x_t.cpu().numpy()
    nnz = int(hip[-1])
    y = np.zeros(rows)
    O.mul_acc_mat_vec_csr(hip, hind, hdat, hx, y)  # warm-up (page-in)
    ts = []
    for _ in range(3):
        y[:] = 0
        t = time.perf_counter()
        O.mul_acc_mat_vec_csr(hip, hind, hdat, hx, y)
        ts.append(time.perf_counter() - t)
    t1 = statistics.median(ts)
    cores = O.num_procs()
    ts = []
    for _ in range(3):
        y[:] = 0
        t = time.perf_counter()
        O.ext_spmv_csr_omp(hip, hind, hdat, hx, y, cores)
        ts.append(time.perf_counter() - t)
    tn = statistics.median(ts)
    return {"value": 2.0 * nnz / t1 / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
            "sample": "%d rows / %d nnz of the same matrix (8 row blocks), 1 thread as in sprs "
                      "(SpMV is single-threaded there), median of 3" % (rows, nnz),
            "ext_all_cores": {"value": 2.0 * nnz / tn / 1e9, "cores": cores,
                              "note": "OpenMP row-chunked extension -- NOT in the reference"},
            "host_cores": cores}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    Rust reference cannot be built here), on the box's host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, n, npr, gen = WORKLOADS[args.workload]
    from oracle import oracle as O
    # inputs: same generator as the GPU arm when a GPU is present, else a numpy stand-in
    sample_nnz = 1 << 26
    try:
        import torch
        import sprs_b200 as sp
        from sprs_b200 import generate as G
        ctx = sp.Context.default(0)
        torch.cuda.set_device(0)
        a = G.make_matrix(ctx, gen, n, npr, SEEDS[args.workload])
        x = G.normal_vector(ctx, n)
        hip, hind, hdat, rows = sample_rows_to_host(a, sample_nnz)
        hx = x.cpu().numpy()
        src = "device-generated matrix, 8 row blocks"
        del a, x
        torch.cuda.empty_cache()
    except Exception as e:  # no GPU: uniform random sample of the same shape
        rng = np.random.default_rng(SEEDS[args.workload])
        rows = sample_nnz // npr
        hind = rng.integers(0, n, size=rows * npr, dtype=np.uint32).reshape(rows, npr)
        hind.sort(axis=1)
        hind = hind.reshape(-1)
        hip = (np.arange(rows + 1, dtype=np.uint64) * npr).astype(np.uint32)
        hdat = rng.standard_normal(rows * npr)
        hx = rng.standard_normal(n)
        src = "numpy uniform stand-in (no GPU for the generator: %s)" % type(e).__name__
    nnz = int(hip[-1])
    y = np.zeros(rows)
    for _ in range(args.warmup):
        O.mul_acc_mat_vec_csr(hip, hind, hdat, hx, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.mul_acc_mat_vec_csr(hip, hind, hdat, hx, y)
    dt = (time.perf_counter() - t0) / args.steps
    val = 2.0 * nnz / dt / 1e9
    line = {"impl": "reference", "metric": "csr_spmv_f64_gflops", "value": val,
            "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": n, "nnz_per_row": npr,
                       "sample_nnz": nnz, "sample_rows": rows, "source": src},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": 1, "kind": "port",
                             "sample": "%d nnz per step; sprs SpMV/SpMM are single-threaded "
                                       "(SURVEY F6), so 1 thread IS all the threads the "
                                       "reference path can use; host has %d cores" %
                                       (nnz, O.num_procs())},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pick_exchange_by_measurement(op, args, ctx, a, bounds, rank, world, n, dist, dev, x, mcast_cls):
    """`--exchange auto` at 6+ GPUs: the exchange measured in round 1 (fused peer stores) runs
    against the NVSwitch-multicast exchanges ON THIS BOX, untimed, before the benchmark proper:
    each candidate must construct on every rank, reproduce the validated operator's y (checked
    from a NaN-filled buffer after one step, so a row that has not landed by the barrier
    fails it) and beat it by more than 2 % over 5 device-timed steps (max over ranks) to be
    selected.  A candidate that is unavailable, raises or disagrees is dropped; the report of
    what was tried goes into the JSON line (config.exchange_trial).  SPRS_B200_AUTO_TRIAL=0
    switches the trial off (the round-1 choice is then used as is)."""
    import torch

    def agree(flag):
        t = torch.tensor([1.0 if flag else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def timed(o, steps=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            o.step(x)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(steps):
            o.step(x)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    report = {}
    base_name = args.exchange
    base_ms = timed(op)
    report[base_name] = round(base_ms, 4)
    y_ref = op.y.clone()
    scale = float(y_ref.abs().max().item()) + 1e-300
    best_op, best_name, best_ms = op, base_name, base_ms
    supported = False
    try:
        import torch.distributed._symmetric_memory as symm
        from torch._C._autograd import DeviceType
        supported = bool(symm._SymmetricMemory.has_multicast_support(DeviceType.CUDA, dev.index))
    except Exception as e:
        report["probe"] = "no symmetric-memory multicast probe: %r" % (e,)
    if not agree(supported):
        report.setdefault("probe", "NVSwitch multicast not supported on every rank")
        return op, base_name, report
    for cand in ("mcast-push", "mcast"):
        c_op, err = None, None
        try:
            c_op = mcast_cls(ctx, a.mirror, bounds, rank, world, n, dist, dev,
                             mode=cand.partition("-")[2] or "fused", barrier=args.barrier)
        except Exception as e:
            err = repr(e)
        if not agree(c_op is not None):
            report[cand] = "unavailable: %s" % (err or "failed on another rank")
            if c_op is not None:
                c_op.close()
            continue
        ok, ms = False, None
        try:
            c_op.y.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            got = c_op.step(x)
            torch.cuda.synchronize()
            ok = bool(((got - y_ref).abs() <= 1e-9 * scale).all().item())
            dist.barrier()
        except Exception as e:
            err = repr(e)
        if not agree(ok):
            report[cand] = "rejected: result differs from the validated exchange" if err is None \
                else "rejected: %s" % err
            c_op.close()
            continue
        ms = timed(c_op)
        report[cand] = round(ms, 4)
        if ms < 0.98 * best_ms:
            if best_op is not op:
                best_op.close()
            best_op, best_name, best_ms = c_op, cand, ms
        else:
            c_op.close()
    if best_op is not op and hasattr(op, "close"):
        op.close()
    report["selected"] = best_name
    return best_op, best_name, report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="spmv_rmat_10m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "overlap", "fused", "push", "stream", "chunked", "mcast", "mcast-push", "mcast-stream", "mcast-chunked", "nccl"],
                    help="N>1 all-gather of y: 'overlap' = row block cut into chunks, each "
                         "chunk's y slice pushed to the peers by DMA copies on a second stream "
                         "while the next chunk computes; 'fused' = the SpMV kernel itself stores "
                         "every finished row into the peers' buffers; 'nccl' = one NCCL "
                         "all_gather after the kernel; 'auto' = push below 6 GPUs, fused from 6 "
                         "up, where an untimed trial may replace it by a multicast exchange that "
                         "reproduces its result and is faster (pick_exchange_by_measurement)")
    ap.add_argument("--chunks", type=int, default=4)
    ap.add_argument("--barrier", default="nccl", choices=["nccl", "symm"],
                    help="mcast modes: barrier after the stores -- 1-element NCCL all_reduce, or "
                         "the signal-pad barrier of the symmetric-memory handle")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import sprs_b200 as sp
    from sprs_b200 import generate as G
    from sprs_b200.dist import (FusedAllGatherSpMV, OverlappedAllGatherSpMV, PushAllGatherSpMV,
                                StreamAllGatherSpMV, ChunkedPushAllGatherSpMV, McastAllGatherSpMV,
                                RowPartitionedSpMV, fit_row_cost, nnz_balanced_bounds)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = sp.Context.default(local)
    kind, n, npr, gen = WORKLOADS[args.workload]
    if kind != "spmv":
        return bench_other.run(args, ctx, kind, n, npr, gen, SEEDS[args.workload])
    peaks, peak_src = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])

    # ---- inputs, generated in HBM (every rank builds the same matrix, keeps its block)
    t_gen = time.time()
    full = G.make_matrix(ctx, gen, n, npr, SEEDS[args.workload])
    nnz = full.nnz
    x = G.normal_vector(ctx, n)
    bounds = nnz_balanced_bounds(full.indptr, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    row_cost = 0.0
    if world > 1:
        # calibrate the partition: time this rank's nnz-balanced block, fit
        # t = alpha*nnz + beta*rows over the ranks, re-cut with rows weighted by beta/alpha
        a = full.slice_rows(r0, r1)
        yt = torch.empty(max(r1 - r0, 1), device=dev, dtype=torch.float64)
        for _ in range(2):
            G.spmv(ctx, a, x, yt)
        ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ce0.record()
        for _ in range(3):
            G.spmv(ctx, a, x, yt)
        ce1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([a.nnz, r1 - r0, ce0.elapsed_time(ce1) / 3e3], device=dev,
                            dtype=torch.float64)
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        row_cost = fit_row_cost([m.tolist() for m in allm])
        del a, yt
        bounds = nnz_balanced_bounds(full.indptr, world, row_cost=row_cost)
        r0, r1 = bounds[rank], bounds[rank + 1]
        a = full.slice_rows(r0, r1)
    else:
        a = full
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # rank 0 at N=1 only
        try:
            cpu_base = cpu_spmv_baseline(full, x, 1 << 27)
        except Exception as e:  # a reported baseline must never take the GPU number down
            cpu_base = {"error": repr(e)}
    t_gen = time.time() - t_gen
    auto_exchange = args.exchange == "auto"
    if args.exchange == "auto":
        # measured (profiles/r1_multi_gpu.md): the own put kernel is the fastest exchange at
        # 2 and 4 GPUs (2.27 / 1.33 ms vs nccl 2.39 / 1.36, fused 2.43 / 1.38); its cost grows
        # with the bytes pushed (~0.14 ms at 40 MB, ~0.22 ms at 60 MB), while at 8 GPUs the
        # fused kernel measured 0.82 ms against 0.61 ms of pure compute -> fused from 6 GPUs up
        args.exchange = "fused" if world >= 6 else "push"
    fused = world > 1 and args.exchange in ("fused", "overlap", "push", "stream", "chunked",
                                            "mcast", "mcast-push", "mcast-stream", "mcast-chunked")

    def make_op(a_blk, bnds):
        if world > 1 and args.exchange == "overlap":
            o = OverlappedAllGatherSpMV(ctx, a_blk, bnds, rank, world, n, dist, dev,
                                        chunks=args.chunks, row_cost=row_cost)
        elif fused and args.exchange.startswith("mcast"):
            o = McastAllGatherSpMV(ctx, a_blk.mirror, bnds, rank, world, n, dist, dev,
                                   mode=args.exchange.partition("-")[2] or "fused",
                                   barrier=args.barrier)
        elif fused:
            cls = {"push": PushAllGatherSpMV, "stream": StreamAllGatherSpMV,
                   "chunked": ChunkedPushAllGatherSpMV}.get(
                args.exchange, FusedAllGatherSpMV)
            o = cls(ctx, a_blk.mirror, bnds, rank, world, n, dist, dev)
        else:
            yb = torch.zeros(n, device=dev, dtype=torch.float64)
            o = RowPartitionedSpMV(bnds, rank, world, yb,
                                   lambda xv, ys: G.spmv(ctx, a_blk, xv, ys),
                                   dist=dist if world > 1 else None)
        return o

    op = make_op(a, bounds)
    rebalanced = 0
    if world > 1:
        # measured re-balancing with the REAL operator (the fused kernel's remote stores
        # change the per-row cost): up to two rounds of equal-time re-cuts
        for _ in range(2):
            for _ in range(2):
                op.step(x)
            ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tsum = 0.0
            for _ in range(4):
                torch.cuda.synchronize()
                dist.barrier()
                ce0.record()
                op.compute(x)
                ce1.record()
                op.exchange()
                torch.cuda.synchronize()
                tsum += ce0.elapsed_time(ce1)
            mine = torch.tensor([tsum / 4], device=dev, dtype=torch.float64)
            allt = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allt, mine)
            times = [float(t.item()) for t in allt]
            if max(times) <= 1.03 * (sum(times) / world):
                break
            from sprs_b200.dist import rebalance_bounds
            nb = rebalance_bounds(full.indptr, bounds, times, row_cost=row_cost)
            if nb == bounds:
                break
            if hasattr(op, "close"):
                op.close()
            del op, a
            torch.cuda.empty_cache()
            bounds = nb
            r0, r1 = bounds[rank], bounds[rank + 1]
            a = full.slice_rows(r0, r1)
            op = make_op(a, bounds)
            rebalanced += 1
        del full
        torch.cuda.empty_cache()
    exchange_trial = None
    trial_min = int(os.environ.get("SPRS_B200_AUTO_TRIAL_MIN_GPUS", "6"))
    if auto_exchange and world >= trial_min and os.environ.get("SPRS_B200_AUTO_TRIAL", "1") != "0":
        op, args.exchange, exchange_trial = pick_exchange_by_measurement(
            op, args, ctx, a, bounds, rank, world, n, dist, dev, x, McastAllGatherSpMV)
    y = op.y
    y_views = [y[bounds[g]:bounds[g + 1]] for g in range(world)]
    local_nnz = a.nnz

    def step():
        op.step(x)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    use_sampler = rank == 0 and not os.environ.get("SPRS_BENCH_NO_SAMPLER")
    if use_sampler:  # started BEFORE the barrier so that no rank enters the timed region late
        sampler.start()
        time.sleep(0.05)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream, max over ranks
    launches0 = ctx.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
            torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    tw0 = time.time()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    cpu_t = []
    for i in range(args.steps):
        cpu_t.append(time.perf_counter())
        evs[i][0].record()
        op.compute(x)
        evs[i][1].record()
        op.exchange()
        evs[i][2].record()
    cpu_t.append(time.perf_counter())
    e_stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tw1 = time.time()
    clocks = sampler.stop(tw0, tw1) if use_sampler else None
    launches = ctx.launches - launches0
    total_ms = e_start.elapsed_time(e_stop)
    kern_ms = [evs[i][0].elapsed_time(evs[i][1]) for i in range(args.steps)]
    coll_ms = [evs[i][1].elapsed_time(evs[i][2]) for i in range(args.steps)]
    if os.environ.get("SPRS_BENCH_DEBUG"):
        gaps = [evs[i][2].elapsed_time(evs[i + 1][0]) for i in range(args.steps - 1)]
        print("[rank %d] kern %s\n[rank %d] coll %s\n[rank %d] gap  %s\n[rank %d] cpu_enqueue_ms %s" % (
            rank, ["%.2f" % v for v in kern_ms], rank, ["%.2f" % v for v in coll_ms], rank,
            ["%.2f" % v for v in gaps], rank,
            ["%.2f" % ((cpu_t[i + 1] - cpu_t[i]) * 1e3) for i in range(args.steps)]),
            file=sys.stderr, flush=True)
    t = torch.tensor([total_ms, statistics.mean(kern_ms), statistics.mean(coll_ms)],
                     device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms_avg, coll_ms_avg = t.tolist()
    ms_per_step = total_ms / args.steps

    # ---- e2e: the reference-facing call with HOST buffers: `&A * &x` through the C ABI,
    # x H2D from pinned memory and y D2H inside the timed region, every step.
    import ctypes as C
    rows_local = r1 - r0
    hx = torch.empty(n, dtype=torch.float64).pin_memory()
    hx.copy_(x)
    hy = torch.empty(max(rows_local, 1), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        ctx.check(ctx.lib.sprs_b200_mul_mat_vec(ctx.h, a.mirror.h, C.c_void_p(hx.data_ptr()), n,
                                                C.c_void_p(hy.data_ptr()), rows_local))
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    # parity spot-check of the e2e result against the device-resident result
    ref_y = y_views[rank]
    got_y = hy[:rows_local].to(dev)
    # same kernel, but the chunked exchange tiles sub-blocks separately: compare to rounding
    ok = bool(((got_y - ref_y).abs() <= 1e-9 * (ref_y.abs().max() + 1e-300)).all())

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra and args.workload == "spmv_rmat_10m":
        try:
            del a, full
            torch.cuda.empty_cache()
            extra = bench_small_spmv(ctx, G, hbm_peak, dev)
        except Exception as e:
            extra = {"error": repr(e)}

    if rank == 0:
        flops = 2.0 * nnz
        alg_bytes = 12.0 * nnz + 8.0 * n
        gflops = flops / (ms_per_step * 1e-3) / 1e9
        # roofline of the dominant kernel (spmv_tile_kernel; the carry fix-up kernel rides in
        # the same event pair and is < 0.5 % of it): algorithmic bytes this rank's launch
        # moves / its mean duration.  Rank 0's block; blocks are nnz-balanced.
        local_bytes = 12.0 * local_nnz + 8.0 * rows_local
        achieved = local_bytes / (kern_ms_avg * 1e-3) / 1e9
        line = {
            "metric": "csr_spmv_f64_gflops", "value": gflops, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n": n, "nnz": nnz, "nnz_per_row": npr,
                       "generator": gen, "index_bytes": 4,
                       "partition": "contiguous row blocks balanced on nnz + %.2f*rows "
                                    "(row cost fitted from per-rank timings), then %d measured "
                                    "equal-time re-cut(s)" % (row_cost, rebalanced),
                       "collective": ("none" if world == 1 else {
                           "overlap": "all-gather of y overlapped with compute: %d row chunks, "
                                      "each slice pushed to the peers by P2P DMA copies on a "
                                      "second stream + 1-element NCCL all_reduce barrier "
                                      "(roofline.kernel_ms then covers the whole step)"
                                      % args.chunks,
                           "fused": "all-gather of y fused into the SpMV kernel (peer stores "
                                    "over NVLink) + 1-element NCCL all_reduce barrier",
                           "push": "SpMV, then one push kernel storing this rank's y slice into "
                                   "every peer buffer (coalesced NVLink stores) + 1-element "
                                   "NCCL all_reduce barrier",
                           "stream": "SpMV publishing its progress + concurrent put kernel "
                                     "copying finished row chunks into the peer buffers "
                                     "(pipelined all-gather over NVLink) + 1-element NCCL "
                                     "all_reduce barrier",
                           "chunked": "SpMV launched in 4 chunks of decreasing size; behind each "
                                      "chunk's event a side stream pushes the rows it completed "
                                      "into the peer buffers (own put kernel over NVLink) + "
                                      "1-element NCCL all_reduce barrier",
                           "mcast": "all-gather of y fused into the SpMV kernel through the "
                                    "NVSwitch multicast address of y (one store per finished row, "
                                    "replicated by the switch) + barrier (%s)" % args.barrier,
                           "mcast-push": "SpMV, then one push kernel storing this rank's y slice "
                                         "to the NVSwitch multicast address of y + barrier (%s)"
                                         % args.barrier,
                           "mcast-stream": "SpMV publishing its progress + concurrent put kernel "
                                           "storing finished row chunks to the NVSwitch multicast "
                                           "address of y + barrier (%s)" % args.barrier,
                           "mcast-chunked": "SpMV in 4 chunks of decreasing size; behind each "
                                            "chunk's event a side stream stores its rows to the "
                                            "NVSwitch multicast address of y + barrier (%s)"
                                            % args.barrier,
                           "nccl": "NCCL all_gather(y), unequal slices"}[args.exchange]),
                       "exchange_trial": exchange_trial,
                       "l2_policy": "inputs (%.1f GB) exceed L2 (126 MB); no flush needed" %
                                    (alg_bytes / 1e9),
                       "gen_seconds": round(t_gen, 1)},
            "achieved_hbm_frac": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / (hbm_peak * world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         "traffic": bench_other.ncu_traffic(args.workload, world),
                         "traffic_source": bench_other.ncu_traffic(args.workload, world, source=True),
                         "kernel": "spmv_warp_kernel (+ spmv_fixup_kernel)",
                         "kernel_ms": kern_ms_avg, "peak_source": peak_src,
                         "algorithmic_bytes": "12*nnz + 8*rows of this rank's block per launch",
                         "variant": os.environ.get("SPRS_B200_SPMV_VARIANT", "default 384,1,8,3 (wt,stages,warps,ctas/SM)")},
            "collective_ms": coll_ms_avg,
            "e2e": {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
                    "ms_per_step": e2e_ms, "h2d_bytes_per_step": 8 * n * world,
                    "d2h_bytes_per_step": 8 * n, "api": "sprs_b200_mul_mat_vec (host x, y; "
                    "A resident as a device mirror)", "matches_device_result": ok},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if fused:
        op.close()
    if world > 1:
        dist.destroy_process_group()


def bench_small_spmv(ctx, G, hbm_peak, dev):
    """BASELINE config 2 (1M x 1M sprs-rand, 32 nnz/row), reported as an extra line item.
    392 MB of inputs > L2, so no flush is needed between iterations."""
    import torch
    n = 1_000_000
    a = G.rand_csr(ctx, n, n, 32, seed=0x5EED0002)
    x = G.normal_vector(ctx, n)
    y = torch.empty(n, device=dev, dtype=torch.float64)
    for _ in range(5):
        G.spmv(ctx, a, x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 50
    e0.record()
    for _ in range(k):
        G.spmv(ctx, a, x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    by = 12.0 * a.nnz + 8.0 * n
    return {"spmv_rand_1m": {"nnz": a.nnz, "ms": ms, "gflops": 2.0 * a.nnz / ms / 1e6,
                             "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / hbm_peak}}


if __name__ == "__main__":
    main()
