"""2+ rank probe: NCCL small all_reduce latency and 80 MB uneven all_gather time, with and
without the NVML sampling thread that bench.py runs on rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
flag = torch.zeros(1, device=dev)
n = 10_000_000
y = torch.zeros(n, device=dev, dtype=torch.float64)
cuts = [0] + [int(n * (g + 1) / world * (0.8 if g % 2 == 0 else 1.0)) for g in range(world - 1)] + [n]
views = [y[cuts[g]:cuts[g + 1]] for g in range(world)]
def timeit(fn, k=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, (time.perf_counter() - t0) * 1e3 / k
def run(tag):
    a = timeit(lambda: dist.all_reduce(flag))
    b = timeit(lambda: dist.all_gather(views, views[rank]))
    if rank == 0:
        print("%s: all_reduce(1) %.3f ms gpu / %.3f ms wall ; uneven all_gather(80MB) %.3f ms gpu / %.3f wall" % (tag, a[0], a[1], b[0], b[1]), flush=True)
run("no sampler")
if rank == 0:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    s = bench.ClockSampler(local); s.start()
run("with NVML sampler on rank 0")
if rank == 0:
    print(s.stop(0, time.time() + 1))
dist.destroy_process_group()
