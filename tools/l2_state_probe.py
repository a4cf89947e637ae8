"""Does the L2 state a config-5 product leaves behind (80 MB of x lines tagged evict_last) slow
the NEXT, unrelated product?  bench.py's config-2 extra (measured after config 5) read 0.217 ms
where a fresh process reads 0.186.  Times config 2 fresh, after config-5 products, after
cudaCtxResetPersistingL2Cache, and after 512 MB of plain writes; per-repetition times of the
first 12 repetitions show how fast any effect decays."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sprs_b200 as sp  # noqa: E402
from sprs_b200 import generate as G  # noqa: E402

ctx = sp.Context.default(0)
cudart = C.CDLL("libcudart.so.12")
n2 = 1_000_000
a2 = G.rand_csr(ctx, n2, n2, 32, seed=0x5EED0002)
x2 = G.normal_vector(ctx, n2)
y2 = torch.empty(n2, device="cuda", dtype=torch.float64)


def time_cfg2(tag, reps=50):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for i in range(reps):
        G.spmv(ctx, a2, x2, y2)
        evs[i + 1].record()
    torch.cuda.synchronize()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]
    print(json.dumps({"state": tag, "mean_ms": round(sum(per) / reps, 4),
                      "mean_last_20": round(sum(per[-20:]) / 20, 4),
                      "first_12": [round(p, 3) for p in per[:12]]}), flush=True)


for _ in range(5):
    G.spmv(ctx, a2, x2, y2)
time_cfg2("fresh process")
n5 = 10_000_000
a5 = G.make_matrix(ctx, "rmat", n5, 100, 0x5EED0005)
x5 = G.normal_vector(ctx, n5)
y5 = torch.empty(n5, device="cuda", dtype=torch.float64)
for _ in range(8):
    G.spmv(ctx, a5, x5, y5)
torch.cuda.synchronize()
time_cfg2("after 8 config-5 products")
for _ in range(8):
    G.spmv(ctx, a5, x5, y5)
torch.cuda.synchronize()
rc = cudart.cudaCtxResetPersistingL2Cache()
time_cfg2("after 8 config-5 products + cudaCtxResetPersistingL2Cache (rc %d)" % rc)
for _ in range(8):
    G.spmv(ctx, a5, x5, y5)
flush = torch.empty(64 << 20, device="cuda", dtype=torch.float64)
flush.fill_(1.0)
torch.cuda.synchronize()
time_cfg2("after 8 config-5 products + 512 MB of plain writes")
for _ in range(8):
    G.spmv(ctx, a5, x5, y5)
del a5, x5, y5, flush
torch.cuda.empty_cache()
torch.cuda.synchronize()
time_cfg2("after 8 config-5 products + freeing them (empty_cache)")
