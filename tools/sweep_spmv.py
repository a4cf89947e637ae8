"""Times the SpMV tile cuts (SPRS_B200_SPMV_VARIANT=w,row_cost) on the
bench workloads; one subprocess per variant because the variant is read once per process."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
import sprs_b200 as sp
from sprs_b200 import generate as G
ctx = sp.Context.default(0)
out = {}
for name, gen, n, npr in %s:
    a = G.make_matrix(ctx, gen, n, npr, 0x5EED0005 if gen == "rmat" else 0x5EED0002)
    x = G.normal_vector(ctx, n); y = torch.empty(n, device="cuda", dtype=torch.float64)
    for _ in range(5): G.spmv(ctx, a, x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 20
    e0.record()
    for _ in range(k): G.spmv(ctx, a, x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    out[name] = {"ms": ms, "gbs": (12.0 * a.nnz + 8.0 * n) / ms / 1e6, "gnnz_s": a.nnz / ms / 1e6}
    del a, x, y; torch.cuda.empty_cache()
print("RESULT " + json.dumps(out))
'''

def main():
    workloads = [("rand_1m_32", "rand", 1_000_000, 32), ("rmat_10m_100", "rmat", 10_000_000, 100)]
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        workloads = [("rand_1m_32", "rand", 1_000_000, 32), ("rmat_1m_100", "rmat", 1_000_000, 100)]
    variants = ["1024,16", "2048,16", "1024,8", "1024,32", "2048,32", "512,16"]  # cost units per tile, row cost (csrc/spmv.cu)
    if len(sys.argv) > 2:
        variants = sys.argv[2:]
    for v in variants:
        env = dict(os.environ, SPRS_B200_SPMV_VARIANT=v)
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, repr(workloads))], env=env,
                           capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(v, "FAILED", r.stderr[-400:])
            continue
        res = json.loads(line[0][7:])
        print(v.ljust(12), "  ".join("%s: %.3f ms %.0f GB/s %.0f Gnnz/s" % (k, d["ms"], d["gbs"], d["gnnz_s"])
                                     for k, d in res.items()), flush=True)

if __name__ == "__main__":
    main()
