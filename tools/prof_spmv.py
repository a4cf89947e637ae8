"""Small driver for ncu: builds one workload and runs a few SpMVs (used under ncu -k regex:spmv)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sprs_b200 as sp
from sprs_b200 import generate as G
ctx = sp.Context.default(0)
gen, n, npr = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("rand", 1_000_000, 32)
a = G.make_matrix(ctx, gen, n, npr, 0x5EED0005 if gen == "rmat" else 0x5EED0002)
x = G.normal_vector(ctx, n); y = torch.empty(n, device="cuda", dtype=torch.float64)
for _ in range(6): G.spmv(ctx, a, x, y)
torch.cuda.synchronize()
print("done", a.nnz)
