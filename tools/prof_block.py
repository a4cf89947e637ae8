"""ncu driver: the SpMV of ONE row block (1/8 of the config-5 matrix by cost) -- blocks differ a lot
in rows per non-zero.  usage: prof_block.py <lo> <hi>  (block indices of an 8-way cut)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sprs_b200 as sp
from sprs_b200 import generate as G
from sprs_b200.dist import nnz_balanced_bounds
ctx = sp.Context.default(0)
n = 10_000_000
full = G.make_matrix(ctx, "rmat", n, 100, 0x5EED0005)
x = G.normal_vector(ctx, n)
b = nnz_balanced_bounds(full.indptr, 8, row_cost=30.0)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
a = full.slice_rows(b[lo], b[hi])
lens = (a.indptr[1:] - a.indptr[:-1]).to(torch.int64)
print("rows", a.rows, "nnz", a.nnz, "mean", a.nnz / a.rows, "empty frac", float((lens == 0).float().mean()),
      "frac rows<=6", float((lens <= 6).float().mean()), "frac nnz in rows<=12",
      float(lens[lens <= 12].sum()) / a.nnz, "frac nnz in rows<=48", float(lens[lens <= 48].sum()) / a.nnz)
y = torch.empty(a.rows, device="cuda", dtype=torch.float64)
for _ in range(6):
    G.spmv(ctx, a, x, y)
torch.cuda.synchronize()
