#!/bin/bash
# Copy the round-2 evidence from the scratch directory into profiles/ (tracked), r2-prefixed.
#   bash tools/r2_collect.sh [tag]
tag=${1:-default}
src=gpurun_out/r2_prof
for f in launches_bench.csv launches_bench_agg.txt ncu_spmv_rmat10m.csv ncu_spmv_rand1m.csv traffic.json; do
  [ -s $src/${tag}_$f ] && cp $src/${tag}_$f profiles/r2_${tag}_$f && echo "profiles/r2_${tag}_$f"
done
for f in gpurun_out/r2_first/bench_*.json gpurun_out/r2_first/dsmem_gather.txt gpurun_out/r2_first/launches_spgemm*_agg.txt gpurun_out/r2_first/bicgstab.json; do
  [ -s $f ] && cp $f profiles/r2_first_$(basename $f) && echo "profiles/r2_first_$(basename $f)"
done
for f in gpurun_out/r2_scale/scale_*.json; do
  [ -s $f ] && mkdir -p profiles/r2_scale && cp $f profiles/r2_scale/ && echo "profiles/r2_scale/$(basename $f)"
done
echo "remember: update profiles/ncu_traffic.json from profiles/r2_${tag}_traffic.json (bench.py reads it)"
