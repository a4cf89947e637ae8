#!/bin/bash
# Round-2 profiling call (one GPU, ~6-8 GPU-minutes): the evidence files profiles/ needs.
#   gpurun --timeout 900 -- 'bash tools/r2_profile.sh'            # default kernel
#   gpurun --timeout 900 -- 'SPRS_B200_SPMV_DYNAMIC=1 bash tools/r2_profile.sh dyn'   # a variant, tagged
# Output: gpurun_out/r2_prof/<tag>_*; copy what is to be judged with tools/r2_collect.sh.
set -u
tag=${1:-default}
out=gpurun_out/r2_prof; mkdir -p $out
# 1. launch list of the bench command itself (share of each kernel in the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra \
  > $out/${tag}_launches_bench.log 2>&1; echo "launch list exit $?"
python tools/agg_launches.py $out/${tag}_launches_bench.csv > $out/${tag}_launches_bench_agg.txt 2>&1
# 2. one full capture of the SpMV kernel on the headline matrix (launch 4 of 6: warm caches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_kernel -s 3 -c 1 \
  -o $out/${tag}_spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/${tag}_ncu_spmv.log 2>&1; echo "ncu full exit $?"
python tools/ncu_summary.py $out/${tag}_spmv_rmat10m.ncu-rep > $out/${tag}_ncu_spmv_rmat10m.csv 2>&1
python tools/ncu_traffic.py $out/${tag}_spmv_rmat10m.ncu-rep > $out/${tag}_traffic.json 2>&1
# 3. the same for config 2 (uniform columns), cheap
timeout 300 ncu --set full --clock-control none -k regex:spmv_warp_kernel -s 3 -c 1 \
  -o $out/${tag}_spmv_rand1m -f python tools/prof_spmv.py rand 1000000 32 > /dev/null 2>&1
python tools/ncu_summary.py $out/${tag}_spmv_rand1m.ncu-rep > $out/${tag}_ncu_spmv_rand1m.csv 2>&1
ls -la $out | tail -n 12; cat $out/${tag}_launches_bench_agg.txt | head -8; cat $out/${tag}_traffic.json
