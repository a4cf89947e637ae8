#!/bin/bash
# Round 2, GPU call 1: information only (no product change yet).
#   gpurun --timeout 1300 -- 'bash tools/r2_call1.sh'
set -u
out=gpurun_out/c1; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
# 1. ceiling sweep on the bench's own matrices
timeout 400 python tools/spmv_lab.py quick > $out/lab.txt 2> $out/lab.err; echo "lab exit $?" >> $out/summary.txt
# 2. full ncu capture of the SHIPPED SpMV kernel on cfg5 and cfg2
timeout 420 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_kernel -s 3 -c 1 \
  -o $out/spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/ncu_spmv.log 2>&1; echo "ncu cfg5 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rmat10m.ncu-rep > $out/ncu_spmv_rmat10m.csv 2>&1
python tools/ncu_traffic.py $out/spmv_rmat10m.ncu-rep > $out/traffic_rmat10m.json 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_kernel -s 3 -c 1 \
  -o $out/spmv_rand1m -f python tools/prof_spmv.py rand 1000000 32 > $out/ncu_spmv2.log 2>&1; echo "ncu cfg2 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rand1m.ncu-rep > $out/ncu_spmv_rand1m.csv 2>&1
python tools/ncu_traffic.py $out/spmv_rand1m.ncu-rep > $out/traffic_rand1m.json 2>&1
# 3. SpMM: default, unroll, panels
timeout 200 python bench.py --workload spmm_rand_1m_k64 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_spmm.json 2> $out/bench_spmm.err; echo "spmm exit $?" >> $out/summary.txt
SPRS_B200_SPMM_UNROLL=4 timeout 200 python bench.py --workload spmm_rand_1m_k64 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_spmm_unroll4.json 2> $out/bench_spmm_unroll4.err
SPRS_B200_SPMM_PANEL=8 timeout 200 python bench.py --workload spmm_rand_1m_k64 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_spmm_panel8.json 2> $out/bench_spmm_panel8.err
# 4. SpGEMM: default and V2, timing then launch breakdown
timeout 300 python bench.py --workload spgemm_rmat_500k --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_spgemm.json 2> $out/bench_spgemm.err; echo "spgemm exit $?" >> $out/summary.txt
SPRS_B200_SPGEMM_V2=1 timeout 300 python bench.py --workload spgemm_rmat_500k --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_spgemm_v2.json 2> $out/bench_spgemm_v2.err; echo "spgemm v2 exit $?" >> $out/summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|num_|nprod|bin_rows|scan_|widen|split_large" -c 600 --csv \
  --log-file $out/launches_spgemm.csv python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline \
  > $out/ncu_spgemm.log 2>&1; echo "ncu spgemm exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches_spgemm.csv > $out/launches_spgemm_agg.txt 2>&1
SPRS_B200_SPGEMM_V2=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|num_|nprod|bin_rows|scan_|widen|split_large" -c 600 --csv \
  --log-file $out/launches_spgemm_v2.csv python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline \
  > $out/ncu_spgemm_v2.log 2>&1; echo "ncu spgemm v2 exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches_spgemm_v2.csv > $out/launches_spgemm_v2_agg.txt 2>&1
cat $out/summary.txt
head -3 $out/lab.txt
