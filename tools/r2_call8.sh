#!/bin/bash
set -u
out=gpurun_out/c8; mkdir -p $out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_rows_kernel -s 3 -c 1 \
  -o $out/spmv_block6 -f python tools/prof_block.py 6 7 > $out/block6.log 2>&1; echo "ncu block6 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_block6.ncu-rep > $out/ncu_spmv_block6.csv 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_rows_kernel -s 3 -c 1 \
  -o $out/spmv_block0 -f python tools/prof_block.py 0 1 > $out/block0.log 2>&1; echo "ncu block0 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_block0.ncu-rep > $out/ncu_spmv_block0.csv 2>&1
cat $out/summary.txt; grep rows $out/block6.log $out/block0.log
