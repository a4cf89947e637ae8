#!/bin/bash
# Round 2, GPU call 4: row-group SpMV kernel; where the SpGEMM panel kernel's time goes.
set -u
out=gpurun_out/c4; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_cpp_host.py tests/test_gpu_comm.py -m gpu -q -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 400 python tools/sweep_spmv.py > $out/sweep.txt 2>&1; echo "sweep exit $?" >> $out/summary.txt
timeout 420 ncu --set full --clock-control none --import-source on -k regex:spmv_rows_kernel -s 3 -c 1 \
  -o $out/spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/ncu_spmv.log 2>&1; echo "ncu cfg5 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rmat10m.ncu-rep > $out/ncu_spmv_rmat10m.csv 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv_rows_kernel -s 3 -c 1 \
  -o $out/spmv_rand1m -f python tools/prof_spmv.py rand 1000000 32 > $out/ncu_spmv2.log 2>&1; echo "ncu cfg2 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rand1m.ncu-rep > $out/ncu_spmv_rand1m.csv 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:num_panel_kernel -c 1 \
  -o $out/num_panel -f python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_panel.log 2>&1; echo "ncu panel exit $?" >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest.txt; cat $out/sweep.txt
