#!/bin/bash
# Round 2, GPU call 2: first hardware run of the pipelined SpMV kernel and of the comm layer.
set -u
out=gpurun_out/c2; mkdir -p $out
# 1. parity first: SpMV/SpMM suites, ABI, C++ KATs, then the 2-rank comm test (ranks share GPU 0)
timeout 600 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_cpp_host.py tests/test_gpu_zz_late.py -m gpu -q -x -k "not l2_blocked and not unrolled_variant" > $out/pytest_spmv.txt 2>&1; echo "pytest spmv exit $?" >> $out/summary.txt
timeout 120 tests/cpp/test_comm_ranks 2 > $out/comm_ranks.txt 2>&1; echo "comm_ranks(2) exit $?" >> $out/summary.txt
timeout 120 tests/cpp/test_comm_ranks 3 > $out/comm_ranks3.txt 2>&1; echo "comm_ranks(3) exit $?" >> $out/summary.txt
# 2. timing of the variants on both matrices
timeout 500 python tools/sweep_spmv.py > $out/sweep.txt 2>&1; echo "sweep exit $?" >> $out/summary.txt
# 3. ncu of the new kernel
timeout 420 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe_kernel -s 3 -c 1 \
  -o $out/spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/ncu_spmv.log 2>&1; echo "ncu cfg5 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rmat10m.ncu-rep > $out/ncu_spmv_rmat10m.csv 2>&1
python tools/ncu_traffic.py $out/spmv_rmat10m.ncu-rep > $out/traffic_rmat10m.json 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe_kernel -s 3 -c 1 \
  -o $out/spmv_rand1m -f python tools/prof_spmv.py rand 1000000 32 > $out/ncu_spmv2.log 2>&1; echo "ncu cfg2 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rand1m.ncu-rep > $out/ncu_spmv_rand1m.csv 2>&1
python tools/ncu_traffic.py $out/spmv_rand1m.ncu-rep > $out/traffic_rand1m.json 2>&1
cat $out/summary.txt; tail -3 $out/pytest_spmv.txt; cat $out/comm_ranks.txt | tail -3; cat $out/sweep.txt
