#!/bin/bash
# Round 2, GPU call 11: tile size / row cost sweep; cold-x cost of a 1/8 row block with and
# without the x prefetch; ncu of the shipped SpMV and SpMM kernels.
set -u
out=gpurun_out/c11; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_spmv_spmm.py -m gpu -q -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
BLOCK_SCALING_FLUSH_MB=0,80,512 SPRS_B200_SPMV_PREFETCH_X=0 timeout 300 python tools/block_scaling.py > $out/block_pf0.txt 2> $out/block_pf0.err; echo "block pf0 exit $?" >> $out/summary.txt
BLOCK_SCALING_FLUSH_MB=0,80,512 SPRS_B200_SPMV_PREFETCH_X=1 timeout 300 python tools/block_scaling.py > $out/block_pf1.txt 2> $out/block_pf1.err; echo "block pf1 exit $?" >> $out/summary.txt
timeout 400 python tools/sweep_spmv.py > $out/sweep.txt 2>&1; echo "sweep exit $?" >> $out/summary.txt
SPRS_B200_SPMV_PREFETCH_X=0 timeout 200 python tools/sweep_spmv.py all 1024,5,4,16 > $out/sweep_pf0.txt 2>&1; echo "sweep pf0 exit $?" >> $out/summary.txt
timeout 420 ncu --set full --clock-control none --import-source on -k regex:spmv_rows_kernel -s 3 -c 1 \
  -o $out/spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/ncu_spmv.log 2>&1; echo "ncu cfg5 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rmat10m.ncu-rep > $out/ncu_spmv_rmat10m.csv 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv_rows_kernel -s 3 -c 1 \
  -o $out/spmv_rand1m -f python tools/prof_spmv.py rand 1000000 32 > $out/ncu_spmv2.log 2>&1; echo "ncu cfg2 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rand1m.ncu-rep > $out/ncu_spmv_rand1m.csv 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_rowmaj_vec_kernel -s 2 -c 1 \
  -o $out/spmm -f python bench.py --workload spmm_rand_1m_k64 --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_spmm.log 2>&1; echo "ncu spmm exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmm.ncu-rep > $out/ncu_spmm.csv 2>&1
cat $out/summary.txt; tail -3 $out/pytest.txt; cat $out/block_pf0.txt $out/block_pf1.txt; cat $out/sweep.txt $out/sweep_pf0.txt
