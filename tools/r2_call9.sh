#!/bin/bash
set -u
out=gpurun_out/c10; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_zz_late.py -m gpu -q -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 400 python tools/sweep_spmv.py > $out/sweep.txt 2>&1; echo "sweep exit $?" >> $out/summary.txt
timeout 300 python tools/block_scaling.py > $out/block_scaling.txt 2> $out/block_scaling.err; echo "block_scaling exit $?" >> $out/summary.txt
cat $out/summary.txt; tail -4 $out/pytest.txt; cat $out/sweep.txt; cat $out/block_scaling.txt
