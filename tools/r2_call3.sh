#!/bin/bash
# Round 2, GPU call 3: lean SpMV hot loop, SpGEMM MLP kernels, new bench.py flow.
set -u
out=gpurun_out/c3; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_spgemm_csc.py tests/test_gpu_cpp_host.py -m gpu -q -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 120 tests/cpp/test_comm_ranks 2 > $out/comm_ranks.txt 2>&1; echo "comm_ranks(2) exit $?" >> $out/summary.txt
timeout 400 python tools/sweep_spmv.py > $out/sweep.txt 2>&1; echo "sweep exit $?" >> $out/summary.txt
timeout 300 python bench.py --workload spgemm_rmat_500k --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_spgemm.json 2> $out/bench_spgemm.err; echo "spgemm exit $?" >> $out/summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|num_|nprod|bin_rows|scan_|widen|split_large" -c 600 --csv \
  --log-file $out/launches_spgemm.csv python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline \
  > $out/ncu_spgemm.log 2>&1; echo "ncu spgemm exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches_spgemm.csv > $out/launches_spgemm_agg.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench exit $?" >> $out/summary.txt
timeout 420 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe_kernel -s 3 -c 1 \
  -o $out/spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/ncu_spmv.log 2>&1; echo "ncu cfg5 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rmat10m.ncu-rep > $out/ncu_spmv_rmat10m.csv 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe_kernel -s 3 -c 1 \
  -o $out/spmv_rand1m -f python tools/prof_spmv.py rand 1000000 32 > $out/ncu_spmv2.log 2>&1; echo "ncu cfg2 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rand1m.ncu-rep > $out/ncu_spmv_rand1m.csv 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref exit $?" >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest.txt; cat $out/sweep.txt; cat $out/launches_spgemm_agg.txt | head -8; tail -c 1500 $out/bench_n1.json; tail -c 600 $out/bench_n1.err; tail -c 400 $out/bench_spgemm.json
