#!/bin/bash
set -u
out=gpurun_out/c5; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_spgemm_csc.py -m gpu -q -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 400 python tools/sweep_spmv.py > $out/sweep.txt 2>&1; echo "sweep exit $?" >> $out/summary.txt
timeout 300 python bench.py --workload spgemm_rmat_500k --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_spgemm.json 2> $out/bench_spgemm.err; echo "spgemm exit $?" >> $out/summary.txt
timeout 300 python bench.py --workload spmm_rand_1m_k64 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_spmm.json 2> $out/bench_spmm.err; echo "spmm exit $?" >> $out/summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|num_|nprod|bin_rows|scan_|widen|split_large" -c 600 --csv \
  --log-file $out/launches_spgemm.csv python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline \
  > $out/ncu_spgemm.log 2>&1; echo "ncu spgemm exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches_spgemm.csv > $out/launches_spgemm_agg.txt 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:spmv_rows_kernel -s 3 -c 1 \
  -o $out/spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/ncu_spmv.log 2>&1; echo "ncu cfg5 exit $?" >> $out/summary.txt
python tools/ncu_summary.py $out/spmv_rmat10m.ncu-rep > $out/ncu_spmv_rmat10m.csv 2>&1
cat $out/summary.txt; tail -3 $out/pytest.txt; cat $out/sweep.txt; head -6 $out/launches_spgemm_agg.txt
python - <<'PY'
import json
for f in ("bench_spgemm","bench_spmm"):
    try:
        d=json.loads(open("gpurun_out/c5/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "ms %.2f"%d["ms_per_step"], "value %.1f"%d["value"], "frac %.4f"%d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
