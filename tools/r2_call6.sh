#!/bin/bash
set -u
out=gpurun_out/c6; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_spgemm_csc.py -m gpu -q -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 400 python tools/sweep_spmv.py > $out/sweep.txt 2>&1; echo "sweep exit $?" >> $out/summary.txt
for h in 496 1024 2048 4096; do
SPRS_B200_SPGEMM_HASH_MAX=$h timeout 300 python bench.py --workload spgemm_rmat_500k --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_spgemm_$h.json 2> $out/bench_spgemm_$h.err; echo "spgemm $h exit $?" >> $out/summary.txt
done
SPRS_B200_SPGEMM_HASH_MAX=496 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|num_|nprod|bin_rows|scan_|widen|split_large" -c 600 --csv \
  --log-file $out/launches_spgemm.csv python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline \
  > $out/ncu_spgemm.log 2>&1; echo "ncu spgemm exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches_spgemm.csv > $out/launches_spgemm_agg.txt 2>&1
cat $out/summary.txt; tail -3 $out/pytest.txt; cat $out/sweep.txt; head -6 $out/launches_spgemm_agg.txt
python - <<'PY'
import json
for h in (496,1024,2048,4096):
    try:
        d=json.loads(open("gpurun_out/c6/bench_spgemm_%d.json"%h).read().strip().splitlines()[-1])
        print(h, "ms %.2f"%d["ms_per_step"], "value %.1f"%d["value"])
    except Exception as e: print(h, "ERR", e)
PY
