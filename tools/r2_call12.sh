#!/bin/bash
# Round 2, GPU call 12: the whole GPU suite, smoke(), the default bench line, and the ncu launch
# list of the same bench command (shares of the step).
set -u
out=gpurun_out/c12; mkdir -p $out
timeout 1500 python -m pytest tests/ -q -m gpu -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/summary.txt
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench ref exit $?" >> $out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > $out/bench_under_ncu.log 2>&1; echo "ncu launches exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches.csv > $out/launches_agg.txt 2>&1
cat $out/summary.txt; tail -5 $out/pytest.txt; tail -2 $out/smoke.txt; tail -c 3000 $out/bench.json; tail -c 600 $out/bench_ref.json; cat $out/launches_agg.txt
