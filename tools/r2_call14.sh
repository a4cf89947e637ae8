#!/bin/bash
# Round 2, GPU call 14: the whole GPU suite and smoke() on the final build.
set -u
out=gpurun_out/c14; mkdir -p $out
timeout 1500 python -m pytest tests/ -q -m gpu -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/summary.txt
cat $out/summary.txt; tail -5 $out/pytest.txt; tail -2 $out/smoke.txt
