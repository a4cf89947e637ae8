#!/bin/bash
# Round 2, GPU call 13: TMA-staged peer stores on hardware (targets on one device, 2-rank
# communicator sharing the device); L2 state probe for the config-2 extra.
set -u
out=gpurun_out/c13; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spmv_spmm.py tests/test_gpu_comm.py tests/test_gpu_cpp_host.py -m gpu -q -x > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/summary.txt
timeout 300 python tools/l2_state_probe.py > $out/l2_probe.txt 2> $out/l2_probe.err; echo "l2 probe exit $?" >> $out/summary.txt
cat $out/summary.txt; tail -6 $out/pytest.txt; cat $out/l2_probe.txt; tail -3 $out/l2_probe.err
