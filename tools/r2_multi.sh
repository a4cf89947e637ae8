#!/bin/bash
# Multi-GPU call: gpurun --gpus N --timeout 1200 -- 'bash tools/r2_multi.sh N'
set -u
N=${1:-2}
out=gpurun_out/m$N; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/gpus.txt 2>&1
nvidia-smi topo -m > $out/topo.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_comm.py -m gpu -q -x > $out/pytest_comm.txt 2>&1; echo "pytest comm exit $?" >> $out/summary.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  tools/scale_modes.py --steps 20 --warmup 5 > $out/scale_modes.txt 2> $out/scale_modes.err; echo "scale_modes exit $?" >> $out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus $N --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_comm.txt 2>/dev/null; cat $out/scale_modes.txt | cut -c1-400; tail -c 1800 $out/bench.json; tail -c 800 $out/bench.err; tail -c 600 $out/scale_modes.err
