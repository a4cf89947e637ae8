#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/r2_multi_bench.sh N [extra bench args]'
set -u
N=${1:-2}; shift
out=gpurun_out/b$N; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus $N --steps 20 --warmup 5 "$@" > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/summary.txt
cat $out/summary.txt; tail -c 2500 $out/bench.json; tail -c 500 $out/bench.err
