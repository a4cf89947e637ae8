#!/bin/bash
# gpurun --gpus N --timeout 1500 -- 'bash tools/r2_multi_final.sh N'
# every exchange mode in one launch (tools/scale_modes.py), then bench.py with the better of
# TMA-staged / per-row peer stores.
set -u
N=${1:-8}
out=gpurun_out/f$N; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  tools/scale_modes.py > $out/scale_modes.txt 2> $out/scale_modes.err; echo "scale_modes exit $?" >> $out/summary.txt
peer=$(python - $out/scale_modes.txt <<'PY'
import json, sys
ms = {}
for l in open(sys.argv[1]):
    try:
        d = json.loads(l)
    except ValueError:
        continue
    if d.get("correct") and "mode" in d:
        ms[d["mode"]] = d["ms_per_step"]
a, b = ms.get("fused+mc"), ms.get("fused+mc+direct")
if a is None and b is not None:      # the TMA form did not survive the multicast address
    print("direct")
else:
    if a is None or b is None:
        a, b = ms.get("fused"), ms.get("fused+direct")
    print("direct" if (b and (a is None or b < 0.99 * a)) else "tma")
PY
)
echo "bench with SPRS_B200_SPMV_PEER_STORES=$peer" >> $out/summary.txt
SPRS_B200_SPMV_PEER_STORES=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus $N --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/summary.txt
cat $out/summary.txt; cat $out/scale_modes.txt; tail -c 2500 $out/bench.json; tail -c 600 $out/bench.err; tail -c 600 $out/scale_modes.err
