#!/bin/bash
# 8-GPU probe of the exchange modes incl. the pipelined put (charged 8x: keep it short).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scale_probe.sh 8 "mcast mcast-push chunked stream fused push"'
# mcast / mcast-push = NVSwitch multicast stores (one store per row instead of 7); a third word
# selects their barrier: BARRIER=symm bash tools/r2_scale_probe.sh 8 "mcast mcast-push"
n=${1:-8}; modes=${2:-"mcast mcast-push mcast-stream mcast-chunked chunked stream fused push"}; barrier=${BARRIER:-nccl}
out=gpurun_out/r2_scale; mkdir -p $out
for ex in $modes; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline \
    --exchange $ex --barrier $barrier 2> $out/err_${n}_$ex.txt | tail -1 > $out/scale_${n}_$ex.json
  python - "$n" "$ex" "$out" <<'PY'
import json, sys
n, ex, out = sys.argv[1:4]
try:
    d = json.load(open("%s/scale_%s_%s.json" % (out, n, ex)))
    print(n, ex, "ms/step %.3f" % d["ms_per_step"], "GFLOP/s %.1f" % d["value"],
          "kernel_ms %.3f" % d["roofline"]["kernel_ms"], "coll_ms %.3f" % d.get("collective_ms", -1), flush=True)
except Exception as e:
    print(n, ex, "FAILED", e, open("%s/err_%s_%s.txt" % (out, n, ex)).read()[-800:], flush=True)
PY
done
