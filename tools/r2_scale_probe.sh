#!/bin/bash
# 8-GPU probe of the exchange modes (charged 8x: keep it short).  Default: ONE launch that
# shares the matrix, the rendezvous and the partition across all modes (tools/scale_modes.py,
# ~3 min at 8 GPUs instead of ~1.2 min per mode):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scale_probe.sh 8'
# Per-mode bench.py runs (the contract's own command line) for the one or two winners:
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scale_probe.sh 8 "mcast-push fused"'
n=${1:-8}; modes=${2:-}; barrier=${BARRIER:-nccl}
out=gpurun_out/r2_scale; mkdir -p $out
if [ -z "$modes" ]; then
  timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $((29600 + RANDOM % 300)) tools/scale_modes.py --steps 20 \
    2> $out/err_${n}_all.txt | tee $out/scale_modes_${n}.jsonl
  echo "scale_modes exit ${PIPESTATUS[0]}"; tail -n 5 $out/err_${n}_all.txt
  exit 0
fi
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
