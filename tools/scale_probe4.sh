#!/bin/bash
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu-baseline --exchange $2 2>gpurun_out/scale_err_$1_$2.txt | tail -1 > gpurun_out/scale_$1_$2.json
  python - "$1" "$2" <<'PY'
import json, sys
n, ex = sys.argv[1], sys.argv[2]
try:
    d = json.load(open("gpurun_out/scale_%s_%s.json" % (n, ex)))
    print(n, ex, "ms/step %.3f" % d["ms_per_step"], "GFLOP/s %.1f" % d["value"], "kernel_ms %.3f" % d["roofline"]["kernel_ms"],
          "coll_ms %.3f" % d["collective_ms"], d["e2e"]["matches_device_result"], d["config"]["partition"][-70:], flush=True)
except Exception as e:
    print(n, ex, "FAILED", e, open("gpurun_out/scale_err_%s_%s.txt" % (n, ex)).read()[-600:], flush=True)
PY
}
run 4 push; run 4 fused; run 4 nccl; run 3 push
