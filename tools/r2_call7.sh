#!/bin/bash
set -u
out=gpurun_out/c7; mkdir -p $out
timeout 400 python tools/block_scaling.py > $out/block_scaling.txt 2> $out/block_scaling.err; echo "block_scaling exit $?" >> $out/summary.txt
for c in 1 4 8; do
SPRS_B200_E2E_CHUNKS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $out/bench_e2e_chunks$c.json 2> $out/bench_e2e_chunks$c.err; echo "bench chunks $c exit $?" >> $out/summary.txt
done
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.txt 2>&1; echo "pytest gpu exit $?" >> $out/summary.txt
cat $out/summary.txt; cat $out/block_scaling.txt; tail -5 $out/pytest_gpu.txt
python - <<'PY'
import json
for c in (1,4,8):
    try:
        d=json.loads(open("gpurun_out/c7/bench_e2e_chunks%d.json"%c).read().strip().splitlines()[-1])
        print("chunks",c,"ms/step %.3f"%d["ms_per_step"],"e2e ms %.3f"%d["e2e"]["ms_per_step"], "ceiling", d["roofline"].get("gather_ceiling"))
    except Exception as e: print(c,"ERR",e)
PY
