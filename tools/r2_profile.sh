#!/bin/bash
# Round-2 profiling call (one GPU, ~6-8 GPU-minutes): the evidence files profiles/ needs.
#   gpurun --timeout 900 -- 'bash tools/r2_profile.sh'            # default kernel
#   gpurun --timeout 900 -- 'SPRS_B200_SPMV_DYNAMIC=1 bash tools/r2_profile.sh dyn'   # a variant, tagged
# Output: gpurun_out/r2_prof/<tag>_*; copy what is to be judged with tools/r2_collect.sh.
set -u
tag=${1:-default}
out=gpurun_out/r2_prof; mkdir -p $out
# 1. launch list of the bench command itself (share of each kernel in the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra \
  > $out/${tag}_launches_bench.log 2>&1; echo "launch list exit $?"
python tools/agg_launches.py $out/${tag}_launches_bench.csv > $out/${tag}_launches_bench_agg.txt 2>&1
# 2. one full capture of the SpMV kernel on the headline matrix (launch 4 of 6: warm caches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_kernel -s 3 -c 1 \
  -o $out/${tag}_spmv_rmat10m -f python tools/prof_spmv.py rmat 10000000 100 > $out/${tag}_ncu_spmv.log 2>&1; echo "ncu full exit $?"
python tools/ncu_summary.py $out/${tag}_spmv_rmat10m.ncu-rep > $out/${tag}_ncu_spmv_rmat10m.csv 2>&1
ncu -i $out/${tag}_spmv_rmat10m.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY' > $out/${tag}_traffic.json
import csv, json, sys
rows = list(csv.reader(sys.stdin))
if len(rows) > 2:
    h = rows[0]
    r = rows[2]
    g = lambda k: float(r[h.index(k)].replace(",", "")) if k in h else None
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    unit = rows[1][h.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in h else ""
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    print(json.dumps({"kernel": r[h.index("Kernel Name")] if "Kernel Name" in h else "", "unit": unit,
                      "read": rd and rd * mult, "write": wr and wr * mult,
                      "bytes": (rd or 0) * mult + (wr or 0) * mult}))
PY
# 3. the same for config 2 (uniform columns), cheap
timeout 300 ncu --set full --clock-control none -k regex:spmv_warp_kernel -s 3 -c 1 \
  -o $out/${tag}_spmv_rand1m -f python tools/prof_spmv.py rand 1000000 32 > /dev/null 2>&1
python tools/ncu_summary.py $out/${tag}_spmv_rand1m.ncu-rep > $out/${tag}_ncu_spmv_rand1m.csv 2>&1
ls -la $out | tail -n 12; cat $out/${tag}_launches_bench_agg.txt | head -8; cat $out/${tag}_traffic.json
