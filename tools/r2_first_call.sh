#!/bin/bash
# First GPU call of round 2 (one box, ~12-15 GPU-minutes): everything written after round 1's
# GPU budget ran out gets its first hardware run, then the numbers the round-2 plan needs.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# Results land in gpurun_out/r2_first/.
set -u
out=gpurun_out/r2_first
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
# 1. the whole GPU suite, then the opt-in pipelined-push tests in their own process (a trap
#    there loses the CUDA context)
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.txt 2>&1; echo "pytest_gpu exit $?" >> $out/summary.txt
SPRS_B200_TEST_STREAM_PUSH=1 timeout 300 python -m pytest tests/test_gpu_zzz_stream_push.py -m gpu -q > $out/pytest_stream_push.txt 2>&1; echo "stream_push exit $?" >> $out/summary.txt
SPRS_B200_TEST_E2E_CHUNKED=1 timeout 600 python -m pytest tests/test_gpu_zzz_e2e_chunked.py -m gpu -q > $out/pytest_e2e_chunked.txt 2>&1; echo "e2e_chunked tests exit $?" >> $out/summary.txt
SPRS_B200_TEST_SPGEMM_V2=1 timeout 600 python -m pytest tests/test_gpu_zzz_spgemm_v2.py -m gpu -q > $out/pytest_spgemm_v2.txt 2>&1; echo "spgemm_v2 tests exit $?" >> $out/summary.txt
# 2. headline bench (N=1) and the secondary workloads (SpGEMM with the panel kernel: first timing)
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench exit $?" >> $out/summary.txt
SPRS_B200_SPMV_DYNAMIC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $out/bench_n1_dynamic.json 2> $out/bench_n1_dynamic.err; echo "bench dynamic exit $?" >> $out/summary.txt
SPRS_B200_E2E_PIPELINE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $out/bench_n1_e2e_pipeline.json 2> $out/bench_n1_e2e_pipeline.err; echo "bench e2e pipeline exit $?" >> $out/summary.txt
SPRS_B200_E2E_PIPELINE=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $out/bench_n1_e2e_chunked.json 2> $out/bench_n1_e2e_chunked.err; echo "bench e2e chunked exit $?" >> $out/summary.txt
for pw in 8 4; do SPRS_B200_SPMM_PANEL=$pw timeout 300 python bench.py --workload spmm_rand_1m_k64 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_spmm_panel$pw.json 2> $out/bench_spmm_panel$pw.err; done
timeout 600 python bench.py --workload spgemm_rmat_500k --steps 3 --warmup 1 > $out/bench_spgemm.json 2> $out/bench_spgemm.err; echo "spgemm exit $?" >> $out/summary.txt
SPRS_B200_SPGEMM_V2=1 timeout 600 python bench.py --workload spgemm_rmat_500k --steps 3 --warmup 1 --no-cpu-baseline > $out/bench_spgemm_v2.json 2> $out/bench_spgemm_v2.err; echo "spgemm v2 exit $?" >> $out/summary.txt
SPRS_B200_SPMM_UNROLL=4 timeout 300 python bench.py --workload spmm_rand_1m_k64 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_spmm_unroll4.json 2> $out/bench_spmm_unroll4.err
timeout 300 python bench.py --workload spmm_rand_1m_k64 --steps 10 --warmup 3 > $out/bench_spmm.json 2> $out/bench_spmm.err; echo "spmm exit $?" >> $out/summary.txt
# 2b. can a cluster's distributed shared memory out-gather L1TEX? (tools/dsmem_gather_bench.cu)
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dsmem_gather_bench.cu -o tools/dsmem_gather_bench > $out/dsmem_build.txt 2>&1 \
  && timeout 120 tools/dsmem_gather_bench > $out/dsmem_gather.txt 2>&1; echo "dsmem bench exit $?" >> $out/summary.txt
# 3. BiCGSTAB: cost of a step next to its two SpMVs (config 5 matrix)
timeout 600 python tools/time_bicgstab.py > $out/bicgstab.json 2> $out/bicgstab.err; echo "bicgstab exit $?" >> $out/summary.txt
# 4. per-kernel launch list of the SpGEMM (where does the time go now?)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/launches_spgemm.csv python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline \
  > $out/ncu_spgemm.log 2>&1; echo "ncu spgemm exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches_spgemm.csv > $out/launches_spgemm_agg.txt 2>&1
SPRS_B200_SPGEMM_V2=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/launches_spgemm_v2.csv python bench.py --workload spgemm_rmat_500k --steps 1 --warmup 1 --no-cpu-baseline \
  > $out/ncu_spgemm_v2.log 2>&1; echo "ncu spgemm v2 exit $?" >> $out/summary.txt
python tools/agg_launches.py $out/launches_spgemm_v2.csv > $out/launches_spgemm_v2_agg.txt 2>&1
cat $out/summary.txt
tail -3 $out/pytest_gpu.txt $out/pytest_stream_push.txt
tail -c 600 $out/bench_n1.json; echo; python - <<'PY'
import json
for f in ("bench_n1", "bench_n1_dynamic", "bench_n1_e2e_pipeline", "bench_n1_e2e_chunked", "bench_spmm", "bench_spmm_unroll4", "bench_spmm_panel8", "bench_spmm_panel4"):
    try:
        d = json.loads(open("gpurun_out/r2_first/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "value %.1f" % d["value"], "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -c 400 $out/bench_spgemm.json; echo; tail -c 400 $out/bench_spgemm_v2.json; echo; cat $out/bicgstab.json
