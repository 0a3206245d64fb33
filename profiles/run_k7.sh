#!/bin/sh
# gpurun command line of the K7 / K6 round: parity tests, stand-alone timing, bench, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/k7_tests.log
NC=100 timeout 300 python profiles/k7_run.py > gpurun_out/k7_run.log 2>&1
AEW_BENCH_CASES=${BENCH_CASES:-1000} timeout 900 python bench.py > gpurun_out/bench_k7.json 2> gpurun_out/bench_k7.err
NC=10 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k7_|k6_' -c 2 \
  -o gpurun_out/k7_r1 -f python profiles/k7_run.py > gpurun_out/k7_ncu.log 2>&1
tail -3 gpurun_out/k7_tests.log; cat gpurun_out/k7_run.log; tail -c 1500 gpurun_out/bench_k7.json
