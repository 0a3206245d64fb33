#!/bin/bash
# how many of a job's last calls should take the low-latency (one warp per record) form of the exact-mask pass
cd /root/repo
for K in 2 3 4 6; do
echo "== last $K calls low-latency"; AEW_BENCH_TRACE=1 AEW_BENCH_LL_CALLS=$K python bench.py --steps 20 --warmup 5 --no-cpu 2> gpurun_out/tail_$K.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['value_fast_masks']['ms_per_step'])"
grep trace gpurun_out/tail_$K.err | tail -2 | cut -c1-400
done
