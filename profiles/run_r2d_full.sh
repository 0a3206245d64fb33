#!/bin/bash
# GPU parity suite + smoke + both bench arms on one B200
cd /root/repo
( time timeout 1200 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_bench_reference.json 2> gpurun_out/r2d_bench_ref.err; echo "ref rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, d['value_fast_masks']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
P
O=gpurun_out/r2d; mkdir -p $O
export AEW_BENCH_CASES=200
python bench.py --steps 2 --warmup 3 --no-cpu > $O/b_small.json 2> $O/b_small.err; echo "small plain rc=$?"
ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_timed.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_l.log 2>&1; echo "ncu (timed steps only) rc=$?"
python profiles/launch_list.py $O/launches_timed.csv > $O/launches_timed.txt; head -12 $O/launches_timed.txt
