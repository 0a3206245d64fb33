#!/bin/bash
# final state: window-stage kernels alone, GPU tests, both bench arms, launch list of the timed steps
cd /root/repo
O=gpurun_out/r2c_final2; mkdir -p $O
NC=1000 timeout 300 python profiles/k3_variants.py > $O/k3_variants.txt 2>&1; tail -3 $O/k3_variants.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
echo "tests rc=$?"; tail -3 $O/gputests.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?"
AEW_BENCH_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; grep trace $O/bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_final2/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, 'fast', d['value_fast_masks']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_fast_masks']['value'])
print('roofline', d['roofline']['frac'], d['roofline']['frac_of_bytes_these_kernels_need'], {k:(round(v['ms_per_launch'],2), round(v['frac'],3)) for k,v in d['roofline_by_kernel'].items()})
PY
export AEW_BENCH_CASES=200
python bench.py --steps 2 --warmup 3 --no-cpu > $O/b_small.json 2> $O/b_small.err; echo "small plain rc=$?"
ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_timed.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_l.log 2>&1; echo "ncu (timed steps only) rc=$?"
python profiles/launch_list.py $O/launches_timed.csv > $O/launches_timed.txt; head -30 $O/launches_timed.txt
