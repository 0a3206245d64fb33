#!/bin/bash
# ncu --set full of the run-based k2_pass (HP fwd / bwd, LP fwd / bwd of 64 one-hour cases) + launch list of the timed steps
cd /root/repo
O=gpurun_out/r2d; mkdir -p $O
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/$name.plain.log 2>&1 || { echo "$name plain run failed"; tail -5 $O/$name.plain.log; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt -f -o $O/$name "$@" > $O/$name.ncu.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  ncu -i $O/$name.ncu-rep --page source --csv --print-source sass > $O/$name.sass.csv 2>/dev/null
  ls -la $O/$name.ncu-rep; rm -f $O/$name.ncu-rep; tail -4 $O/$name.plain.log
}
NC=64 cap k2 'k2_pass' 4 4 python profiles/k2_run.py
export AEW_BENCH_CASES=200
python bench.py --steps 2 --warmup 3 --no-cpu > $O/b_small.json 2> $O/b_small.err; echo "small plain rc=$?"
ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_timed.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_l.log 2>&1; echo "ncu (timed steps only) rc=$?"
python profiles/launch_list.py $O/launches_timed.csv > $O/launches_timed.txt; head -30 $O/launches_timed.txt
