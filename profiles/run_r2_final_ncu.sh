#!/bin/bash
# round-2 ncu --set full captures of the CURRENT k3_windows / k3_median and k7_delineate / k6_features
cd /root/repo
O=gpurun_out/r2final; mkdir -p $O
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/$name.plain.log 2>&1 || { echo "$name plain run failed"; tail -5 $O/$name.plain.log; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt -f -o $O/$name "$@" > $O/$name.ncu.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  ncu -i $O/$name.ncu-rep --page source --csv --print-source sass > $O/$name.sass.csv 2>/dev/null
  rm -f $O/$name.ncu-rep; tail -2 $O/$name.plain.log
}
NO_TICKS=1 NC=100 cap k3 'k3_windows|k3_median' 4 2 python profiles/k3_phase_timing.py
NC=10 REPS=1 cap k7 'k7_|k6_' 0 2 python profiles/k7_run.py
