#!/bin/bash
# K1 with the guessed record look-up and the shift taken from the block's own loads; k2_pass with runs of 64 tiles;
# ncu --set full of both at 500 cases (enough runs for every block)
cd /root/repo
O=gpurun_out/r2d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "zscore or filter or butterworth or pipeline_e2e" 2>&1 | tail -3
NC=1000 timeout 300 python profiles/k1_run.py 2>&1 | tail -4
NC=1000 timeout 300 python profiles/k2_run.py 2>&1 | tail -4
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/$name.plain.log 2>&1 || { echo "$name plain run failed"; tail -5 $O/$name.plain.log; return; }
  ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c $cnt -f -o $O/$name "$@" > $O/$name.ncu.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  ncu -i $O/$name.ncu-rep --page source --csv --print-source sass > $O/$name.sass.csv 2>/dev/null
  ls -la $O/$name.ncu-rep; rm -f $O/$name.ncu-rep; tail -2 $O/$name.plain.log
}
NC=500 cap k2 'k2_pass' 4 4 python profiles/k2_run.py
NC=500 cap k1f 'k1_fast' 2 2 python profiles/k1_run.py
