#!/bin/bash
# ncu --set full capture of the window stage's three kernels (after the same command has run plain)
cd /root/repo
O=gpurun_out/r2b; mkdir -p $O
export NO_TICKS=1 NC=100
python profiles/k3_phase_timing.py > $O/k3_plain.log 2>&1 || { tail -5 $O/k3_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:k3_windows|k3_beats|k3_median" --launch-skip 6 -c 3 -f -o $O/k3 python profiles/k3_phase_timing.py > $O/k3_ncu.log 2>&1
ncu -i $O/k3.ncu-rep --page raw --csv > $O/k3.raw.csv 2>/dev/null
ncu -i $O/k3.ncu-rep --page source --csv --print-source sass > $O/k3.sass.csv 2>/dev/null
rm -f $O/k3.ncu-rep; tail -3 $O/k3_plain.log
