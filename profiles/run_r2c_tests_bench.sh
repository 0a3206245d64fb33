# round 2, session 5: GPU tests, the bench line, an ncu capture of the fast_stats window kernels
cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
echo "tests rc=$?"; tail -3 $O/gputests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, 'fast', d['value_fast_masks']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_fast_masks']['value'], 'narrow', d['e2e_narrow_source']['value'], 'feat', d['with_features']['ms_per_step'])
print('roofline', d['roofline']['frac'], {k:(round(v['ms_per_launch'],2), round(v['frac'],3)) for k,v in d['roofline_by_kernel'].items()})
print('parity', d['parity']['benchmarked_configuration'])
PY
export NO_TICKS=1 NC=100 K3_FAST=1
python profiles/k3_phase_timing.py > $O/k3_plain.log 2>&1 || { tail -5 $O/k3_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:k3_windows|k3_beats|k3_median" --launch-skip 6 -c 3 -f -o $O/k3 python profiles/k3_phase_timing.py > $O/k3_ncu.log 2>&1
ncu -i $O/k3.ncu-rep --page raw --csv > $O/k3.raw.csv 2>/dev/null
ncu -i $O/k3.ncu-rep --page source --csv --print-source sass > $O/k3.sass.csv 2>/dev/null
rm -f $O/k3.ncu-rep; tail -3 $O/k3_plain.log
cuobjdump -xelf k3_windows ae_waveforms_b200/csrc/libaewave.so > /dev/null 2>&1
