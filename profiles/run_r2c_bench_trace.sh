# the bench line with the per-call trace on stderr; then ncu of the one-sweep K1 (after its plain run)
cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
AEW_BENCH_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench2.json 2> $O/bench2.err
echo "bench rc=$?"; grep trace $O/bench2.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c/bench2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, 'fast', d['value_fast_masks']['ms_per_step'])
print('roofline', d['roofline']['frac'], d['roofline']['frac_of_bytes_these_kernels_need'], {k:(round(v['ms_per_launch'],2), round(v['frac'],3)) for k,v in d['roofline_by_kernel'].items()})
PY
NC=200 python profiles/k1_run.py > $O/k1_plain.log 2>&1 || { tail -5 $O/k1_plain.log; exit 1; }
NC=200 ncu --set full --clock-control none --import-source on -k "regex:k1_fast" --launch-skip 2 -c 2 -f -o $O/k1f python profiles/k1_run.py > $O/k1f_ncu.log 2>&1
ncu -i $O/k1f.ncu-rep --page raw --csv > $O/k1f.raw.csv 2>/dev/null
rm -f $O/k1f.ncu-rep; tail -3 $O/k1_plain.log
