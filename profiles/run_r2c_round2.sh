# staged runs in k3_beats + key-based median: kernels alone, GPU tests, the bench line
cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
NC=1000 timeout 300 python profiles/k3_variants.py > $O/k3_variants2.txt 2>&1; tail -7 $O/k3_variants2.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/gputests2.log 2>&1
echo "tests rc=$?"; tail -3 $O/gputests2.log
AEW_BENCH_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench3.json 2> $O/bench3.err
echo "bench rc=$?"; grep trace $O/bench3.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c/bench3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, 'fast', d['value_fast_masks']['ms_per_step'], 'e2e', d['e2e']['value'])
print('roofline', d['roofline']['frac'], d['roofline']['frac_of_bytes_these_kernels_need'], {k:(round(v['ms_per_launch'],2), round(v['frac'],3)) for k,v in d['roofline_by_kernel'].items()})
print('parity', d['parity']['benchmarked_configuration'])
PY
