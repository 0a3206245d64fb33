cd /root/repo; mkdir -p gpurun_out/r2c
AEW_BENCH_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c/bench5.json 2> gpurun_out/r2c/bench5.err
echo "bench rc=$?"; grep trace gpurun_out/r2c/bench5.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c/bench5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, 'fast', d['value_fast_masks']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_fast_masks']['value'], d['e2e_narrow_source'])
PY
AEW_BENCH_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c/bench6.json 2> gpurun_out/r2c/bench6.err
echo "bench rc=$?"; grep trace gpurun_out/r2c/bench6.err | cut -c1-300
