cd /root/repo
mkdir -p gpurun_out/r2b
timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
echo "bench rc=$?"; grep "\[mem\]" gpurun_out/r2b/bench.err | head -30; tail -3 gpurun_out/r2b/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, 'fast', d['value_fast_masks']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_fast_masks']['value'], 'narrow', d['e2e_narrow_source']['value'], 'feat', d['with_features']['ms_per_step'])
PY
