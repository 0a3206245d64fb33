cd /root/repo
mkdir -p gpurun_out/r2b
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b/gputests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2b/gputests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, d['value_fast_masks']['ms_per_step'], d['e2e']['value'], d['e2e_fast_masks']['value'])
PY
