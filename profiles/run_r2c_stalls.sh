cd /root/repo; mkdir -p gpurun_out/r2c
EXACT=1 DEPTH=7 timeout 300 python profiles/loop_stalls.py > gpurun_out/r2c/loop_stalls2.txt 2>&1
grep -A2 "timed\|warm" gpurun_out/r2c/loop_stalls2.txt | cut -c1-260; grep "reserved" gpurun_out/r2c/loop_stalls2.txt | cut -c1-200
AEW_BENCH_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c/bench4.json 2> gpurun_out/r2c/bench4.err
echo "bench rc=$?"; grep trace gpurun_out/r2c/bench4.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c/bench4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','stage_ms')}, 'fast', d['value_fast_masks']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_fast_masks']['value'])
PY
timeout 600 python -m pytest tests -m gpu -x -q -k "parity_64 or pipeline or host_pipelined or exact" 2>&1 | tail -3
