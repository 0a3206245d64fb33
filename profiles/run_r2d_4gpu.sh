cd /root/repo; mkdir -p gpurun_out/r2d_4gpu
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2d_4gpu/bench_4gpu.json 2> gpurun_out/r2d_4gpu/bench_4gpu.err
echo "bench rc=$?"; tail -2 gpurun_out/r2d_4gpu/bench_4gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_4gpu/bench_4gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'fast', d['value_fast_masks']['value'], d['value_fast_masks']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_fast_masks']['value'])
PY
