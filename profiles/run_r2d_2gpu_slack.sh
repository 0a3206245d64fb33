cd /root/repo; mkdir -p gpurun_out/r2d_2gpu_b
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_2gpu_b/bench_2gpu.json 2> gpurun_out/r2d_2gpu_b/bench_2gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/r2d_2gpu_b/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_2gpu_b/bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'fast', d['value_fast_masks'], 'e2e', d.get('e2e'), d.get('e2e_fast_masks'), d.get('sigma_search_sharded'))
PY
