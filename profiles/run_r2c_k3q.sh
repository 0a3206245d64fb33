cd /root/repo; mkdir -p gpurun_out/r2c
NC=1000 timeout 300 python profiles/k3_variants.py 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q -k "screen or pipeline or parity_64 or golden" 2>&1 | tail -2
