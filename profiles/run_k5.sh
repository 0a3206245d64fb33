#!/bin/bash
# K5 check: GPU tests for the sigma search + its bench section
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sigma or recon" 2>&1 | tail -5
python - <<'PY'
import json, torch, bench
print(json.dumps(bench.bench_sigma_search(torch.device("cuda", 0)), indent=1))
PY
