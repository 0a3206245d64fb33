#!/bin/bash
# K3 check: GPU parity tests, then the per-phase cycle split
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
python profiles/k3_phase_timing.py 2>&1 | tail -4
