#!/bin/bash
# K2 check: GPU parity tests that involve the filter, then its phase timing and isolated pass time
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "butterworth or filter or pipeline or host_pipelined" 2>&1 | tail -3
NC=${NC:-200} timeout 600 python profiles/k2_phase_timing.py 2>&1 | tail -3
