#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "zscore or pipeline or filter_many_nan" 2>&1 | tail -3
python profiles/k1_run.py
