#!/bin/bash
# K1 one-sweep: persistent blocks with double-buffered asynchronous copies (k1_fast_stream, AEW_K1_FORM unset) against
# one block per unit (k1_fast_leaf, AEW_K1_FORM=0).  Result: 4.35 against 3.63 ms per 1,000 cases; the persistent form
# was not kept (the script is the record of the measurement: the switch no longer exists)
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "zscore or pipeline_e2e or ragged" 2>&1 | tail -2
echo "== k1_fast_stream"; NC=1000 timeout 300 python profiles/k1_run.py 2>&1 | grep "one sweep"
echo "== k1_fast_leaf"; AEW_K1_FORM=0 NC=1000 timeout 300 python profiles/k1_run.py 2>&1 | grep "one sweep"
