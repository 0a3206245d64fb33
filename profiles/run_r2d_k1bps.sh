#!/bin/bash
# k1_fast_leaf with 4 / 5 / 6 resident blocks per SM asked for (64 / 48 / 40 registers, the latter two with spills)
cd /root/repo
for B in 4 5 6; do echo "== AEW_K1_BPS=$B"; AEW_K1_BPS=$B NC=1000 timeout 300 python profiles/k1_run.py 2>&1 | grep "one sweep"; done
