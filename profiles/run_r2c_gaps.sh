cd /root/repo; mkdir -p gpurun_out/r2c
PRE_EXACT=1 FAST=1 timeout 300 python profiles/step_gaps.py 2>&1 | grep -v "^   " > gpurun_out/r2c/step_gaps_after_exact.txt
cat gpurun_out/r2c/step_gaps_after_exact.txt
