cd /root/repo; mkdir -p gpurun_out/r2c
for d in 0 1 2; do echo "AEW_K2_DIAG=$d (0 normal, 1 no stores, 2 no loads)"; AEW_K2_DIAG=$d NC=1000 timeout 300 python profiles/k2_run.py 2>&1 | tail -3; done | tee gpurun_out/r2c/k2_diag.txt
