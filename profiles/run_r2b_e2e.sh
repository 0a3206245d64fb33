cd /root/repo; mkdir -p gpurun_out/r2b
timeout 600 python profiles/e2e_exact_probe.py 2>&1 | tail -5 | tee gpurun_out/r2b/e2e_probe.txt
