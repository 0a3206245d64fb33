cd /root/repo; mkdir -p gpurun_out/r2c
for d in 7 10 13; do DEPTH=$d LL=2 K=40 VARIANTS=1:1 timeout 600 python profiles/exact_masks_steady.py 2>&1 | grep -v "^   " | tail -1; done | tee gpurun_out/r2c/steady_depth.txt
