cd /root/repo; mkdir -p gpurun_out/r2b
for ll in 0 2; do LL=$ll K=20 VARIANTS=1:1 timeout 600 python profiles/exact_masks_steady.py 2>&1 | tail -2; done | tee gpurun_out/r2b/steady_zon.txt
