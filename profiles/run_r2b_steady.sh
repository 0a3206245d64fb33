cd /root/repo; mkdir -p gpurun_out/r2b
MEM=1 LL=2 K=24 VARIANTS=1:1,1:0 timeout 600 python profiles/exact_masks_steady.py 2>&1 | grep -v drain | tee gpurun_out/r2b/steady_mem.txt | tail -40
