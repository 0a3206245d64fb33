# window stage timing only (screen + beat + median kernels), 1,000 cases
cd /root/repo
mkdir -p gpurun_out/r2b
NC=1000 timeout 300 python profiles/k3_variants.py > gpurun_out/r2b/k3_variants.txt 2>&1; tail -6 gpurun_out/r2b/k3_variants.txt
