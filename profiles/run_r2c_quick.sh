# round 2, session 5: the one-sweep statistics (K1 fast, K3 fast_stats) -- kernels alone, then the GPU tests
cd /root/repo
mkdir -p gpurun_out/r2c
NC=1000 timeout 300 python profiles/k1_run.py > gpurun_out/r2c/k1_run.txt 2>&1; tail -4 gpurun_out/r2c/k1_run.txt
NC=1000 timeout 300 python profiles/k3_variants.py > gpurun_out/r2c/k3_variants.txt 2>&1; tail -8 gpurun_out/r2c/k3_variants.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c/gputests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2c/gputests.log
