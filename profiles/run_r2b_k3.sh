cd /root/repo
mkdir -p gpurun_out/r2b
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b/gputests.log 2>&1
echo "tests rc=$?" 
tail -3 gpurun_out/r2b/gputests.log
NC=1000 timeout 300 python profiles/k3_variants.py > gpurun_out/r2b/k3_variants.txt 2>&1; tail -6 gpurun_out/r2b/k3_variants.txt
NC=200 timeout 300 python profiles/k3_phase_timing.py > gpurun_out/r2b/k3_phase.txt 2>&1; tail -8 gpurun_out/r2b/k3_phase.txt
