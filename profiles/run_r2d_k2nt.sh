#!/bin/bash
# K2 as runs of tiles (no look-back between blocks), geometry A/B on one B200: 128-thread blocks / 1,664-sample
# tiles / 4 blocks per SM (the build in the tree) against 256 / 3,328 / 2 (alt/libaewave_nt256.so, same sources with
# -DK2_NT=256), and the run length
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter or butterworth or pipeline_e2e or k2" 2>&1 | tail -3
for R in "" 32 64 256; do
echo "== nt128 run=$R"; env ${R:+AEW_K2_RUN=$R} AEW_K2_OCC=1 NC=1000 timeout 300 python profiles/k2_run.py 2>&1 | sort | uniq -c | tail -20
done
cp ae_waveforms_b200/csrc/libaewave.so /tmp/lib128.so
cp alt/libaewave_nt256.so ae_waveforms_b200/csrc/libaewave.so
for R in "" 16 32 128; do
echo "== nt256 run=$R"; env ${R:+AEW_K2_RUN=$R} AEW_K2_OCC=1 NC=1000 timeout 300 python profiles/k2_run.py 2>&1 | sort | uniq -c | tail -20
done
cp /tmp/lib128.so ae_waveforms_b200/csrc/libaewave.so
