#!/bin/bash
# k2_pass with the correction in the same iteration (outputs wait in registers, every warp adds up its own entering
# state, two barriers per tile, copies two tiles ahead) against the committed form (alt/libaewave_v2.so)
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter or butterworth or pipeline_e2e or k2" 2>&1 | tail -3
echo "== v3"; AEW_K2_OCC=1 NC=1000 timeout 300 python profiles/k2_run.py 2>&1 | sort | uniq -c | tail -20
echo "== v3 bench"; python bench.py --steps 10 --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['stage_ms'], d['value_fast_masks']['ms_per_step'], d['ms_per_step'])"
cp ae_waveforms_b200/csrc/libaewave.so /tmp/libv3.so
cp alt/libaewave_v2.so ae_waveforms_b200/csrc/libaewave.so
echo "== v2"; AEW_K2_OCC=1 NC=1000 timeout 300 python profiles/k2_run.py 2>&1 | sort | uniq -c | tail -20
cp /tmp/libv3.so ae_waveforms_b200/csrc/libaewave.so
