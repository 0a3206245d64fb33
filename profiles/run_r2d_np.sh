#!/bin/bash
# exact-mask pass: deeper deferral (calls finished 9 / 11 late) and two chain / mover warp pairs per block
cd /root/repo
one() { python bench.py --steps 20 --warmup 5 --no-cpu 2> gpurun_out/np_$1.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['value_fast_masks']['ms_per_step'])"; grep trace gpurun_out/np_$1.err | head -1 | cut -c1-300; }
echo "== depth 7 (default)"; AEW_BENCH_TRACE=1 one d7
echo "== depth 9"; AEW_BENCH_TRACE=1 AEW_BENCH_DEPTH=9 AEW_SIDE_STREAMS=14 one d9
echo "== depth 11"; AEW_BENCH_TRACE=1 AEW_BENCH_DEPTH=11 AEW_SIDE_STREAMS=16 one d11
cp ae_waveforms_b200/csrc/libaewave.so /tmp/lib_np1.so
cp alt/libaewave_np2.so ae_waveforms_b200/csrc/libaewave.so
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exact or parity_64 or remove_samples" 2>&1 | tail -2
echo "== two pairs per block, depth 7"; AEW_BENCH_TRACE=1 one np2d7
echo "== two pairs per block, depth 9"; AEW_BENCH_TRACE=1 AEW_BENCH_DEPTH=9 AEW_SIDE_STREAMS=14 one np2d9
cp /tmp/lib_np1.so ae_waveforms_b200/csrc/libaewave.so
