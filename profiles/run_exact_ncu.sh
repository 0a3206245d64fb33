#!/bin/bash
# ncu capture of the sequential filter kernel inside the exact-mask pass (200 cases: ~20 hazard records)
cd /root/repo
O=gpurun_out/r2exact; mkdir -p $O
cat > /tmp/exact_small.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda'); n_cases = 200; ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
batch = batched.RecordBatch(flat, np.arange(n_cases + 1, dtype=np.int64) * ns)
pre = batched.Preprocessor(exact_masks=True)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); res = pre.run(batch); e1.record(); torch.cuda.synchronize()
    print(rep, e0.elapsed_time(e1), res['mask_hazard_windows'])
PY
python /tmp/exact_small.py > $O/plain.log 2>&1 || { tail -5 $O/plain.log; exit 1; }
cat $O/plain.log
ncu --set full --clock-control none --import-source on -k "regex:k2_exact" --launch-skip 2 -c 1 -f -o $O/k2e python /tmp/exact_small.py > $O/ncu.log 2>&1
ncu -i $O/k2e.ncu-rep --page raw --csv > $O/k2e.raw.csv 2>/dev/null
ncu -i $O/k2e.ncu-rep --page source --csv --print-source sass > $O/k2e.sass.csv 2>/dev/null
rm -f $O/k2e.ncu-rep
