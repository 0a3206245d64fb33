"""Runs K1 + the tiled filter on NC synthetic cases (used under ncu); prints the filter-stage time."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "100")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
pre = batched.Preprocessor()
beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
mean, sd, cnt = pre.zscore_stats(flat, beg_h, end_h, d_beg, d_end)
y = torch.empty_like(flat); tmp = torch.empty_like(flat)
for name, p in (("HP + LP (product)", pre), ("HP + LP, two launches each", None), ("HP alone", batched.Preprocessor(lp=None)),
                ("LP alone", batched.Preprocessor(hp=None))):
    if p is None:
        p = batched.Preprocessor(); p.fuse_passes = False
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.filter_records(flat, beg_h, end_h, d_beg, d_end, mean, sd, cnt, out=y, tmp=tmp)
        e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1):.3f} ms filter stage for {n_cases} cases")
