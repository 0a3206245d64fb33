"""Runs the whole-record z-score statistics (K1) on NC synthetic cases; used under ncu."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "200")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
pre = batched.Preprocessor()
beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mean, sd, cnt = pre.zscore_stats(flat, beg_h, end_h, d_beg, d_end)
    e1.record()
    torch.cuda.synchronize()
    print(rep, f"{e0.elapsed_time(e1):.3f} ms for {n_cases} cases, {2 * flat.numel() * 8 / e0.elapsed_time(e1) / 1e6:.0f} GB/s (two reads)")
for rep in range(3):                                     # the one-sweep form (aew_zscore_stats_fast_f64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mean_f, sd_f, cnt_f = pre.zscore_stats(flat, beg_h, end_h, d_beg, d_end, fast=True)
    e1.record()
    torch.cuda.synchronize()
    print(rep, f"one sweep: {e0.elapsed_time(e1):.3f} ms for {n_cases} cases, {flat.numel() * 8 / e0.elapsed_time(e1) / 1e6:.0f} GB/s (one read); "
          f"max rel diff mean {float(((mean_f - mean).abs() / sd).max()):.2e} sd {float(((sd_f - sd).abs() / sd).max()):.2e} counts equal {bool((cnt_f == cnt).all())}")
