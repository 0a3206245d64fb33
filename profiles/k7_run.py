"""Runs filter -> screen -> delineation (K7) -> features (K6) on NC synthetic 1 h cases and times the
K7 + K6 stage with CUDA events; used stand-alone and under ncu (-k regex:k7_|k6_)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "20")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
pre = batched.Preprocessor()
scratch = {"y": torch.empty_like(flat), "tmp": torch.empty_like(flat)}
res = pre.run(batch, scratch=scratch)
beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
win_start, _, _ = pre.window_table(beg_h, end_h, batch.case_ids)
y = scratch["y"]
scr = pre.screen_windows(y, win_start, zwin_scratch=scratch["tmp"])
n_win = len(win_start)
for rep in range(int(os.environ.get("REPS", "3"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    feats, flags, fid, st = pre.window_features(scratch["tmp"], n_win, scr)
    e1.record()
    torch.cuda.synchronize()
    kept = int((flags != 3).sum().item())
    ms = e0.elapsed_time(e1)
    print(rep, f"{ms:.3f} ms for {kept} delineated windows of {n_win} ({kept / ms * 1e3:.0f} windows/s, "
               f"{kept * 10000 / ms / 1e3:.1f} Msamples/s); flags 0/1/2: "
               f"{int((flags == 0).sum())}/{int((flags == 1).sum())}/{int((flags == 2).sum())}; "
               f"NaN share of features {float(torch.isnan(feats[flags == 0]).double().mean()):.4f}")
