"""Per-phase cycle split of k7_delineate (AEW_K7_TIMING=1): thread 0 of every block adds the cycles between phase
boundaries (barrier waits included) to device counters.  NC cases of 1.8 M samples; prints the share of each phase
and the kernel time measured with CUDA events (NO_TICKS=1: the same run without the counters)."""
import os, sys, ctypes, numpy as np
if not os.environ.get("NO_TICKS"):
    os.environ["AEW_K7_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched, _cabi
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "100")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
pre = batched.Preprocessor()
scratch = {"y": torch.empty_like(flat), "tmp": torch.empty_like(flat)}
pre.run(batched.RecordBatch(flat, rec_off), scratch=scratch)
beg_h, end_h, d_beg, d_end = pre._ranges(batched.RecordBatch(flat, rec_off), trim=False)
win_start, _, _ = pre.window_table(beg_h, end_h, np.arange(n_cases))
scr = pre.screen_windows(scratch["y"], win_start, zwin_scratch=scratch["tmp"])
lib = _cabi.load()
names = ["A heart rate (PCHIP + np.mean)", "B Q / S peaks", "C spline prefilter", "D staging + 2 kHz samples",
         "D cascade + rows", "E T / P peaks", "E six bounds per beat", "F wrapper rules"]
out = np.zeros(8, np.uint64)
lib.aew_delineate_timing_dump(ctypes.c_void_p(out.ctypes.data))
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    feats, flags, fid, st = pre.window_features(scratch["tmp"], len(win_start), scr)
    e1.record()
    torch.cuda.synchronize()
    lib.aew_delineate_timing_dump(ctypes.c_void_p(out.ctypes.data))
    kept = int((flags != 3).sum().item())
    print(f"rep {rep}: {e0.elapsed_time(e1):.3f} ms for {kept} windows")
    tot = float(out.sum())
    if tot:
        print(f"  cycles per window (thread 0 of the block): {tot / kept:.0f}")
        for n, v in zip(names, out):
            print(f"    {n:34s} {100 * float(v) / tot:5.1f} %   {float(v) / kept:9.0f} cycles")
