"""Per-phase cycle split of the window kernel (AEW_K3_TIMING=1): thread 0 of every block adds the cycles between
phase boundaries (barrier waits included) to device counters.  NC cases of 1.8 M samples; prints the share of each
phase, the cycles per window and the kernel time measured with CUDA events."""
import os, sys, ctypes, numpy as np
if not os.environ.get("NO_TICKS"):
    os.environ["AEW_K3_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched, _cabi
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "200")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
pre = batched.Preprocessor()
beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
mean, sd, cnt = pre.zscore_stats(flat, beg_h, end_h, d_beg, d_end)
y, _ = pre.filter_records(flat, beg_h, end_h, d_beg, d_end, mean, sd, cnt)
win_start, _, _ = pre.window_table(beg_h, end_h, batch.case_ids)
lib = _cabi.load()
print('input checksum', int(flat.view(torch.int64).sum().item()), 'mean/sd checksum', int(mean.view(torch.int64).sum().item()), int(sd.view(torch.int64).sum().item()))
print('y checksum', int(y.view(torch.int64).sum().item()), 'hazard windows and kept by mode:')
for mode in (0, 1):
    pre.screen_mode = mode
    scr = pre.screen_windows(y, win_start)
    print('  mode', mode, 'kept', int(scr['keep'].sum()), 'hazards', int(scr['hazards'].item()), 'peaks checksum', int(scr['peaks'].to(torch.int64).sum().item()), 'median checksum', int(torch.nan_to_num(scr['median'][scr['keep'].bool()]).view(torch.int64).sum().item()))
pre.screen_mode = 0
out = np.zeros(16, np.uint64)
names = ["load (+ nan + sum)", "zscore sweeps", "t2", "|gradient| prefix (+ one-sweep statistics)", "smoothgrad prefix", "indicator + hazard vote", "seq fallback", "bits out",
         "min_len", "peak per run", "(unused)", "warp0 chain+stats, peaks out", "median", "t13 leaf table", "t14 load pass", "t15 combine"]
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fast = bool(int(os.environ.get("K3_FAST", "1")))        # as Preprocessor.run calls it behind the tiled filter
    scr = pre.screen_windows(y, win_start, zwin_scratch=pre._filter_tmp, lazy_sd=True, fast_stats=fast)
    e1.record()
    e2 = torch.cuda.Event(enable_timing=True)
    pre.screen_windows(y, win_start, want_median=False, lazy_sd=True, fast_stats=fast)
    e2.record()
    torch.cuda.synchronize()
    if os.environ.get('NO_TICKS'):
        print(rep, f'{e0.elapsed_time(e1):.2f} ms (screen alone {e1.elapsed_time(e2):.2f} ms)')
        continue
    lib.aew_windows_timing_dump(ctypes.c_void_p(out.ctypes.data))
    tot = float(out[:16].sum())
    nw = len(win_start)
    print(rep, f"{e0.elapsed_time(e1):.2f} ms (screen alone {e1.elapsed_time(e2):.2f} ms), {nw} windows, {tot / nw / 2:.0f} cycles/window, kept {int(scr['keep'].sum())}, "
          f"n_peaks>32: {int((scr['n_peaks'] > 32).sum())}")
    print("   ", {n: f"{100 * float(v) / tot:.1f}%" for n, v in zip(names, out[:16])})
