"""Window-stage variants on NC cases of 1 h: the z-scored windows parked for the median kernel or re-derived by it,
np.std always or only where it decides (lazy_sd).  CUDA events around screen + median."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "1000")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
pre = batched.Preprocessor()
beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
mean, sd, cnt = pre.zscore_stats(flat, beg_h, end_h, d_beg, d_end)
y, _ = pre.filter_records(flat, beg_h, end_h, d_beg, d_end, mean, sd, cnt)
win_start, _, _ = pre.window_table(beg_h, end_h, batch.case_ids)
del flat
ref = None
for name, kw in (("parked, full sd", dict(zwin_scratch=pre._filter_tmp)), ("parked, lazy sd", dict(zwin_scratch=pre._filter_tmp, lazy_sd=True)),
                 ("derived, full sd", dict()), ("derived, lazy sd", dict(lazy_sd=True)), ("screen only, lazy sd", dict(lazy_sd=True, want_median=False)),
                 ("fast_stats (no NumPy-order sweeps, nothing parked)", dict(lazy_sd=True, fast_stats=True)),
                 ("fast_stats, screen only", dict(lazy_sd=True, fast_stats=True, want_median=False))):
    ts = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); scr = pre.screen_windows(y, win_start, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    chk = (int(scr['keep'].sum()), int(scr['peaks'].to(torch.int64).sum()),
           int(torch.nan_to_num(scr['median'][scr['keep'].bool()]).view(torch.int64).sum()) if scr['median'] is not None else None)
    print(f"{name:52s} {min(ts[1:]):7.2f} ms per {len(win_start)} windows   checks {chk}", flush=True)
