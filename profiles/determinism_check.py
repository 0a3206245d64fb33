"""Is the pipeline run-to-run deterministic?  Generates the synthetic input twice, runs K1 and the tiled filter
five times on the same input and the window kernels twice, and compares bit-level checksums."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "100")); ns = 1800000
cs = lambda t: int(t.contiguous().view(torch.int64).sum().item())
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
flat2 = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
print("input generated twice: identical =", bool(torch.equal(flat.view(torch.int64), flat2.view(torch.int64))))
del flat2
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
pre = batched.Preprocessor()
beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
ys = []
for rep in range(5):
    mean, sd, cnt = pre.zscore_stats(flat, beg_h, end_h, d_beg, d_end)
    y, _ = pre.filter_records(flat, beg_h, end_h, d_beg, d_end, mean, sd, cnt)
    ys.append((cs(mean), cs(sd), cs(y)))
    if rep == 0: y0 = y.clone()
    else:
        diff = (y.view(torch.int64) != y0.view(torch.int64))
        nd = int(diff.sum().item())
        if nd:
            idx = torch.nonzero(diff).flatten()
            print(f"  rep {rep}: {nd} samples differ, first at {int(idx[0])} (tile {int(idx[0]) % ns // 3328} of its case), "
                  f"max |diff| {float((y - y0).abs().max()):.3e}")
print("K1 mean/sd and filter output checksums per repetition:", ys)
win_start, _, _ = pre.window_table(beg_h, end_h, batch.case_ids)
a = pre.screen_windows(y0, win_start, zwin_scratch=pre._filter_tmp)
b = pre.screen_windows(y0, win_start, zwin_scratch=pre._filter_tmp)
print("window kernels twice on the same filtered signal: keep equal", bool(torch.equal(a['keep'], b['keep'])),
      "peaks equal", bool(torch.equal(a['peaks'], b['peaks'])),
      "waves equal", bool(torch.equal(a['waves'][a['keep'].bool()].view(torch.int32), b['waves'][b['keep'].bool()].view(torch.int32))))
