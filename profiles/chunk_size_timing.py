"""Device time of Preprocessor.run on resident sub-batches of different sizes: does the per-case cost depend on the
chunk size the end-to-end feeder uses?"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "1000")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
scr_y, scr_t = torch.empty_like(flat), torch.empty_like(flat)
pre = batched.Preprocessor()
ev = lambda: torch.cuda.Event(enable_timing=True)
k7 = []
orig = pre.window_features
def wf(*a, **k):
    a0, a1 = ev(), ev(); a0.record(); r = orig(*a, **k); a1.record(); k7.append((a0, a1)); return r
pre.window_features = wf
for feat in (False, True):
    for chunk in (1000, 500, 250, 100, 40, 12):
        if chunk > n_cases: continue
        n = chunk * ns
        batch = batched.RecordBatch(flat[:n], np.arange(chunk + 1, dtype=np.int64) * ns, np.arange(chunk))
        sc = {"y": scr_y[:n], "tmp": scr_t[:n]}
        ts = []
        for rep in range(5):
            k7.clear()
            e0, e1 = ev(), ev(); torch.cuda.synchronize(); e0.record()
            r = pre.run(batch, scratch=sc, features=feat)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        k7ms = k7[0][0].elapsed_time(k7[0][1]) if k7 else 0.0
        t = min(ts[2:])
        print(f"features={feat} {chunk:5d} cases: {t:7.2f} ms = {t * 1000 / chunk:6.1f} ms per 1,000 cases "
              f"(K7+K6 {k7ms:6.2f} = {k7ms * 1000 / chunk:6.1f} per 1,000)  runs {[round(x, 1) for x in ts]}", flush=True)
