"""Steady state of the deferred exact-mask loop: K calls at a given depth, per-call time from CUDA events placed
every 10 calls (so the pipeline's fill and drain are visible separately)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "1000")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
scratch = {"y": torch.empty_like(flat), "tmp": torch.empty_like(flat)}
pre = batched.Preprocessor(exact_masks=True)
K = int(os.environ.get("K", "40")); depth = int(os.environ.get("DEPTH", "7"))
for rnd in range(2):
    marks = []
    q = []
    t0 = time.perf_counter()
    for i in range(K):
        if i % 10 == 0:
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
        q.append(pre.run(batch, scratch=scratch, defer=True))
        while len(q) > depth:
            pre.finish(q.pop(0))
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    while q:
        pre.finish(q.pop(0))
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    torch.cuda.synchronize()
    seg = [marks[j].elapsed_time(marks[j + 1]) for j in range(len(marks) - 1)]
    print(f"round {rnd} depth {depth} hog {os.environ.get('AEW_K2E_HOG_KB', 'default')}: per 10 calls {[round(s / 10, 1) for s in seg[:-1]]} ms/call, drain {seg[-1]:.0f} ms, "
          f"total {sum(seg) / K:.1f} ms/call, wall {1e3 * (time.perf_counter() - t0) / K:.1f}", flush=True)
