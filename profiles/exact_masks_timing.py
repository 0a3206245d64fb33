"""Cost of exact keep / remove masks under the tiled filter on NC cases of 1 h: Preprocessor.run with and without
exact_masks, call by call (the sequential re-filtering waits in the call) and deferred (finish() two calls late, the
re-filtering overlaps the next calls' kernels); then the end-to-end feeder on pinned host buffers."""
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
for exact in (False, True):
    pre = batched.Preprocessor(exact_masks=exact)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); res = pre.run(batch, scratch=scratch); e1.record(); torch.cuda.synchronize()
        print(f"exact_masks={exact} call {rep}: {e0.elapsed_time(e1):.1f} ms, rows {res['np_waves'].shape[0]}, "
              f"20-peak windows re-decided {res['mask_hazard_windows']}, decisions changed {res['mask_flips_resolved']}", flush=True)
    if exact:
        for depth in (3, 5, 7):
            K = 16
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            q = []
            for _ in range(K):
                q.append(pre.run(batch, scratch=scratch, defer=True))
                while len(q) > depth:
                    res = pre.finish(q.pop(0))
            while q:
                res = pre.finish(q.pop(0))
            e1.record(); torch.cuda.synchronize()
            print(f"exact_masks deferred by {depth}: {e0.elapsed_time(e1) / K:.1f} ms per call over {K} calls, rows {res['np_waves'].shape[0]}", flush=True)
host = torch.empty(flat.numel(), dtype=torch.float64, pin_memory=True); host.copy_(flat); torch.cuda.synchronize()
del flat, batch, scratch; torch.cuda.empty_cache()
for exact in (False, True):
    pre = batched.Preprocessor(exact_masks=exact)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = batched.run_host_pipelined(pre, host, rec_off, chunk_cases=max(1, n_cases // 10), to_host=True, pinned_results=True)
        torch.cuda.synchronize()
        print(f"end to end exact_masks={exact} rep {rep}: {1e3 * (time.perf_counter() - t0):.1f} ms, rows {r['np_waves'].shape[0]}, "
              f"re-decided {r['mask_hazard_windows']}, changed {r['mask_flips_resolved']}", flush=True)
