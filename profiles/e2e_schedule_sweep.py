"""End-to-end call (1,000 cases x 1 h, pinned float64 host buffers) under different chunk schedules."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "1000")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
host = torch.empty(flat.numel(), dtype=torch.float64, pin_memory=True); host.copy_(flat); torch.cuda.synchronize()
del flat; torch.cuda.empty_cache()
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
pre = batched.Preprocessor()
for feat in (False, True):
    for chunk, taper in ((100, False), (100, True), (50, True), (150, True), (200, True), (250, True)):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = batched.run_host_pipelined(pre, host, rec_off, None, chunk_cases=chunk, to_host=True, pinned_results=True,
                                           features=feat, taper=taper)
            torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
        print(f"features={feat} chunk_cases={chunk} taper={taper}: {min(ts[1:]):.1f} ms (runs {[round(t, 1) for t in ts]}), "
              f"{len(batched.chunk_schedule(n_cases, chunk, taper))} chunks, rows {r['np_waves'].shape[0]}", flush=True)
