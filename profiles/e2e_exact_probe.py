import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda'); n_cases = 1000; ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
host = torch.empty(flat.numel(), dtype=torch.float64, pin_memory=True); host.copy_(flat); torch.cuda.synchronize()
del flat; torch.cuda.empty_cache()
pre = batched.Preprocessor()
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = batched.run_host_pipelined(pre, host, rec_off, chunk_cases=100, to_host=True, pinned_results=True)
    torch.cuda.synchronize()
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.1f} ms; hazard windows per chunk {r['mask_hazard_windows_per_chunk']}", flush=True)
