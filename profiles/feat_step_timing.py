"""Per-call times of Preprocessor.run(features=True) on a device-resident batch (NC cases x 1 h)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "1000")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off, np.arange(n_cases))
scratch = {"y": torch.empty_like(flat), "tmp": torch.empty_like(flat)}
pre = batched.Preprocessor()
for feat in (False, True, False, True):
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
        r = pre.run(batch, scratch=scratch, features=feat)
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"features={feat} call {rep}: device {e0.elapsed_time(e1):7.1f} ms, host {1e3 * (t1 - t0):7.1f} ms, "
              f"rows {r['np_waves'].shape[0]}, reserved {torch.cuda.memory_reserved() / 1e9:.1f} GB")
