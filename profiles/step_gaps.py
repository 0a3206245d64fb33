"""Where a step's time goes besides its three kernel stages: CUDA events and host clocks around every method of
Preprocessor.run / finish in the deferred loop of bench.py (no exact-mask pass).  FAST=0/1: K1 / K3 one-sweep forms."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
os.environ["AEW_K3_FAST_STATS"] = os.environ.get("FAST", "1")
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "1000")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
scratch = {"y": torch.empty_like(flat), "tmp": torch.empty_like(flat)}
K = int(os.environ.get("K", "16")); depth = int(os.environ.get("DEPTH", "7"))
pre = batched.Preprocessor(exact_masks=False)
log = []


def wrap(name):
    fn = getattr(pre, name)

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        out = fn(*a, **k)
        e1.record(); t1 = time.perf_counter()
        log.append((name, e0, e1, t0, t1))
        return out
    setattr(pre, name, w)


for n in ("_ranges", "zscore_stats", "filter_records", "window_table", "screen_windows", "finish"):
    wrap(n)
if os.environ.get("PRE_EXACT"):                          # as bench.py: the exact-mask loop runs first
    pe = batched.Preprocessor(exact_masks=True)
    q = []
    for i in range(12):
        q.append(pe.run(batch, scratch=scratch, defer=True, lazy_hazards=True, low_latency=i >= 10))
        while len(q) > depth:
            pe.finish(q.pop(0))
    pe.flush_hazards()
    while q:
        pe.finish(q.pop(0))
    print(f"after the exact loop: allocated {torch.cuda.memory_allocated() / 2**30:.1f} GiB, reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB")
for rnd in range(2):
    log.clear()
    q = []
    start = torch.cuda.Event(enable_timing=True); start.record(); t_start = time.perf_counter()
    for i in range(K):
        q.append(pre.run(batch, scratch=scratch, defer=True, lazy_hazards=True))
        while len(q) > depth:
            pre.finish(q.pop(0))
    while q:
        pre.finish(q.pop(0))
    end = torch.cuda.Event(enable_timing=True); end.record()
    torch.cuda.synchronize()
    print(f"FAST={os.environ['FAST']} round {rnd}: {start.elapsed_time(end) / K:.2f} ms per call; calls started at "
          f"{[round(start.elapsed_time(e0)) for name, e0, e1, t0, t1 in log if name == '_ranges']} ms, "
          f"host {[round(1e3 * (t0 - t_start)) for name, e0, e1, t0, t1 in log if name == '_ranges']}; stats {torch.cuda.memory_stats()['num_alloc_retries']} retries")
    # one line per method call of calls 8..10: device time inside, device time since the previous method's end, host ms
    prev = None
    shown = 0
    for name, e0, e1, t0, t1 in log:
        if name == "_ranges":
            shown += 1
        if 9 <= shown <= 11:
            gap = prev.elapsed_time(e0) if prev is not None else 0.0
            print(f"   {name:16s} at {start.elapsed_time(e0):8.2f} ms  device {e0.elapsed_time(e1):7.2f} ms  since previous end {gap:6.2f} ms   "
                  f"host {1e3 * (t1 - t0):6.2f} ms (host clock {1e3 * (t0 - t_start):8.2f})")
        prev = e1
