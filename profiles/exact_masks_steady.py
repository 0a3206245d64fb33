"""Steady state of the deferred exact-mask loop: K calls at a given depth, per-call time from CUDA events placed
every 10 calls (so the pipeline's fill and drain are visible separately), and the host's time inside run() / finish().
EXACT=0: the same loop without the exact-mask pass.  LAZY=1: run(lazy_hazards=True) (bench.py's loop)."""
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
K = int(os.environ.get("K", "40")); depth = int(os.environ.get("DEPTH", "7")); LL = int(os.environ.get("LL", "0"))
for exact, lazy in [(int(a), int(b)) for a, b in (v.split(":") for v in os.environ.get("VARIANTS", "0:0,1:0,1:1").split(","))]:
    pre = batched.Preprocessor(exact_masks=bool(exact))
    for rnd in range(2):
        marks, q, t_run, t_fin = [], [], 0.0, 0.0
        t0 = time.perf_counter()
        sampler = bench.ClockSampler(0) if os.environ.get("NVML") else None
        if sampler: sampler.__enter__()
        for i in range(K):
            if i % 10 == 0:
                e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
            a = time.perf_counter()
            q.append(pre.run(batch, scratch=scratch, defer=True, lazy_hazards=bool(lazy), low_latency=i >= K - LL))
            b = time.perf_counter()
            while len(q) > depth:
                pre.finish(q.pop(0))
            t_run += b - a; t_fin += time.perf_counter() - b
            if os.environ.get("MEM") and i % 4 == 3:
                print(f"      call {i}: allocated {torch.cuda.memory_allocated() / 2**30:.1f} GiB, reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB", flush=True)
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
        torch.cuda.current_stream().synchronize()          # (probe only: T = the last call's kernels are done)
        tT = time.perf_counter()
        pre.flush_hazards()
        done_at = [round(1e3 * (time.perf_counter() - tT))]
        while q:
            pre.finish(q.pop(0))
            done_at.append(round(1e3 * (time.perf_counter() - tT)))
        print("   drain: flush returned / calls finished at (ms after the last kernels):", done_at, flush=True)
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
        torch.cuda.synchronize()
        if sampler: sampler.__exit__()
        seg = [marks[j].elapsed_time(marks[j + 1]) for j in range(len(marks) - 1)]
        print(f"exact {exact} lazy {lazy} LL {LL} K {K} round {rnd} depth {depth} hog {os.environ.get('AEW_K2E_HOG_KB', 'default')}: per 10 calls "
              f"{[round(s / 10, 1) for s in seg[:-1]]} ms/call, drain {seg[-1]:.0f} ms, total {sum(seg) / K:.1f} ms/call, "
              f"wall {1e3 * (time.perf_counter() - t0) / K:.1f}; host in run() {1e3 * t_run / K:.1f}, in finish() {1e3 * t_fin / K:.1f} ms/call",
              flush=True)
    del pre
    torch.cuda.empty_cache()
