"""Looks for what stalls bench.py's deferred loop now and then: per call the time its first event is reached on the
stream, the host clock at that point, cudaMalloc calls so far (caching allocator), ring wrap-arounds of the upload ring.
EXACT=0/1, DEPTH, K, WARM (calls of an untimed loop before, as bench.py's warm-up)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = 1000; ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
batch = batched.RecordBatch(flat, rec_off)
scratch = {"y": torch.empty_like(flat), "tmp": torch.empty_like(flat)}
K = int(os.environ.get("K", "20")); depth = int(os.environ.get("DEPTH", "9")); warm = int(os.environ.get("WARM", "5"))
pre = batched.Preprocessor(exact_masks=bool(int(os.environ.get("EXACT", "1"))))
wraps = [0]
orig_sync = torch.cuda.Event.synchronize


def loop(steps, label):
    q, marks, rows = [], [], []
    t0 = time.perf_counter()
    for i in range(steps):
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
        st = torch.cuda.memory_stats()
        rows.append((1e3 * (time.perf_counter() - t0), st["num_device_alloc"], st["num_device_free"], st["reserved_bytes.all.current"] / 2**30,
                     getattr(pre, "_up_pos", 0) >> 20))
        q.append(pre.run(batch, scratch=scratch, defer=True, low_latency=i >= steps - 2, lazy_hazards=True))
        while len(q) > depth:
            pre.finish(q.pop(0))
    pre.flush_hazards()
    while q:
        pre.finish(q.pop(0))
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
    torch.cuda.synchronize()
    starts = [round(marks[0].elapsed_time(m)) for m in marks]
    print(f"{label}: exact {int(pre.exact_masks)} depth {depth}: {starts[-1] / steps:.1f} ms per call")
    print("   device time of call starts", starts)
    print("   gaps", [starts[i + 1] - starts[i] for i in range(len(starts) - 1)])
    print("   host clock  ", [round(r[0]) for r in rows])
    print("   cudaMalloc  ", [r[1] for r in rows], "cudaFree", [r[2] for r in rows][-1])
    print("   reserved GiB", [round(r[3]) for r in rows])
    print("   ring pos MiB", [r[4] for r in rows], flush=True)


loop(warm, "warm-up")
loop(K, "timed")
loop(K, "timed again")
