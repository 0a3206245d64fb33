"""Where does the end-to-end step spend its time?  Replays run_host_pipelined with events around every chunk's
host->device copy and kernels."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from ae_waveforms_b200 import batched
dev = torch.device('cuda')
n_cases = int(os.environ.get("NC", "400")); ns = 1800000
flat = bench.synth_cases_gpu(np.arange(n_cases), ns, dev)
host = torch.empty(flat.numel(), dtype=torch.float64, pin_memory=True); host.copy_(flat); torch.cuda.synchronize()
del flat; torch.cuda.empty_cache()
rec_off = np.arange(n_cases + 1, dtype=np.int64) * ns
pre = batched.Preprocessor()
FEAT = os.environ.get("FEAT", "0") == "1"          # FEAT=1: delineation + features switched on
for rep in range(8 if not FEAT else 0):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = batched.run_host_pipelined(pre, host, rec_off, None, chunk_cases=n_cases // 10, to_host=False)
    w = r["np_waves"].cpu(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"step {rep}: {1e3 * (t1 - t0):.1f} ms for {host.numel() * 8 / 1e9:.2f} GB -> {host.numel() * 8 / (t1 - t0) / 1e9:.1f} GB/s "
          f"(pure copy at 55.5 GB/s: {host.numel() * 8 / 55.5e6:.1f} ms)")
# instrumented replay
chunk = n_cases // 10
chunks = [(c0, min(n_cases, c0 + chunk)) for c0 in range(0, n_cases, chunk)]
cap = chunk * ns
bufs = [torch.empty(cap, dtype=torch.float64, device=dev) for _ in range(2)]
scratch = {"y": torch.empty(cap, dtype=torch.float64, device=dev), "tmp": torch.empty(cap, dtype=torch.float64, device=dev)}
cs = torch.cuda.Stream(); main = torch.cuda.current_stream()
ev = lambda: torch.cuda.Event(enable_timing=True)
c_beg = [ev() for _ in chunks]; c_end = [ev() for _ in chunks]; k_beg = [ev() for _ in chunks]; k_end = [ev() for _ in chunks]
freed = [torch.cuda.Event() for _ in chunks]
host_t = []
def start_copy(k):
    c0, c1 = chunks[k]
    with torch.cuda.stream(cs):
        if k >= 2: cs.wait_event(freed[k - 2])
        c_beg[k].record(cs)
        bufs[k % 2][:(c1 - c0) * ns].copy_(host[c0 * ns:c1 * ns], non_blocking=True)
        c_end[k].record(cs)
torch.cuda.synchronize(); t0 = time.perf_counter(); origin = ev(); origin.record(main)
start_copy(0)
for k, (c0, c1) in enumerate(chunks):
    if k + 1 < len(chunks): start_copy(k + 1)
    main.wait_event(c_end[k])
    k_beg[k].record(main)
    sub = batched.RecordBatch(bufs[k % 2][:(c1 - c0) * ns], rec_off[c0:c1 + 1] - rec_off[c0], np.arange(c0, c1))
    th = time.perf_counter()
    res = pre.run(sub, scratch={"y": scratch["y"][:(c1 - c0) * ns], "tmp": scratch["tmp"][:(c1 - c0) * ns]}, features=FEAT)
    host_t.append(1e3 * (time.perf_counter() - th))
    k_end[k].record(main); freed[k].record(main)
torch.cuda.synchronize()
print("chunk  copy[beg..end] ms          kernels[beg..end] ms        host time in run() ms")
for k in range(len(chunks)):
    print(f"{k:3d}   {origin.elapsed_time(c_beg[k]):7.1f} .. {origin.elapsed_time(c_end[k]):7.1f} ({c_beg[k].elapsed_time(c_end[k]):5.1f})   "
          f"{origin.elapsed_time(k_beg[k]):7.1f} .. {origin.elapsed_time(k_end[k]):7.1f} ({k_beg[k].elapsed_time(k_end[k]):5.1f})   {host_t[k]:6.1f}")

if FEAT:
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = batched.run_host_pipelined(pre, host, rec_off, None, chunk_cases=n_cases // 10, to_host=True, pinned_results=True, features=True)
        torch.cuda.synchronize(); print(f"library call with features: {1e3 * (time.perf_counter() - t0):.1f} ms")
    sys.exit(0)
# the library function with different copy piece sizes
for piece in (1 << 40, 4 * 1024 * 1024, 1024 * 1024):
    batched.COPY_PIECE = piece
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = batched.run_host_pipelined(pre, host, rec_off, None, chunk_cases=n_cases // 10, to_host=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"run_host_pipelined, copy pieces of {piece} samples: {1e3 * (t1 - t0):.1f} ms")
