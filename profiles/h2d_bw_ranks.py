"""Host -> device bandwidth with 1 / 2 / 4 / 8 GPUs copying AT THE SAME TIME from page-locked host memory, one
process per GPU (as bench.py --gpus N runs): what bounds the end-to-end number at N > 1 (e2e
6.9 -> 13.3 -> 13.9 -> 21.6 Gsamples/s at N = 1 / 2 / 4 / 8).  Also: the box's topology, and the same copies with each
process confined to its own slice of the host cores (first-touch places its buffer near them, if there is more
than one NUMA node)."""
import os, subprocess, sys, time
import torch, torch.multiprocessing as mp

GIB = 1 << 30


def worker(rank, n, barrier, out, pin_cores, gib, reps):
    if pin_cores:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // n)
        os.sched_setaffinity(0, set(cores[rank * per:(rank + 1) * per]))
    torch.cuda.set_device(rank)
    h = torch.empty(gib * GIB, dtype=torch.uint8, pin_memory=True)
    h.zero_()                                                # first touch by this process
    d = torch.empty(gib * GIB, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    barrier.wait()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    out[rank] = (reps * gib * GIB / e0.elapsed_time(e1) / 1e6, time.perf_counter() - t0)
    barrier.wait()


def main():
    for cmd in ("nvidia-smi topo -m", "lscpu | grep -i -E 'numa|socket|model name|^cpu\\(s\\)'", "nproc",
                "ls /sys/devices/system/node | grep node", "free -g | head -2"):
        print("$", cmd); print(subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout, flush=True)
    ngpu = torch.cuda.device_count()
    gib, reps = 2, 12
    for pin in (False, True):
        for n in (1, 2, 4, 8):
            if n > ngpu:
                continue
            ctx = mp.get_context("spawn")
            barrier, out = ctx.Barrier(n), ctx.Manager().dict()
            ps = [ctx.Process(target=worker, args=(r, n, barrier, out, pin, gib, reps)) for r in range(n)]
            [p.start() for p in ps]; [p.join() for p in ps]
            rates = [out[r][0] for r in range(n)]
            print(f"{n} GPUs at once, cores {'split' if pin else 'free'}: per GPU {[round(x, 1) for x in rates]} GB/s, "
                  f"sum {sum(rates):.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
