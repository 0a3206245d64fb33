"""Runs the direct Gaussian-reconstruction kernel (K5) on 16,384 beats x 64 maps; used under ncu."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, bench
from scipy.stats import qmc
from ae_waveforms_b200.sigma_search import SigmaSearch
dev = torch.device('cuda')
nb, nc = int(os.environ.get("NB", "16384")), int(os.environ.get("NCAND", "64"))
target, fx, fy, bl = bench.synth_fiducials_gpu(nb, dev)
s = SigmaSearch(fx, fy, bl, target, device=dev)
cand = np.floor(qmc.LatinHypercube(d=12, seed=1).random(nc) * 5).clip(0, 4) + 1.0
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sse = s.sse(cand, method="direct"); e1.record(); torch.cuda.synchronize()
    print(rep, f"{e0.elapsed_time(e1):.3f} ms direct, {nb * nc * 664 / e0.elapsed_time(e1) / 1e6:.1f} G exp/s")
