"""Dumps K7's fiducials for the parity-test windows to gpurun_out/k7_dump.npz (debugging aid)."""
import os, sys, numpy as np, warnings
warnings.simplefilter("ignore")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ae_waveforms_b200 import batched as B
import test_gpu_parity as T
wins, pks = T._delineation_cases()
pre = B.Preprocessor()
n, L, dev = len(wins), 10000, pre.device
maxp = pre.lib.aew_windows_max_peaks()
pk = np.full((n, maxp), -1, np.int32)
for i, p in enumerate(pks):
    pk[i, :len(p)] = p
sig = torch.from_numpy(np.ascontiguousarray(np.stack(wins).reshape(-1))).to(dev)
off = torch.arange(n, dtype=torch.int64, device=dev) * L
out = {}
for rep in range(3):
    fid, flags = pre.delineate_windows(sig, off, torch.from_numpy(pk).to(dev),
                                       torch.tensor([len(p) for p in pks], dtype=torch.int32, device=dev))
    out[f"fid{rep}"] = fid.cpu().numpy(); out[f"flags{rep}"] = flags.cpu().numpy()
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "k7_dump.npz"), pk=pk, **out)
print("dumped", n)
