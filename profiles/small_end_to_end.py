"""A small end-to-end run (3 short cases with NaN gaps, every screen mode, both sigma-search formulations):
exercises every kernel on little data.  (compute-sanitizer is not available on this GPU pool; races are looked
for with determinism_check.py and the bit-exact parity tests instead.)"""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from ae_waveforms_b200 import batched
from ae_waveforms_b200.sigma_search import SigmaSearch
from oracle import ecg_synth
cases = [ecg_synth.synth_case(1, 65000), ecg_synth.synth_case(2, 41000, nan_gaps=2), ecg_synth.synth_case(3, 30011)]
for mode in ("auto", "exact", "forced_fallback"):
    pre = batched.Preprocessor(screen_mode=mode)
    res = pre.run(batched.RecordBatch.from_host(cases, [1, 2, 3]))
    print(mode, res["np_waves"].shape, res["n_windows"], res["hazard_windows"])
tt, fx, fy, bl = ecg_synth.synth_fiducials(64, seed=5, nan_frac=0.03)
s = SigmaSearch(fx, fy, bl, tt)
c = ecg_synth.sigma_candidates(40, "integer")
print(float(s.sse(c, method="direct").sum()), float(s.sse(c, method="tables").sum()))
torch.cuda.synchronize()
print("done")
