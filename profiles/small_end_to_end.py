"""A small end-to-end run (short cases with NaN gaps and a 60 bpm case, every screen mode, exact masks, features,
both sigma-search formulations): exercises every kernel on little data.  Used under `compute-sanitizer --tool
racecheck` / `memcheck` (profiles/run_sanitizer.sh); races are also looked for with determinism_check.py and the
bit-exact parity tests."""
import os, sys, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from ae_waveforms_b200 import batched
from ae_waveforms_b200.sigma_search import SigmaSearch
from oracle import ecg_synth
cases = [ecg_synth.synth_case(1, 45000), ecg_synth.synth_case(2, 41000, nan_gaps=2), ecg_synth.synth_case(3, 30011),
         ecg_synth.synth_case(7, 40001, hr_bpm=60.0)]
for mode in ("auto", "exact", "forced_fallback"):
    pre = batched.Preprocessor(screen_mode=mode)
    res = pre.run(batched.RecordBatch.from_host(cases, [1, 2, 3, 7]), features=(mode == "auto"))
    print(mode, res["np_waves"].shape, res["n_windows"], res["hazard_windows"], res["mask_hazard_windows"])
res = batched.Preprocessor(filter_mode="exact").run(batched.RecordBatch.from_host(cases, [1, 2, 3, 7]))
print("exact filter", res["np_waves"].shape)
tt, fx, fy, bl = ecg_synth.synth_fiducials(64, seed=5, nan_frac=0.03)
s = SigmaSearch(fx, fy, bl, tt)
c = ecg_synth.sigma_candidates(40, "integer")
print(float(s.sse(c, method="direct").sum()), float(s.sse(c, method="tables").sum()))
torch.cuda.synchronize()
print("done")
