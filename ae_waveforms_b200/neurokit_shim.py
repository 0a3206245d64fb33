"""Per-call stand-in for the one neurokit2 delineation call on the path,
``nk.ecg_delineate(ecg_sample, r_peaks, sampling_rate=fs)[1]`` (data_extract_vitaldb.py:131),
running the K7 kernel on one window.  Same return shape as the library call: a tuple whose
second element maps the ten ``ECG_*`` names to lists of ints with ``np.nan`` (the singleton)
for missing points; entries that are neither > 0 nor NaN are dropped like the library's
wrapper does, and the library's exceptions (too few peaks ...) surface as ``ValueError`` so the
caller's bare ``except`` (:132-135) behaves the same.  Throughput work goes through
``batched.Preprocessor.run(features=True)``; this exists for drop-in correctness."""
from __future__ import annotations

import numpy as np

from . import _cabi
from .batched import Preprocessor

KEYS = ["ECG_P_Peaks", "ECG_P_Onsets", "ECG_P_Offsets", "ECG_Q_Peaks", "ECG_R_Onsets",
        "ECG_R_Offsets", "ECG_S_Peaks", "ECG_T_Peaks", "ECG_T_Onsets", "ECG_T_Offsets"]
FID_NAN = -2147483648


def ecg_delineate(ecg_cleaned, rpeaks=None, sampling_rate=1000, method="dwt", **kwargs):
    if method != "dwt":
        raise _cabi.AewError("only the default 'dwt' method is implemented")
    if isinstance(rpeaks, dict):
        rpeaks = rpeaks["ECG_R_Peaks"]
    torch, lib = _cabi.require_cuda()
    sig = np.ascontiguousarray(ecg_cleaned, dtype=np.float64)
    L, maxp = len(sig), lib.aew_windows_max_peaks()
    if rpeaks is None or len(rpeaks) == 0 or len(rpeaks) > maxp:
        raise ValueError("ecg_delineate: between 1 and %d R-peaks are needed" % maxp)
    pre = Preprocessor(fs=sampling_rate, sample_length=L / sampling_rate)
    pk = np.full((1, maxp), -1, np.int32)
    pk[0, :len(rpeaks)] = np.asarray(rpeaks, dtype=np.int64)
    d_sig = torch.from_numpy(sig).to(pre.device)
    fid, flags = pre.delineate_windows(d_sig, torch.zeros(1, dtype=torch.int64, device=pre.device),
                                       torch.from_numpy(pk).to(pre.device),
                                       torch.tensor([len(rpeaks)], dtype=torch.int32, device=pre.device), L=L)
    if int(flags[0].item()) == 1:
        raise ValueError("ecg_delineate: heart rate cannot be estimated from these peaks")
    f = fid[0, :, :len(rpeaks)].cpu().numpy()
    waves = {k: [np.nan if v == FID_NAN else int(v) for v in f[i]] for i, k in enumerate(KEYS)}
    return None, {k: [x for x in v if x > 0 or x is np.nan] for k, v in waves.items()}


def ecg_findpeaks(ecg_cleaned, sampling_rate=1000, method="neurokit", **kwargs):
    """``nk.ecg_findpeaks(x, sampling_rate)["ECG_R_Peaks"]`` (remove_samples.py:50-51) on the window kernel
    (K3): gradient-energy QRS runs, most prominent maximum per run, 0.3 s refractory."""
    if method != "neurokit":
        raise _cabi.AewError("only the default 'neurokit' method is implemented")
    torch, _ = _cabi.require_cuda()
    sig = np.ascontiguousarray(ecg_cleaned, dtype=np.float64)
    L = len(sig)
    pre = Preprocessor(fs=sampling_rate, sample_length=L / sampling_rate)
    scr = pre.screen_windows(torch.from_numpy(sig).to(pre.device), np.zeros(1, np.int64), want_median=False,
                             zscore_windows=False, L=L)
    _cabi.check_device_status(int(scr["status"].item()), "ecg_findpeaks")
    n = int(scr["n_peaks"][0].item())
    return {"ECG_R_Peaks": scr["peaks"][0, :n].cpu().numpy().astype(np.int64)}
