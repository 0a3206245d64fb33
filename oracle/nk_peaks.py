"""Restatement of ``neurokit2.ecg_findpeaks(x, sampling_rate, method="neurokit")``.

TEST INFRASTRUCTURE.  PARITY UNPINNED: neurokit2 is a third-party dependency of
the reference (requirements.txt:10, no version pin), it is not vendored under
/root/reference, not installed in this image and cannot be fetched.  The only
call site on the path is data_load_preprocess/remove_samples.py:50-51.  This
file restates the published default detector on top
of the SciPy primitives it delegates to, which ARE present, so the only part
written from the published description is the ~40 lines of glue below.

Defaults: smoothwindow 0.1 s, avgwindow 0.75 s, gradthreshweight 1.5,
minlenweight 0.4, mindelay 0.3 s.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter1d
from scipy.signal import find_peaks


def qrs_mask(signal: np.ndarray, fs: int = 500):
    """Gradient-energy QRS indicator: smoothgrad > 1.5 * avggrad."""
    absgrad = np.abs(np.gradient(signal))
    k_smooth = int(np.rint(0.1 * fs))
    k_avg = int(np.rint(0.75 * fs))
    smoothgrad = uniform_filter1d(absgrad, k_smooth, mode="nearest")
    avggrad = uniform_filter1d(smoothgrad, k_avg, mode="nearest")
    return smoothgrad > 1.5 * avggrad, smoothgrad, avggrad


def qrs_runs(qrs: np.ndarray):
    """Paired (beg, end) indices of QRS runs, ends before the first beg dropped."""
    beg = np.flatnonzero(~qrs[:-1] & qrs[1:])
    end = np.flatnonzero(qrs[:-1] & ~qrs[1:])
    if beg.size == 0:
        return beg, end[:0]
    end = end[end > beg[0]]
    n = min(beg.size, end.size)
    return beg[:n], end[:n]


def ecg_findpeaks_neurokit(signal, sampling_rate: int = 500) -> np.ndarray:
    signal = np.asarray(signal, dtype=np.float64)
    fs = sampling_rate
    qrs, _, _ = qrs_mask(signal, fs)
    beg, end = qrs_runs(qrs)
    if beg.size == 0:
        return np.zeros(0, dtype=np.int64)
    refractory = int(np.rint(0.3 * fs))
    min_len = np.mean(end - beg) * 0.4
    accepted = [0]
    for b, e in zip(beg, end):
        if e - b < min_len:
            continue
        loc, props = find_peaks(signal[b:e], prominence=(None, None))
        if loc.size == 0:
            continue
        cand = b + loc[np.argmax(props["prominences"])]
        if cand - accepted[-1] > refractory:
            accepted.append(cand)
    return np.asarray(accepted[1:], dtype=np.int64)


def ecg_findpeaks(ecg_cleaned, sampling_rate=1000, method="neurokit", **_):
    """Shim with the neurokit2 call shape used at remove_samples.py:51."""
    if method != "neurokit":
        raise NotImplementedError("only the default 'neurokit' detector is restated")
    return {"ECG_R_Peaks": ecg_findpeaks_neurokit(ecg_cleaned, sampling_rate)}
