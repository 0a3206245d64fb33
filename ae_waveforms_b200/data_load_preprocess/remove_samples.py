"""Drop-in for data_load_preprocess/remove_samples.py (reference lines 8-121)."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _cabi
from ..batched import Preprocessor

# "auto": parallel box filters with the exact sequential fallback on hazard windows
# (integer outputs identical to the sequential evaluation); "exact": sequential everywhere
SCREEN_MODE = "auto"


def remove_samples(samples_wave_tracks, indices_samples, track_names, sampling_frequency=500, pvc_threshold=150,
                   sd_threshold_ecg=0.05):
    """remove_samples.py:8-82 -- drops windows containing NaN, detects R-peaks, and removes
    windows with HR < 30 or > 180 bpm, NaN HR, RMSSD > pvc_threshold or sd/n_peaks >
    sd_threshold_ecg.  Returns the reference's DataFrame (same columns, cell types, order)."""
    import pandas as pd
    torch, _ = _cabi.require_cuda()
    names = [t for t in track_names if t == "SNUADC/ECG_II"]
    if not names:
        raise ValueError("no 'SNUADC/ECG_II' track")
    samples = np.asarray(samples_wave_tracks)
    col = names.index("SNUADC/ECG_II")
    n_win, L = samples.shape[0], samples.shape[1]
    # remove_samples.py:31-33 drops a window if ANY of its wave tracks holds a NaN: the tracks next to the ECG (ART, CVP)
    # are looked through on the device as well
    nan_other = np.zeros(n_win, dtype=bool)
    others = [k for k in range(samples.shape[2]) if k != col]
    if others and n_win:
        rest = torch.from_numpy(np.ascontiguousarray(samples[:, :, others], dtype=np.float64)).cuda()
        nan_other = torch.isnan(rest).any(dim=2).any(dim=1).cpu().numpy().astype(bool)
    win = np.ascontiguousarray(samples[:, :, col], dtype=np.float64)
    pre = Preprocessor(fs=sampling_frequency, sample_length=L / sampling_frequency, pvc_threshold=pvc_threshold,
                       sd_threshold=sd_threshold_ecg, screen_mode=SCREEN_MODE)
    flat = torch.from_numpy(win.reshape(-1)).cuda()
    scr = pre.screen_windows(flat, np.arange(n_win, dtype=np.int64) * L, want_median=False, zscore_windows=False,
                             L=L)
    _cabi.check_device_status(int(scr["status"].item()), "remove_samples")
    keep = scr["keep"].cpu().numpy().astype(bool) & ~nan_other
    n_peaks = scr["n_peaks"].cpu().numpy()
    peaks = scr["peaks"].cpu().numpy()
    stats = scr["stats"].cpu().numpy()
    sel = np.flatnonzero(keep)
    if n_win == 0 or (scr["has_nan"].cpu().numpy().astype(bool) | nan_other).all():
        return pd.DataFrame({"ecg_sample": [], "indices_samples": []})
    return pd.DataFrame({
        "ecg_sample": [samples[i, :, col] for i in sel],
        "indices_samples": [indices_samples[i] for i in sel],
        "ECG_R_Peaks": [[int(v) for v in peaks[i, :n_peaks[i]]] for i in sel],
        "ecg_hr": stats[sel, 0], "ecg_hrv_rmssd": stats[sel, 1], "ecg_sd": stats[sel, 2], "ecg_sd/hr": stats[sel, 3]})


def _peak_stats(min_peaks, max_peaks, sampling_frequency, mode):
    """(hr, rmssd) of peak lists on the device (``aew_peak_stats_f64``: the arithmetic the window kernel uses)."""
    torch, lib = _cabi.require_cuda()
    lo = np.ascontiguousarray(np.asarray(min_peaks, dtype=np.int64).reshape(-1))
    hi = np.ascontiguousarray(np.asarray(max_peaks, dtype=np.int64).reshape(-1))
    d_lo = torch.from_numpy(lo).cuda() if lo.size else None
    d_hi = torch.from_numpy(hi).cuda() if hi.size else None
    scratch = torch.empty(2 * (lo.size + hi.size) + 8, dtype=torch.float64, device="cuda")
    out = torch.empty(2, dtype=torch.float64, device="cuda")
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    _cabi.check(lib.aew_peak_stats_f64(p(d_lo), lo.size, p(d_hi), hi.size, mode, float(60 * sampling_frequency),
                                       float(sampling_frequency * 1000), p(scratch), p(out),
                                       _cabi.stream_ptr(torch)), "peak_stats")
    return out.cpu().numpy()


def estimate_hr(min_peaks, max_peaks, sampling_frequency, mod='mean'):
    """remove_samples.py:85-107 -- 60 * fs / median of the peak-to-peak intervals of the chosen list(s):
    'min' -> min_peaks, 'mean' -> both lists pooled, anything else -> max_peaks (the reference prints a notice for
    an unknown ``mod`` and goes on with 'max'; so does this).  NaN below two peaks."""
    mode = {'min': 0, 'mean': 1, 'max': 2}.get(mod)
    if mode is None:
        print(f'estimate_hr: unknown mod {mod!r}, using "max"')
        mode = 2
    return np.float64(_peak_stats(min_peaks, max_peaks, sampling_frequency, mode)[0])


def rmssd(peaks, sampling_frequency=500):
    """remove_samples.py:110-121 -- sqrt(mean(diff(diff(peaks) / (fs * 1000)) ** 2)); the divisor is the
    reference's.  NaN below three peaks."""
    return np.float64(_peak_stats([], peaks, sampling_frequency, 2)[1])
