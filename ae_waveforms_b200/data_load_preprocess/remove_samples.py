"""Drop-in for data_load_preprocess/remove_samples.py (reference lines 8-121)."""
from __future__ import annotations

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
    nan_other = np.zeros(n_win, dtype=bool)
    for k in range(samples.shape[2]):
        if k != col:
            nan_other |= np.isnan(samples[:, :, k]).any(axis=1)
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


def estimate_hr(min_peaks, max_peaks, sampling_frequency, mod='mean'):
    """remove_samples.py:85-107 -- 60*fs / median(peak intervals); scalar host arithmetic
    (the batched path computes it inside the window kernel)."""
    lo, hi = np.diff(min_peaks), np.diff(max_peaks)
    if mod == 'min':
        med = np.median(lo)
    elif mod == 'mean':
        med = np.median(np.concatenate([lo, hi]))
    else:
        if not mod == 'max':
            print('mod selection got wrong parameter: {}, Use "max" as default settings'.format(mod))
        med = np.median(hi)
    return 60 * sampling_frequency / med


def rmssd(peaks, sampling_frequency=500):
    """remove_samples.py:110-121 (divisor fs*1000 as in the reference)."""
    rr = np.diff(peaks) / (sampling_frequency * 1000)
    return np.sqrt(np.mean(np.diff(rr) ** 2))
