"""Drop-in for data_load_preprocess/median_sample.py (reference lines 6-82)."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _cabi
from ..batched import BEAT_LEN, BEAT_POST, BEAT_PRE, _ptr


def compute_median_sample_ecg(sample_ecg, r_peaks, sampling_frequency, resample_ecg=False, return_aligned=False):
    """median_sample.py:6-54 -- 320-sample sections around each R-peak (short slices
    left-aligned in a NaN row), column-wise nanmedian on the GPU."""
    if resample_ecg:
        raise NotImplementedError("resample_ecg=True is dead code on the reference path")
    section = int(0.8 * (60 / 75) * sampling_frequency)
    pre_len = int(1 / 3 * section)
    if section != BEAT_LEN or pre_len != BEAT_PRE:
        raise NotImplementedError("the kernel is built for 500 Hz (320-sample beats)")
    peaks = np.asarray(list(r_peaks), dtype=np.int64)
    if peaks.size == 0:
        med = np.float64(np.nan)                       # reference: nanmedian of an empty array
        return (med, np.array([])) if return_aligned else med
    if peaks.size > 128:
        raise _cabi.AewError("at most 128 beats per window")
    torch, lib = _cabi.require_cuda()
    x = torch.from_numpy(np.ascontiguousarray(sample_ecg, dtype=np.float64)).cuda()
    d_peaks = torch.from_numpy(peaks.astype(np.int32)).cuda()
    d_off = torch.tensor([0, peaks.size], dtype=torch.int64, device="cuda")
    med = torch.empty(BEAT_LEN, dtype=torch.float64, device="cuda")
    aligned = torch.empty((peaks.size, BEAT_LEN), dtype=torch.float64, device="cuda") if return_aligned else None
    _cabi.check(lib.aew_median_beats_f64(_ptr(x), 1, x.numel(), _ptr(d_peaks), _ptr(d_off), int(peaks.size),
                                         BEAT_PRE, BEAT_POST, _ptr(med), _ptr(aligned), _cabi.stream_ptr(torch)),
                "median_beats")
    if return_aligned:
        return med.cpu().numpy(), aligned.cpu().numpy()
    return med.cpu().numpy()


def compute_median_features_ecg(ecg_features, ecg_features_morphology):
    """median_sample.py:57-82 -- nanmedian over beats of every fiducial x / y and of every
    morphology feature (a handful of scalars per window: host arithmetic)."""
    out = {}
    for feat in ecg_features.keys():
        out[feat] = {"x": np.nanmedian(ecg_features[feat]["x"]), "y": np.nanmedian(ecg_features[feat]["y"])}
    for feat in ecg_features_morphology.keys():
        out[feat] = np.nanmedian(ecg_features_morphology[feat])
    return out


def plot_median_ecg(*args, **kwargs):
    """Imported by data_extract_vitaldb.py:14 (only called from commented-out code)."""
    raise NotImplementedError("plotting is outside the preprocessing hot path")
