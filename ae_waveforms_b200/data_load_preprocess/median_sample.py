"""Drop-in for data_load_preprocess/median_sample.py (reference lines 6-82)."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _cabi
from ..batched import _ptr, beat_geometry


def compute_median_sample_ecg(sample_ecg, r_peaks, sampling_frequency, resample_ecg=False, return_aligned=False):
    """median_sample.py:6-54 -- 320-sample sections around each R-peak (short slices
    left-aligned in a NaN row), column-wise nanmedian on the GPU."""
    if resample_ecg:
        raise NotImplementedError("resample_ecg=True is dead code on the reference path")
    section, pre_len, post_len = beat_geometry(sampling_frequency)      # 320 / 106 / 214 at 500 Hz
    if not 1 <= section <= 512:
        raise _cabi.AewError("beat sections of 1..512 samples (sampling rates up to 800 Hz)")
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
    med = torch.empty(section, dtype=torch.float64, device="cuda")
    aligned = torch.empty((peaks.size, section), dtype=torch.float64, device="cuda") if return_aligned else None
    _cabi.check(lib.aew_median_beats_f64(_ptr(x), 1, x.numel(), _ptr(d_peaks), _ptr(d_off), int(peaks.size),
                                         pre_len, post_len, _ptr(med), _ptr(aligned), _cabi.stream_ptr(torch)),
                "median_beats")
    if return_aligned:
        return med.cpu().numpy(), aligned.cpu().numpy()
    return med.cpu().numpy()


def compute_median_features_ecg(ecg_features, ecg_features_morphology):
    """median_sample.py:57-82 -- np.nanmedian over the beats of every fiducial's x and y and of every morphology
    feature: all rows go to the device as one (rows, n_beats) array, one warp per row (``aew_nanmedian_rows_f64``).
    Returns ``{fiducial: {'x': m, 'y': m}, ..., morphology feature: m, ...}`` with NumPy float64 scalars."""
    torch, lib = _cabi.require_cuda()
    rows, slots = [], []
    for name, xy in ecg_features.items():
        for axis in ("x", "y"):
            rows.append(np.asarray(xy[axis], dtype=np.float64).reshape(-1))
            slots.append((name, axis))
    for name, vals in ecg_features_morphology.items():
        rows.append(np.asarray(vals, dtype=np.float64).reshape(-1))
        slots.append((name, None))
    n = max((len(r) for r in rows), default=0)
    mat = np.full((len(rows), max(n, 1)), np.nan)
    for i, r in enumerate(rows):                       # NaN padding is ignored by nanmedian (ragged rows allowed)
        mat[i, :len(r)] = r
    med = np.full(len(rows), np.nan)
    if rows:
        d_mat = torch.from_numpy(mat).cuda()
        d_out = torch.empty(len(rows), dtype=torch.float64, device="cuda")
        d_scratch = torch.empty_like(d_mat)
        _cabi.check(lib.aew_nanmedian_rows_f64(_ptr(d_mat), len(rows), mat.shape[1], _ptr(d_scratch),
                                               _ptr(d_out), _cabi.stream_ptr(torch)), "nanmedian_rows")
        med = d_out.cpu().numpy()
    out = {}
    for (name, axis), m in zip(slots, med):
        if axis is None:
            out[name] = np.float64(m)
        else:
            out.setdefault(name, {})[axis] = np.float64(m)
    return out


def plot_median_ecg(*args, **kwargs):
    """Imported by data_extract_vitaldb.py:14 (only called from commented-out code)."""
    raise NotImplementedError("plotting is outside the preprocessing hot path")
