"""Drop-in for data_load_preprocess/feature_extraction/features_ecg.py (reference lines 56-136): the two
functions the caller imports (data_extract_vitaldb.py:15), same arguments and return structure, the arithmetic in
the stage kernels of K6 (``aew_fiducial_relative_f64``, ``aew_morphology_f64``).  The batched path computes the
same numbers for all windows at once (``aew_fiducial_features_f64``)."""
from __future__ import annotations

import ctypes

import numpy as np

from ... import _cabi

FID_NAN = -2147483648
# the ten delineation outputs in np_features order (data_extract_vitaldb.py:109-112); slot 10 = the inserted R peak
FIDUCIALS = ["ECG_P_Peaks", "ECG_P_Onsets", "ECG_P_Offsets", "ECG_Q_Peaks", "ECG_R_Onsets", "ECG_R_Offsets",
             "ECG_S_Peaks", "ECG_T_Peaks", "ECG_T_Onsets", "ECG_T_Offsets"]
MORPHOLOGY = ["ECG_P_Duration", "ECG_Q_Duration", "ECG_R_Duration", "ECG_S_Duration", "ECG_T_Duration",
              "ECG_PR_Interval", "ECG_QRS_Complex", "ECG_ST_Interval", "ECG_ST_Segment", "ECG_QT_Interval",
              "ECG_P_Prominence", "ECG_Q_Prominence", "ECG_R_Prominence", "ECG_S_Prominence", "ECG_T_Prominence"]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def ecg_features_relative_r(ecg_sample, ecg_features, r_peaks, sampling_frequency):
    """features_ecg.py:56-86.  ``ecg_features`` maps names to per-beat sample indices with the ``np.nan``
    SINGLETON for a missing point (the reference tests identity, not isnan); returns ``{name: {'x': [...],
    'y': [...]}}``: milliseconds and amplitude relative to the beat's R peak, NaN where the point is missing."""
    torch, lib = _cabi.require_cuda()
    names = list(ecg_features.keys())
    nb = len(r_peaks)
    if not names or nb == 0:
        return {k: {"x": [], "y": []} for k in names}
    idx = np.empty((len(names), nb), dtype=np.int32)
    for f, k in enumerate(names):
        col = ecg_features[k]
        idx[f] = [FID_NAN if v is np.nan else int(v) for v in (col[b] for b in range(nb))]
    sig = np.ascontiguousarray(ecg_sample, dtype=np.float64)
    L = len(sig)
    rp = np.asarray(r_peaks, dtype=np.int64)
    neg = idx < 0
    neg &= idx != FID_NAN
    idx = np.where(neg, idx + L, idx).astype(np.int32)            # Python's negative indexing of ecg_sample[idx]
    rp = np.where(rp < 0, rp + L, rp).astype(np.int32)
    if ((idx != FID_NAN) & ((idx < 0) | (idx >= L))).any() or ((rp < 0) | (rp >= L)).any():
        raise IndexError("index out of bounds for the ECG sample")
    d_sig, d_idx, d_rp = torch.from_numpy(sig).cuda(), torch.from_numpy(idx).cuda(), torch.from_numpy(rp).cuda()
    xy = torch.empty((len(names), 2, nb), dtype=torch.float64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    # x is (idx - r) / fs * 1000 with the ORIGINAL (possibly negative) index difference: wrapped indices only
    # address the sample, so the time axis is computed from the unwrapped values below when any index wrapped
    _cabi.check(lib.aew_fiducial_relative_f64(_p(d_sig), L, _p(d_rp), _p(d_idx), len(names), nb,
                                              float(sampling_frequency), _p(xy), _p(status),
                                              _cabi.stream_ptr(torch)), "fiducial_relative")
    out = xy.cpu().numpy()
    if neg.any() or (np.asarray(r_peaks) < 0).any():
        raw = np.array([[np.nan if v is np.nan else float(v) for v in (ecg_features[k][b] for b in range(nb))]
                        for k in names])
        out[:, 0, :] = (raw - np.asarray(r_peaks, dtype=np.float64)[None, :]) / sampling_frequency * 1000
    return {k: {"x": [float(v) for v in out[f, 0]], "y": [float(v) for v in out[f, 1]]} for f, k in enumerate(names)}


def compute_ecg_morphology_features(ecg_features, nr_of_beats):
    """features_ecg.py:89-136 -- ten durations and five prominences per beat from the relative coordinates.  Like
    the reference it ADDS ``ecg_features['ECG_R_Peaks'] = {'x': zeros, 'y': zeros}`` to its argument (:111; the key
    then flows through compute_median_features_ecg)."""
    torch, lib = _cabi.require_cuda()
    ecg_features['ECG_R_Peaks'] = {'x': [0.0] * nr_of_beats, 'y': [0.0] * nr_of_beats}
    order = FIDUCIALS + ['ECG_R_Peaks']
    xs = [np.asarray(ecg_features[k]['x'], dtype=np.float64) for k in order]
    ys = [np.asarray(ecg_features[k]['y'], dtype=np.float64) for k in order]
    if any(a.shape != (nr_of_beats,) for a in xs + ys):
        raise ValueError("operands could not be broadcast together: every fiducial needs one entry per beat")
    if nr_of_beats == 0:
        return {k: np.zeros(0) for k in MORPHOLOGY}
    d_x, d_y = torch.from_numpy(np.stack(xs)).cuda(), torch.from_numpy(np.stack(ys)).cuda()
    out = torch.empty((15, nr_of_beats), dtype=torch.float64, device="cuda")
    _cabi.check(lib.aew_morphology_f64(_p(d_x), _p(d_y), nr_of_beats, _p(out), _cabi.stream_ptr(torch)), "morphology")
    res = out.cpu().numpy()
    return {k: res[i].copy() for i, k in enumerate(MORPHOLOGY)}
