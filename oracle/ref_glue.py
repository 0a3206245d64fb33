"""The REFERENCE'S OWN functions driven by the glue of its caller (TEST INFRASTRUCTURE / CPU baseline).

``process_case`` walks data_extract_vitaldb.py:53-106 (and :108-172 with ``features``) exactly as the
reference's ``process_case_id`` does, but every arithmetic step is a call into the reference's modules as
imported by ``oracle.ref_import`` (from /root/reference in the build container, from the byte-compiled
``oracle/_ref`` on the GPU box).  Only what the reference itself delegates to absent third parties is not its own
code: ``neurokit2`` (``ecg_findpeaks`` / ``ecg_delineate``) resolves to the restatements in ``oracle/nk_peaks.py``
/ ``oracle/nk_delineate.py`` -- PARITY UNPINNED at that boundary -- and record loading (vitaldb, network) is
replaced by the caller handing over the track.  ``bench.py``'s CPU arm times this (``cpu_baseline.kind =
"reference"``); the tests compare it with ``oracle.ref_port`` and with the CUDA path.
"""
from __future__ import annotations

import numpy as np

from . import ref_import

TRACKS = ["SNUADC/ECG_II"]
FEATURES = [
    "ECG_P_Peaks", "ECG_P_Onsets", "ECG_P_Offsets", "ECG_Q_Peaks", "ECG_R_Onsets", "ECG_R_Offsets", "ECG_S_Peaks",
    "ECG_T_Peaks", "ECG_T_Onsets", "ECG_T_Offsets",
    "ECG_P_Duration", "ECG_Q_Duration", "ECG_R_Duration", "ECG_S_Duration", "ECG_T_Duration",
    "ECG_PR_Interval", "ECG_QRS_Complex", "ECG_ST_Interval", "ECG_ST_Segment", "ECG_QT_Interval",
    "ECG_P_Prominence", "ECG_Q_Prominence", "ECG_R_Prominence", "ECG_S_Prominence", "ECG_T_Prominence"]

_ref = None


def ref():
    global _ref
    if _ref is None:
        _ref = ref_import.load()
    return _ref


def process_case(track, case_id=0, fs=500, sample_length=20, hp=1, lp=30, order=4, pvc_threshold=150,
                 features=False):
    """One single-lead record through the reference's functions.  Returns ``np_waves`` (n,320) f32, ``np_info``
    (n,2) f64 (+ ``np_features`` (n,35) f64 with ``features``) as data_create_arrays.py:34-46 stacks them, plus
    ``n_beats`` and the frame of kept windows."""
    import sys
    R = ref()
    record = np.asarray(track, dtype=np.float64).reshape(-1, 1)
    non_nan = ~np.isnan(record).all(axis=1)                                           # :53-57
    if not non_nan.any():
        return dict(np_waves=np.zeros((0, 320), np.float32), np_info=np.zeros((0, 2)), n_beats=0, frame=None)
    record = record[np.argmax(non_nan):len(non_nan) - np.argmax(non_nan[::-1])].copy()
    z = R.filters.zscore(record[:, 0])                                                # :64
    f = R.filters.butterworth(data=z, btype="high", cutoff_frequency=hp, sampling_frequency=fs,
                              filter_order=order) if hp else z                        # :66-70
    f = R.filters.butterworth(data=f, btype="low", cutoff_frequency=lp, sampling_frequency=fs,
                              filter_order=order) if lp else f                        # :71-76
    record[:, 0] = f                                                                  # :78
    samples, idx = R.sample_from_record.sample_from_record(record=record, track_names=TRACKS,
                                                           sampling_frequency=fs, sample_length=sample_length)
    if len(samples) == 0:                                                             # :89-91
        return dict(np_waves=np.zeros((0, 320), np.float32), np_info=np.zeros((0, 2)), n_beats=0, frame=None)
    samples[:, :, 0] = np.apply_along_axis(R.filters.zscore, 1, samples[:, :, 0])     # :93
    df = R.remove_samples.remove_samples(samples_wave_tracks=samples, indices_samples=idx, track_names=TRACKS,
                                         sampling_frequency=fs, pvc_threshold=pvc_threshold)   # :96-97
    if len(df) == 0 or "ECG_R_Peaks" not in df:
        return dict(np_waves=np.zeros((0, 320), np.float32), np_info=np.zeros((0, 2)), n_beats=0, frame=df)
    med = [R.median_sample.compute_median_sample_ecg(s, p, fs, return_aligned=True)[0]
           for s, p in zip(df["ecg_sample"], df["ECG_R_Peaks"])]                      # :104-106
    valid = np.ones(len(df), dtype=bool)
    rows = None
    if features:                                                                      # :108-172
        nk = sys.modules["neurokit2"]
        rows = np.full((len(df), 35), np.nan)
        raised = not (hp and lp)                                                      # :166-170: NaN for every window
        for i, (s, p) in enumerate(zip(df["ecg_sample"], df["ECG_R_Peaks"])):
            if not (hp and lp):
                continue
            try:
                fid = nk.ecg_delineate(s, p, sampling_rate=fs)[1]
            except Exception:                                                         # bare except, :132-135
                raised = True
                continue
            if not all(len(fid[k]) == len(p) for k in fid.keys()) or len(p) == 0:     # :140-144
                valid[i] = False
                continue
            rel = R.features_ecg.ecg_features_relative_r(s, fid, p, fs)
            morph = R.features_ecg.compute_ecg_morphology_features(rel, len(p))
            m = R.median_sample.compute_median_features_ecg(rel, morph)
            rows[i, :20] = np.array([[m[k]["x"], m[k]["y"]] for k in FEATURES[:10]]).reshape(-1)   # dca :34-35
            rows[i, 20:] = [m[k] for k in FEATURES[10:]]                              # dca :36
        if raised:               # data_create_arrays.py:37-40: a NaN cell among the dicts -> TypeError -> zeros for the case
            rows[valid] = 0.0
    sel = np.flatnonzero(valid)
    waves = np.stack([med[i] for i in sel]).astype(np.float32) if sel.size else np.zeros((0, 320), np.float32)
    info = np.stack([np.full(sel.size, float(case_id)),
                     np.array([float(np.asarray(df["indices_samples"].iloc[i]).reshape(-1)[0]) for i in sel])],
                    axis=1) if sel.size else np.zeros((0, 2))
    out = dict(np_waves=waves, np_info=info, frame=df,
               n_beats=int(sum(len(df["ECG_R_Peaks"].iloc[i]) for i in sel)))
    if features:
        out["np_features"] = rows[sel]
    return out
