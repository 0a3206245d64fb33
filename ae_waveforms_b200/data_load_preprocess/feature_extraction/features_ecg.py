"""Drop-in for data_load_preprocess/feature_extraction/features_ecg.py (reference
lines 56-136).  A few dozen scalars per window; kept as host arithmetic until the wave
delineator produces fiducials on the device."""
from __future__ import annotations

import numpy as np


def ecg_features_relative_r(ecg_sample, ecg_features, r_peaks, sampling_frequency):
    """features_ecg.py:56-86 -- fiducial index -> (ms, amplitude) relative to its R-peak;
    missing points are the np.nan singleton (identity test, as in the reference)."""
    rel = {}
    for feat in ecg_features.keys():
        xs, ys = [], []
        for beat_nr, r_idx in enumerate(r_peaks):
            idx = ecg_features[feat][beat_nr]
            if idx is np.nan:
                xs.append(np.nan)
                ys.append(np.nan)
            else:
                xs.append((idx - r_idx) / sampling_frequency * 1000)
                ys.append(ecg_sample[idx] - ecg_sample[r_idx])
        rel[feat] = {"x": xs, "y": ys}
    return rel


def compute_ecg_morphology_features(ecg_features, nr_of_beats):
    """features_ecg.py:89-136 -- ten durations and five prominences per beat.  Like the
    reference it ADDS an 'ECG_R_Peaks' entry to its argument."""
    f = ecg_features
    f['ECG_R_Peaks'] = {'x': [0.0] * nr_of_beats, 'y': [0.0] * nr_of_beats}

    def dt(a, b):
        return np.array(f[b]['x']) - np.array(f[a]['x'])

    def dy(peak, base):
        if type(base) is list:
            ref = np.nanmax(np.stack([f[base[0]]['y'], f[base[1]]['y']]), axis=0)
        else:
            ref = f[base]['y']
        return np.array(f[peak]['y']) - np.array(ref)

    return {
        "ECG_P_Duration": dt("ECG_P_Onsets", "ECG_P_Offsets"),
        "ECG_Q_Duration": dt("ECG_R_Onsets", "ECG_R_Peaks"),
        "ECG_R_Duration": dt("ECG_Q_Peaks", "ECG_S_Peaks"),
        "ECG_S_Duration": dt("ECG_R_Peaks", "ECG_R_Offsets"),
        "ECG_T_Duration": dt("ECG_T_Onsets", "ECG_T_Offsets"),
        "ECG_PR_Interval": dt("ECG_P_Onsets", "ECG_R_Peaks"),
        "ECG_QRS_Complex": dt("ECG_R_Onsets", "ECG_R_Offsets"),
        "ECG_ST_Interval": dt("ECG_R_Offsets", "ECG_T_Offsets"),
        "ECG_ST_Segment": dt("ECG_R_Offsets", "ECG_T_Onsets"),
        "ECG_QT_Interval": dt("ECG_R_Onsets", "ECG_T_Offsets"),
        "ECG_P_Prominence": dy("ECG_P_Peaks", ["ECG_P_Onsets", "ECG_P_Offsets"]),
        "ECG_Q_Prominence": dy("ECG_Q_Peaks", "ECG_R_Onsets"),
        "ECG_R_Prominence": dy("ECG_R_Peaks", ["ECG_Q_Peaks", "ECG_S_Peaks"]),
        "ECG_S_Prominence": dy("ECG_S_Peaks", "ECG_R_Offsets"),
        "ECG_T_Prominence": dy("ECG_T_Peaks", ["ECG_T_Onsets", "ECG_T_Offsets"]),
    }
