"""Restatement of ``neurokit2.ecg_delineate(ecg, rpeaks, sampling_rate)[1]`` (default method "dwt").

TEST INFRASTRUCTURE.  PARITY UNPINNED: neurokit2 is a third-party dependency of the
reference (requirements.txt:10, no version pin), it is not vendored under
/root/reference, not installed in this image and cannot be fetched.  The only call
site on the path is data_extract_vitaldb.py:131; the reference holds no test or golden
vector for it.  This file restates the published method of neurokit2 0.2.x
(``ecg_delineate`` -> ``_dwt_ecg_delineator``) on the SciPy /
NumPy primitives it delegates to, which ARE present (``scipy.ndimage.zoom``,
``scipy.signal.find_peaks`` / ``peak_prominences``, ``scipy.interpolate.PchipInterpolator``).
Written from the published description, not from a copy of the library; the points that
are version-sensitive are marked [V] and listed in DESIGN.md.

Structure (function names of the library in brackets):
  1. Q and S peaks at the input rate from HR-sized epochs around each R
     [ecg_segment, _ecg_segment_window, epochs_create, _ecg_delineator_peak_Q/_S].
  2. cubic-spline resampling to 2,000 Hz [signal_resample, method "interpolation"].
  3. 9-level a-trous quadratic-spline wavelet bank [_dwt_compute_multiscales].
  4. T and P peaks from zero crossings between significant wavelet extrema [_dwt_delineate_tp_peaks].
  5. QRS on/offsets and P/T on/offsets from threshold crossings of the wavelet
     [_dwt_delineate_qrs_bounds, _dwt_delineate_tp_onsets_offsets].
  6. indices mapped back to the input rate by truncation [_dwt_resample_points];
     the wrapper blanks a last entry >= len(signal) and drops entries <= 0.
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import PchipInterpolator
from scipy.ndimage import zoom
from scipy.signal import find_peaks, peak_prominences

ANALYSIS_RATE = 2000
KEYS = ["ECG_P_Peaks", "ECG_P_Onsets", "ECG_P_Offsets", "ECG_Q_Peaks", "ECG_R_Onsets",
        "ECG_R_Offsets", "ECG_S_Peaks", "ECG_T_Peaks", "ECG_T_Onsets", "ECG_T_Offsets"]
NAN = np.nan          # the singleton: the reference tests `idx is np.nan` (features_ecg.py:75)


class DelineationError(Exception):
    """Stands for any exception neurokit2 would raise; the reference catches it with a
    bare except and stores NaN features (data_extract_vitaldb.py:130-135)."""


# ---------------------------------------------------------------- heart rate helpers
def signal_period(peaks, fs, desired_length=None):
    """[signal_period] diff / fs with the mean of the others prepended; optionally interpolated to every
    sample with a monotone cubic (PCHIP), constant outside the first / last peak."""
    peaks = np.asarray(peaks)
    if peaks.size <= 3:
        raise DelineationError("too few peaks")          # np.full(None / n, nan) -> int(nan) / empty shape
    period = np.ediff1d(peaks, to_begin=0) / fs
    period[0] = np.mean(period[1:])
    if desired_length is None:
        return period
    if desired_length <= peaks[-1]:
        raise DelineationError("desired_length")
    x_new = np.arange(desired_length)
    out = PchipInterpolator(peaks, period, extrapolate=True)(x_new)
    out[:int(peaks[0])] = out[int(peaks[0])]
    out[int(peaks[-1]):] = out[int(peaks[-1])]
    return out


def signal_rate(peaks, fs, desired_length=None):
    return 60 / signal_period(peaks, fs, desired_length)


# ---------------------------------------------------------------- 1. Q and S peaks
def segment_window(rpeaks, fs, n):
    """[_ecg_segment_window] [V: 0.2.8+ form with ratio_pre = 0.35] -> (start_s, end_s)."""
    heart_rate = np.mean(signal_rate(rpeaks, fs, desired_length=n))
    window_size = 60 / heart_rate
    return -(0.35 * window_size), (1 - 0.35) * window_size


def epoch_geometry(r, start_s, end_s, fs):
    """[epochs_create] sample range of the epoch around onset r and the positions that the
    label slices [:0] / [0:] select.  Returns (s0, n, n_left, right0): the epoch is samples
    s0 .. s0 + n of the signal; its time labels are np.linspace(start_s, end_s, n);
    `heartbeat[:0]` = the first n_left samples (labels <= 0), `heartbeat[0:]` starts at right0
    (labels >= 0)."""
    buf = int((end_s - start_s) * fs)                    # epoch_max_duration, the NaN/zero buffer
    a = int(r + buf + start_s * fs)
    b = int(r + buf + end_s * fs)
    n = b - a
    labels = np.linspace(start_s, end_s, n) if n > 0 else np.zeros(0)
    return a - buf, n, int(np.count_nonzero(labels <= 0)), int(np.count_nonzero(labels < 0))


def _epoch_samples(sig, s0, n, last):
    """Samples of an epoch; outside the signal the buffer is 0 [V: the buffer is created as
    zeros and its NaN replacement acts on a copy]; the last beat's samples past the end are
    NaN [ecg_segment], which for peak finding equals cutting the segment there."""
    idx = s0 + np.arange(n)
    out = np.zeros(n)
    ok = (idx >= 0) & (idx < len(sig))
    out[ok] = sig[idx[ok]]
    if last:
        out = out[idx < len(sig)]
    return out


def _first_or_last_prominent(seg, which):
    """[signal_findpeaks(-seg, height_min = 0.05 * range)] 'Height' is the prominence."""
    if len(seg) == 0:
        return None
    neg = -seg
    pk, _ = find_peaks(neg)
    if len(pk) == 0:
        return None
    prom = peak_prominences(neg, pk)[0]
    pk = pk[~(prom < 0.05 * (seg.max() - seg.min()))]
    if len(pk) == 0:
        return None
    return int(pk[-1] if which == "last" else pk[0])


def qs_peaks(sig, rpeaks, fs):
    """[_dwt_ecg_delineator, first loop] Q = right-most prominent minimum left of R,
    S = left-most prominent minimum from R on; NaN when there is none."""
    if len(sig) < fs * 4:
        raise DelineationError("too short")
    start_s, end_s = segment_window(rpeaks, fs, len(sig))
    q, s = [], []
    for i, r in enumerate(rpeaks):
        s0, n, n_left, right0 = epoch_geometry(int(r), start_s, end_s, fs)
        ep = _epoch_samples(sig, s0, n, last=(i == len(rpeaks) - 1))
        if n_left >= n:
            raise DelineationError("no label above 0")   # np.min of an empty array
        qi = _first_or_last_prominent(ep[:n_left], "last")
        q.append(NAN if qi is None else int(r) - (n_left - qi))
        si = _first_or_last_prominent(ep[right0:], "first")
        s.append(NAN if si is None else int(r) + si)
    return q, s


# ---------------------------------------------------------------- 2./3. resample + wavelet bank
def resample_2k(sig, fs):
    """[signal_resample(method='interpolation')] scipy.ndimage.zoom, cubic spline."""
    want = int(np.round(len(sig) * ANALYSIS_RATE / fs))
    out = zoom(sig, want / len(sig))
    if len(out) < want:
        out = np.concatenate([out, np.full(want - len(out), out[-1])])
    return out[:want]


def multiscales(ecg, max_degree=9):
    """[_dwt_compute_multiscales] row k = G_k(H_{k-1} ... H_0 ecg); every filter is a 'full'
    convolution followed by a left shift of 2**k that leaves the last 2**k entries in place;
    the low-pass intermediate keeps its growing length; rows are cut to len(ecg).  The
    library convolves with scipy.signal.convolve (which may pick an FFT for the long sparse
    kernels); here the sparse taps are applied directly, equal up to rounding."""
    n = len(ecg)
    rows = []
    cur = np.array(ecg, dtype=np.float64)
    for k in range(max_degree):
        d = 2 ** k
        g = np.zeros(len(cur) + d)
        g[:len(cur)] += 2 * cur
        g[d:] -= 2 * cur
        g[:-d] = g[d:].copy()
        h = np.zeros(len(cur) + 3 * d)
        for j, c in enumerate((0.125, 0.375, 0.375, 0.125)):
            h[j * d:j * d + len(cur)] += c * cur
        h[:-d] = h[d:].copy()
        rows.append(g[:n])
        cur = h
    return np.array(rows)


def resample_points(points, src, dst):
    """[_dwt_resample_points] scale and truncate; NaN stays the np.nan singleton."""
    arr = np.array(points)
    if arr.dtype.kind in "iu":
        arr = arr.astype(np.int64)
    scaled = arr * dst / src
    return [NAN if np.isnan(x) else int(x) for x in scaled.tolist()]


def adjust_degree(rpeaks2k):
    average_rate = np.median(signal_rate(rpeaks2k, ANALYSIS_RATE))
    return int(np.log2((ANALYSIS_RATE / 250) / (average_rate / 60)))


def adjust_duration(rpeaks2k, duration):
    average_rate = np.median(signal_rate(rpeaks2k, ANALYSIS_RATE))
    return np.round(duration * (60 / average_rate), 3)


# ---------------------------------------------------------------- 4. T and P peaks
def _zero_crossing_candidates(dwt_local, ecg_local, height, score_fn):
    peaks, _ = find_peaks(np.abs(dwt_local), height=height)
    top = max(dwt_local)
    peaks = [int(p) for p in peaks if np.abs(dwt_local[p]) > 0.025 * top]
    if dwt_local[0] > 0:
        peaks = [0] + peaks
    best, best_score = None, None
    for a, b in zip(peaks[:-1], peaks[1:]):
        if dwt_local[a] > 0 and dwt_local[b] < 0:
            sgn = np.sign(dwt_local[a:b + 1])
            z = int(np.nonzero(np.abs(np.diff(sgn)) > 0)[0][0]) + a
            sc = score_fn(z, ecg_local[z])
            if best is None or sc > best_score:          # np.argmax: first maximum
                best, best_score = z, sc
    return best


def tp_peaks(ecg2k, rpeaks2k, dwt, qrs_width=0.13, p2r_duration=0.2, rt_duration=0.25):
    fs = ANALYSIS_RATE
    bndry = int(0.5 * qrs_width * fs)
    deg = adjust_degree(rpeaks2k)
    p2r = adjust_duration(rpeaks2k, p2r_duration)
    rt = adjust_duration(rpeaks2k, rt_duration)
    if not (0 <= 3 + deg < dwt.shape[0] and 0 <= 2 + deg):
        raise DelineationError("wavelet scale out of range")
    tpk, ppk = [], []
    for r in rpeaks2k:
        a, b = r + bndry, r + 2 * int(rt * fs)
        loc = dwt[3 + deg, a:b]
        if len(loc) == 0:
            tpk.append(NAN)
            continue
        height = 0.25 * np.sqrt(np.mean(np.square(loc)))
        z = _zero_crossing_candidates(loc, ecg2k[a:b], height,
                                      lambda i, y: y - (float(i) / fs - (rt - 0.5 * qrs_width)))
        tpk.append(NAN if z is None else z + a)
    for r in rpeaks2k:
        a, b = r - 2 * int(p2r * fs), r - bndry
        loc = dwt[2 + deg, a:b]
        if len(loc) == 0:
            ppk.append(NAN)
            continue
        height = 0.02 * np.sqrt(np.mean(np.square(loc)))
        z = _zero_crossing_candidates(loc, ecg2k[a:b], height, lambda i, y: y - abs(float(i) / fs - p2r))
        ppk.append(NAN if z is None else z + a)
    return tpk, ppk


# ---------------------------------------------------------------- 5. onsets / offsets
def _onset(row, a, b, sign, weight):
    """last sample before the last local maximum of sign*row[a:b] where sign*row < weight * that maximum"""
    if a is NAN or b is NAN:
        return NAN
    loc = sign * row[a:b]
    pk, _ = find_peaks(loc)
    if len(pk) == 0:
        return NAN
    below = np.where(loc[:pk[-1]] < weight * loc[pk[-1]])[0]
    return NAN if len(below) == 0 else int(below[-1]) + a


def _offset(row, a, b, sign, weight):
    """first sample from the first local maximum of sign*row[a:b] on where sign*row < weight * that maximum"""
    if a is NAN or b is NAN:
        return NAN
    loc = sign * row[a:b]
    pk, _ = find_peaks(loc)
    if len(pk) == 0:
        return NAN
    below = np.where(loc[pk[0]:] < weight * loc[pk[0]])[0]
    return NAN if len(below) == 0 else int(below[0]) + int(pk[0]) + a


def qrs_bounds(rpeaks2k, dwt, ppeaks, tpeaks, qpeaks2k):
    """[_dwt_delineate_qrs_bounds] [V: onset search ends at the Q peak]"""
    row = dwt[2 + adjust_degree(rpeaks2k)]
    on = [_onset(row, ppeaks[i], qpeaks2k[i], -1.0, 0.5) for i in range(len(qpeaks2k))]
    off = [_offset(row, rpeaks2k[i], tpeaks[i], 1.0, 0.5) for i in range(len(rpeaks2k))]
    return on, off


def tp_bounds(peaks, rpeaks2k, dwt, duration_onset=0.3, duration_offset=0.3, onset_weight=0.4, offset_weight=0.4):
    """[_dwt_delineate_tp_onsets_offsets]"""
    fs = ANALYSIS_RATE
    d_on = adjust_duration(rpeaks2k, duration_onset)
    d_off = adjust_duration(rpeaks2k, duration_offset)
    row = dwt[2 + adjust_degree(rpeaks2k)]
    on = [_onset(row, NAN if p is NAN else p - int(d_on * fs), p, 1.0, onset_weight) for p in peaks]
    off = [_offset(row, p, NAN if p is NAN else p + int(d_off * fs), -1.0, offset_weight) for p in peaks]
    return on, off


# ---------------------------------------------------------------- the call the reference makes
def dwt_delineator(sig, rpeaks, fs):
    sig = np.asarray(sig, dtype=np.float64)
    rpeaks = [int(r) for r in rpeaks]
    qpk, spk = qs_peaks(sig, rpeaks, fs)
    ecg2k = resample_2k(sig, fs)
    dwt = multiscales(ecg2k, 9)
    r2k = resample_points(rpeaks, fs, ANALYSIS_RATE)
    q2k = resample_points(qpk, fs, ANALYSIS_RATE)
    tpk, ppk = tp_peaks(ecg2k, r2k, dwt)
    ron, roff = qrs_bounds(r2k, dwt, ppk, tpk, q2k)
    pon, poff = tp_bounds(ppk, r2k, dwt)
    ton, toff = tp_bounds(tpk, r2k, dwt, onset_weight=0.6, duration_onset=0.6)
    back = lambda v: resample_points(v, ANALYSIS_RATE, fs)     # noqa: E731
    return dict(ECG_P_Peaks=back(ppk), ECG_P_Onsets=back(pon), ECG_P_Offsets=back(poff), ECG_Q_Peaks=qpk,
                ECG_R_Onsets=back(ron), ECG_R_Offsets=back(roff), ECG_S_Peaks=spk, ECG_T_Peaks=back(tpk),
                ECG_T_Onsets=back(ton), ECG_T_Offsets=back(toff))


def ecg_delineate(ecg_cleaned, rpeaks, sampling_rate=500):
    """[ecg_delineate(...)[1]] the sanitised dict the reference consumes: a last entry
    >= len(signal) becomes NaN, entries that are neither > 0 nor np.nan are DROPPED (which is
    what makes list lengths differ from len(rpeaks), data_extract_vitaldb.py:140)."""
    if len(rpeaks) == 0:
        raise DelineationError("no peaks")
    waves = dwt_delineator(ecg_cleaned, rpeaks, sampling_rate)
    for v in waves.values():
        if v[-1] >= len(ecg_cleaned):
            v[-1] = NAN
    return {k: [x for x in v if x > 0 or x is NAN] for k, v in waves.items()}


def window_fiducials(ecg_sample, r_peaks, fs=500):
    """Glue of data_extract_vitaldb.py:129-144 for one window -> (status, fid) with
    status 'ok' | 'nan' (exception: NaN features, window kept) | 'drop' (list lengths differ:
    window removed); fid (10, n_beats) float64 with NaN = missing, only for 'ok'."""
    try:
        feats = ecg_delineate(ecg_sample, r_peaks, fs)
    except DelineationError:
        return "nan", None
    if not all(len(feats[k]) == len(r_peaks) for k in KEYS) or len(r_peaks) == 0:
        return "drop", None
    return "ok", np.array([[np.nan if x is NAN else float(x) for x in feats[k]] for k in KEYS])
