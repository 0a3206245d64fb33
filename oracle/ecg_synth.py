"""Seeded synthetic 500 Hz ECG (TEST INFRASTRUCTURE / bench input generator).

Signal model of SURVEY.md section 8d: every beat is a sum of five Gaussians
(P, Q, R, S, T) placed relative to the R time; per-case heart rate is drawn
from U(50, 110) bpm with 3 % Gaussian RR jitter; white noise, baseline wander
and mains hum are added.  VitalDB / MIMIC data are not available offline, so
this generator feeds the parity tests, the CPU baseline and the GPU bench.
"""
from __future__ import annotations

import numpy as np

FS = 500
# offsets [s], widths [s], amplitudes [mV] for P, Q, R, S, T
WAVE_OFFSETS = np.array([-0.20, -0.025, 0.0, 0.030, 0.25])
WAVE_WIDTHS = np.array([0.025, 0.010, 0.012, 0.012, 0.05])
WAVE_AMPS = np.array([0.15, -0.12, 1.0, -0.25, 0.30])
NOISE_SD = 0.02
WANDER_AMP, WANDER_HZ = 0.1, 0.25
MAINS_AMP, MAINS_HZ = 0.01, 50.0


def case_rng(case_id: int) -> np.random.Generator:
    return np.random.default_rng(0xEC6 + int(case_id))


def r_times(n_samples: int, rng: np.random.Generator, hr_bpm: float | None = None,
            jitter: float = 0.03) -> np.ndarray:
    """R-peak times in seconds covering [-1 s, duration + 1 s]."""
    if hr_bpm is None:
        hr_bpm = rng.uniform(50.0, 110.0)
    rr = 60.0 / hr_bpm
    duration = n_samples / FS
    n_beats = int(np.ceil((duration + 2.0) / (rr * 0.8))) + 2
    intervals = rr * (1.0 + jitter * rng.standard_normal(n_beats))
    intervals = np.maximum(intervals, 0.25)
    t = -1.0 + rng.uniform(0.0, rr) + np.cumsum(intervals)
    return t[t < duration + 1.0]


def synth_case(case_id: int, n_samples: int, hr_bpm: float | None = None,
               nan_gaps: int = 0, jitter: float = 0.03) -> np.ndarray:
    """One single-lead record, float64, shape (n_samples,)."""
    rng = case_rng(case_id)
    t_r = r_times(n_samples, rng, hr_bpm, jitter)
    sig = np.zeros(n_samples, dtype=np.float64)
    for off, wid, amp in zip(WAVE_OFFSETS, WAVE_WIDTHS, WAVE_AMPS):
        half = int(np.ceil(5.0 * wid * FS))
        centres = (t_r + off) * FS
        for c in centres:
            lo = max(0, int(np.floor(c)) - half)
            hi = min(n_samples, int(np.floor(c)) + half + 1)
            if hi <= lo:
                continue
            k = np.arange(lo, hi)
            sig[lo:hi] += amp * np.exp(-0.5 * ((k - c) / (wid * FS)) ** 2)
    t = np.arange(n_samples) / FS
    sig += WANDER_AMP * np.sin(2 * np.pi * WANDER_HZ * t + rng.uniform(0, 2 * np.pi))
    sig += MAINS_AMP * np.sin(2 * np.pi * MAINS_HZ * t + rng.uniform(0, 2 * np.pi))
    sig += NOISE_SD * rng.standard_normal(n_samples)
    if nan_gaps:
        for _ in range(nan_gaps):
            glen = int(FS * np.exp(rng.uniform(np.log(0.01), np.log(30.0))))
            glen = max(1, min(glen, n_samples // 4))
            start = int(rng.integers(0, max(1, n_samples - glen)))
            sig[start:start + glen] = np.nan
    return sig


def synth_fiducials(n_beats: int, seed: int = 0, nan_frac: float = 0.02):
    """Synthetic inputs of the sigma search.

    Returns ``test_true`` (n,320) float32 (median beat minus its row maximum),
    ``features_x`` / ``features_y`` (n,11) float64 in the reordered layout of
    analysis_performance_traditional_vs_ae.py:40-44 (R inserted at slot 5 as
    (0, -1e-10)) and ``baselines`` (n,) float64.
    """
    rng = np.random.default_rng(seed)
    ms = np.arange(-106, 214) / FS * 1000.0                       # (320,) ms rel. to R
    # per-beat morphology jitter
    amp = WAVE_AMPS * (1.0 + 0.1 * rng.standard_normal((n_beats, 5)))
    off = WAVE_OFFSETS * 1000.0 * (1.0 + 0.05 * rng.standard_normal((n_beats, 5)))
    wid = WAVE_WIDTHS * 1000.0 * (1.0 + 0.05 * rng.standard_normal((n_beats, 5)))
    beat = np.zeros((n_beats, 320))
    for w in range(5):
        beat += amp[:, w:w + 1] * np.exp(-0.5 * ((ms[None, :] - off[:, w:w + 1]) / wid[:, w:w + 1]) ** 2)
    beat += 0.01 * rng.standard_normal(beat.shape)
    beat /= beat.std(axis=1, keepdims=True)
    test_true = (beat - beat.max(axis=1, keepdims=True)).astype(np.float32)

    def at(x_ms):
        idx = np.clip(np.rint(x_ms / 2.0).astype(int) + 106, 0, 319)
        return np.take_along_axis(beat, idx[:, None], axis=1)[:, 0]

    r_amp = at(np.zeros(n_beats))
    # order: P_on, P_pk, P_off, R_on, Q_pk, [R], S_pk, R_off, T_on, T_pk, T_off
    xs = np.stack([
        off[:, 0] - 2.2 * wid[:, 0], off[:, 0], off[:, 0] + 2.2 * wid[:, 0],
        off[:, 1] - 1.8 * wid[:, 1], off[:, 1],
        off[:, 3], off[:, 3] + 1.8 * wid[:, 3],
        off[:, 4] - 2.2 * wid[:, 4], off[:, 4], off[:, 4] + 2.2 * wid[:, 4]], axis=1)
    xs = np.rint(xs / 2.0) * 2.0                                  # fiducials live on the 2 ms sample grid
    ys = np.stack([at(xs[:, j]) for j in range(10)], axis=1) - r_amp[:, None]
    nan_mask = rng.random(xs.shape) < nan_frac
    xs = np.where(nan_mask, np.nan, xs)
    ys = np.where(nan_mask, np.nan, ys)
    features_x = np.insert(xs, 5, 0.0, axis=1)
    features_y = np.insert(ys, 5, -1e-10, axis=1)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            baselines = np.nanmean(features_y[:, [0, 2, 3, 7, 8, 10]], axis=1)
    return test_true, features_x, features_y, baselines


def sigma_candidates(n: int = 4096, mode: str = "integer", seed: int = 1) -> np.ndarray:
    """Latin-hypercube sigma-map candidates.

    ``integer``: int(s*5)+1 in {1..5} (the reference's Integer(1,5) space,
    analysis_performance_traditional_vs_ae.py:58); ``continuous``: 1 + 4 s.
    The reference's sigma_map0 (:55) and final sigma_map (:121) are appended so
    a known-good point is always in the list.
    """
    from scipy.stats import qmc
    s = qmc.LatinHypercube(d=12, seed=seed).random(n)
    if mode == "integer":
        c = np.floor(s * 5).astype(np.int64) + 1
        c = np.clip(c, 1, 5).astype(np.float64)
    else:
        c = 1.0 + 4.0 * s
    extra = np.array([[1, 1, 1, 2, 2, 3, 3, 3, 4, 2, 2, 1],
                      [1, 2, 2, 3, 2, 3, 3, 2, 3, 2, 2, 1]], dtype=np.float64)
    return np.concatenate([c, extra], axis=0)
