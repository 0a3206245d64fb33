"""The CUDA kernel SOURCES, executed on host threads (tests/emul_harness.py), against
the golden vectors frozen from the reference and against the oracle.  CPU only.
The same assertions run on the real GPU in tests/test_gpu_parity.py."""
import warnings

import numpy as np
import pytest

import emul_harness as eh
from conftest import unragged
from oracle import ecg_synth, ref_port as rp

warnings.simplefilter("ignore", RuntimeWarning)


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_k1_stats_bit_exact():
    rng = np.random.default_rng(0)
    lens = [1, 5, 8, 100, 128, 129, 1000, 4999, 10000, 30011]
    recs = [rng.standard_normal(n) * rng.uniform(0.5, 50) + rng.uniform(-5, 5) for n in lens]
    recs[6][3:40] = np.nan
    recs[8][::7] = np.nan
    recs[4][:] = np.nan
    mean, sd, cnt = eh.zscore_stats(recs)
    for r, m, s, c in zip(recs, mean, sd, cnt):
        assert same(m, np.nanmean(r)) and same(s, np.nanstd(r)) and c == (~np.isnan(r)).sum()


def test_k1_one_sweep_statistics_within_rounding():
    """aew_zscore_stats_fast_f64: one sweep, (n, mean, M2) per block around a local shift, merged per record in a
    fixed order: within a few ulp of np.nanmean / np.nanstd, exact counts, an all-NaN record gives NaN as NumPy does,
    a record a long way from zero loses nothing, and two runs give the same bits."""
    rng = np.random.default_rng(1)
    lens = [1, 5, 8, 100, 128, 129, 1000, 4999, 10000, 30011, 200003]
    recs = [rng.standard_normal(n) * rng.uniform(0.5, 50) + rng.uniform(-5, 5) for n in lens]
    recs[6][3:40] = np.nan
    recs[8][::7] = np.nan
    recs[9][0] = np.nan; recs[9][-1] = np.nan; recs[9][15000:15100] = np.nan
    recs[4][:] = np.nan
    recs[10] += 1e6                                          # sd ~ 1e-5 of the mean
    mean, sd, cnt = eh.zscore_stats(recs, fast=True)
    mean2, sd2, _ = eh.zscore_stats(recs, fast=True)
    assert same(mean, mean2) and same(sd, sd2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r, m, s, c in zip(recs, mean, sd, cnt):
            assert c == (~np.isnan(r)).sum()
            wm, ws = np.nanmean(r), np.nanstd(r)
            if np.isnan(wm):
                assert np.isnan(m) and np.isnan(s)
            else:
                assert abs(m - wm) <= 1e-14 * max(abs(wm), ws) and abs(s - ws) <= 1e-13 * ws, (len(r), m, wm, s, ws)


def test_trim_edges():
    a = np.arange(50.0); a[:7] = np.nan; a[-3:] = np.nan
    b = np.full(20, np.nan)
    c = np.arange(600.0); c[:300] = np.nan; c[577:] = np.nan; c[400] = np.nan
    t = eh.trim([a, b, c, np.arange(5.0)])
    assert t.tolist() == [[7, 47], [0, 0], [300, 577], [0, 5]]


def test_k2_exact_mode_is_bit_identical_to_scipy(golden):
    g = golden("filters")
    recs = [g["gaps_z"], g["clean_z"][:9000]]
    outs, hdr = eh.filtfilt(recs, [("high", 1, 4)], mode=1)
    assert same(outs[0], g["gaps_hp"]) and same(outs[1], rp.butterworth(recs[1], "high", 1))
    assert hdr[0] == 8 and hdr[4] == 0
    outs, _ = eh.filtfilt([g["gaps_hp"]], [("low", 30, 4)], mode=1)
    assert same(outs[0], g["gaps_lp"])


def test_k2_exact_warp_form_has_the_same_bits(golden):
    """mode 3 (one warp per segment, the low-latency form of the exact-mask pass) == mode 1 == SciPy."""
    g = golden("filters")
    recs = [g["gaps_z"], g["clean_z"][:9000]]
    outs, hdr = eh.filtfilt(recs, [("high", 1, 4)], mode=3)
    assert same(outs[0], g["gaps_hp"]) and same(outs[1], rp.butterworth(recs[1], "high", 1))
    outs, _ = eh.filtfilt([g["gaps_hp"]], [("low", 30, 4)], mode=3)
    assert same(outs[0], g["gaps_lp"])
    outs, _ = eh.filtfilt([g["short_x"]], [("low", 40, 3)], mode=3)
    assert same(outs[0], g["short_lp40_o3"])


def test_k2_exact_forms_pack_many_segments_per_block():
    """The sequential kernels pack 128 segments (four chain / mover warp pairs with a barrier each, mode 1) resp.
    four segments (one warp each, mode 3) into a block: 150 records of unequal length, some too short to filter,
    come out bit-identical to SciPy from both."""
    rng = np.random.default_rng(11)
    recs = [rng.standard_normal(int(n)) for n in rng.integers(5, 700, 150)]
    want = [rp.butterworth(r, "low", 30.0) for r in recs]
    for mode in (1, 3):
        outs, hdr = eh.filtfilt(recs, [("low", 30, 4)], mode=mode)
        assert hdr[0] == len(recs) and hdr[4] == 0
        for k, (o, w) in enumerate(zip(outs, want)):
            assert same(o, w), (mode, k, len(w))
    # z-scored on the way in (first stage of the exact-mask pass): the chain warp z-scores a round's samples in bulk
    raw = [r * 3.0 + 0.7 for r in recs[:40] if len(r) > 40]
    wantz = [rp.butterworth(rp.zscore(r), "high", 1.0) for r in raw]
    for mode in (1, 3):
        outs, _ = eh.filtfilt(raw, [("high", 1, 4)], mode=mode, zscore=True)
        for k, (o, w) in enumerate(zip(outs, wantz)):
            assert same(o, w), ("zscore", mode, k, len(w))


def test_k2_one_launch_form_within_tolerance(golden):
    """mode 2: forward + backward of a short-memory filter in one launch over tiles with a 416-sample halo (segment
    starts / ends inside a region, segments shorter than a tile, many tiles, NaN gaps, short islands)."""
    g = golden("filters")
    x = ecg_synth.synth_case(5, 30000)
    hp = rp.butterworth(rp.zscore(x), "high", 1)
    recs = [g["gaps_hp"], hp, hp[:3328 + 416], hp[:3328 + 417], hp[:4160], hp[:4161], hp[:2 * 3328], hp[:16], hp[:15]]
    outs, hdr = eh.filtfilt(recs, [("low", 30, 4)], mode=2)
    assert hdr[4] == 0
    for out, r in zip(outs, recs):
        ref = rp.butterworth(r, "low", 30)
        assert same(np.isnan(out), np.isnan(ref))
        m = ~np.isnan(ref)
        assert np.max(np.abs(out[m] - ref[m])) <= 1e-7, (len(r), np.max(np.abs(out[m] - ref[m])))
    outs, _ = eh.filtfilt([g["short_x"]], [("low", 40, 3)], mode=2)
    assert np.max(np.abs(outs[0] - g["short_lp40_o3"])) <= 1e-6 * np.max(np.abs(g["short_lp40_o3"]))
    L = eh.lib()
    tab, _ = eh.filter_tables("high", 1, 500, 4)
    assert L.aew_filt_fused_ok(eh.ptr(tab)) == 0           # the 1 Hz high-pass remembers ~8,000 samples: two launches


def test_k2_tiled_scan_within_tolerance(golden):
    g = golden("filters")
    recs = [g["gaps_x"], g["clean_x"][:17000]]
    outs, hdr = eh.filtfilt(recs, [("high", 1, 4), ("low", 30, 4)], mode=0, zscore=True)
    for out, ref in zip(outs, (g["gaps_lp"], rp.butterworth(rp.butterworth(rp.zscore(recs[1]), "high", 1), "low", 30))):
        assert same(np.isnan(out), np.isnan(ref))
        m = ~np.isnan(ref)
        assert np.max(np.abs(out[m] - ref[m])) <= 1e-6 * np.max(np.abs(ref[m]))   # north_star tolerance
        assert np.max(np.abs(out[m] - ref[m])) <= 1e-7                              # what is actually achieved
    assert hdr[4] == 0


def test_k2_runs_of_tiles_and_their_warm_ups(monkeypatch):
    """k2_pass hands a block RUNS of a segment's tiles and starts a run behind a warm-up over the `lookback` tiles in
    front of it.  With short runs (AEW_K2_RUN = 1, 3: every tile resp. every third starts from a warm-up, chunks of
    the table that hold the tail of another block's run) the output stays within the tolerance, and a record's output
    does not depend on what else is in the batch (runs are cut relative to the segment)."""
    x = rp.zscore(ecg_synth.synth_case(3, 24000))
    y = rp.zscore(ecg_synth.synth_case(4, 26000)); y[11000:11010] = np.nan; y[23000] = np.nan
    z = np.random.default_rng(5).standard_normal(12)
    recs = [x, y, z]
    for run, st in (("1", ("low", 30, 4)), ("3", ("high", 1, 4))):
        monkeypatch.setenv("AEW_K2_RUN", run)
        if True:
            outs, hdr = eh.filtfilt(recs, [st], mode=0)
            assert hdr[4] == 0
            for r, o in zip(recs, outs):
                ref = rp.butterworth(r, st[0], st[1])
                assert same(np.isnan(o), np.isnan(ref))
                m = ~np.isnan(ref)
                assert np.max(np.abs(o[m] - ref[m])) <= 1e-7
            alone, _ = eh.filtfilt([y], [st], mode=0)
            assert same(alone[0], outs[1])


def test_k2_runs_across_many_short_segments(monkeypatch):
    """Chunks of the tile table that hold several segments, runs that end at a segment's end or reach into the next
    chunk, tiles of 1 .. 1,665 samples (AEW_K2_RUN = 2), NaN gaps: every sample comes out, within the tolerance."""
    rng = np.random.default_rng(21)
    lens = list(rng.integers(5, 7000, 24)) + [1664, 1663, 1665, 3328, 3329, 4992, 16, 15, 1]
    recs = [rng.standard_normal(int(n)) for n in lens]
    big = int(np.argmax(lens))
    recs[big][lens[big] // 3:lens[big] // 3 + 7] = np.nan
    recs[big][lens[big] // 2] = np.nan
    monkeypatch.setenv("AEW_K2_RUN", "2")
    outs, hdr = eh.filtfilt(recs, [("high", 1, 4)], mode=0)
    assert hdr[4] == 0
    for r, o in zip(recs, outs):
        ref = rp.butterworth(r, "high", 1)
        assert same(np.isnan(o), np.isnan(ref))
        m = ~np.isnan(ref)
        assert not m.any() or np.max(np.abs(o[m] - ref[m])) <= 1e-7


def test_k2_other_orders(golden):
    g = golden("filters")
    xs = g["short_x"]
    for key, st in (("short_hp05_o2", ("high", 0.5, 2)), ("short_lp40_o3", ("low", 40, 3)), ("short_lp15_o1", ("low", 15, 1))):
        exact, _ = eh.filtfilt([xs], [st], mode=1)
        assert same(exact[0], g[key])
        fast, _ = eh.filtfilt([xs], [st], mode=0)
        assert np.max(np.abs(fast[0] - g[key])) <= 1e-6 * np.max(np.abs(g[key]))


def test_k3_e2e_golden(golden):
    g = golden("e2e")
    nw = int(g["n_windows"])
    for mode in (1, 0):
        o = eh.screen(g["filtered"], np.arange(nw) * 10000, mode=mode)
        kept = np.flatnonzero(o["keep"])
        assert same((kept + 1) * 10000, g["kept_idx_end"])
        for got, want in zip([o["peaks"][i, :o["n_peaks"][i]] for i in kept], unragged(g["kept_peaks"], g["kept_peaks_off"])):
            assert same(got, want)
        assert same(o["median"][kept], g["medians"]) and same(o["waves"][kept], g["np_waves"])
        assert o["status"][0] == 0


def test_k3_screen_golden(golden):
    g = golden("screen")
    raw = g["windows_unscaled"]
    sel = [0, 4, 6, 8, 10, 12, 13]                     # one window of every fate
    o = eh.screen(raw[sel].reshape(-1), np.arange(len(sel)) * 10000, mode=0, want_z=True)
    want_peaks = unragged(g["all_peaks"], g["all_peaks_off"])
    for j, w in enumerate(sel):
        assert same(o["peaks"][j, :o["n_peaks"][j]], want_peaks[w])
        assert bool(o["keep"][j]) == (w in g["kept_rows"])
        if not np.isnan(g["all_sd"][w]):
            assert same(o["zwin"][j], g["windows"][w])
            assert same(o["stats"][j, :3], [g["all_hr"][w], g["all_rmssd"][w], g["all_sd"][w]])
        else:
            assert o["has_nan"][j] == 1
    # pre-scaled windows, as the remove_samples() drop-in passes them
    o2 = eh.screen(g["windows"][sel].reshape(-1), np.arange(len(sel)) * 10000, mode=0, zscore_windows=False)
    assert same(o2["keep"], o["keep"]) and same(o2["peaks"], o["peaks"])


def test_k3_window_comes_back_after_the_in_place_prefix_sums(golden):
    """The indicator's prefix sums overwrite the z-scored window; it returns from the parked copy (one bulk copy) or
    is formed again from the filter output.  mode 3 runs the prefix passes and THEN the sequential evaluation on the
    restored samples; np.std at the end of the kernel sweeps them once more.  Both restore paths, an odd window
    length (no 16-byte granules: plain copies), pre-scaled windows: everything equals the sequential mode."""
    g = golden("e2e")
    nw = min(int(g["n_windows"]), 6)
    y = g["filtered"]
    ref = eh.screen(y, np.arange(nw) * 10000, mode=1, want_z=True)
    for want_z in (True, False):
        for mode in (0, 3):
            o = eh.screen(y, np.arange(nw) * 10000, mode=mode, want_z=want_z)
            for k in ("keep", "n_peaks", "peaks", "stats", "median", "waves"):
                assert same(o[k], ref[k]), (want_z, mode, k)
            if want_z:
                assert same(o["zwin"], ref["zwin"])
    Lw = 9999
    ref = eh.screen(y, np.arange(3) * Lw, Lw=Lw, mode=1, want_z=True)
    for mode in (0, 3):
        o = eh.screen(y, np.arange(3) * Lw, Lw=Lw, mode=mode, want_z=True)
        for k in ("keep", "n_peaks", "peaks", "stats", "median", "zwin"):
            assert same(o[k], ref[k]), (Lw, mode, k)
    pre = eh.screen(ref["zwin"].reshape(-1), np.arange(3) * Lw, Lw=Lw, mode=3, zscore_windows=False, want_z=True)
    assert same(pre["keep"], ref["keep"]) and same(pre["peaks"], ref["peaks"]) and same(pre["median"], ref["median"])


def test_k3_lazy_sd_takes_the_same_decisions():
    """lazy_sd: np.std is evaluated only where n_peaks * 0.05 == 1; every keep / remove decision, the peaks and the
    median beats are those of the full evaluation, sd is reported only for the 20-peak windows."""
    x = ecg_synth.synth_case(7, 8 * 10000 + 5, hr_bpm=60.0)          # 19 / 20 / 21 peaks per window
    y = rp.butterworth(rp.butterworth(rp.zscore(x), "high", 1.0), "low", 30.0)
    full = eh.screen(y, np.arange(8) * 10000)
    lazy = eh.screen(y, np.arange(8) * 10000, lazy_sd=1)
    assert same(full["keep"], lazy["keep"]) and same(full["peaks"], lazy["peaks"])
    assert same(full["stats"][:, :2], lazy["stats"][:, :2])
    k = full["keep"].astype(bool)
    assert same(full["median"][k], lazy["median"][k])
    n20 = full["n_peaks"] == 20
    assert n20.any() and (~n20).any()
    assert same(lazy["stats"][n20, 2:], full["stats"][n20, 2:]) and np.isnan(lazy["stats"][~n20, 2:]).all()
    want = rp.window_screen(np.stack([rp.zscore(y[i * 10000:(i + 1) * 10000]) for i in range(8)]))
    assert same(lazy["keep"].astype(bool), want["keep"])


def test_k3_zscore_sweep_general_form():
    """Samples that equal the window mean exactly (zero quotients) send a warp of the z-score sweep through its
    checked form after the branch-free one has already overwritten part of the window: same bits as NumPy."""
    x = ecg_synth.synth_case(9, 30000)
    y = rp.butterworth(rp.butterworth(rp.zscore(x), "high", 1.0), "low", 30.0)
    half = y[10000:15000].copy()
    half[[7, 1234, 4999]] = 0.0
    win = np.concatenate([half, -half])                   # NumPy's pairwise sum of this is exactly 0
    assert np.add.reduce(win) == 0.0
    flat = np.concatenate([win, y[:10000]])
    o = eh.screen(flat, np.array([0, 10000]), want_z=True)
    want = rp.window_screen(np.stack([rp.zscore(win), rp.zscore(y[:10000])]))
    assert same(o["zwin"][0], rp.zscore(win)) and same(o["zwin"][1], rp.zscore(y[:10000]))
    assert same(o["keep"].astype(bool), want["keep"])
    for j in range(2):
        assert same(o["peaks"][j, :o["n_peaks"][j]], want["peaks"][j])
        assert same(o["stats"][j, :3], [want["hr"][j], want["rmssd"][j], want["sd"][j]])


def test_k3_other_window_lengths():
    """The segmented box filters, the bit-packed indicator and the run tables at window lengths / sampling rates
    other than 10,000 @ 500 Hz, against the oracle (fast and exact evaluation)."""
    for fs, Lw, seed in ((250, 5000, 3), (360, 7200, 4), (500, 6000, 5), (100, 2000, 6), (500, 10240, 7)):
        x = ecg_synth.synth_case(seed, 4 * Lw + 17)
        y = rp.butterworth(rp.butterworth(rp.zscore(x), "high", 1.0, fs), "low", 30.0, fs)
        wins = np.stack([rp.zscore(y[k * Lw:(k + 1) * Lw]) for k in range(4)])
        want = rp.window_screen(wins, fs=fs)
        for mode in (0, 1):
            o = eh.screen(y[:4 * Lw], np.arange(4) * Lw, Lw=Lw, mode=mode, fs=fs)
            for j in range(4):
                assert same(o["peaks"][j, :o["n_peaks"][j]], want["peaks"][j]), (fs, Lw, mode, j)
                assert bool(o["keep"][j]) == bool(want["keep"][j])
                assert same(o["stats"][j, :3], [want["hr"][j], want["rmssd"][j], want["sd"][j]])
                if want["keep"][j]:                        # the section scales with the rate (median_sample.py:20-25)
                    m = rp.compute_median_sample_ecg(wins[j], want["peaks"][j], fs)
                    assert m.shape == (int(0.8 * (60 / 75) * fs),) and same(o["median"][j], m), (fs, Lw, mode, j)


def test_k4_median_golden(golden):
    g = golden("median")
    for name in ("edge", "even", "odd", "one", "far_right"):
        med, al = eh.median_beats(g["window"], g[f"{name}_peaks"])
        assert same(med, g[f"{name}_median"]) and same(al, g[f"{name}_aligned"])


def test_k4_median_all_beat_counts(golden):
    """One thread per column (<= 32 beats), lane pairs (33..64) and the counting fallback (> 64) against the oracle,
    with edge peaks (short, left-aligned slices) and ties (a coarsely quantised window)."""
    g = golden("median")
    rng = np.random.default_rng(3)
    quant = np.round(g["window"] * 4) / 4
    for n in (2, 31, 32, 33, 34, 47, 63, 64, 65, 70):
        for win in (g["window"], quant):
            peaks = np.sort(rng.choice(np.arange(120, len(win) - 230), n - 2, replace=False))
            peaks = np.concatenate([[40], peaks, [len(win) - 30]])          # both edges clipped
            med, _ = eh.median_beats(win, peaks, want_aligned=False)
            assert same(med, rp.compute_median_sample_ecg(win, list(peaks), 500)), n


def test_k5_recon_golden(golden):
    g = golden("recon")
    for k, smap in enumerate(g["sigma_maps"]):
        _, pred = eh.recon(g["fx"], g["fy"], g["baselines"], g["test_true"], g["x_t"], smap, want_pred=True)
        ref = g["recon"][k]
        assert same(np.isnan(pred), np.isnan(ref))
        assert np.nanmax(np.abs(pred - ref)) <= 1e-6 * np.nanmax(np.abs(ref))
    sse, _ = eh.recon(g["fx"], g["fy"], g["baselines"], g["test_true"], g["x_t"], g["sigma_maps"])
    want = np.array([rp.sigma_objective(g["fx"], g["fy"], g["baselines"], g["test_true"], m, g["x_t"])[0] for m in g["sigma_maps"]])
    assert np.all(np.abs(sse - want) <= 1e-6 * np.abs(want))
    mse = sse / g["test_true"].size
    assert np.all(np.abs(mse - g["mse_torch_f32"]) <= 1e-5 * g["mse_torch_f32"])
    assert np.argmin(sse) == np.argmin(want)


def test_k5_grid_tables_match_bruteforce(golden):
    """Separable tables for a discrete sigma grid == the per-candidate kernel == the oracle objective."""
    g = golden("recon")
    rng = np.random.default_rng(5)
    vals = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    idx = rng.integers(0, 5, size=(24, 12)).astype(np.uint8)
    idx[0], idx[1] = 0, 4
    cands = vals[idx]
    grid = eh.recon_grid(g["fx"], g["fy"], g["baselines"], g["test_true"], g["x_t"], vals, idx)
    brute, _ = eh.recon(g["fx"], g["fy"], g["baselines"], g["test_true"], g["x_t"], cands)
    assert same(np.isnan(grid), np.isnan(brute))
    assert np.all(np.abs(grid - brute) <= 1e-12 * np.abs(brute))
    want = np.array([rp.sigma_objective(g["fx"], g["fy"], g["baselines"], g["test_true"], m, g["x_t"])[0] for m in cands[:6]])
    assert np.all(np.abs(grid[:6] - want) <= 1e-6 * np.abs(want))
    # non-integer grids work the same way (three values)
    vals3 = np.array([0.75, 2.5, 6.0])
    idx3 = rng.integers(0, 3, size=(8, 12)).astype(np.uint8)
    grid3 = eh.recon_grid(g["fx"], g["fy"], g["baselines"], g["test_true"], g["x_t"], vals3, idx3)
    brute3, _ = eh.recon(g["fx"], g["fy"], g["baselines"], g["test_true"], g["x_t"], vals3[idx3])
    assert np.all(np.abs(grid3 - brute3) <= 1e-12 * np.abs(brute3))


def _random_fiducials(rng, peaks, L, p_nan=0.15):
    """plausible fiducial indices around each R peak, with missing entries"""
    offs = np.array([-100, -125, -75, -12, -20, 22, 15, 125, 90, 170])
    f = np.asarray(peaks)[None, :] + offs[:, None] + rng.integers(-6, 7, size=(10, len(peaks)))
    f = np.clip(f, 1, L - 1).astype(np.float64)
    f[rng.random(f.shape) < p_nan] = np.nan
    return f


def _oracle_row(window, peaks, f):
    fid = {k: [np.nan if np.isnan(v) else int(v) for v in f[i]] for i, k in enumerate(rp.FIDUCIALS)}
    return rp.features_row_from_fiducials(window, [int(p) for p in peaks], fid)


def test_k6_features_golden_and_oracle(golden):
    g = golden("features")
    gf = np.where(g["fiducials"] < 0, np.nan, g["fiducials"])       # the fixture stores np.nan as -1
    rows, st = eh.fiducial_features(g["window"][None, :], [g["r_peaks"]], [gf])
    assert st == 0 and same(rows[0], g["row35"])           # frozen from the reference's own three functions
    rng = np.random.default_rng(5)
    wins, pks, fids = [], [], []
    for n_beats in (1, 2, 3, 10, 24, 33, 47, 64, 72):
        w = rng.standard_normal(10000)
        p = np.sort(rng.choice(np.arange(150, 9800), n_beats, replace=False))
        f = _random_fiducials(rng, p, 10000, p_nan=0.3 if n_beats % 2 else 0.1)
        if n_beats == 10:
            f[1, :] = np.nan                               # an all-NaN fiducial: NaN medians, nanmax of one side
            f[4, ::2] = f[4, 1]                            # ties
        wins.append(w); pks.append(p); fids.append(f)
    rows, st = eh.fiducial_features(np.stack(wins), pks, fids, valid=[1] * 8 + [0])
    assert st == 0
    for i in range(8):
        assert same(rows[i], _oracle_row(wins[i], pks[i], fids[i])), i
    assert np.isnan(rows[8]).all()                         # delineation failed: NaN row, window kept


def test_k6_stage_entries_golden(golden):
    """The per-call drop-ins' kernels against the outputs of the reference's own three functions."""
    g = golden("features")
    gf = np.where(g["fiducials"] < 0, np.nan, g["fiducials"])
    x, y, morph, med = eh.feature_stages(g["window"], g["r_peaks"], gf)
    assert same(x, g["rel_x"]) and same(y, g["rel_y"]) and same(morph, g["morph"]) and same(med, g["row35"])
    rng = np.random.default_rng(11)
    for n in (1, 2, 5, 37, 90, 200):                       # any number of beats, NaN-heavy
        w = rng.standard_normal(10000)
        p = np.sort(rng.choice(np.arange(150, 9800), n, replace=False))
        f = _random_fiducials(rng, p, 10000, p_nan=0.4)
        _, _, _, med = eh.feature_stages(w, p, f)
        assert same(med, _oracle_row(w, p, f)), n


def test_k6_peak_stats_match_the_reference_formulas():
    import warnings
    rng = np.random.default_rng(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n in (0, 1, 2, 3, 4, 9, 26, 71, 140, 300):
            hi = np.cumsum(rng.integers(160, 700, n)) if n else np.zeros(0, np.int64)
            lo = np.cumsum(rng.integers(160, 700, max(0, n - 1))) if n > 1 else np.zeros(0, np.int64)
            for mode, mod in ((0, "min"), (1, "mean"), (2, "max")):
                got = eh.peak_stats(lo, hi, 500, mode)
                assert same(got[0], rp.estimate_hr(lo, hi, 500, mod)), (n, mod)
                assert same(got[1], rp.rmssd(hi, 500)), n


def _delineation_cases():
    """(windows, peak lists): three heart rates (wavelet scales 2+d, 3+d for d = 1, 2, 3), too few peaks
    (the library raises -> NaN row), a peak next to the window start (negative slice bounds -> dropped
    window) and next to its end (entries >= L blanked)."""
    wins, pks = [], []
    for cid, hr in ((12, 52.0), (14, 95.0), (15, 140.0)):
        out = rp.process_case(ecg_synth.synth_case(cid, 20001, hr_bpm=hr), cid)
        wins.append(out["windows"][0]); pks.append([int(p) for p in out["peaks"][0]])
    wins.append(wins[0]); pks.append(pks[0][:3])
    wins.append(wins[1]); pks.append([20] + [p for p in pks[1] if p > 300])
    wins.append(wins[0]); pks.append([p for p in pks[0] if p < 9700] + [9990])
    return wins, pks


def test_k7_delineation_matches_restatement():
    from oracle import nk_delineate as nd
    wins, pks = _delineation_cases()
    fid, flags = eh.delineate(np.stack(wins), pks, select=[1, 1, 1, 1, 1, 1])
    want_flag = {"ok": 0, "nan": 1, "drop": 2}
    seen = set()
    for i, (w, p) in enumerate(zip(wins, pks)):
        st, ref = nd.window_fiducials(w, p)
        seen.add(st)
        assert flags[i] == want_flag[st], (i, st, flags[i])
        if st == "ok":
            got = fid[i, :, :len(p)].astype(np.float64)
            got[got == eh.FID_NAN] = np.nan
            assert same(got, ref), i                       # every fiducial index identical
            rows, status = eh.fiducial_features(w[None, :], [p], [got])
            assert status == 0 and same(rows[0], rp.window_features(w, p)[1])
    assert seen == {"ok", "nan", "drop"}
    _, fl = eh.delineate(np.stack(wins[:2]), pks[:2], select=[0, 1])
    assert fl.tolist() == [3, 0]                           # unselected windows are left alone


def test_k8_fetch_host_copies_bytes():
    rng = np.random.default_rng(81)
    for n in (0, 1, 15, 16, 17, 808, 4099, 300000):
        raw = rng.integers(0, 256, n).astype(np.uint8)
        assert np.array_equal(eh.fetch_host(raw), raw)
    t = rng.standard_normal((1790, 2))
    assert np.array_equal(eh.fetch_host(t).view(np.float64).reshape(t.shape), t)


def test_k8_widen_matches_numpy():
    rng = np.random.default_rng(8)
    for n in (0, 1, 7, 8, 1000, 4099):
        x = (rng.standard_normal(n) * 3).astype(np.float32)
        if n > 5:
            x[3] = np.nan
        assert same(eh.widen(x), x.astype(np.float64))
        v = rng.integers(-32768, 32767, n).astype(np.int16)
        if n > 5:
            v[2] = -32768
        want = v.astype(np.float64) * 0.00244 + (-1.25)
        want[v == -32768] = np.nan
        assert same(eh.widen(v, 0.00244, -1.25, sentinel=-32768), want)
        assert same(eh.widen(v, 0.5, 0.0), v.astype(np.float64) * 0.5 + 0.0)


def test_k9_beat_metrics_match_torch():
    rng = np.random.default_rng(9)
    med = rng.standard_normal((37, 320)).cumsum(axis=1) * 0.1
    y = med - med.max(axis=1, keepdims=True) + 0.05 * rng.standard_normal((37, 320))
    y[5, 17] = np.nan                                     # a NaN reconstruction poisons that beat only
    rmse, prd, qs = eh.beat_metrics(y, med)
    w_rmse, w_prd, w_qs = rp.beat_metrics(y, med)
    for got, want in ((rmse, w_rmse), (prd, w_prd), (qs, w_qs)):
        assert same(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.all(np.abs(got[ok] - want[ok]) <= 1e-5 * np.abs(want[ok]))     # float32 outputs: SURVEY.md C.5


def _other_geometry_windows():
    """10 s strips at 500 Hz (the configs[2] adaptation, L = 5,000) and 20 s windows at 250 Hz."""
    from oracle import nk_peaks
    case = ecg_synth.synth_case(31, 40000, hr_bpm=80.0)
    f = rp.butterworth(rp.butterworth(rp.zscore(case), "high", 1, 500, 4), "low", 30, 500, 4)
    out = []
    for sig, fs in ((f[10000:15000], 500), (f[::2][5000:10000], 250)):
        w = rp.zscore(sig)
        p = [int(v) for v in nk_peaks.ecg_findpeaks(w, fs)["ECG_R_Peaks"] if 0 < v < len(w)]
        out.append((w, p, fs))
    return out


def test_k7_other_window_lengths_and_rates():
    from oracle import nk_delineate as nd
    for w, p, fs in _other_geometry_windows():
        fid, flags = eh.delineate(w[None, :], [p], fs=float(fs))
        st, ref = nd.window_fiducials(w, p, fs)
        assert st == "ok" and flags[0] == 0
        got = fid[0, :, :len(p)].astype(np.float64)
        got[got == eh.FID_NAN] = np.nan
        assert same(got, ref), fs


def test_k3_fast_stats_agrees_with_the_numpy_order_chain(golden):
    """fast_stats: window mean / sd from one pass of shifted sums, R peaks from the unscaled samples, the median beat
    z-scored after the column medians, nothing parked.  Against the bit-exact chain on the same windows: the same
    peaks, decisions and integer-derived statistics; mean / sd / medians to rounding (1e-12); NaN and constant
    windows removed; np.std reported for the 20-beat windows only (lazy_sd)."""
    g = golden("e2e")
    nw = int(g["n_windows"])
    y = g["filtered"]
    ref = eh.screen(y, np.arange(nw) * 10000, lazy_sd=1)
    for mode in (0, 3):                                     # 3: prefix passes, then the sequential evaluation
        o = eh.screen(y, np.arange(nw) * 10000, mode=mode, lazy_sd=1, fast_stats=1)
        for k in ("keep", "has_nan", "n_peaks", "peaks"):
            assert same(o[k], ref[k]), (mode, k)
        assert same(o["stats"][:, :2], ref["stats"][:, :2])
        live = ref["has_nan"] == 0
        assert np.allclose(o["zparams"][live], ref["zparams"][live], rtol=1e-12, atol=1e-14)
        k = ref["keep"].astype(bool)
        assert k.any()
        assert np.max(np.abs(o["median"][k] - ref["median"][k])) < 1e-12
        assert np.max(np.abs(o["waves"][k] - ref["waves"][k])) < 1e-6
        assert o["status"][0] == 0
    # 20-beat windows (np.std decides), a window with a NaN, a constant window, a window far from zero
    x = ecg_synth.synth_case(7, 8 * 10000 + 5, hr_bpm=60.0)
    y = rp.butterworth(rp.butterworth(rp.zscore(x), "high", 1.0), "low", 30.0)[:80000].copy()
    y[10000 + 4321] = np.nan
    y[20000:30000] = 0.25
    y[30000:40000] += 1000.0
    ref = eh.screen(y, np.arange(8) * 10000, lazy_sd=1)
    o = eh.screen(y, np.arange(8) * 10000, lazy_sd=1, fast_stats=1)
    assert list(o["has_nan"][:4]) == [0, 1, 1, 0] and same(o["has_nan"], ref["has_nan"])
    assert same(o["n_peaks"], ref["n_peaks"]) and same(o["peaks"], ref["peaks"])
    n20 = (ref["n_peaks"] == 20) & (ref["has_nan"] == 0)
    assert n20.any()
    # where np.std's last bit decides, the beat kernel runs the NumPy-order chain after all: same value, same decision
    assert same(o["stats"], ref["stats"]) and np.isnan(o["stats"][~n20, 2]).all()
    assert same(o["keep"], ref["keep"])
    assert np.allclose(o["zparams"][3], ref["zparams"][3], rtol=1e-11)       # the window shifted by 1000
    k = (ref["keep"] & o["keep"]).astype(bool)
    assert np.max(np.abs(o["median"][k] - ref["median"][k])) < 1e-11
    # the entry refuses what fast_stats does not serve
    with pytest.raises(AssertionError):
        eh.screen(y, np.arange(8) * 10000, lazy_sd=0, fast_stats=1)
    with pytest.raises(AssertionError):
        eh.screen(y, np.arange(8) * 10000, mode=1, lazy_sd=1, fast_stats=1)


def test_k3_median_from_keys_is_exact_with_ties_and_edge_beats():
    """k3_median_keys sorts 32-bit keys (high words + slot) and orders only the four middle values as doubles; a
    window whose middle values agree in their keys goes back to the general kernel.  Coarsely quantised windows
    (many exactly equal samples), the same plus 1e-10 noise (equal keys, different doubles: the hand-back), window
    starts that cut the first / last beat short: every median equals the oracle's bit for bit."""
    rng = np.random.default_rng(11)
    x = ecg_synth.synth_case(21, 6 * 10000 + 700, hr_bpm=71.0)
    y = rp.butterworth(rp.butterworth(rp.zscore(x), "high", 1.0), "low", 30.0)
    x2 = ecg_synth.synth_case(22, 6 * 10000 + 700, hr_bpm=121.0)       # ~40 beats per window: the 64-slot instance
    y2 = rp.butterworth(rp.butterworth(rp.zscore(x2), "high", 1.0), "low", 30.0)
    for name, sig in (("plain", y), ("quantised", np.round(y * 64) / 64),
                      ("quantised + noise", np.round(y * 64) / 64 + 1e-10 * rng.standard_normal(len(y))),
                      ("fast", y2), ("fast, quantised + noise", np.round(y2 * 64) / 64 + 1e-10 * rng.standard_normal(len(y2)))):
        for shift in (0, 137, 333, 611) if not name.startswith("fast") else (0, 211):   # moves the beats against the window edges
            starts = shift + np.arange(6) * 10000
            o = eh.screen(sig, starts, lazy_sd=0)
            wins = np.stack([rp.zscore(sig[s:s + 10000]) for s in starts])
            want = rp.window_screen(wins)
            checked = 0
            for j in range(6):
                assert same(o["peaks"][j, :o["n_peaks"][j]], want["peaks"][j]), (name, shift, j)
                if o["keep"][j]:
                    m = rp.compute_median_sample_ecg(wins[j], want["peaks"][j])
                    assert same(o["median"][j], m), (name, shift, j)
                    checked += 1
            assert checked >= 3, (name, shift)
            if name.startswith("fast"):
                assert (o["n_peaks"][o["keep"].astype(bool)] > 32).all()
