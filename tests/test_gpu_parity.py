"""Parity of the CUDA path (through the C ABI, on a real B200) with the oracle, the golden
vectors frozen from the reference, and size-independent properties.  Bar: integer / index
outputs and NumPy-ordered reductions bit-exact; filtered signals, medians under the fast
filter and reconstruction errors within 1e-6 relative (north_star)."""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import unragged
from oracle import ecg_synth, ref_port as rp

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore", RuntimeWarning)
RTOL = 1e-6          # north_star tolerance for float64 outputs of the fast path


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if not same(np.isnan(a), np.isnan(b)):
        return False
    m = ~np.isnan(b)
    return (not m.any()) or np.max(np.abs(a[m] - b[m])) <= rtol * max(np.max(np.abs(b[m])), 1e-300)


@pytest.fixture(scope="module")
def gpu():
    import torch
    from ae_waveforms_b200 import _cabi
    from ae_waveforms_b200 import batched
    _cabi.require_cuda()
    return torch, batched


# ------------------------------------------------------------------------- K1
def test_zscore_stats_bit_exact(gpu):
    torch, B = gpu
    rng = np.random.default_rng(0)
    lens = [1, 5, 8, 100, 128, 129, 1000, 4999, 10000, 30011, 300001, 1800000]
    recs = [rng.standard_normal(n) * rng.uniform(0.5, 50) + rng.uniform(-5, 5) for n in lens]
    recs[6][3:40] = np.nan
    recs[9][::7] = np.nan
    recs[4][:] = np.nan
    batch = B.RecordBatch.from_host(recs)
    pre = B.Preprocessor()
    beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
    mean, sd, cnt = pre.zscore_stats(batch.flat, beg_h, end_h, d_beg, d_end)
    mean, sd, cnt = mean.cpu().numpy(), sd.cpu().numpy(), cnt.cpu().numpy()
    for r, m, s, c in zip(recs, mean, sd, cnt):
        assert same(m, np.nanmean(r)) and same(s, np.nanstd(r)) and c == (~np.isnan(r)).sum()


def test_zscore_dropin(gpu, golden):
    from ae_waveforms_b200.data_load_preprocess.filters import zscore
    g = golden("filters")
    for tag in ("clean", "gaps"):
        out = zscore(g[f"{tag}_x"])
        assert same(out, g[f"{tag}_z"]) and out.flags.writeable and out.flags.c_contiguous


# ------------------------------------------------------------------------- K2
def test_butterworth_dropin_exact_mode_bit_identical(gpu, golden):
    from ae_waveforms_b200.data_load_preprocess import filters
    g = golden("filters")
    saved, filters.FILTER_MODE = filters.FILTER_MODE, "exact"
    try:
        for tag in ("clean", "gaps"):
            hp = filters.butterworth(g[f"{tag}_z"], "high", 1, 500, 4)
            assert same(hp, g[f"{tag}_hp"])
            assert same(filters.butterworth(hp, "low", 30, 500, 4), g[f"{tag}_lp"])
        assert same(filters.butterworth(g["short_x"], "high", 0.5, 500, 2), g["short_hp05_o2"])
        assert same(filters.butterworth(g["short_x"], "low", 40, 500, 3), g["short_lp40_o3"])
        assert same(filters.butterworth(g["short_x"], "low", 15, 500, 1), g["short_lp15_o1"])
    finally:
        filters.FILTER_MODE = saved


def test_butterworth_dropin_fast_mode(gpu, golden, monkeypatch):
    from ae_waveforms_b200.data_load_preprocess import filters
    from ae_waveforms_b200.data_load_preprocess.filters import butterworth
    monkeypatch.setattr(filters, "FILTER_MODE", "fast")
    g = golden("filters")
    for tag in ("clean", "gaps"):
        hp = butterworth(g[f"{tag}_z"], "high", 1)
        assert close(hp, g[f"{tag}_hp"])
        assert close(butterworth(g[f"{tag}_hp"], "low", 30), g[f"{tag}_lp"])
    assert close(butterworth(g["short_x"], "high", 0.5, 500, 2), g["short_hp05_o2"])
    assert close(butterworth(g["short_x"], "low", 40, 500, 3), g["short_lp40_o3"])
    assert close(butterworth(g["short_x"], "low", 15, 500, 1), g["short_lp15_o1"])
    f32 = butterworth(g["short_x"].astype(np.float32), "low", 30)
    assert f32.dtype == np.float32                                  # appendix A.1: output dtype = input dtype


def test_filter_ragged_batch_long_records(gpu):
    """Many tiles per segment (several runs with their warm-ups), NaN gaps, short islands, both filters."""
    torch, B = gpu
    recs = [ecg_synth.synth_case(40, 1800000), ecg_synth.synth_case(41, 250000, nan_gaps=3),
            ecg_synth.synth_case(42, 8448 * 3), ecg_synth.synth_case(43, 8449), ecg_synth.synth_case(44, 8447),
            ecg_synth.synth_case(45, 16), ecg_synth.synth_case(46, 15), ecg_synth.synth_case(47, 33000, nan_gaps=2)]
    # lengths around the tile (1,664 samples) and the run (64 tiles = 106,496 samples) of the tiled pass: a run's last
    # tile one sample long, a run that is all warm-up but one tile, segments ending exactly on either boundary
    recs += [ecg_synth.synth_case(50 + i, n) for i, n in enumerate((1663, 1664, 1665, 3 * 1664, 106495, 106496, 106497,
                                                                    106496 + 1664, 2 * 106496 + 1, 3 * 106496))]
    batch = B.RecordBatch.from_host(recs)
    for mode in ("fast", "exact"):
        pre = B.Preprocessor(filter_mode=mode)
        beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
        mean, sd, cnt = pre.zscore_stats(batch.flat, beg_h, end_h, d_beg, d_end)
        y, hdr = pre.filter_records(batch.flat, beg_h, end_h, d_beg, d_end, mean, sd, cnt)
        y = y.cpu().numpy()
        assert int(hdr[4].item()) == 0
        for r, s, e in zip(recs, batch.rec_off[:-1], batch.rec_off[1:]):
            ref = rp.butterworth(rp.butterworth(rp.zscore(r), "high", 1), "low", 30)
            assert same(y[s:e], ref) if mode == "exact" else close(y[s:e], ref)


def test_filter_properties_full_length(gpu):
    """Linearity and time-reversal symmetry of the zero-phase filter at case length."""
    torch, B = gpu
    n = 1800000
    rng = np.random.default_rng(9)
    x1, x2 = ecg_synth.synth_case(50, n), rng.standard_normal(n)
    recs = [x1, x2, 0.75 * x1 - 2.5 * x2, x1[::-1].copy()]
    batch = B.RecordBatch.from_host(recs)
    pre = B.Preprocessor()
    beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
    y, _ = pre.filter_records(batch.flat, beg_h, end_h, d_beg, d_end)
    y = y.cpu().numpy().reshape(4, n)
    scale = np.abs(y[2]).max()
    assert np.max(np.abs(y[2] - (0.75 * y[0] - 2.5 * y[1]))) <= 1e-7 * scale
    # forward-backward == backward-forward away from the (differently extended) ends
    inner = slice(20000, n - 20000)
    assert np.max(np.abs(y[3][::-1][inner] - y[0][inner])) <= 1e-7 * np.abs(y[0]).max()


# ------------------------------------------------------------------------- K3 / K4
def _screen(B, flat, starts, **kw):
    import torch
    pre = B.Preprocessor(screen_mode=kw.pop("screen_mode", "auto"))
    out = pre.screen_windows(torch.from_numpy(np.ascontiguousarray(flat)).cuda(), np.asarray(starts, np.int64), **kw)
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}


def test_screen_e2e_golden(gpu, golden):
    torch, B = gpu
    g = golden("e2e")
    nw = int(g["n_windows"])
    for mode in ("exact", "auto", "fast"):
        o = _screen(B, g["filtered"], np.arange(nw) * 10000, screen_mode=mode)
        kept = np.flatnonzero(o["keep"])
        assert same((kept + 1) * 10000, g["kept_idx_end"])
        for got, want in zip([o["peaks"][i, :o["n_peaks"][i]] for i in kept], unragged(g["kept_peaks"], g["kept_peaks_off"])):
            assert same(got, want)
        assert same(o["median"][kept], g["medians"]) and same(o["waves"][kept], g["np_waves"])
        assert o["status"][0] == 0


def test_screen_golden_all_fates(gpu, golden):
    torch, B = gpu
    g = golden("screen")
    raw = g["windows_unscaled"]
    n = len(raw)
    o = _screen(B, raw.reshape(-1), np.arange(n) * 10000, want_zwin=True)
    want_peaks = unragged(g["all_peaks"], g["all_peaks_off"])
    assert same(np.flatnonzero(o["keep"]), g["kept_rows"])
    for w in range(n):
        assert same(o["peaks"][w, :o["n_peaks"][w]], want_peaks[w])
        if np.isnan(g["all_sd"][w]):
            assert o["has_nan"][w] == 1
        else:
            assert same(o["zwin"][w], g["windows"][w])
            assert same(o["stats"][w, :3], [g["all_hr"][w], g["all_rmssd"][w], g["all_sd"][w]])
    k = g["kept_rows"]
    assert same(o["stats"][k, 3], g["kept_sd_per_beat"]) and same(o["median"][k], g["kept_median"])


def test_screen_many_windows_against_oracle(gpu):
    """Seeded synthetic cases of different heart rates: R-peaks, statistics, masks and
    medians of every window, parallel box filters (with fallback) and sequential mode."""
    torch, B = gpu
    wins = []
    for cid, hr in ((60, None), (61, 48.0), (62, 118.0), (63, 171.0), (64, 60.0), (65, 33.0)):
        f = rp.butterworth(rp.butterworth(rp.zscore(ecg_synth.synth_case(cid, 100000, hr_bpm=hr)), "high", 1), "low", 30)
        wins.append(f[:90000].reshape(9, 10000))
    rng = np.random.default_rng(4)
    wins.append(rng.standard_normal((3, 10000)))                     # noise: many QRS runs
    wins.append(np.tile(np.sin(np.arange(10000) / 40.0), (1, 1)))    # smooth periodic: ties / plateaus
    const = np.ones((1, 10000)); wins.append(const)                  # constant: z-score gives NaN -> dropped
    raw = np.concatenate(wins, axis=0)
    n = len(raw)
    zw = np.apply_along_axis(rp.zscore, 1, raw)
    ref = rp.window_screen(zw)
    for mode in ("auto", "exact"):
        o = _screen(B, raw.reshape(-1), np.arange(n) * 10000, screen_mode=mode)
        assert same(o["keep"].astype(bool), ref["keep"])
        assert same(o["has_nan"].astype(bool), ref["has_nan"])
        for w in range(n):
            assert same(o["peaks"][w, :o["n_peaks"][w]], ref["peaks"][w]), (mode, w)
        ok = ~ref["has_nan"]
        assert same(o["stats"][ok, 0], ref["hr"][ok]) and same(o["stats"][ok, 1], ref["rmssd"][ok])
        assert same(o["stats"][ok, 2], ref["sd"][ok]) and same(o["stats"][ok, 3], ref["sd_per_beat"][ok])
        for w in np.flatnonzero(ref["keep"]):
            assert same(o["median"][w], rp.compute_median_sample_ecg(zw[w], ref["peaks"][w], 500))
        assert o["status"][0] == 0


def test_screen_other_window_lengths(gpu):
    """Window lengths / sampling rates other than 10,000 @ 500 Hz (segmented box filters, bit-packed indicator,
    run tables, both median kernels' addressing) against the oracle."""
    import torch
    torch_, B = gpu
    for fs, secs, seed in ((250, 20, 3), (360, 20, 4), (500, 12, 5), (100, 20, 6), (512, 20, 7)):
        Lw = fs * secs
        x = ecg_synth.synth_case(seed, 6 * Lw + 17)
        y = rp.butterworth(rp.butterworth(rp.zscore(x), "high", 1.0, fs), "low", 30.0, fs)
        wins = np.stack([rp.zscore(y[k * Lw:(k + 1) * Lw]) for k in range(6)])
        want = rp.window_screen(wins, fs=fs)
        for mode in ("auto", "exact"):
            pre = B.Preprocessor(fs=fs, sample_length=secs, screen_mode=mode)
            out = pre.screen_windows(torch.from_numpy(np.ascontiguousarray(y[:6 * Lw])).cuda(), np.arange(6, dtype=np.int64) * Lw)
            o = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
            assert same(o["keep"].astype(bool), want["keep"]), (fs, mode)
            for j in range(6):
                assert same(o["peaks"][j, :o["n_peaks"][j]], want["peaks"][j]), (fs, mode, j)
                assert same(o["stats"][j, :3], [want["hr"][j], want["rmssd"][j], want["sd"][j]])
            for j in np.flatnonzero(want["keep"]):
                assert same(o["median"][j], rp.compute_median_sample_ecg(wins[j], want["peaks"][j], fs))


def test_median_dropin_golden(gpu, golden):
    from ae_waveforms_b200.data_load_preprocess.median_sample import compute_median_sample_ecg
    g = golden("median")
    for name in ("edge", "even", "odd", "one", "far_right"):
        med, al = compute_median_sample_ecg(g["window"], list(g[f"{name}_peaks"]), 500, return_aligned=True)
        assert same(med, g[f"{name}_median"]) and same(al, g[f"{name}_aligned"])
        assert same(compute_median_sample_ecg(g["window"], list(g[f"{name}_peaks"]), 500), g[f"{name}_median"])
    assert np.isnan(compute_median_sample_ecg(g["window"], [], 500))
    rng = np.random.default_rng(1)                                   # permutation invariance, many beats
    peaks = np.sort(rng.choice(np.arange(200, 9700), 70, replace=False))
    a = compute_median_sample_ecg(g["window"], list(peaks), 500)
    b = compute_median_sample_ecg(g["window"], list(rng.permutation(peaks)), 500)
    assert same(a, b) and same(a, rp.compute_median_sample_ecg(g["window"], list(peaks), 500))
    quant = np.round(g["window"] * 4) / 4                            # ties; every code path of the column median
    for n in (2, 31, 32, 33, 34, 47, 63, 64, 65, 100):
        for win in (g["window"], quant):
            peaks = np.sort(rng.choice(np.arange(120, len(win) - 230), n - 2, replace=False))
            peaks = np.concatenate([[40], peaks, [len(win) - 30]])
            assert same(compute_median_sample_ecg(win, list(peaks), 500),
                        rp.compute_median_sample_ecg(win, list(peaks), 500)), n


def test_remove_samples_dropin(gpu, golden):
    from ae_waveforms_b200.data_load_preprocess.remove_samples import remove_samples
    from ae_waveforms_b200.data_load_preprocess.sample_from_record import sample_from_record
    g = golden("screen")
    wins = g["windows"]
    idx = (np.arange(len(wins)) + 1).reshape(-1, 1) * 10000
    df = remove_samples(wins[:, :, None].copy(), idx, ["SNUADC/ECG_II"], 500, 150)
    ref = rp.remove_samples(wins[:, :, None].copy(), idx, ["SNUADC/ECG_II"], 500, 150)
    assert list(df.columns) == list(ref.columns) and len(df) == len(ref) == len(g["kept_rows"])
    for c in df.columns:
        for u, v in zip(df[c], ref[c]):
            assert same(u, v), c
    assert isinstance(df["ECG_R_Peaks"][0], list) and isinstance(df["ECG_R_Peaks"][0][0], int)
    w = golden("windows")
    rec = (np.arange(35000, dtype=np.float64) * 0.5)[:, None]
    s, i = sample_from_record(rec, ["SNUADC/ECG_II"])
    assert same(s[:, ::997, 0], w["rec35000_samples"]) and same(i, w["rec35000_idx"]) and i.dtype == np.int64
    assert s.flags.writeable
    with pytest.raises(ValueError):
        sample_from_record(rec, ["SNUADC/PLETH"])


def test_remove_samples_dropin_three_wave_tracks(gpu, golden):
    """remove_samples.py:31-33: a window goes if ANY wave track (ECG_II, ART, CVP) holds a NaN -- the configured path
    is single-lead, the reference's function is not."""
    from ae_waveforms_b200.data_load_preprocess.remove_samples import remove_samples
    g = golden("screen")
    wins = g["windows"]
    rng = np.random.default_rng(3)
    # (the reference reads the ECG from column `wave_track_names.index(name)` = 0: ECG first, as sample_from_record
    # delivers it for the configured track order)
    three = np.stack([wins, rng.standard_normal(wins.shape), rng.standard_normal(wins.shape)], axis=2)
    kept = [int(k) for k in g["kept_rows"]]
    three[kept[0], 17, 1] = np.nan                       # ART gap in a window the ECG alone would keep
    three[kept[-1], 9999, 2] = np.nan                    # CVP gap in another
    names = ["SNUADC/ECG_II", "SNUADC/ART", "SNUADC/CVP"]
    idx = (np.arange(len(wins)) + 1).reshape(-1, 1) * 10000
    df = remove_samples(three.copy(), idx, names, 500, 150)
    ref = rp.remove_samples(three.copy(), idx, names, 500, 150)
    assert len(df) == len(ref) == len(kept) - 2
    assert list(df.columns) == list(ref.columns)
    for c in df.columns:
        for u, v in zip(df[c], ref[c]):
            assert same(u, v), c


# ------------------------------------------------------------------------- whole path
def test_pipeline_e2e_golden_and_oracle(gpu, golden):
    torch, B = gpu
    g = golden("e2e")
    cases = [g["case"], ecg_synth.synth_case(70, 130000, nan_gaps=2), ecg_synth.synth_case(71, 95000, hr_bpm=60.0),
             ecg_synth.synth_case(72, 9000), np.full(25000, np.nan)]
    cases[1][:13] = np.nan
    cases[1][-5:] = np.nan
    ids = [9, 70, 71, 72, 73]
    refs = [rp.process_case(c, case_id=i) for c, i in zip(cases, ids)]
    want_info = np.concatenate([r["np_info"] for r in refs])
    want_med = np.concatenate([r["medians"] for r in refs])
    # exact filter: the whole chain is bit-identical, including the C.1 ulp-hazard windows
    pre = B.Preprocessor(filter_mode="exact")
    res = pre.run(B.RecordBatch.from_host(cases, ids), keep_intermediates=True)
    assert same(res["np_info"].cpu().numpy(), want_info)
    assert same(res["medians"].cpu().numpy(), want_med)
    assert same(res["np_waves"].cpu().numpy(), want_med.astype(np.float32))
    assert same(res["np_waves"].cpu().numpy()[:len(g["np_waves"])], g["np_waves"])
    y = res["filtered"].cpu().numpy()
    for r, b, e in zip(refs, res["rec_beg"], res["rec_end"]):
        assert same(y[b:e], r["filtered"])
    assert res["n_windows"] == sum(len(r["indices"]) for r in refs)
    # fast filter: floats within tolerance; integers identical except where np.std(window) sits on
    # the sd/n_peaks == 0.05 tie (n_peaks == 20), which is counted, not hidden
    res2 = B.Preprocessor(filter_mode="fast").run(B.RecordBatch.from_host(cases, ids))
    keep2 = res2["keep"].cpu().numpy().astype(bool)
    keep1 = res["keep"].cpu().numpy().astype(bool)
    npk = res["n_peaks"].cpu().numpy()
    diff = np.flatnonzero(keep1 != keep2)
    assert np.all(npk[diff] == 20), "mask differences outside the documented ulp hazard"
    assert same(res2["peaks"].cpu().numpy(), res["peaks"].cpu().numpy())
    both = keep1 & keep2
    m1 = res["medians"].cpu().numpy()[np.searchsorted(np.flatnonzero(keep1), np.flatnonzero(both))]
    m2 = res2["medians"].cpu().numpy()[np.searchsorted(np.flatnonzero(keep2), np.flatnonzero(both))]
    assert close(m2, m1)
    print(f"ulp-hazard mask flips under the fast filter: {len(diff)} of {len(keep1)} windows")


def test_process_cases_host_api(gpu):
    from ae_waveforms_b200.batched import process_cases
    cases = [ecg_synth.synth_case(80, 60000), ecg_synth.synth_case(81, 45000)]
    out = process_cases(cases, [80, 81], filter_mode="exact")
    refs = [rp.process_case(c, case_id=i) for c, i in zip(cases, (80, 81))]
    assert same(out["np_waves"], np.concatenate([r["np_waves"] for r in refs]))
    assert same(out["np_info"], np.concatenate([r["np_info"] for r in refs]))
    assert out["np_waves"].dtype == np.float32 and out["np_features"].shape == (len(out["np_waves"]), 35)


# ------------------------------------------------------------------------- K5
def test_recon_golden_and_dropins(gpu, golden):
    from ae_waveforms_b200.functions_traditional.signal_reconstruction import (extract_coordinates, gaussian_line,
                                                                              reconstruct_gaussian)
    from ae_waveforms_b200.sigma_search import SigmaSearch
    g = golden("recon")
    s = SigmaSearch(g["fx"], g["fy"], g["baselines"], g["test_true"], g["x_t"])
    for k, smap in enumerate(g["sigma_maps"]):
        assert close(s.reconstruct(smap).cpu().numpy(), g["recon"][k])
    sse = s.sse(g["sigma_maps"]).cpu().numpy()
    want = np.array([rp.sigma_objective(g["fx"], g["fy"], g["baselines"], g["test_true"], m, g["x_t"])[0]
                     for m in g["sigma_maps"]])
    assert np.all(np.abs(sse - want) <= RTOL * np.abs(want))
    mse = s.mse(g["sigma_maps"]).cpu().numpy()
    assert np.all(np.abs(mse - g["mse_torch_f32"]) <= 1e-5 * g["mse_torch_f32"])
    assert s.best(g["sigma_maps"])[0] == int(np.argmin(want))
    for b in (0, 1, 2, 5, 7):
        assert close(reconstruct_gaussian(g["fx"][b], g["fy"][b], g["x_t"], g["baselines"][b], g["sigma_maps"][3]),
                     g["recon"][3][b])
    cx, cy = unragged(g["coord_x"], g["coord_off"]), unragged(g["coord_y"], g["coord_off"])
    for b in range(10):
        x, y, m = extract_coordinates(g["fx"][b], g["fy"][b], g["x_t"], g["baselines"][b])
        assert same(x, cx[b]) and same(y, cy[b]) and same(m, g["coord_nanmask"][b])
    gl = np.concatenate([gaussian_line(-80.0, -0.4, -38.0, -1.2, 21.0), gaussian_line(-38.0, -1.2, 0.0, -1e-10, 12.5),
                         gaussian_line(10.5, -0.3, 61.9, -0.9, 8.0)])
    assert close(gl, g["gaussian_lines"], 1e-12)


def test_sigma_search_against_oracle(gpu):
    from ae_waveforms_b200.sigma_search import SigmaSearch
    tt, fx, fy, bl = ecg_synth.synth_fiducials(192, seed=5, nan_frac=0.03)
    cands = np.concatenate([ecg_synth.sigma_candidates(20, "integer"), ecg_synth.sigma_candidates(10, "continuous")[:10]])
    s = SigmaSearch(fx, fy, bl, tt)
    sse = s.sse(cands).cpu().numpy()
    want = np.array([rp.sigma_objective(fx, fy, bl, tt, c)[0] for c in cands])
    assert same(np.isnan(sse), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.all(np.abs(sse[ok] - want[ok]) <= RTOL * np.abs(want[ok]))
    order_gpu = np.argsort(np.where(ok, sse, np.inf), kind="stable")
    order_ref = np.argsort(np.where(ok, want, np.inf), kind="stable")
    assert order_gpu[0] == order_ref[0], "optimal sigma map differs"


def test_sigma_grid_tables_against_direct_and_oracle(gpu):
    """The separable-table formulation for discrete grids gives the direct kernel's sums (different float64
    summation order only), the oracle's objective and the same winner; sse() routes to it automatically."""
    from ae_waveforms_b200 import _cabi
    from ae_waveforms_b200.sigma_search import SigmaSearch
    torch, _ = gpu
    tt, fx, fy, bl = ecg_synth.synth_fiducials(700, seed=9, nan_frac=0.04)
    cands = ecg_synth.sigma_candidates(300, "integer")
    s = SigmaSearch(fx, fy, bl, tt)
    direct = s.sse(cands, method="direct").cpu().numpy()
    tables = s.sse(cands, method="tables").cpu().numpy()
    auto = s.sse(cands).cpu().numpy()
    assert same(auto, tables)
    assert same(tables, s.sse(cands, method="tables").cpu().numpy()), "table sums must be run-to-run deterministic"
    assert same(np.isnan(direct), np.isnan(tables))
    ok = ~np.isnan(direct)
    assert ok.any() and np.all(np.abs(tables[ok] - direct[ok]) <= 1e-12 * np.abs(direct[ok]))
    want = np.array([rp.sigma_objective(fx, fy, bl, tt, c)[0] for c in cands[:12]])
    assert np.all(np.abs(tables[:12] - want) <= RTOL * np.abs(want))
    assert s.best(cands)[0] == int(np.argmin(np.where(ok, direct, np.inf)))
    # fractional three-value grid, explicit (values, index) entry
    vals = np.array([0.5, 2.25, 7.0])
    idx = np.random.default_rng(2).integers(0, 3, size=(40, 12))
    g3 = s.sse_grid(vals, idx).cpu().numpy()
    d3 = s.sse(vals[idx], method="direct").cpu().numpy()
    assert np.all(np.abs(g3 - d3) <= 1e-12 * np.abs(d3))
    with pytest.raises(_cabi.AewError):
        s.sse(ecg_synth.sigma_candidates(40, "continuous"), method="tables")


def test_host_pipelined_equals_single_batch(gpu):
    """The chunked, double-buffered host feeder returns exactly what one resident batch returns."""
    torch, B = gpu
    cases = [ecg_synth.synth_case(90 + i, 40000 + 3000 * i, nan_gaps=(i % 3 == 0)) for i in range(7)]
    ids = list(range(90, 97))
    batch = B.RecordBatch.from_host(cases, ids)
    pre = B.Preprocessor()
    want = pre.run(batch)
    host = torch.from_numpy(np.concatenate(cases)).pin_memory()
    for chunk in (2, 3, 100):
        got = B.run_host_pipelined(pre, host, batch.rec_off, ids, chunk_cases=chunk)
        assert same(got["np_waves"].numpy(), want["np_waves"].cpu().numpy())
        assert same(got["np_info"].numpy(), want["np_info"].cpu().numpy())
        assert got["n_windows"] == want["n_windows"]


def test_screen_forced_fallback_transition(gpu, golden):
    """mode 3 runs the parallel box filters and then the sequential fallback for every window
    (the path a hazard window takes); results must equal both other modes."""
    torch, B = gpu
    g = golden("e2e")
    nw = int(g["n_windows"])
    ref = _screen(B, g["filtered"], np.arange(nw) * 10000, screen_mode="exact")
    got = _screen(B, g["filtered"], np.arange(nw) * 10000, screen_mode="forced_fallback")
    for k in ("keep", "has_nan", "n_peaks", "peaks", "stats"):
        assert same(got[k], ref[k]), k
    kept = np.flatnonzero(ref["keep"])
    assert same(got["median"][kept], ref["median"][kept]) and int(got["hazards"][0]) == int((ref["has_nan"] == 0).sum())


def test_filter_many_nan_islands(gpu):
    """1 % random NaNs: thousands of tiny segments (pass-through and filtered), both modes; and the
    segment-table overflow is reported, not ignored."""
    torch, B = gpu
    from ae_waveforms_b200 import _cabi
    rng = np.random.default_rng(12)
    x = ecg_synth.synth_case(95, 120000)
    x[rng.random(x.size) < 0.01] = np.nan
    x[40000:40100] = np.nan
    ref = rp.butterworth(rp.butterworth(rp.zscore(x), "high", 1), "low", 30)
    batch = B.RecordBatch.from_host([x])
    for mode in ("exact", "fast"):
        pre = B.Preprocessor(filter_mode=mode, max_segments=4096)
        beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
        mean, sd, cnt = pre.zscore_stats(batch.flat, beg_h, end_h, d_beg, d_end)
        y, hdr = pre.filter_records(batch.flat, beg_h, end_h, d_beg, d_end, mean, sd, cnt)
        assert int(hdr[4].item()) == 0 and int(hdr[0].item()) > 1000
        y = y.cpu().numpy()
        assert same(y, ref) if mode == "exact" else close(y, ref)
    small = B.Preprocessor(max_segments=64)
    with pytest.raises(_cabi.AewError):
        small.run(batch)


def test_empty_and_windowless_batches(gpu):
    torch, B = gpu
    pre = B.Preprocessor()
    res = pre.run(B.RecordBatch.from_host([ecg_synth.synth_case(1, 9000), np.full(300, np.nan)], [5, 6]))
    assert res["np_waves"].shape == (0, 320) and res["np_info"].shape == (0, 2) and res["n_windows"] == 0


# ------------------------------------------------------------------------- K6 / K7: features
FID_NAN = -2147483648


def _gpu_features(gpu, windows, peaks_list, fids):
    torch, B = gpu
    pre = B.Preprocessor()
    n, L = windows.shape
    maxp = pre.lib.aew_windows_max_peaks()
    pk = np.full((n, maxp), -1, np.int32)
    fid = np.full((n, 10, maxp), FID_NAN, np.int32)
    for i, (p, f) in enumerate(zip(peaks_list, fids)):
        pk[i, :len(p)] = p
        f = np.asarray(f, dtype=np.float64).reshape(10, -1)
        fid[i, :, :f.shape[1]] = np.where(np.isnan(f), FID_NAN, f).astype(np.int64)
    dev = pre.device
    sig = torch.from_numpy(np.ascontiguousarray(windows.reshape(-1))).to(dev)
    off = torch.arange(n, dtype=torch.int64, device=dev) * L
    npk = torch.tensor([len(p) for p in peaks_list], dtype=torch.int32, device=dev)
    out, status = pre.fiducial_features(sig, off, torch.from_numpy(pk).to(dev), npk, torch.from_numpy(fid).to(dev), L=L)
    return out.cpu().numpy(), int(status.item())


def test_fiducial_features_bit_exact(gpu, golden):
    g = golden("features")
    gf = np.where(g["fiducials"] < 0, np.nan, g["fiducials"])
    rows, st = _gpu_features(gpu, g["window"][None, :], [g["r_peaks"]], [gf])
    assert st == 0 and same(rows[0], g["row35"])            # frozen from the reference's own functions
    rng = np.random.default_rng(5)
    wins, pks, fids = [], [], []
    offs = np.array([-100, -125, -75, -12, -20, 22, 15, 125, 90, 170])
    for n_beats in (1, 2, 3, 10, 24, 33, 47, 64, 72) * 4:
        w = rng.standard_normal(10000)
        p = np.sort(rng.choice(np.arange(150, 9800), n_beats, replace=False))
        f = np.clip(p[None, :] + offs[:, None] + rng.integers(-6, 7, size=(10, n_beats)), 1, 9999).astype(float)
        f[rng.random(f.shape) < 0.2] = np.nan
        if n_beats == 10:
            f[1, :] = np.nan
            f[4, ::2] = f[4, 1]
        wins.append(w); pks.append(p); fids.append(f)
    rows, st = _gpu_features(gpu, np.stack(wins), pks, fids)
    assert st == 0
    for i in range(len(wins)):
        fid = {k: [np.nan if np.isnan(v) else int(v) for v in fids[i][j]] for j, k in enumerate(rp.FIDUCIALS)}
        assert same(rows[i], rp.features_row_from_fiducials(wins[i], [int(p) for p in pks[i]], fid)), i


def _delineation_cases():
    wins, pks = [], []
    for cid, hr in ((11, 38.0), (12, 52.0), (13, 60.0), (14, 95.0), (15, 130.0), (16, 170.0), (17, None),
                    (18, 57.7), (19, 62.5), (20, 60.0), (21, 60.0)):   # 0.35 (n - 1) integral: an epoch label is exactly 0
        out = rp.process_case(ecg_synth.synth_case(cid, 50001, hr_bpm=hr), cid)
        for w, p in zip(out["windows"], out["peaks"]):
            wins.append(w); pks.append([int(v) for v in p])
    wins.append(wins[0]); pks.append(pks[0][:3])                                    # the library raises
    wins.append(wins[1]); pks.append([20] + [p for p in pks[1] if p > 300])         # negative slice bounds
    wins.append(wins[2]); pks.append([p for p in pks[2] if p < 9700] + [9990])      # entries >= L
    rng = np.random.default_rng(3)
    wins.append(rng.standard_normal(10000))                                          # pure noise
    pks.append(sorted(rng.choice(np.arange(5, 9995, 160), 40, replace=False).tolist()))
    return wins, pks


def test_delineation_matches_restatement(gpu):
    """K7 against oracle/nk_delineate.py (restatement of neurokit2's DWT delineator; parity unpinned at
    the library).  Every fiducial index must be identical: the two sides differ by float64 rounding only
    (chunked spline prefilter, summation order), which can flip an index only at an exact tie."""
    from oracle import nk_delineate as nd
    torch, B = gpu
    wins, pks = _delineation_cases()
    pre = B.Preprocessor()
    n, L, dev = len(wins), 10000, pre.device
    maxp = pre.lib.aew_windows_max_peaks()
    pk = np.full((n, maxp), -1, np.int32)
    for i, p in enumerate(pks):
        pk[i, :len(p)] = p
    sig = torch.from_numpy(np.ascontiguousarray(np.stack(wins).reshape(-1))).to(dev)
    off = torch.arange(n, dtype=torch.int64, device=dev) * L
    d_pk = torch.from_numpy(pk).to(dev)
    npk = torch.tensor([len(p) for p in pks], dtype=torch.int32, device=dev)
    fid, flags = pre.delineate_windows(sig, off, d_pk, npk)
    feats, st = pre.fiducial_features(sig, off, d_pk, npk, fid, (flags == 0).to(torch.uint8))
    fid, flags, feats = fid.cpu().numpy(), flags.cpu().numpy(), feats.cpu().numpy()
    assert int(st.item()) == 0
    want_flag = {"ok": 0, "nan": 1, "drop": 2}
    total = 0
    for i, (w, p) in enumerate(zip(wins, pks)):
        status, ref = nd.window_fiducials(w, p)
        assert flags[i] == want_flag[status], (i, status, flags[i])
        if status == "ok":
            got = fid[i, :, :len(p)].astype(np.float64)
            got[got == FID_NAN] = np.nan
            assert same(got, ref), i
            assert same(feats[i], rp.window_features(w, p)[1]), i
            total += ref.size
        else:
            assert np.isnan(feats[i]).all()
    assert total > 5000


def test_pipeline_with_features_exact_filter(gpu):
    """np_waves, np_features, np_info of the whole path (exact filter mode) against the oracle's process_case."""
    torch, B = gpu
    cases = [ecg_synth.synth_case(c, 120000 + 977 * c, hr_bpm=hr, nan_gaps=g)
             for c, hr, g in ((0, None, 0), (1, 58.0, 0), (2, 118.0, 1), (3, None, 2))]
    refs = [rp.process_case(c, case_id=i, features=True) for i, c in enumerate(cases)]
    res = B.Preprocessor(filter_mode="exact").run(B.RecordBatch.from_host(cases, list(range(len(cases)))),
                                                  features=True)
    assert same(res["np_info"].cpu().numpy(), np.concatenate([r["np_info"] for r in refs]))
    assert same(res["np_waves"].cpu().numpy(), np.concatenate([r["np_waves"] for r in refs]))
    assert same(res["np_features"].cpu().numpy(), np.concatenate([r["np_features"] for r in refs]))
    assert res["np_features"].shape[1] == 35 and res["np_features"].shape[0] > 15
    fast = B.process_cases(cases, list(range(len(cases))))
    ref_f = np.concatenate([r["np_features"] for r in refs])
    assert fast["np_features"].shape == ref_f.shape
    # fast filter: windows differ by ~1e-9 from SciPy's; no row of these cases changes, and the entries are within
    # 1e-6 of the reference's on the scale of their row wherever no fiducial moved (a moved fiducial changes a
    # median over beats by a sample step and is counted by test_parity_64_cases_against_the_oracle, not tolerated here)
    scale = np.nanmax(np.abs(ref_f), axis=1, keepdims=True)
    err = np.abs(fast["np_features"] - ref_f) / scale
    moved = np.nanmax(err, axis=1) > 1e-6
    assert same(np.isnan(fast["np_features"]), np.isnan(ref_f))
    assert moved.sum() <= max(1, len(ref_f) // 100), (int(moved.sum()), len(ref_f))
    # the same call with exact_masks re-decides the 20-peak windows from SciPy-order filter output: same rows
    fast_x = B.process_cases(cases, list(range(len(cases))), exact_masks=True)
    assert same(fast_x["np_info"], np.concatenate([r["np_info"] for r in refs]))


def test_parity_64_cases_against_the_oracle(gpu):
    """>= 64 HOST-generated 1 h cases (BASELINE configs[1] subsample) through the benchmarked
    configuration (tiled filter) against the oracle, with hard bounds.  Every tenth case is pinned to 60 bpm so that
    the n_peaks == 20 rounding hazard is well represented.
      R peaks identical in every window; keep / remove differs only on 20-peak windows, and not at all with
      exact_masks; median beats within 1e-6; fiducials identical in > 99.9 % of the entries; the 35 features
      within 1e-6 (row scale) on windows whose fiducials are identical."""
    torch, B = gpu
    from oracle import parity
    n_cases, n_samp = 64, 1800000
    ids = list(range(3000, 3000 + n_cases))
    hr = [60.0 if k % 10 == 0 else float(np.random.default_rng(k).uniform(50, 110)) for k in range(n_cases)]
    exp = parity.expected(ids, n_samp, features=True, hr=hr)
    cases = [e["x"] for e in exp]
    for exact_masks in (False, True):
        pre = B.Preprocessor(exact_masks=exact_masks)
        res = pre.run(B.RecordBatch.from_host(cases, ids), features=True)
        rep = parity.compare(exp, parity.gpu_result_to_numpy(res))
        print("parity", "exact_masks" if exact_masks else "fast masks", rep,
              {k: res[k] for k in ("mask_hazard_windows", "mask_flips_resolved")})
        assert rep["windows"] == 64 * 179 and rep["windows_rpeaks_differ"] == 0
        assert rep["mask_flips_off_hazard"] == 0
        assert rep["median_max_rel_err"] <= 1e-6
        assert rep["delineation_fate_differs"] == 0
        assert rep["fiducials_identical_fraction"] > 0.999
        assert rep["features_max_rel_err"] <= 1e-6 and rep["features_compared"] > 10000
        assert rep["hazard_windows"] > 200
        if exact_masks:
            assert rep["mask_flips"] == 0 and rep["rows_gpu"] == rep["rows_reference"]
            assert res["mask_hazard_windows"] >= rep["hazard_windows"] * 0.9


def test_feature_dropins_match_golden(gpu, golden):
    """ecg_features_relative_r -> compute_ecg_morphology_features -> compute_median_features_ecg (the stage kernels of
    K6 behind the reference's three signatures) against arrays frozen from the reference's own functions; estimate_hr /
    rmssd (aew_peak_stats_f64) against the reference formulas."""
    from ae_waveforms_b200.data_load_preprocess.feature_extraction.features_ecg import (
        compute_ecg_morphology_features, ecg_features_relative_r)
    from ae_waveforms_b200.data_load_preprocess.median_sample import compute_median_features_ecg
    g = golden("features")
    fid = {n: [np.nan if v < 0 else int(v) for v in g["fiducials"][k]] for k, n in enumerate(rp.FIDUCIALS)}
    rel = ecg_features_relative_r(g["window"], fid, list(g["r_peaks"]), 500)
    morph = compute_ecg_morphology_features(rel, len(g["r_peaks"]))
    assert "ECG_R_Peaks" in rel
    med = compute_median_features_ecg(rel, morph)
    row = [v for n in rp.FIDUCIALS for v in (med[n]["x"], med[n]["y"])] + [med[n] for n in rp.MORPHOLOGY]
    assert np.array_equal(np.array(row), g["row35"], equal_nan=True)
    assert np.array_equal(np.array([rel[n]["x"] for n in rp.FIDUCIALS]), g["rel_x"], equal_nan=True)
    assert np.array_equal(np.array([morph[n] for n in rp.MORPHOLOGY]), g["morph"], equal_nan=True)
    from ae_waveforms_b200.data_load_preprocess.remove_samples import estimate_hr, rmssd
    assert estimate_hr([], [100, 600, 1100], 500, mod="max") == 60.0
    pk = list(g["r_peaks"])
    for mod in ("min", "mean", "max"):
        assert same(estimate_hr(pk[::2], pk, 500, mod=mod), rp.estimate_hr(pk[::2], pk, 500, mod))
    assert same(rmssd(pk, 500), rp.rmssd(pk, 500)) and np.isnan(rmssd(pk[:2], 500)) and np.isnan(estimate_hr([], [5], 500))


def test_neurokit_shim(gpu):
    from ae_waveforms_b200 import neurokit_shim as nk
    from oracle import nk_delineate as nd
    out = rp.process_case(ecg_synth.synth_case(21, 30001, hr_bpm=72.0), 21)
    w, p = out["windows"][0], [int(v) for v in out["peaks"][0]]
    got = nk.ecg_delineate(w, p, sampling_rate=500)[1]
    want = nd.ecg_delineate(w, p, 500)
    assert list(got) == nd.KEYS
    for k in nd.KEYS:
        assert len(got[k]) == len(want[k])
        assert all((a is np.nan and b is np.nan) or a == b for a, b in zip(got[k], want[k])), k
    with pytest.raises(ValueError):
        nk.ecg_delineate(w, p[:3], sampling_rate=500)
    from oracle import nk_peaks
    assert nk.ecg_findpeaks(w, sampling_rate=500)["ECG_R_Peaks"].tolist() == \
        [int(v) for v in nk_peaks.ecg_findpeaks(w, sampling_rate=500)["ECG_R_Peaks"] if 0 < v < len(w)]


# ------------------------------------------------------------------------- K8: input feeder
def test_widen_and_narrow_source_feeder(gpu):
    torch, B = gpu
    rng = np.random.default_rng(8)
    for n in (1, 7, 8, 1000, 4099, 1 << 20):
        x = (rng.standard_normal(n) * 3).astype(np.float32)
        x[n // 2] = np.nan
        assert same(B.widen_to_f64(torch.from_numpy(x).cuda()).cpu().numpy(), x.astype(np.float64))
        v = rng.integers(-32768, 32767, n).astype(np.int16)
        v[n // 3] = -32768
        want = v.astype(np.float64) * 0.00244 + (-1.25)
        want[v == -32768] = np.nan
        got = B.widen_to_f64(torch.from_numpy(v).cuda(), gain=0.00244, offset=-1.25, sentinel=-32768)
        assert same(got.cpu().numpy(), want)
    # the end-to-end feeder on float32 / int16 host buffers == the device-resident path on the widened values
    cases = [ecg_synth.synth_case(c, 90000 + 1000 * c) for c in range(5)]
    off = np.concatenate([[0], np.cumsum([len(c) for c in cases])]).astype(np.int64)
    flat32 = np.concatenate(cases).astype(np.float32)
    pre = B.Preprocessor()
    ref = pre.run(B.RecordBatch.from_host([flat32[off[i]:off[i + 1]].astype(np.float64) for i in range(5)]))
    got = B.run_host_pipelined(pre, torch.from_numpy(flat32).pin_memory(), off, chunk_cases=2)
    assert same(got["np_waves"].numpy(), ref["np_waves"].cpu().numpy())
    assert same(got["np_info"].numpy(), ref["np_info"].cpu().numpy())
    gotf = B.run_host_pipelined(pre, torch.from_numpy(flat32).pin_memory(), off, chunk_cases=2, features=True)
    reff = pre.run(B.RecordBatch.from_host([flat32[off[i]:off[i + 1]].astype(np.float64) for i in range(5)]),
                   features=True)
    assert same(gotf["np_waves"].numpy(), reff["np_waves"].cpu().numpy())
    assert same(gotf["np_features"].numpy(), reff["np_features"].cpu().numpy()) and gotf["np_features"].shape[1] == 35
    gain = 1.0 / 4096
    flat16 = np.clip(np.rint(np.concatenate(cases) / gain), -32767, 32767).astype(np.int16)
    ref = pre.run(B.RecordBatch.from_host([flat16[off[i]:off[i + 1]].astype(np.float64) * gain + 0.0
                                           for i in range(5)]))
    got = B.run_host_pipelined(pre, torch.from_numpy(flat16).pin_memory(), off, chunk_cases=2, gain=gain)
    assert same(got["np_waves"].numpy(), ref["np_waves"].cpu().numpy()) and len(got["np_waves"]) > 10


def test_host_tables_reach_the_device_without_the_copy_engine(gpu):
    """Preprocessor._upload: bytes identical for every dtype / size, across ring wrap-arounds, and with a bulk
    host->device transfer in flight (the situation it exists for)."""
    torch, B = gpu
    pre = B.Preprocessor()
    pre.UPLOAD_RING_BYTES = 1 << 20                                  # small ring: many wraps
    rng = np.random.default_rng(12)
    big_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    big_d = torch.empty_like(big_h, device="cuda")
    side = torch.cuda.Stream()
    for i in range(60):
        if i % 10 == 0:
            with torch.cuda.stream(side):
                big_d.copy_(big_h, non_blocking=True)
        n = int(rng.integers(0, 40000))
        for a in (rng.integers(-2 ** 62, 2 ** 62, n), rng.standard_normal((n // 2, 2)), rng.integers(0, 255, n).astype(np.uint8),
                  rng.integers(-9, 9, n).astype(np.int32)):
            d = pre._upload(a)
            assert d.shape == a.shape and same(d.cpu().numpy(), a)
    large = rng.standard_normal(200000)                               # > half the ring: the ordinary copy
    assert same(pre._upload(large).cpu().numpy(), large)
    torch.cuda.synchronize()


# ------------------------------------------------------------------------- K9: post-search metrics
def test_beat_metrics_match_reference_torch(gpu):
    from ae_waveforms_b200.sigma_search import SigmaSearch
    torch, B = gpu
    tt, fx, fy, bl = ecg_synth.synth_fiducials(300, seed=4)
    rng = np.random.default_rng(4)
    med = tt.astype(np.float64) + rng.uniform(0.5, 3.0, size=(len(tt), 1))      # medians whose row max is not 0
    sigma_map = [1, 2, 2, 3, 2, 3, 3, 2, 3, 2, 2, 1]                            # analysis_...:121
    s = SigmaSearch(fx, fy, bl, tt)
    m = s.metrics(sigma_map, med)
    y = np.array([rp.reconstruct_gaussian(fx[i], fy[i], rp.beat_time_axis(), bl[i], sigma_map) for i in range(len(tt))])
    assert close(m["recon"].cpu().numpy(), y)
    want = rp.beat_metrics(y, med)
    for key, w in zip(("rmse", "prd", "qs"), want):
        g = m[key].cpu().numpy()
        assert same(np.isnan(g), np.isnan(w))
        ok = ~np.isnan(w)
        assert np.all(np.abs(g[ok] - w[ok]) <= 1e-5 * np.abs(w[ok])), key
    ok = ~np.isnan(want[0])
    assert abs(m["summary"]["rmse"]["median"] - np.median(want[0][ok])) <= 1e-5 * np.median(want[0][ok])
    assert abs(m["summary"]["qs"]["q3"] - np.percentile(want[2][ok], 75)) <= 1e-5 * np.percentile(want[2][ok], 75)


def test_delineation_other_window_lengths_and_rates(gpu):
    """10 s strips at 500 Hz (configs[2] adaptation, L = 5,000) and 20 s windows at 250 Hz."""
    from oracle import nk_delineate as nd, nk_peaks
    torch, B = gpu
    case = ecg_synth.synth_case(31, 90000, hr_bpm=80.0)
    f = rp.butterworth(rp.butterworth(rp.zscore(case), "high", 1, 500, 4), "low", 30, 500, 4)
    for fs, sig in ((500, f[10000:]), (250, f[::2][5000:])):
        n = len(sig) // 5000
        wins = [rp.zscore(sig[k * 5000:(k + 1) * 5000]) for k in range(n)]
        pks = [[int(v) for v in nk_peaks.ecg_findpeaks(w, fs)["ECG_R_Peaks"] if 0 < v < 5000] for w in wins]
        pre = B.Preprocessor(fs=fs, sample_length=5000 / fs)
        maxp = pre.lib.aew_windows_max_peaks()
        pk = np.full((n, maxp), -1, np.int32)
        for i, p in enumerate(pks):
            pk[i, :len(p)] = p
        dev = pre.device
        d_sig = torch.from_numpy(np.ascontiguousarray(np.stack(wins).reshape(-1))).to(dev)
        off = torch.arange(n, dtype=torch.int64, device=dev) * 5000
        fid, flags = pre.delineate_windows(d_sig, off, torch.from_numpy(pk).to(dev),
                                           torch.tensor([len(p) for p in pks], dtype=torch.int32, device=dev), L=5000)
        fid, flags = fid.cpu().numpy(), flags.cpu().numpy()
        for i, (w, p) in enumerate(zip(wins, pks)):
            st, ref = nd.window_fiducials(w, p, fs)
            assert flags[i] == {"ok": 0, "nan": 1, "drop": 2}[st], (fs, i)
            if st == "ok":
                got = fid[i, :, :len(p)].astype(np.float64)
                got[got == FID_NAN] = np.nan
                assert same(got, ref), (fs, i)


# ------------------------------------------------------------------------- full-size properties (BASELINE configs[1])
def test_full_size_properties_1000_cases(gpu):
    """configs[1] at its full size (1,000 cases x 1 h, 1.8e9 samples, 14.4 GB): properties that do not need
    the oracle -- the result must not depend on how the cases are batched or ordered, rows are case-major,
    every case contributes at most 179 windows, R peaks sit where the beat window puts them, and the
    delineation + feature stage is reproducible and complete on every kept window."""
    import bench
    torch, B = gpu
    if torch.cuda.mem_get_info()[1] < 90e9:
        pytest.skip("needs a full B200 (about 60 GB of HBM)")
    n_cases, ns = 1000, 1800000
    dev = torch.device("cuda")
    ids = np.arange(n_cases)
    flat = bench.synth_cases_gpu(ids, ns, dev)
    off = np.arange(n_cases + 1, dtype=np.int64) * ns
    pre = B.Preprocessor()
    scratch = {"y": torch.empty_like(flat), "tmp": torch.empty_like(flat)}
    res = pre.run(B.RecordBatch(flat, off, ids), scratch=scratch, features=True)
    waves, info, feats = res["np_waves"].clone(), res["np_info"].clone(), res["np_features"].clone()
    n = waves.shape[0]
    assert 100000 < n <= 179 * n_cases and feats.shape == (n, 35) and res["n_windows"] == 179 * n_cases
    case, end = info[:, 0].to(torch.int64), info[:, 1].to(torch.int64)
    key = case * (1 << 40) + end
    assert bool((key[1:] > key[:-1]).all())                                  # case-major, idx_end ascending
    assert bool((end % 10000 == 0).all()) and int(end.min()) >= 10000 and int(end.max()) <= 1790000
    assert int(torch.bincount(case, minlength=n_cases).max()) <= 179
    assert bool((waves.argmax(dim=1) == 106).float().mean() > 0.999)         # the R peak's column (median_sample.py:20-25)
    assert not bool(torch.isnan(waves).any())
    assert float(torch.isnan(feats).double().mean()) < 1e-3                  # nearly every fiducial found on clean ECG
    ok = ~torch.isnan(feats[:, 0])
    assert bool((feats[ok, 0] < 0).all())                                    # P peak before R ...
    ok = ~torch.isnan(feats[:, 14])
    assert bool((feats[ok, 14] > 0).all())                                   # ... T peak after it
    # batching independence: the second half of the cases alone, and 100-case chunks, give the same rows bit for bit
    h = n_cases // 2
    sub = pre.run(B.RecordBatch(flat[h * ns:], off[h:] - off[h], ids[h:]),
                  scratch={"y": scratch["y"][:(n_cases - h) * ns], "tmp": scratch["tmp"][:(n_cases - h) * ns]},
                  features=True)
    sel = case >= h
    assert torch.equal(sub["np_waves"], waves[sel]) and torch.equal(sub["np_info"], info[sel])
    assert torch.equal(torch.nan_to_num(sub["np_features"], nan=-1e300), torch.nan_to_num(feats[sel], nan=-1e300))
    del sub
    # without the feature stage the window statistics come from the one-sweep kernels (K3Params.fast_stats, behind
    # the tiled filter): the same windows are kept, the beats agree to rounding -- and THAT path, too, must not
    # depend on the batching
    plain = pre.run(B.RecordBatch(flat, off, ids), scratch=scratch)
    assert torch.equal(plain["np_info"], info) and torch.equal(plain["keep"], res["keep"])
    assert torch.equal(plain["peaks"], res["peaks"])
    assert float((plain["np_waves"] - waves).abs().max()) < 1e-6
    waves = plain["np_waves"].clone()
    del plain
    rows = 0
    for c0 in range(0, n_cases, 100):
        part = pre.run(B.RecordBatch(flat[c0 * ns:(c0 + 100) * ns], off[c0:c0 + 101] - off[c0], ids[c0:c0 + 100]),
                       scratch={"y": scratch["y"][:100 * ns], "tmp": scratch["tmp"][:100 * ns]})
        k = part["np_waves"].shape[0]
        assert torch.equal(part["np_waves"], waves[rows:rows + k]) and torch.equal(part["np_info"], info[rows:rows + k])
        rows += k
    assert rows == n
    # affine invariance of the z-scored path: 3 x + 7 gives the same windows and (to rounding) the same beats
    flat.mul_(3.0).add_(7.0)
    aff = pre.run(B.RecordBatch(flat, off, ids), scratch=scratch)
    # 3 x + 7 rounds differently inside the z-score, so np.std of a 20-peak window may land on the other side of
    # 1.0 -- in the reference just the same; every OTHER window must keep its decision and its beat
    same_fate = (aff["keep"] == res["keep"]) | (res["n_peaks"] == 20)
    assert bool(same_fate.all())
    if int(res["keep"].sum()) == n:                            # (no window dropped by the delineation stage)
        both = (aff["keep"] & res["keep"]).bool()
        a_rows = aff["np_waves"][both[aff["keep"].bool()]]
        r_rows = waves[both[res["keep"].bool()]]
        assert a_rows.shape == r_rows.shape and float((a_rows - r_rows).abs().max()) < 1e-5


# ------------------------------------------------------------------------- configs[2]: 12-lead 10 s strips
def test_strips_mode_12_lead_records(gpu):
    """BASELINE configs[2] shape (records x 12 leads x 5,000 samples): every (record, lead) strip is filtered on
    its own and taken as ONE window (the reference's windowing cuts none from a record of exactly L samples).
    The sd/n_peaks threshold is 0.05 for 20 s windows (n_peaks >= 20); a 10 s strip needs 0.1 for the same rate."""
    torch, B = gpu
    n_rec, n_lead, L = 24, 12, 5000
    strips, ids = [], []
    for r in range(n_rec):
        base = ecg_synth.synth_case(400 + r, L, hr_bpm=66.0 + 4 * r)
        for l in range(n_lead):
            x = base * (0.4 + 0.1 * l) + 0.01 * np.random.default_rng(r * 12 + l).standard_normal(L)
            if r == 5 and l == 3:
                x[1000:1040] = np.nan                      # a gap: the strip is removed (NaN window)
            strips.append(x); ids.append(r * n_lead + l)
    pre = B.Preprocessor(sample_length=10, sd_threshold=0.1, filter_mode="exact")
    res = pre.run(B.RecordBatch.from_host(strips, ids), strips=True, features=True)
    refs = [rp.process_case(x, case_id=i, sample_length=10, strip=True, sd_threshold_ecg=0.1, features=True)
            for x, i in zip(strips, ids)]
    want_w = np.concatenate([r["np_waves"] for r in refs])
    want_i = np.concatenate([r["np_info"] for r in refs])
    want_f = np.concatenate([r["np_features"] for r in refs])
    assert len(want_w) > 200
    assert same(res["np_info"].cpu().numpy(), want_i) and same(res["np_waves"].cpu().numpy(), want_w)
    assert same(res["np_features"].cpu().numpy(), want_f)
    with pytest.raises(Exception):
        pre.run(B.RecordBatch.from_host([strips[0][:4000]], [0]), strips=True)


# ------------------------------------------------------------------------- on-disk store + resume
def test_process_to_store_resumes(gpu, tmp_path):
    from ae_waveforms_b200.store import ArrayStore, process_to_store
    torch, B = gpu
    cases = {c: ecg_synth.synth_case(c, 60000 + 500 * c) for c in (3, 5, 6, 9, 10)}
    pre = B.Preprocessor()
    calls = []

    def records_of(ids):
        calls.append(list(ids))
        return [cases[c] for c in ids]

    root = str(tmp_path / "store")
    process_to_store(ArrayStore(root), pre, records_of, [3, 5], chunk_cases=2)          # a first, interrupted run
    process_to_store(ArrayStore(root), pre, records_of, sorted(cases), chunk_cases=2)   # the resumed run
    assert calls == [[3, 5], [6, 9], [10]]                                              # finished cases are not redone
    out = ArrayStore(root).finalize()
    ref = pre.run(B.RecordBatch.from_host([cases[c] for c in sorted(cases)], sorted(cases)), features=True)
    assert same(out["np_waves"], ref["np_waves"].cpu().numpy()) and same(out["np_info"], ref["np_info"].cpu().numpy())
    assert same(out["np_features"], ref["np_features"].cpu().numpy())
    assert same(np.load(os.path.join(root, "np_features.npy")), out["np_features"])
