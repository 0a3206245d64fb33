"""The oracle (oracle/ref_port.py) against the golden vectors frozen from the
reference's own functions (tests/golden/make_golden.py).  CPU only."""
import warnings

import numpy as np
import pytest
from scipy.signal import filtfilt

from conftest import unragged
from oracle import ecg_synth, np_pairwise, ref_import, ref_port as rp

warnings.simplefilter("ignore", RuntimeWarning)


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_filters_golden(golden):
    g = golden("filters")
    for tag in ("clean", "gaps"):
        z = rp.zscore(g[f"{tag}_x"])
        assert same(z, g[f"{tag}_z"])
        hp = rp.butterworth(z, "high", 1, 500, 4)
        assert same(hp, g[f"{tag}_hp"])
        assert same(rp.butterworth(hp, "low", 30, 500, 4), g[f"{tag}_lp"])
    xs = g["short_x"]
    assert same(rp.butterworth(xs, "high", 0.5, 500, 2), g["short_hp05_o2"])
    assert same(rp.butterworth(xs, "low", 40, 500, 3), g["short_lp40_o3"])
    assert same(rp.butterworth(xs, "low", 15, 500, 1), g["short_lp15_o1"])


def test_filtfilt_restatement_is_bit_exact(golden):
    g = golden("filters")
    x = g["clean_z"][:6000]
    for btype, fc in (("high", 1), ("low", 30)):
        b, a = rp.butter_ba(btype, fc)
        assert np.array_equal(rp.filtfilt_df2t(b, a, x), filtfilt(b, a, x))


def test_window_counts_golden(golden):
    g = golden("windows")
    for T, cnt, last in zip(g["lens"], g["counts"], g["last_end"]):
        ends = rp.window_ends(int(T))
        assert len(ends) == cnt
        assert (ends[-1] if len(ends) else -1) == last
    rec = (np.arange(35000, dtype=np.float64) * 0.5)[:, None]
    s, idx = rp.sample_from_record(rec, ["SNUADC/ECG_II"])
    assert same(s[:, ::997, 0], g["rec35000_samples"]) and same(idx, g["rec35000_idx"])
    assert idx.dtype == np.int64 and idx.shape == (3, 1)
    with pytest.raises(ValueError):
        rp.sample_from_record(rec, ["SNUADC/PLETH"])


def test_screen_golden(golden):
    g = golden("screen")
    scr = rp.window_screen(g["windows"])
    assert same(np.flatnonzero(scr["keep"]), g["kept_rows"])
    for got, want in zip(scr["peaks"], unragged(g["all_peaks"], g["all_peaks_off"])):
        assert same(got, want)
    assert same(scr["hr"], g["all_hr"]) and same(scr["rmssd"], g["all_rmssd"]) and same(scr["sd"], g["all_sd"])
    k = g["kept_rows"]
    assert same(scr["sd_per_beat"][k], g["kept_sd_per_beat"])
    df = rp.remove_samples(g["windows"][:, :, None], (np.arange(14) + 1).reshape(-1, 1) * 10000,
                           ["SNUADC/ECG_II"])
    assert same(df["ecg_hr"], g["kept_hr"]) and list(df.columns) == [
        "ecg_sample", "indices_samples", "ECG_R_Peaks", "ecg_hr", "ecg_hrv_rmssd", "ecg_sd", "ecg_sd/hr"]
    med = np.stack([rp.compute_median_sample_ecg(s, p, 500) for s, p in zip(df["ecg_sample"], df["ECG_R_Peaks"])])
    assert same(med, g["kept_median"])


def test_median_golden(golden):
    g = golden("median")
    w = g["window"]
    for name in ("edge", "even", "odd", "one", "far_right"):
        m, al = rp.compute_median_sample_ecg(w, list(g[f"{name}_peaks"]), 500, return_aligned=True)
        assert same(m, g[f"{name}_median"]) and same(al, g[f"{name}_aligned"])
    assert np.isnan(rp.compute_median_sample_ecg(w, [], 500)) and np.isnan(g["none_median"])


def test_features_golden(golden):
    g = golden("features")
    fid = {}
    for k, name in enumerate(rp.FIDUCIALS):
        fid[name] = [np.nan if v < 0 else int(v) for v in g["fiducials"][k]]
    rel = rp.ecg_features_relative_r(g["window"], fid, list(g["r_peaks"]), 500)
    assert same([rel[n]["x"] for n in rp.FIDUCIALS], g["rel_x"])
    assert same([rel[n]["y"] for n in rp.FIDUCIALS], g["rel_y"])
    morph = rp.compute_ecg_morphology_features(rel, len(g["r_peaks"]))
    assert "ECG_R_Peaks" in rel                      # the reference mutates its argument
    assert same([morph[n] for n in rp.MORPHOLOGY], g["morph"])
    assert same(rp.features_row_from_fiducials(g["window"], list(g["r_peaks"]), fid), g["row35"])


def test_recon_golden(golden):
    g = golden("recon")
    x_t = rp.beat_time_axis()
    assert same(x_t, g["x_t"])
    coords = unragged(g["coord_x"], g["coord_off"]), unragged(g["coord_y"], g["coord_off"])
    for s in range(len(g["fx"])):
        cx, cy, m = rp.extract_coordinates(g["fx"][s], g["fy"][s], x_t, g["baselines"][s])
        assert same(cx, coords[0][s]) and same(cy, coords[1][s]) and same(m, g["coord_nanmask"][s])
    for k, smap in enumerate(g["sigma_maps"]):
        rec = rp.reconstruct_all(g["fx"], g["fy"], g["baselines"], smap, x_t)
        assert same(rec, g["recon"][k])
        sse, mse = rp.sigma_objective(g["fx"], g["fy"], g["baselines"], g["test_true"], smap, x_t)
        want = g["mse_torch_f32"][k]
        assert (np.isnan(mse) and np.isnan(want)) or abs(mse - want) <= 1e-5 * abs(want)
    gl = np.concatenate([rp.gaussian_line(-80.0, -0.4, -38.0, -1.2, 21.0),
                         rp.gaussian_line(-38.0, -1.2, 0.0, -1e-10, 12.5),
                         rp.gaussian_line(10.5, -0.3, 61.9, -0.9, 8.0)])
    assert same(gl, g["gaussian_lines"])


def test_e2e_golden(golden):
    g = golden("e2e")
    out = rp.process_case(g["case"], case_id=9)
    assert same(out["filtered"], g["filtered"])
    assert len(out["indices"]) == int(g["n_windows"])
    assert same(out["np_info"][:, 1], g["kept_idx_end"]) and np.all(out["np_info"][:, 0] == 9)
    kept = [out["peaks"][i] for i in np.flatnonzero(out["keep"])]
    for got, want in zip(kept, unragged(g["kept_peaks"], g["kept_peaks_off"])):
        assert same(got, want)
    assert same(out["medians"], g["medians"]) and same(out["np_waves"], g["np_waves"])
    assert out["np_waves"].dtype == np.float32


def test_pairwise_restatement_matches_numpy():
    rng = np.random.default_rng(5)
    for n in (1, 7, 8, 9, 127, 128, 129, 1000, 4999, 5000, 10000, 65537):
        v = rng.standard_normal(n) * 10 ** rng.uniform(-3, 3)
        assert np_pairwise.pairwise_sum(v) == np.add.reduce(v)
        assert np_pairwise.slot_tree_sum(v) == np.add.reduce(v)


def test_synth_passes_reference_pipeline():
    out = rp.process_case(ecg_synth.synth_case(0, 300000))
    assert out["keep"].sum() == 29 and np.all(np.argmax(out["medians"], axis=1) == 106)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_port_against_live_reference():
    ref = ref_import.load()
    x = ecg_synth.synth_case(21, 45000, nan_gaps=2)
    z = ref.filters.zscore(x)
    assert same(z, rp.zscore(x))
    f = ref.filters.butterworth(ref.filters.butterworth(z, "high", 1), "low", 30)
    assert same(f, rp.butterworth(rp.butterworth(z, "high", 1), "low", 30))


def test_parity_report_selfcheck():
    """oracle.parity.compare on a GPU-shaped result built from the oracle itself: nothing differs; a flipped
    hazard window and a moved fiducial are counted where they belong."""
    from oracle import parity
    exp = parity.expected([7, 3], 60000, features=True, cores=2)
    keep, npk, pk, med, feat, fid, flags = [], [], [], [], [], [], []
    for e in exp:
        ks = 0
        for j, k in enumerate(e["keep"]):
            p = e["peaks"][j]
            row = np.full(72, -1, np.int32)
            row[:len(p)] = p
            f, fl = np.full((10, 72), -2147483648, np.int64), 3
            if k:
                fr = e["fid"][ks]
                if fr is None:
                    fl = 1 if e["valid"][ks] else 2
                else:
                    fl = 0
                    f[:, :len(p)] = np.where(np.isnan(fr), -2147483648, fr).astype(np.int64)
                if fl != 2:
                    med.append(e["medians"][ks]); feat.append(e["rows"][ks])
                ks += 1
            keep.append(k); npk.append(len(p)); pk.append(row); fid.append(f); flags.append(fl)
    res = dict(keep=np.array(keep), n_peaks=np.array(npk), peaks=np.stack(pk), medians=np.stack(med),
               np_features=np.stack(feat), fiducials=np.stack(fid), delineation_flags=np.array(flags))
    rep = parity.compare(exp, res)
    assert rep["windows"] == 10 and rep["mask_flips"] == 0 and rep["windows_rpeaks_differ"] == 0
    assert rep["median_max_rel_err"] == 0 and rep["features_max_rel_err"] == 0
    assert rep["fiducials_identical_fraction"] == 1.0 and rep["delineation_fate_differs"] == 0
    assert rep["windows_fiducials_identical"] == rep["windows_delineated"] > 0
    w = int(np.flatnonzero(res["delineation_flags"] == 0)[0])
    res["fiducials"][w, 3, 1] += 1                          # one Q peak moved by a sample
    rep = parity.compare(exp, res)
    assert rep["fiducials_identical"] == rep["fiducials_compared"] - 1
    assert rep["windows_fiducials_identical"] == rep["windows_delineated"] - 1
