"""Freeze golden vectors by running the REFERENCE'S OWN functions.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

Imports the reference modules through ``oracle.ref_import`` (matplotlib stub,
neurokit2 -> oracle.nk_peaks shim) and stores inputs and
outputs as small ``.npz`` files next to this script.  The reference has no
tests or fixtures of its own, so these files are the
pinned known answers for the oracle AND for the CUDA path.  Every array below
is produced by reference code, never by oracle/ref_port.py.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ecg_synth, ref_import  # noqa: E402

warnings.simplefilter("ignore")
TRACKS = ["SNUADC/ECG_II"]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def ragged(list_of_arrays, dtype):
    lens = np.array([len(a) for a in list_of_arrays], dtype=np.int64)
    flat = np.concatenate([np.asarray(a, dtype=dtype) for a in list_of_arrays]) if len(lens) else np.zeros(0, dtype)
    return flat, np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)


def main():
    ref = ref_import.load()
    F, S, R, M = ref.filters, ref.sample_from_record, ref.remove_samples, ref.median_sample
    FE, SR = ref.features_ecg, ref.signal_reconstruction

    # ---- filters: clean record, and a record with NaN gaps and short islands (appendix A.1)
    x = ecg_synth.synth_case(1, 30000)
    xn = x.copy()
    xn[:7] = np.nan                       # leading NaNs
    xn[1000:1003] = np.nan                # 3-sample gap
    xn[1003 + 5:1003 + 5 + 40] = np.nan   # leaves a 5-sample island  -> passthrough
    xn[5000:5000 + 15] = np.nan
    xn[5015 + 15:5015 + 15 + 9] = np.nan  # 15-sample island -> passthrough (<= 15)
    xn[5039 + 16:5039 + 16 + 9] = np.nan  # 16-sample island -> filtered
    xn[12000:19000] = np.nan              # long gap
    xn[-3:] = np.nan                      # trailing NaNs
    out = {}
    for tag, sig in (("clean", x), ("gaps", xn)):
        z = F.zscore(sig)
        hp = F.butterworth(z, "high", 1, 500, 4)
        lp = F.butterworth(hp, "low", 30, 500, 4)
        out.update({f"{tag}_x": sig, f"{tag}_z": z, f"{tag}_hp": hp, f"{tag}_lp": lp})
    # other orders / cut-offs on a short clean piece
    xs = x[:4000]
    out["short_x"] = xs
    out["short_hp05_o2"] = F.butterworth(xs, "high", 0.5, 500, 2)
    out["short_lp40_o3"] = F.butterworth(xs, "low", 40, 500, 3)
    out["short_lp15_o1"] = F.butterworth(xs, "low", 15, 500, 1)
    save("filters", **out)

    # ---- windowing (appendix A.2)
    lens = [5000, 10000, 10001, 19999, 20000, 20001, 30000, 45000]
    counts, firsts = [], []
    for T in lens:
        rec = np.arange(T, dtype=np.float64)[:, None]
        s, idx = S.sample_from_record(rec, TRACKS, 500, 20)
        counts.append(len(idx))
        firsts.append(-1 if len(idx) == 0 else int(idx[-1, 0]))
        assert s.shape[1:] == (10000, 1)
    rec = (np.arange(35000, dtype=np.float64) * 0.5)[:, None]
    s, idx = S.sample_from_record(rec, TRACKS, 500, 20)
    save("windows", lens=np.array(lens), counts=np.array(counts), last_end=np.array(firsts),
         rec35000_samples=s[:, ::997, 0], rec35000_idx=idx)

    # ---- screening (remove_samples) on windows with different fates
    rng = np.random.default_rng(7)
    wins, raws = [], []
    a = ecg_synth.synth_case(2, 50000)                      # regular case
    b = ecg_synth.synth_case(3, 30000, hr_bpm=60.0, jitter=0.0)   # 20 beats / window hazard (appendix C.1)
    c = ecg_synth.synth_case(4, 30000, hr_bpm=27.0)         # HR < 30  -> removed
    d = ecg_synth.synth_case(5, 30000, hr_bpm=170.0)        # fast
    e = 0.05 * rng.standard_normal(30000)                    # noise only
    g = ecg_synth.synth_case(6, 30000, hr_bpm=95.0)
    g[14000:14010] = np.nan                                  # NaN window
    filtered = []
    for sig in (a, b, c, d, e, g):
        z = F.zscore(sig)
        f = F.butterworth(F.butterworth(z, "high", 1), "low", 30)
        filtered.append(f)
        s, idx = S.sample_from_record(f[:, None], TRACKS, 500, 20)
        raws.append(s[:, :, 0].copy())
        s[:, :, 0] = np.apply_along_axis(F.zscore, 1, s[:, :, 0])
        wins.append(s[:, :, 0])
    wins = np.concatenate(wins, axis=0)
    raws = np.concatenate(raws, axis=0)
    idx = (np.arange(len(wins)) + 1).reshape(-1, 1) * 10000
    df = R.remove_samples(wins[:, :, None].copy(), idx, TRACKS, 500, 150)
    kept_rows = (np.array([int(v[0]) for v in df["indices_samples"]]) // 10000) - 1
    pk, pk_off = ragged(list(df["ECG_R_Peaks"]), np.int64)
    # statistics for ALL non-NaN windows (the frame only holds the kept ones)
    all_hr, all_rm, all_sd, all_pk = [], [], [], []
    import neurokit2 as nk  # the shim installed by ref_import
    for w in wins:
        if np.isnan(w).any():
            all_hr.append(np.nan); all_rm.append(np.nan); all_sd.append(np.nan); all_pk.append([])
            continue
        p = [v for v in nk.ecg_findpeaks(w, sampling_rate=500)["ECG_R_Peaks"] if 0 < v < len(w)]
        all_pk.append(p)
        all_hr.append(R.estimate_hr([], p, 500, mod="max"))
        all_rm.append(R.rmssd(p, 500))
        all_sd.append(np.std(w))
    apk, apk_off = ragged(all_pk, np.int64)
    med = np.stack([M.compute_median_sample_ecg(s_, p_, 500) for s_, p_ in zip(df["ecg_sample"], df["ECG_R_Peaks"])])
    save("screen", windows=wins, windows_unscaled=raws, kept_rows=kept_rows, kept_peaks=pk, kept_peaks_off=pk_off,
         kept_hr=df["ecg_hr"].to_numpy(), kept_rmssd=df["ecg_hrv_rmssd"].to_numpy(),
         kept_sd=df["ecg_sd"].to_numpy(), kept_sd_per_beat=df["ecg_sd/hr"].to_numpy(),
         all_peaks=apk, all_peaks_off=apk_off, all_hr=np.array(all_hr), all_rmssd=np.array(all_rm),
         all_sd=np.array(all_sd), kept_median=med)

    # ---- median beats with edge peaks (appendix A.4)
    w0 = wins[0]
    cases = {
        "edge": [50, 400, 800, 1200, 9900],           # left- and right-truncated beats
        "even": [300, 700, 1100, 1500],
        "odd": [300, 700, 1100, 1500, 1900],
        "one": [5000],
        "far_right": [9990, 9995],                    # columns >= 10 are all-NaN
    }
    out = {"window": w0}
    for name, p in cases.items():
        m, al = M.compute_median_sample_ecg(w0, p, 500, return_aligned=True)
        out[f"{name}_peaks"] = np.array(p)
        out[f"{name}_median"] = m
        out[f"{name}_aligned"] = al
    out["none_median"] = np.asarray(M.compute_median_sample_ecg(w0, [], 500))
    save("median", **out)

    # ---- features from given fiducials (a16-a18)
    r_peaks = [v for v in all_pk[0]]
    rngf = np.random.default_rng(11)
    names = ["ECG_P_Peaks", "ECG_P_Onsets", "ECG_P_Offsets", "ECG_Q_Peaks", "ECG_R_Onsets",
             "ECG_R_Offsets", "ECG_S_Peaks", "ECG_T_Peaks", "ECG_T_Onsets", "ECG_T_Offsets"]
    rel_off = [-100, -125, -75, -12, -25, 30, 15, 125, 80, 170]
    fid = {}
    fid_arr = np.full((10, len(r_peaks)), -1, dtype=np.int64)
    for k, (nm, off) in enumerate(zip(names, rel_off)):
        lst = []
        for j, r in enumerate(r_peaks):
            v = int(r + off + rngf.integers(-3, 4))
            if rngf.random() < 0.15 or not (0 <= v < len(w0)):
                lst.append(np.nan)
            else:
                lst.append(v); fid_arr[k, j] = v
        fid[nm] = lst
    rel = FE.ecg_features_relative_r(w0, fid, r_peaks, 500)
    morph = FE.compute_ecg_morphology_features(rel, len(r_peaks))
    medf = M.compute_median_features_ecg(rel, morph)
    row = []
    for nm in names:
        row += [medf[nm]["x"], medf[nm]["y"]]
    morph_names = list(morph.keys())
    row += [medf[nm] for nm in morph_names]
    save("features", window=w0, r_peaks=np.array(r_peaks), fiducials=fid_arr,
         rel_x=np.array([rel[nm]["x"] for nm in names]), rel_y=np.array([rel[nm]["y"] for nm in names]),
         morph=np.array([morph[nm] for nm in morph_names]), row35=np.array(row, dtype=np.float64))

    # ---- Gaussian reconstruction (a11-a14, appendix A.6 / A.7)
    x_t = np.linspace(-106, 213, 320) / 500 * 1000
    tt, fx, fy, bl = ecg_synth.synth_fiducials(48, seed=3, nan_frac=0.08)
    # hand-made hazards: duplicates, out-of-order, out-of-range fiducials, NaN runs
    fx[0, 3] = fx[0, 4]                      # duplicate x
    fx[1, 1] = fx[1, 2] = fx[1, 3]           # triple duplicate
    fx[2, 3], fx[2, 4] = fx[2, 4], fx[2, 3]  # out of order
    fx[3, 0] = -300.0                        # left of x_t[0]
    fx[4, 10] = 500.0                        # right of x_t[-1]
    fx[5, 1:4] = np.nan                      # NaN run
    fy[6, 9] = np.nan
    fx[7, 6] = 0.5                           # non-integer, delta x <= 1 against R
    maps = np.array([[1, 1, 1, 2, 2, 3, 3, 3, 4, 2, 2, 1],
                     [1, 2, 2, 3, 2, 3, 3, 2, 3, 2, 2, 1],
                     [5, 4, 3, 2, 1, 1, 2, 3, 4, 5, 5, 5],
                     [1.5, 2.25, 3.1, 4.9, 1.01, 2.0, 3.3, 2.7, 1.9, 4.4, 3.6, 2.2]], dtype=np.float64)
    recon = np.array([[SR.reconstruct_gaussian(fx[s], fy[s], x_t, bl[s], m) for s in range(len(fx))]
                      for m in maps])
    coords = [SR.extract_coordinates(fx[s], fy[s], x_t, bl[s]) for s in range(len(fx))]
    cx, cx_off = ragged([c_[0] for c_ in coords], np.float64)
    cy, _ = ragged([c_[1] for c_ in coords], np.float64)
    cm = np.array([c_[2] for c_ in coords])
    gl = np.concatenate([SR.gaussian_line(-80.0, -0.4, -38.0, -1.2, 21.0),
                         SR.gaussian_line(-38.0, -1.2, 0.0, -1e-10, 12.5),
                         SR.gaussian_line(10.5, -0.3, 61.9, -0.9, 8.0)])
    import torch
    mse32 = np.array([torch.nn.MSELoss()(torch.tensor(r_, dtype=torch.float32), torch.tensor(tt)).item()
                      for r_ in recon])
    save("recon", x_t=x_t, test_true=tt, fx=fx, fy=fy, baselines=bl, sigma_maps=maps, recon=recon,
         coord_x=cx, coord_y=cy, coord_off=cx_off, coord_nanmask=cm, gaussian_lines=gl, mse_torch_f32=mse32)

    # ---- end to end: one short case through the reference functions with the glue of
    #      data_extract_vitaldb.py:53-106
    case = ecg_synth.synth_case(9, 65000, nan_gaps=1)
    case[:11] = np.nan
    good = ~np.isnan(case)
    first, last = int(np.argmax(good)), len(good) - int(np.argmax(good[::-1]))
    tr = case[first:last]
    f = F.butterworth(F.butterworth(F.zscore(tr), "high", 1, 500, 4), "low", 30, 500, 4)
    s, idx = S.sample_from_record(f[:, None].copy(), TRACKS, 500, 20)
    s[:, :, 0] = np.apply_along_axis(F.zscore, 1, s[:, :, 0])
    df = R.remove_samples(s, idx, TRACKS, 500, 150)
    meds = np.stack([M.compute_median_sample_ecg(s_, p_, 500) for s_, p_ in zip(df["ecg_sample"], df["ECG_R_Peaks"])])
    pk, pk_off = ragged(list(df["ECG_R_Peaks"]), np.int64)
    save("e2e", case=case, filtered=f, n_windows=np.array(len(idx)),
         kept_idx_end=np.array([int(v[0]) for v in df["indices_samples"]]),
         kept_peaks=pk, kept_peaks_off=pk_off, medians=meds, np_waves=meds.astype(np.float32))


if __name__ == "__main__":
    main()
