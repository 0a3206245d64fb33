"""Parity report of the CUDA path against the oracle on host-generated cases (TEST INFRASTRUCTURE).

SURVEY.md section 8d: parity is checked on HOST-generated inputs (host exp / sin differ from the device's bitwise)
-- C1 and a >= 64-case subsample of C2.  ``expected`` runs the oracle (``ref_port.process_case`` + the delineator
restatement) over the cases on a process pool; ``compare`` lines the windows of one GPU result up with it and
counts, per window: R-peak lists that differ, keep / remove decisions that differ (and whether the window sits on
the ``n_peaks == 20`` rounding hazard of SURVEY.md C.1), the largest relative error of the median beats, the
fraction of identical fiducial indices and the largest relative error of the 35 features where the fiducials agree.
Used by ``tests/test_gpu_parity.py`` (hard bounds) and by ``bench.py`` (the ``parity`` block of the JSON line).
"""
from __future__ import annotations

import os
import warnings

import numpy as np

FS = 500


def _worker(args):
    case_id, n_samples, features, hr = args
    warnings.simplefilter("ignore")
    from . import ecg_synth, nk_delineate as nd, ref_port as rp
    x = ecg_synth.synth_case(case_id, n_samples, hr_bpm=hr)
    r = rp.process_case(x, case_id=case_id)
    out = dict(case_id=case_id, x=x, keep=np.asarray(r["keep"], bool), peaks=[np.asarray(p) for p in r["peaks"]],
               medians=r["medians"], sd=np.asarray(r.get("sd", [])), hr=np.asarray(r.get("hr", [])),
               rmssd=np.asarray(r.get("rmssd", [])))
    if features:
        fids, rows, valid = [], [], []
        for i in np.flatnonzero(out["keep"]):
            win, pk = r["windows"][i], [int(p) for p in r["peaks"][i]]
            status, fid = nd.window_fiducials(win, pk, FS)           # (as ref_port.window_features, fiducials kept)
            if status == "ok":
                named = {k: [np.nan if np.isnan(v) else int(v) for v in fid[q]] for q, k in enumerate(rp.FIDUCIALS)}
                rows.append(rp.features_row_from_fiducials(win, pk, named, FS))
                fids.append(np.asarray(fid, dtype=np.float64))
            else:
                rows.append(np.full(35, np.nan))
                fids.append(None)
            valid.append(status != "drop")
        out.update(fid=fids, rows=np.stack(rows) if rows else np.zeros((0, 35)), valid=np.asarray(valid, bool))
    return out


def expected(case_ids, n_samples, features=True, cores=None, hr=None):
    """Oracle outputs per case (list of dicts, in ``case_ids`` order), on ``cores`` processes."""
    import multiprocessing as mp
    cores = cores or max(1, min(len(case_ids), len(os.sched_getaffinity(0))))
    jobs = [(int(c), int(n_samples), bool(features), None if hr is None else float(hr[k]))
            for k, c in enumerate(case_ids)]
    if cores == 1:
        return [_worker(j) for j in jobs]
    with mp.get_context("spawn").Pool(cores) as pool:
        return pool.map(_worker, jobs, chunksize=1)


def compare(exp, res, hazard_n=20, features=True, fid_nan=-2147483648):
    """``exp``: output of ``expected``; ``res``: dict of NumPy arrays of ONE Preprocessor.run over the same cases
    in the same order (keep, n_peaks, peaks, medians of the kept rows in order, np_features, fiducials,
    delineation_flags).  Returns the counters described in the module docstring."""
    keep_g = np.asarray(res["keep"]).astype(bool)
    n_peaks_g, peaks_g = np.asarray(res["n_peaks"]), np.asarray(res["peaks"])
    med_g = np.asarray(res["medians"])
    final = keep_g                                       # window -> row of the compacted GPU arrays
    w = 0
    out = dict(cases=len(exp), windows=0, windows_rpeaks_differ=0, mask_flips=0, mask_flips_off_hazard=0,
               hazard_windows=0, rows_reference=0, rows_gpu=int(keep_g.sum()), median_max_rel_err=0.0,
               fiducials_compared=0, fiducials_identical=0, windows_fiducials_identical=0, windows_delineated=0,
               features_max_rel_err=0.0, features_compared=0, delineation_fate_differs=0)
    feat_g = fid_g = flags_g = None
    if features and "np_features" in res:
        flags_g = np.asarray(res["delineation_flags"])
        fid_g = np.asarray(res["fiducials"])
        feat_g = np.asarray(res["np_features"])
        final = keep_g & (flags_g != 2)                  # rows follow the FINAL keep (screen keep minus dropped windows)
    row_of = frow_of = np.cumsum(final) - 1
    for e in exp:
        k_ref = e["keep"]
        kept_seen = 0
        for j in range(len(k_ref)):
            pr = e["peaks"][j]
            pg = peaks_g[w, :n_peaks_g[w]]
            same_peaks = len(pr) == len(pg) and np.array_equal(pr, pg)
            nan_window = len(pr) == 0 and n_peaks_g[w] == 0
            out["windows"] += 1
            if not same_peaks and not nan_window:
                out["windows_rpeaks_differ"] += 1
            on_hazard = len(pr) == hazard_n
            out["hazard_windows"] += int(on_hazard)
            if bool(k_ref[j]) != bool(keep_g[w]):
                out["mask_flips"] += 1
                out["mask_flips_off_hazard"] += int(not on_hazard)
            if k_ref[j]:
                out["rows_reference"] += 1
                if keep_g[w] and not final[w]:
                    f_ref = e["fid"][kept_seen] if feat_g is not None else None
                    out["delineation_fate_differs"] += int(feat_g is not None and
                                                           (f_ref is not None or e["valid"][kept_seen]))
                elif keep_g[w]:
                    m_ref, m_gpu = e["medians"][kept_seen], med_g[row_of[w]]
                    scale = np.nanmax(np.abs(m_ref))
                    err = np.nanmax(np.abs(m_gpu - m_ref)) / scale if scale > 0 else 0.0
                    out["median_max_rel_err"] = max(out["median_max_rel_err"], float(err))
                    if feat_g is not None:
                        f_ref = e["fid"][kept_seen]
                        fate_ref = 0 if f_ref is not None else (1 if e["valid"][kept_seen] else 2)
                        if fate_ref != int(flags_g[w]):
                            out["delineation_fate_differs"] += 1
                        elif f_ref is not None and same_peaks:
                            out["windows_delineated"] += 1
                            fg = fid_g[w, :, :len(pr)].astype(np.float64)
                            fg[fid_g[w, :, :len(pr)] == fid_nan] = np.nan
                            eq = (fg == f_ref) | (np.isnan(fg) & np.isnan(f_ref))
                            out["fiducials_compared"] += eq.size
                            out["fiducials_identical"] += int(eq.sum())
                            if eq.all():
                                out["windows_fiducials_identical"] += 1
                                a, b = feat_g[frow_of[w]], e["rows"][kept_seen]
                                both = ~(np.isnan(a) | np.isnan(b))
                                if (np.isnan(a) != np.isnan(b)).any():
                                    out["features_max_rel_err"] = float("inf")
                                if both.any():
                                    # durations are differences of exact multiples of 2 ms: compare on the scale of
                                    # the row, not of an entry that may be exactly zero
                                    den = np.maximum(np.abs(b[both]), 1e-3 * np.abs(b[both]).max() + 1e-300)
                                    out["features_max_rel_err"] = max(out["features_max_rel_err"],
                                                                      float(np.max(np.abs(a[both] - b[both]) / den)))
                                    out["features_compared"] += int(both.sum())
                kept_seen += 1
            w += 1
    assert w == len(keep_g), (w, len(keep_g))
    out["fiducials_identical_fraction"] = (out["fiducials_identical"] / out["fiducials_compared"]
                                           if out["fiducials_compared"] else None)
    return out


def gpu_result_to_numpy(res, features=True):
    keys = ["keep", "n_peaks", "peaks", "medians"] + (["np_features", "fiducials", "delineation_flags"]
                                                     if features else [])
    return {k: res[k].cpu().numpy() for k in keys if res.get(k) is not None}
