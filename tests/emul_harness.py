"""ctypes harness for the HOST EMULATION build of the kernel sources
(tests/emul/libaewave_emul.so, built from ae_waveforms_b200/csrc/*.cu with -DAEW_EMUL).

Test infrastructure only: it lets the CPU test-suite execute the very kernel code that
nvcc compiles for sm_100a (one host thread per CUDA thread) against the golden vectors,
so that logic errors are caught before a GPU is involved.  "Device" memory is NumPy."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
sys.path.insert(0, ROOT)
from ae_waveforms_b200 import _cabi  # noqa: E402

LL = ctypes.c_longlong
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run([os.path.join(HERE, "emul", "build_emul.sh")], check=True, capture_output=True)
        _lib = _cabi.load(os.path.join(HERE, "emul", "libaewave_emul.so"))
    return _lib


def ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def flatten(records):
    lens = [len(r) for r in records]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = np.concatenate([np.asarray(r, dtype=np.float64) for r in records]) if records else np.zeros(0)
    return np.ascontiguousarray(flat), off


def zscore_stats(records, fast=False):
    L = lib()
    flat, off = flatten(records)
    n = len(records)
    beg, end = np.ascontiguousarray(off[:-1]), np.ascontiguousarray(off[1:])
    blk = np.zeros(n + 1, np.int64)
    nb = L.aew_zscore_plan(ptr(beg), ptr(end), n, ptr(blk))
    wsb = L.aew_zscore_workspace_bytes(n, nb)
    ws = np.zeros(wsb, np.uint8)
    mean, sd, cnt = np.zeros(n), np.zeros(n), np.zeros(n, np.int64)
    entry = L.aew_zscore_stats_fast_f64 if fast else L.aew_zscore_stats_f64
    rc = entry(ptr(flat), ptr(beg), ptr(end), ptr(blk), n, nb, ptr(mean), ptr(sd), ptr(cnt), ptr(ws), wsb, None)
    assert rc == 0, rc
    return mean, sd, cnt


def trim(records):
    L = lib()
    flat, off = flatten(records)
    n = len(records)
    out = np.zeros(2 * n, np.int64)
    rc = L.aew_trim_nan_edges_f64(ptr(flat), ptr(np.ascontiguousarray(off[:-1])), ptr(np.ascontiguousarray(off[1:])),
                                  n, ptr(out), None)
    assert rc == 0
    return out.reshape(n, 2) - off[:-1, None]


def filter_tables(btype, fc, fs=500, order=4):
    from scipy.signal import butter, lfilter_zi
    L = lib()
    b, a = butter(order, fc / (0.5 * fs), btype)
    b, a = np.ascontiguousarray(b), np.ascontiguousarray(a)
    zi = np.ascontiguousarray(lfilter_zi(b, a))
    tab = np.zeros(L.aew_filt_tables_len())
    rc = L.aew_filt_tables(ptr(b), ptr(a), ptr(zi), order, ptr(tab))
    return tab, rc


def filtfilt(records, stages, mode, zscore=False, max_seg=None):
    """stages: list of (btype, fc, order).  Returns (list of outputs, header)."""
    L = lib()
    flat, off = flatten(records)
    n, total = len(records), len(flat)
    beg, end = np.ascontiguousarray(off[:-1]), np.ascontiguousarray(off[1:])
    max_seg = max_seg or n + 64
    wsb = L.aew_filt_workspace_bytes(n, total, max_seg)
    ws = np.zeros(wsb + 256, np.uint8)
    out, tmp = np.full(total, -777.0), np.zeros(total)
    mean = sd = cnt = None
    if zscore:
        mean, sd, cnt = zscore_stats(records)
    rc = L.aew_segments_find_f64(ptr(flat), ptr(beg), ptr(end), n, total, ptr(cnt), ptr(out), ptr(ws), wsb, max_seg,
                                 None)
    assert rc == 0, rc
    src, m, s = flat, mean, sd
    for btype, fc, order in stages:
        tab, trc = filter_tables(btype, fc, 500, order)
        assert trc == 0
        if mode == 2:                                      # one launch: src -> a buffer that is not src
            assert L.aew_filt_fused_ok(ptr(tab)) == 1
            dst = tmp if src is out else out
            nanmask = np.isnan(flat)
            rc = L.aew_filtfilt_f64(ptr(src), ptr(m), ptr(s), ptr(dst), ptr(out if dst is tmp else tmp), ptr(tab),
                                    ptr(tab), n, total, ptr(ws), wsb, max_seg, 2, None)
            assert rc == 0, rc
            dst[nanmask] = np.nan                          # (the caller plans the NaN fill into the final buffer)
            if dst is tmp:
                out, tmp = tmp, out
        else:
            rc = L.aew_filtfilt_f64(ptr(src), ptr(m), ptr(s), ptr(out), ptr(tmp), ptr(tab), ptr(tab), n, total, ptr(ws),
                                    wsb, max_seg, mode, None)
            assert rc == 0, rc
        src, m, s = out, None, None
    hdr = ws[(-ws.ctypes.data) % 128:][:64].view(np.int64).copy()
    return [out[off[i]:off[i + 1]] for i in range(n)], hdr


def screen(flat, starts, Lw=10000, mode=0, zscore_windows=True, want_z=False, fs=500, pvc=150.0, sdt=0.05,
           lazy_sd=0, fast_stats=0):
    L = lib()
    flat = np.ascontiguousarray(flat, dtype=np.float64)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    n, maxp = len(starts), L.aew_windows_max_peaks()
    sec = int(0.8 * (60 / 75) * fs)                        # median_sample.py:20-25
    pre = int(1 / 3 * sec)
    o = dict(keep=np.zeros(n, np.uint8), has_nan=np.zeros(n, np.uint8), n_peaks=np.zeros(n, np.int32),
             peaks=np.zeros((n, maxp), np.int32), stats=np.zeros((n, 4)), median=np.full((n, sec), -5.0),
             waves=np.zeros((n, sec), np.float32), zwin=np.zeros((n, Lw)), zparams=np.zeros((n, 2)),
             status=np.zeros(1, np.int32),
             hazards=np.zeros(1, np.uint64))
    wsb = int(L.aew_windows_workspace_bytes(n))
    ws = np.zeros(wsb + 16, np.uint8)
    zpark = o["zwin"] if want_z else np.zeros((n, Lw))    # the z-scored windows are always parked (beat / median kernels)
    p = _cabi.K3Params(
        y=flat.ctypes.data, win_start=starts.ctypes.data, n_win=n, L=Lw, smooth_k=int(np.rint(0.1 * fs)),
        avg_k=int(np.rint(0.75 * fs)), refractory=int(np.rint(0.3 * fs)), pre=pre, post=sec - pre, hr_num=60.0 * fs,
        rr_div=fs * 1000.0, hr_lo=30.0, hr_hi=180.0, pvc_threshold=pvc, sd_threshold=sdt, mode=mode,
        zscore_windows=1 if zscore_windows else 0, keep=o["keep"].ctypes.data, has_nan=o["has_nan"].ctypes.data,
        n_peaks=o["n_peaks"].ctypes.data, peaks=o["peaks"].ctypes.data, stats=o["stats"].ctypes.data,
        median=o["median"].ctypes.data, waves=o["waves"].ctypes.data,
        zwin=None if fast_stats else (zpark.ctypes.data if zscore_windows else (o["zwin"].ctypes.data if want_z else None)),
        zparams=o["zparams"].ctypes.data,
        status=o["status"].ctypes.data, hazard_count=o["hazards"].ctypes.data, lazy_sd=lazy_sd,
        workspace=ws.ctypes.data, workspace_bytes=wsb, fast_stats=fast_stats)
    rc = L.aew_windows_beats_f64(ctypes.byref(p), None)
    assert rc == 0, rc
    return o


def median_beats(window, peaks, want_aligned=True):
    L = lib()
    w = np.ascontiguousarray(window, dtype=np.float64)
    pk = np.ascontiguousarray(peaks, dtype=np.int32)
    off = np.array([0, len(pk)], np.int64)
    med = np.zeros(320)
    al = np.zeros((len(pk), 320)) if want_aligned else None
    rc = L.aew_median_beats_f64(ptr(w), 1, len(w), ptr(pk), ptr(off), len(pk), 106, 214, ptr(med), ptr(al), None)
    assert rc == 0, rc
    return med, al


def recon(fx, fy, bl, target, xt, cands, want_pred=False):
    L = lib()
    fx, fy, bl, xt = [np.ascontiguousarray(a, dtype=np.float64) for a in (fx, fy, bl, xt)]
    target = None if target is None else np.ascontiguousarray(target, dtype=np.float32)
    cands = np.ascontiguousarray(cands, dtype=np.float64).reshape(-1, 12)
    nb, nc = len(fx), len(cands)
    sse = np.zeros(nc)
    pred = np.zeros((nb, len(xt))) if want_pred else None
    wsb = L.aew_recon_workspace_bytes(nb, nc)
    ws = np.zeros(wsb, np.uint8)
    rc = L.aew_recon_sse_f64(ptr(fx), ptr(fy), ptr(bl), ptr(target), nb, ptr(xt), len(xt), ptr(cands), nc, ptr(sse),
                             ptr(pred), ptr(ws), wsb, None)
    assert rc == 0, rc
    return sse, pred


def recon_grid(fx, fy, bl, target, xt, vals, idx):
    L = lib()
    fx, fy, bl, xt, vals = [np.ascontiguousarray(a, dtype=np.float64) for a in (fx, fy, bl, xt, vals)]
    target = np.ascontiguousarray(target, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.uint8).reshape(-1, 12)
    nb, nc = len(fx), len(idx)
    sse = np.zeros(nc)
    wsb = L.aew_recon_tables_workspace_bytes(nb)
    ws = np.zeros(wsb, np.uint8)
    rc = L.aew_recon_sse_grid_f64(ptr(fx), ptr(fy), ptr(bl), ptr(target), nb, ptr(xt), len(xt), ptr(vals), len(vals),
                                  ptr(idx), nc, ptr(sse), ptr(ws), wsb, None)
    assert rc == 0, rc
    return sse


FID_NAN = -2147483648


def fiducial_features(windows, peaks_list, fid_list, valid=None, fs=500.0):
    """windows (n, L); peaks_list: per window int list; fid_list: per window (10, n_beats) float array, NaN = missing."""
    L = lib()
    w = np.ascontiguousarray(windows, dtype=np.float64)
    n, Lw = w.shape
    maxp = L.aew_windows_max_peaks()
    pk = np.full((n, maxp), -1, np.int32)
    npk = np.zeros(n, np.int32)
    fid = np.full((n, 10, maxp), FID_NAN, np.int32)
    for i, (p, f) in enumerate(zip(peaks_list, fid_list)):
        npk[i] = len(p)
        pk[i, :len(p)] = p
        f = np.asarray(f, dtype=np.float64).reshape(10, -1)
        fid[i, :, :f.shape[1]] = np.where(np.isnan(f), FID_NAN, f).astype(np.int64)
    off = (np.arange(n, dtype=np.int64) * Lw)
    out = np.full((n, 35), -9.0)
    status = np.zeros(1, np.int32)
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    rc = L.aew_fiducial_features_f64(ptr(w), ptr(off), n, Lw, ptr(pk), ptr(npk), maxp, ptr(fid), ptr(v), float(fs),
                                     ptr(out), ptr(status), None)
    assert rc == 0, rc
    return out, int(status[0])


def delineate(windows, peaks_list, fs=500.0, select=None):
    """K7 on (n, L) windows -> fid (n, 10, max_peaks) int32 (FID_NAN = missing), flags (n,) uint8."""
    L = lib()
    w = np.ascontiguousarray(windows, dtype=np.float64)
    n, Lw = w.shape
    maxp = L.aew_windows_max_peaks()
    pk = np.full((n, maxp), -1, np.int32)
    npk = np.zeros(n, np.int32)
    for i, p in enumerate(peaks_list):
        npk[i] = len(p)
        pk[i, :len(p)] = p
    off = np.arange(n, dtype=np.int64) * Lw
    fid = np.zeros((n, 10, maxp), np.int32)
    flags = np.full(n, 3, np.uint8)
    wsb = L.aew_delineate_workspace_bytes(n, Lw, float(fs))
    ws = np.zeros(wsb, np.uint8)
    sel = None if select is None else np.ascontiguousarray(select, dtype=np.uint8)
    rc = L.aew_delineate_f64(ptr(w), ptr(off), n, Lw, float(fs), ptr(pk), ptr(npk), maxp, ptr(sel), ptr(fid),
                             ptr(flags), ptr(ws), wsb, None)
    assert rc == 0, rc
    return fid, flags


def widen(src, gain=1.0, offset=0.0, sentinel=None):
    L = lib()
    src = np.ascontiguousarray(src)
    code = {np.dtype(np.float32): 0, np.dtype(np.int16): 1}[src.dtype]
    out = np.full(len(src), -7.0)
    rc = L.aew_widen_f64(ptr(src), code, len(src), float(gain), float(offset), 0 if sentinel is None else 1,
                         0 if sentinel is None else int(sentinel), ptr(out), None)
    assert rc == 0, rc
    return out


def fetch_host(src):
    """aew_fetch_host under emulation: a 16-byte aligned copy of ``src``'s bytes."""
    L = lib()
    raw = np.ascontiguousarray(src).reshape(-1).view(np.uint8)
    buf = np.zeros(len(raw) + 32, np.uint8)
    a = (-buf.ctypes.data) % 16
    h = buf[a:a + len(raw)]
    h[:] = raw
    obuf = np.full(len(raw) + 32, 0xAB, np.uint8)
    b = (-obuf.ctypes.data) % 16
    out = obuf[b:b + len(raw)]
    rc = L.aew_fetch_host(h.ctypes.data if len(raw) else None, len(raw), out.ctypes.data if len(raw) else None, None)
    assert rc == 0, rc
    assert np.all(obuf[b + len(raw):] == 0xAB)
    return out.copy()


def beat_metrics(recon, median, cr=16.0):
    L = lib()
    recon = np.ascontiguousarray(recon, dtype=np.float64)
    median = np.ascontiguousarray(median, dtype=np.float64)
    n, nt = recon.shape
    out = [np.zeros(n, np.float32) for _ in range(3)]
    rc = L.aew_beat_metrics_f32(ptr(recon), ptr(median), n, nt, float(cr), ptr(out[0]), ptr(out[1]), ptr(out[2]), None)
    assert rc == 0, rc
    return out


def feature_stages(window, peaks, fid, fs=500.0):
    """The three K6 stage entries (the per-call drop-ins' kernels): returns rel x, y (10, n), morph (15, n),
    medians (35,) in np_features order."""
    L = lib()
    w = np.ascontiguousarray(window, dtype=np.float64)
    pk = np.ascontiguousarray(peaks, dtype=np.int32)
    nb = len(pk)
    f = np.asarray(fid, dtype=np.float64).reshape(10, nb)
    idx = np.ascontiguousarray(np.where(np.isnan(f), FID_NAN, f).astype(np.int64).astype(np.int32))
    xy = np.zeros((10, 2, nb))
    status = np.zeros(1, np.int32)
    assert L.aew_fiducial_relative_f64(ptr(w), len(w), ptr(pk), ptr(idx), 10, nb, float(fs), ptr(xy), ptr(status),
                                       None) == 0 and status[0] == 0
    X = np.ascontiguousarray(np.concatenate([xy[:, 0, :], np.zeros((1, nb))]))
    Y = np.ascontiguousarray(np.concatenate([xy[:, 1, :], np.zeros((1, nb))]))
    morph = np.zeros((15, nb))
    assert L.aew_morphology_f64(ptr(X), ptr(Y), nb, ptr(morph), None) == 0
    rows = np.ascontiguousarray(np.concatenate([xy.reshape(20, nb), morph]))
    med, scratch = np.zeros(35), np.zeros_like(rows)
    assert L.aew_nanmedian_rows_f64(ptr(rows), 35, nb, ptr(scratch), ptr(med), None) == 0
    return xy[:, 0, :], xy[:, 1, :], morph, med


def peak_stats(min_peaks, max_peaks, fs=500, mode=2):
    L = lib()
    lo = np.ascontiguousarray(min_peaks, dtype=np.int64)
    hi = np.ascontiguousarray(max_peaks, dtype=np.int64)
    scratch = np.zeros(2 * (len(lo) + len(hi)) + 8)
    out = np.zeros(2)
    assert L.aew_peak_stats_f64(ptr(lo) if len(lo) else None, len(lo), ptr(hi) if len(hi) else None, len(hi), mode,
                                float(60 * fs), float(fs * 1000), ptr(scratch), ptr(out), None) == 0
    return out
