"""Drop-in for functions_traditional/signal_reconstruction.py (reference lines 5-81).
Every function runs the K5 device code; use
``ae_waveforms_b200.sigma_search.SigmaSearch`` for whole test sets and candidate lists."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _cabi
from ..sigma_search import SigmaSearch


def degenerate_beat_error(x_coordinates, nan_mask, x_t, n_sigma):
    """What the reference's reconstruct_gaussian RAISES for a beat, or None (signal_reconstruction.py:23-42).
    Its segment loop walks an index through ``nan_mask`` and ``sigma_map`` and hands the collected samples to
    ``np.interp``; the index can run past either list (IndexError) and the collection can stay empty (np.interp's
    ValueError).  Only that index walk is retraced here -- integers and three comparisons per segment, no arithmetic
    of the reconstruction -- so that a caller catching the reference's exceptions sees the same ones."""
    xc = np.asarray(x_coordinates, dtype=np.float64)
    mask = np.asarray(nan_mask, dtype=bool)
    at, drew = 0, False
    for seg in range(len(xc) - 1):
        if at >= n_sigma:
            return IndexError("list index out of range")
        while True:
            if at + 1 >= len(mask):
                return IndexError(f"index {at + 1} is out of bounds for axis 0 with size {len(mask)}")
            if not mask[at + 1]:
                break
            at += 1
        lo, hi = xc[seg], xc[seg + 1]
        drew = drew or bool(hi - lo > 1 and hi > x_t[0] and lo < x_t[-1])
        at += 1
    return None if drew else ValueError("array of sample points is empty")


def reconstruct_gaussian(x, y, x_t, baseline, sigma_map):
    """signal_reconstruction.py:17-42 -- half-Gaussian segments between consecutive
    fiducial points, resampled to len(x_t) samples.  Degenerate inputs raise what the
    reference raises (IndexError / ValueError, ``degenerate_beat_error``); the batched
    ``SigmaSearch`` gives such a beat an all-NaN row instead."""
    x_t = np.asarray(x_t, dtype=np.float64)
    xc, _, mask = extract_coordinates(x, y, x_t, baseline)
    err = degenerate_beat_error(xc, mask, x_t, len(sigma_map))
    if err is not None:
        raise err
    s = SigmaSearch(np.asarray(x, dtype=np.float64).reshape(1, 11), np.asarray(y, dtype=np.float64).reshape(1, 11),
                    np.asarray([baseline], dtype=np.float64), None, x_t)
    return s.reconstruct(np.asarray(sigma_map, dtype=np.float64)).cpu().numpy()[0]


def extract_coordinates(x, y, x_t, baseline):
    """signal_reconstruction.py:45-81 -- 13 points (baseline end points + fiducials), NaN
    drop, stable sort by x, duplicates bumped by one step until unique (no re-sort)."""
    torch, lib = _cabi.require_cuda()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    fx, fy, xt = dev(x), dev(y), dev(x_t)
    if fx.numel() != 11 or fy.numel() != 11:
        raise ValueError("expected 11 fiducial coordinates")
    coords = torch.zeros(26, dtype=torch.float64, device="cuda")
    meta = torch.zeros(14, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _cabi.check(lib.aew_extract_coordinates_f64(p(fx), p(fy), float(baseline), p(xt), xt.numel(), p(coords), p(meta),
                                                _cabi.stream_ptr(torch)), "extract_coordinates")
    c, m = coords.cpu().numpy(), meta.cpu().numpy()
    n = int(m[0])
    return c[:n].copy(), c[13:13 + n].copy(), m[1:14].astype(bool)


def gaussian_line(x1, y1, x2, y2, s):
    """signal_reconstruction.py:5-14 -- one half-Gaussian segment on the integer grid
    arange(int(x1), int(x2)), affinely rescaled to hit both end points."""
    torch, lib = _cabi.require_cuda()
    n = int(x2) - int(x1)
    if n <= 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")   # as the reference does
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    _cabi.check(lib.aew_gaussian_line_f64(float(x1), float(y1), float(x2), float(y2), float(s),
                                          ctypes.c_void_p(out.data_ptr()), n, _cabi.stream_ptr(torch)),
                "gaussian_line")
    return out.cpu().numpy()
