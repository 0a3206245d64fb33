"""Drop-in for functions_traditional/signal_reconstruction.py (reference lines 5-81).
Every function runs the K5 device code; use
``ae_waveforms_b200.sigma_search.SigmaSearch`` for whole test sets and candidate lists."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _cabi
from ..sigma_search import SigmaSearch


def reconstruct_gaussian(x, y, x_t, baseline, sigma_map):
    """signal_reconstruction.py:17-42 -- half-Gaussian segments between consecutive
    fiducial points, resampled to len(x_t) samples.  Degenerate inputs for which the
    reference raises IndexError / ValueError yield an all-NaN beat."""
    x_t = np.asarray(x_t, dtype=np.float64)
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
