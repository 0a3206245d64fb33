"""Drop-in for data_load_preprocess/filters.py (reference lines 5-55): same names,
argument order, defaults and NaN behaviour; the arithmetic runs on the GPU."""
from __future__ import annotations

import os

import numpy as np

from .. import _cabi
from ..batched import Preprocessor, RecordBatch, _ptr

# "exact" (default of the per-call drop-in): SciPy's sequential DF2T order, one warp per segment -- bit-identical to
# scipy.signal.filtfilt, so the reference's scripts produce the reference's arrays (tests/test_callers.py);
# "fast": the tiled block scan of the batched path (within 1e-6 relative of SciPy, measured ~1e-9)
FILTER_MODE = os.environ.get("AEW_FILTER_MODE", "exact")


def _as_record(data):
    arr = np.asarray(data)
    if arr.ndim != 1:
        raise ValueError("expected a 1-D track")
    return arr, RecordBatch.from_host([np.asarray(arr, dtype=np.float64)])


def zscore(data):
    """filters.py:5-12 -- (data - nanmean) / nanstd, statistics bit-identical to NumPy."""
    arr, batch = _as_record(data)
    pre = Preprocessor()
    torch, lib = pre.torch, pre.lib
    if batch.total == 0:
        return np.asarray(arr, dtype=np.float64).copy()
    beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
    mean, sd, _ = pre.zscore_stats(batch.flat, beg_h, end_h, d_beg, d_end)
    out = torch.empty_like(batch.flat)
    _cabi.check(lib.aew_zscore_apply_f64(_ptr(batch.flat), _ptr(d_beg), _ptr(d_end), 1, _ptr(mean), _ptr(sd),
                                         _ptr(out), 0, batch.total, _cabi.stream_ptr(torch)), "zscore_apply")
    return out.cpu().numpy()


def butterworth(data, btype, cutoff_frequency, sampling_frequency=500, filter_order=4):
    """filters.py:15-55 -- zero-phase Butterworth per non-NaN segment; segments of at
    most 3*max(len(a), len(b)) samples pass through; output dtype = input dtype."""
    arr, batch = _as_record(data)
    pre = Preprocessor(fs=sampling_frequency, order=filter_order, filter_mode=FILTER_MODE,
                       hp=cutoff_frequency if btype in ("high", "highpass", "hp") else None,
                       lp=cutoff_frequency if btype in ("low", "lowpass", "lp") else None)
    if not (pre.hp or pre.lp):
        raise ValueError(f"unsupported btype {btype!r} (the path uses 'low' and 'high')")
    if batch.total == 0:
        return np.full_like(arr, np.nan)
    beg_h, end_h, d_beg, d_end = pre._ranges(batch, trim=False)
    y, header = pre.filter_records(batch.flat, beg_h, end_h, d_beg, d_end)
    res = y.cpu().numpy()
    _cabi.check_device_status(int(header[4].item()), "butterworth")
    out = np.full_like(arr, np.nan)
    out[...] = res                        # SciPy filters in float64; assignment casts back (appendix A.1)
    return out
