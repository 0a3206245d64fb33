"""Drop-in for data_load_preprocess/sample_from_record.py (reference lines 4-40)."""
from __future__ import annotations

import numpy as np

from .. import _cabi

WAVE_TRACKS = ("SNUADC/ECG_II", "SNUADC/ART", "SNUADC/CVP")


def sample_from_record(record, track_names, sampling_frequency=500, sample_length=20):
    """Non-overlapping windows of the wave tracks: window k covers rows [k*L, (k+1)*L),
    end indices np.arange(L, T, L) (exclusive stop: a record of exactly k*L rows yields
    k-1 windows).  The gather is a strided device copy."""
    torch, _ = _cabi.require_cuda()
    L = int(sampling_frequency * sample_length)
    ends = np.arange(L, record.shape[0], L)
    n = len(ends)
    picked = [i for i, name in enumerate(track_names) if name in WAVE_TRACKS]
    if not picked:
        raise ValueError("not enough values to unpack (expected 2, got 0)")
    rec = np.ascontiguousarray(record)
    if n == 0:
        return np.zeros((0, L, len(picked)), dtype=rec.dtype), ends.reshape(0, 1)
    dev = torch.from_numpy(rec[:n * L]).cuda()
    win = dev[:, picked].reshape(n, L, len(picked)).contiguous()
    return win.cpu().numpy(), ends.reshape(n, 1)
