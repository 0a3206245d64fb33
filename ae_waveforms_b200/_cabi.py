"""ctypes binding of libaewave.so (include/aewave.h).

There is no CPU fallback: if the library is missing, or no sm_100 device is
current, every entry point raises.  Tensors are passed as raw device pointers
(``tensor.data_ptr()``) plus the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libaewave.so")

ERRORS = {-1: "bad argument", -2: "workspace too small", -3: "unsupported parameter", -4: "no sm_100 CUDA device"}
DEV_STATUS = {1: "more NaN-delimited segments than max_segments", 2: "more filter tiles than the workspace holds",
              4: "filter: a tile's bulk copy did not complete", 8: "more than max_peaks R-peaks in a window",
              16: "a fiducial index lies outside its window"}


class AewError(RuntimeError):
    pass


class K3Params(ctypes.Structure):
    """Mirror of ``struct K3Params`` in include/aewave.h."""
    _fields_ = [
        ("y", c_void_p), ("win_start", c_void_p), ("n_win", c_longlong), ("L", c_int),
        ("smooth_k", c_int), ("avg_k", c_int), ("refractory", c_int), ("pre", c_int), ("post", c_int),
        ("hr_num", c_double), ("rr_div", c_double), ("hr_lo", c_double), ("hr_hi", c_double),
        ("pvc_threshold", c_double), ("sd_threshold", c_double), ("mode", c_int), ("zscore_windows", c_int),
        ("keep", c_void_p), ("has_nan", c_void_p), ("n_peaks", c_void_p), ("peaks", c_void_p),
        ("stats", c_void_p), ("median", c_void_p), ("waves", c_void_p), ("zwin", c_void_p),
        ("zparams", c_void_p), ("status", c_void_p), ("hazard_count", c_void_p), ("lazy_sd", c_int),
        ("workspace", c_void_p), ("workspace_bytes", ctypes.c_ulonglong), ("fast_stats", c_int)]


# name -> (restype, argtypes); every symbol include/aewave.h declares
SIGNATURES = {
    "aew_version": (c_int, []),
    "aew_device_check": (c_int, []),
    "aew_zscore_plan": (c_longlong, [c_void_p, c_void_p, c_longlong, c_void_p]),
    "aew_zscore_workspace_bytes": (c_size_t, [c_longlong, c_longlong]),
    "aew_zscore_stats_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aew_zscore_stats_fast_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aew_zscore_apply_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                     c_longlong, c_longlong, c_void_p]),
    "aew_trim_nan_edges_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    "aew_filt_tables_len": (c_longlong, []),
    "aew_filt_tables": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "aew_filt_max_tiles": (c_longlong, [c_longlong, c_longlong, c_longlong]),
    "aew_filt_workspace_bytes": (c_size_t, [c_longlong, c_longlong, c_longlong]),
    "aew_segments_find_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_longlong, c_void_p]),
    "aew_filtfilt_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong,
                                 c_longlong, c_void_p, c_size_t, c_longlong, c_int, c_void_p]),
    "aew_filt_fused_ok": (c_int, [c_void_p]),
    "aew_filt_header": (c_int, [c_void_p, c_void_p]),
    "aew_filt_timing_dump": (c_int, [c_void_p]),
    "aew_windows_max_peaks": (c_int, []),
    "aew_windows_max_len": (c_int, []),
    "aew_windows_workspace_bytes": (c_size_t, [c_longlong]),
    "aew_windows_beats_f64": (c_int, [c_void_p, c_void_p]),
    "aew_windows_timing_dump": (c_int, [c_void_p]),
    "aew_median_beats_f64": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    "aew_beat_metrics_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "aew_fetch_host": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p]),
    "aew_copy_heads_h2d": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "aew_widen_f64": (c_int, [c_void_p, c_int, c_longlong, c_double, c_double, c_int, c_int, c_void_p, c_void_p]),
    "aew_delineate_blocks": (c_int, [c_longlong]),
    "aew_delineate_workspace_bytes": (c_size_t, [c_longlong, c_int, c_double]),
    "aew_delineate_f64": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_double, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aew_delineate_timing_dump": (c_int, [c_void_p]),
    "aew_fiducial_features_f64": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                          c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "aew_fiducial_relative_f64": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_double, c_void_p,
                                          c_void_p, c_void_p]),
    "aew_morphology_f64": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "aew_nanmedian_rows_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "aew_peak_stats_f64": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_double, c_void_p, c_void_p,
                                   c_void_p]),
    "aew_recon_workspace_bytes": (c_size_t, [c_longlong, c_longlong]),
    "aew_recon_sse_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_int, c_void_p,
                                  c_longlong, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aew_recon_max_grid_values": (c_int, []),
    "aew_recon_tables_workspace_bytes": (c_size_t, [c_longlong]),
    "aew_recon_sse_grid_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_int, c_void_p,
                                       c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_size_t, c_void_p]),
    "aew_gaussian_line_f64": (c_int, [c_double, c_double, c_double, c_double, c_double, c_void_p, c_int, c_void_p]),
    "aew_extract_coordinates_f64": (c_int, [c_void_p, c_void_p, c_double, c_void_p, c_int, c_void_p, c_void_p,
                                            c_void_p]),
}

_lib = None


def load(path: str | None = None) -> ctypes.CDLL:
    """Load the shared library and attach prototypes (no GPU needed for this)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise AewError(f"{p} is missing: build it with `python -m ae_waveforms_b200.build` "
                       "(there is no CPU fallback)")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise AewError(f"{what}: {ERRORS.get(rc, rc)}")
    raise AewError(f"{what}: CUDA error {rc}")


def check_device_status(bits: int, what: str) -> None:
    if bits:
        msgs = [m for b, m in DEV_STATUS.items() if bits & b]
        raise AewError(f"{what}: " + "; ".join(msgs))


def require_cuda():
    """Import torch, insist on a CUDA device of the right architecture."""
    import torch
    if not torch.cuda.is_available():
        raise AewError("ae_waveforms_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    lib = load()
    check(lib.aew_device_check(), "device check")
    return torch, lib


def stream_ptr(torch) -> int:
    """The current stream of the CURRENT device: the C ABI launches on the current device, so every caller runs
    inside ``on_device`` (below), which makes the object's device current first."""
    return torch.cuda.current_stream().cuda_stream


def on_device(fn):
    """Method decorator: run with ``self.device`` as the current CUDA device.  The library never calls
    cudaSetDevice; tensors of ``device='cuda:1'`` with cuda:0 current would otherwise meet kernels launched on
    cuda:0's stream."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        torch, dev = self.torch, self.device
        if dev.index is None or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper
