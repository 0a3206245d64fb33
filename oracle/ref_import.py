"""Import the REAL reference modules from /root/reference (TEST INFRASTRUCTURE).

Only works in the build container, where the reference tree is mounted
read-only; the GPU box has no such tree, so nothing that runs there may call
this.  Used by tests/golden/make_golden.py to freeze golden vectors and by the
CPU tests (skipped when the tree is absent) to cross-check ``oracle.ref_port``.

The reference imports matplotlib (absent here) at module level and
``neurokit2`` (absent, un-vendored) inside remove_samples.py; both are replaced
in ``sys.modules`` before import: matplotlib by an empty stub, neurokit2 by
``oracle.nk_peaks``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "data_load_preprocess"))


def _install_stubs():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "neurokit2" not in sys.modules:
        from . import nk_peaks
        nk = types.ModuleType("neurokit2")
        nk.ecg_findpeaks = nk_peaks.ecg_findpeaks
        sys.modules["neurokit2"] = nk


class _Ref:
    """Namespace holding the reference's own functions."""


def load() -> _Ref:
    if not available():
        raise RuntimeError("reference tree not present (only exists in the build container)")
    _install_stubs()
    # the drop-in packages of this repo use the same dotted names; make sure the
    # reference's win for this import and are not cached under those names after
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in ("data_load_preprocess", "functions_traditional")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ref = _Ref()
        ref.filters = importlib.import_module("data_load_preprocess.filters")
        ref.sample_from_record = importlib.import_module("data_load_preprocess.sample_from_record")
        ref.remove_samples = importlib.import_module("data_load_preprocess.remove_samples")
        ref.median_sample = importlib.import_module("data_load_preprocess.median_sample")
        ref.features_ecg = importlib.import_module("data_load_preprocess.feature_extraction.features_ecg")
        ref.signal_reconstruction = importlib.import_module("functions_traditional.signal_reconstruction")
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules
                  if k.split(".")[0] in ("data_load_preprocess", "functions_traditional")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
    return ref
