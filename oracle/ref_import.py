"""Import the REAL reference modules (TEST INFRASTRUCTURE).

In the build container the reference tree is mounted read-only at /root/reference.  The GPU box has no such
tree: there the modules come from ``oracle/_ref/`` -- the reference's own files byte-compiled by
``oracle/build_ref.py`` (git-ignored build output that travels with the snapshot, like the repo's ``.so``).
Used by tests/golden/make_golden.py to freeze golden vectors, by the tests to cross-check ``oracle.ref_port``
and the drop-ins, and by ``bench.py``'s CPU arm (``cpu_baseline.kind = "reference"``).

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

SOURCE_ROOT = "/root/reference"
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_modules.zip")


def root() -> str | None:
    """The live tree if present, else the byte-compiled copy, else None."""
    if os.path.isdir(os.path.join(SOURCE_ROOT, "data_load_preprocess")):
        return SOURCE_ROOT
    if os.path.exists(COMPILED_ROOT):                    # importable as a sys.path entry (zipimport)
        return COMPILED_ROOT
    return None


def available() -> bool:
    return root() is not None


def install_stubs():
    """matplotlib stub + neurokit2 -> oracle restatements (both absent from this image)."""
    _install_stubs()


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
        try:
            from . import nk_delineate
            nk.ecg_delineate = lambda sig, rpeaks=None, sampling_rate=1000, **kw: (
                None, nk_delineate.ecg_delineate(sig, rpeaks, sampling_rate))
        except Exception:                                   # pragma: no cover
            pass
        sys.modules["neurokit2"] = nk


class _Ref:
    """Namespace holding the reference's own functions."""


def load() -> _Ref:
    ref_root = root()
    if ref_root is None:
        raise RuntimeError("reference modules not present: neither /root/reference nor oracle/_ref "
                           "(run oracle/build_ref.py in the build container)")
    _install_stubs()
    # the drop-in packages of this repo use the same dotted names; make sure the
    # reference's win for this import and are not cached under those names after
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in ("data_load_preprocess", "functions_traditional")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, ref_root)
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
