"""Recipe for ``oracle/_ref`` (TEST INFRASTRUCTURE): the reference's own modules, compiled.

The reference is pure Python, so its "build" is byte-compilation: every module on the hot path (and the two
caller scripts) is compiled FROM WHERE IT LIES under ``/root/reference`` to a sourceless ``.pyc`` under
``oracle/_ref/`` -- no reference source is copied into the repo, ``oracle/_ref/`` is git-ignored (so it stays out
of history) but not gpurun-ignored (so it travels to the GPU box like the repo's own built ``.so``).  There the
CPU arm of ``bench.py`` times the reference's OWN functions (``cpu_baseline.kind = "reference"``) and the tests
cross-check the oracle port and the drop-ins against them.

``python oracle/build_ref.py`` (or ``__graft_entry__.build()``).  A no-op when ``/root/reference`` is absent.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref")

# hot-path modules + the two callers that must keep working unchanged (section 8b)
MODULES = [
    "data_load_preprocess/__init__.py",
    "data_load_preprocess/filters.py",
    "data_load_preprocess/load_record.py",
    "data_load_preprocess/sample_from_record.py",
    "data_load_preprocess/remove_samples.py",
    "data_load_preprocess/median_sample.py",
    "data_load_preprocess/feature_extraction/__init__.py",
    "data_load_preprocess/feature_extraction/features_ecg.py",
    "functions_traditional/__init__.py",
    "functions_traditional/signal_reconstruction.py",
    "data_extract_vitaldb.py",
    "data_create_arrays.py",
]


def build(verbose: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "data_load_preprocess")):
        return OUT if os.path.isdir(OUT) else None
    for rel in MODULES:
        src = os.path.join(REFERENCE_ROOT, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(OUT, rel[:-3] + ".pyc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        # dfile: the path tracebacks show; UNCHECKED_HASH: the .pyc is valid without its source next to it
        py_compile.compile(src, cfile=dst, dfile=src, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", rel, "->", os.path.relpath(dst, HERE))
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write(sys.version.split()[0] + "\n")
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
