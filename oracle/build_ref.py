"""Recipe for ``oracle/_ref`` (TEST INFRASTRUCTURE): the reference's own modules, compiled.

The reference is pure Python, so its "build" is byte-compilation: every module on the hot path (and the two
caller scripts) is compiled FROM WHERE IT LIES under ``/root/reference`` to sourceless bytecode, packed into
``oracle/_ref/reference_modules.zip`` (importable as it is: zipimport) -- no reference source is copied into the
repo, ``oracle/_ref/`` is git-ignored (so it stays out of history) but not gpurun-ignored (so it travels to the
GPU box like the repo's own built ``.so``; loose ``*.pyc`` files do not travel, an archive does).  There the
CPU arm of ``bench.py`` times the reference's OWN functions (``cpu_baseline.kind = "reference"``) and the tests
cross-check the oracle port and the drop-ins against them.

``python oracle/build_ref.py`` (or ``__graft_entry__.build()``).  A no-op when ``/root/reference`` is absent.
"""
from __future__ import annotations

import os
import py_compile
import sys
import tempfile
import zipfile

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


ARCHIVE = os.path.join(OUT, "reference_modules.zip")


def build(verbose: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "data_load_preprocess")):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    srcs = [os.path.join(REFERENCE_ROOT, rel) for rel in MODULES if os.path.exists(os.path.join(REFERENCE_ROOT, rel))]
    if os.path.exists(ARCHIVE) and all(os.path.getmtime(ARCHIVE) >= os.path.getmtime(p) for p in srcs + [__file__]):
        return ARCHIVE
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ARCHIVE + ".tmp", "w", zipfile.ZIP_DEFLATED) as z:
        for rel in MODULES:
            src = os.path.join(REFERENCE_ROOT, rel)
            if not os.path.exists(src):
                continue
            pyc = os.path.join(tmp, rel.replace("/", "_") + "c")
            # dfile: the path tracebacks show; UNCHECKED_HASH: valid without its source next to it
            py_compile.compile(src, cfile=pyc, dfile=src, doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            z.write(pyc, rel[:-3] + ".pyc")
            if verbose:
                print("compiled", rel)
        z.writestr("PYTHON_VERSION", sys.version.split()[0] + "\n")
    os.replace(ARCHIVE + ".tmp", ARCHIVE)
    return ARCHIVE


if __name__ == "__main__":
    print(build(verbose=True))
