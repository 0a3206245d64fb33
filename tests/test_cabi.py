"""The C-ABI library loads and exports every symbol include/aewave.h declares (no
compute calls: this runs without a GPU), and the product path fails loudly without one."""
import os
import re

import pytest

from ae_waveforms_b200 import _cabi, build

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    text = open(os.path.join(ROOT, "include", "aewave.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aew_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = _cabi.load()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in aewave.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names, "ctypes prototypes and header out of sync"
    want = int(re.search(r"#define AEW_VERSION (\d+)", open(os.path.join(ROOT, "include", "aewave.h")).read()).group(1))
    assert lib.aew_version() == want


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_path_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ae_waveforms_b200.data_load_preprocess.filters import zscore
    with pytest.raises(_cabi.AewError):
        zscore([1.0, 2.0, 3.0])
    from ae_waveforms_b200.batched import Preprocessor
    with pytest.raises(_cabi.AewError):
        Preprocessor()
