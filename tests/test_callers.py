"""The reference's UNMODIFIED caller scripts (data_extract_vitaldb.py, data_create_arrays.py) end to end.

CPU: the scripts over the reference's own modules pin the oracle's restatement of their glue
(``oracle.ref_port.process_case`` / ``oracle.ref_glue.process_case``).  GPU: the same scripts with ``dropin/`` and
``dropin_neurokit/`` first on the path -- "switch unchanged" (BASELINE.json north_star) -- must write
the same three ``.npy`` arrays, and those must equal what the batched API returns for the same records.
"""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

from oracle import ecg_synth, ref_import, ref_port as rp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HELPER = os.path.join(ROOT, "tests", "run_reference_callers.py")
N_SAMPLES = 100000
CASES = [3, 4, 12]

needs_ref = pytest.mark.skipif(not ref_import.available(),
                               reason="needs /root/reference or oracle/_ref (oracle/build_ref.py)")


def run_callers(out_dir, pure, filter_mode=None):
    env = dict(os.environ)
    env.pop("AEW_CALLERS_PURE_REFERENCE", None)
    if pure:
        env["AEW_CALLERS_PURE_REFERENCE"] = "1"
    if filter_mode:
        env["AEW_FILTER_MODE"] = filter_mode
    r = subprocess.run([sys.executable, HELPER, str(out_dir), str(N_SAMPLES)] + [str(c) for c in CASES],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return {n: np.load(os.path.join(out_dir, n + ".npy"), allow_pickle=True)
            for n in ("np_waves", "np_features", "np_info")}


def records():
    return [ecg_synth.synth_case(c, N_SAMPLES, nan_gaps=1 if c % 2 else 0) for c in CASES]


@needs_ref
def test_reference_callers_pin_the_glue(tmp_path):
    """process_case_id + data_create_arrays as the reference ships them == the oracle's glue, bit for bit."""
    warnings.simplefilter("ignore")
    got = run_callers(tmp_path, pure=True)
    assert got["np_waves"].dtype == np.float32 and got["np_waves"].shape[1] == 320
    assert got["np_features"].shape[1] == 35 and got["np_info"].shape[1] == 2
    from oracle import ref_glue
    for fn in (rp.process_case, ref_glue.process_case):
        outs = [fn(x, case_id=c, features=True) for c, x in zip(CASES, records())]
        assert np.array_equal(np.concatenate([o["np_waves"] for o in outs]), got["np_waves"])
        assert np.array_equal(np.concatenate([o["np_info"] for o in outs]), got["np_info"])
        assert np.array_equal(np.concatenate([o["np_features"] for o in outs]), got["np_features"], equal_nan=True)


@needs_ref
@pytest.mark.gpu
def test_unchanged_callers_on_the_dropin(tmp_path):
    """The scripts, untouched, on the CUDA drop-ins.  Exact filter mode (the drop-in's default): the three arrays
    are bit-identical to the run over the reference's own modules.  Tiled filter: identical rows, np_waves within
    1e-6 relative, and the batched API gives the same arrays on the same records."""
    import torch
    from ae_waveforms_b200 import batched
    for d in ("ref", "exact", "fast"):
        os.makedirs(tmp_path / d, exist_ok=True)
    want = run_callers(tmp_path / "ref", pure=True)
    exact = run_callers(tmp_path / "exact", pure=False, filter_mode="exact")
    for n in ("np_waves", "np_features", "np_info"):
        assert exact[n].dtype == want[n].dtype and exact[n].shape == want[n].shape, n
        assert np.array_equal(exact[n], want[n], equal_nan=True), n
    fast = run_callers(tmp_path / "fast", pure=False, filter_mode="fast")
    assert np.array_equal(fast["np_info"], want["np_info"])
    scale = np.abs(want["np_waves"]).max()
    assert np.abs(fast["np_waves"].astype(np.float64) - want["np_waves"]).max() <= 2e-6 * scale
    # the batched API on the same records: bit-identical in exact filter mode; with the tiled filter the batched path
    # z-scores the record inside the filter's load ((x - mean) * (1 / sd), one rounding apart from the division the
    # caller's own zscore() call performs), so float64 amplitudes agree to ~1e-15 and the float32 rows exactly
    res = batched.Preprocessor(filter_mode="exact").run(batched.RecordBatch.from_host(records(), CASES), features=True)
    for n in ("np_waves", "np_features", "np_info"):
        assert np.array_equal(res[n].cpu().numpy(), exact[n], equal_nan=True), n
    res = batched.Preprocessor().run(batched.RecordBatch.from_host(records(), CASES), features=True)
    torch.cuda.synchronize()
    assert np.array_equal(res["np_waves"].cpu().numpy(), fast["np_waves"])
    assert np.array_equal(res["np_info"].cpu().numpy(), fast["np_info"])
    a, b = res["np_features"].cpu().numpy(), fast["np_features"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) <= 1e-9 * np.nanmax(np.abs(b))
