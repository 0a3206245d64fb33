"""Runs the reference's UNMODIFIED caller scripts against this repo's drop-in packages (test helper).

    python tests/run_reference_callers.py OUT_DIR N_SAMPLES CASE_ID [CASE_ID ...]

``data_extract_vitaldb.py`` (its ``__main__`` block: the case loop, ``process_case_id``, ``data_stats.pkl``) and then
``data_create_arrays.py`` are executed as they are -- from /root/reference in the build container, from the
byte-compiled ``oracle/_ref`` on the GPU box -- with ``dropin/`` and ``dropin_neurokit/`` FIRST on ``sys.path``, so
that every ``data_load_preprocess.*`` / ``neurokit2`` name they import resolves to the CUDA implementations
(INTEGRATION.md section 2).  What the scripts need from outside the hot path is stubbed:

  vitaldb      no network: ``find_cases`` returns the case IDs given on the command line, ``vital_recs`` returns the
               synthetic track of oracle/ecg_synth.py as the (Time, track) frame the reference's own
               ``load_record`` (unmodified, found BEHIND dropin/ on the path) merges
  config       the constants of config.py:17-33 without the ``vars.env`` it wants; ``path_data`` = OUT_DIR + '/'
  to_numpy     returns a writable copy (pandas 3 returns read-only views; the reference writes in place, :78)
  to_parquet   pyarrow cannot write the 2-D ``ecg_samples_aligned`` cells: a frame pyarrow refuses
               is pickled under the same file name, and ``read_parquet`` reads it back

With ``AEW_CALLERS_PURE_REFERENCE=1`` the drop-ins are left off the path: the same scripts then run over the
reference's own modules on the CPU and produce the arrays the drop-in run is compared with.

Leaves ``np_waves.npy``, ``np_features.npy``, ``np_info.npy`` in OUT_DIR, as the reference does.
"""
from __future__ import annotations

import os
import runpy
import sys
import types
import warnings

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def main():
    out_dir, n_samples = sys.argv[1], int(sys.argv[2])
    case_ids = [int(c) for c in sys.argv[3:]]
    warnings.simplefilter("ignore")
    sys.path.insert(0, ROOT)
    from oracle import ecg_synth, ref_import
    ref_root = ref_import.root()
    if ref_root is None:
        raise SystemExit("no reference modules (neither /root/reference nor oracle/_ref)")
    pure = bool(os.environ.get("AEW_CALLERS_PURE_REFERENCE"))
    if pure:
        # the expected side: the same scripts over the reference's OWN modules on the CPU (neurokit2 = the oracle's
        # restatements, matplotlib stubbed)
        ref_import.install_stubs()
        sys.path[:0] = [ref_root]
    else:
        # dropin/ first, the reference tree behind it
        sys.path[:0] = [os.path.join(ROOT, "dropin"), os.path.join(ROOT, "dropin_neurokit"), ROOT, ref_root]

    import numpy as np
    import pandas as pd

    vitaldb = types.ModuleType("vitaldb")
    vitaldb.find_cases = lambda track_names: list(case_ids)

    def vital_recs(ipath, track_names, interval, return_timestamp=True, return_pandas=True):
        x = ecg_synth.synth_case(int(ipath), n_samples, nan_gaps=1 if int(ipath) % 2 else 0)
        return pd.DataFrame({"Time": np.arange(len(x)) * interval, track_names: x})
    vitaldb.vital_recs = vital_recs
    sys.modules["vitaldb"] = vitaldb

    cfg = types.ModuleType("config")
    cfg.path_data = out_dir.rstrip("/") + "/"
    cfg.parallel_preprocessing = False
    cfg.track_names = ["SNUADC/ECG_II"]
    cfg.use_pickle = True
    cfg.sampling_frequency = 500
    cfg.pvc_threshold = 150
    cfg.butter_order = 4
    cfg.butter_hp_ecg = 1
    cfg.butter_lp_ecg = 30
    cfg.sample_length = 20
    sys.modules["config"] = cfg

    real_to_parquet, real_read_parquet = pd.DataFrame.to_parquet, pd.read_parquet

    def to_parquet(self, path, *a, **k):
        try:
            return real_to_parquet(self, path, *a, **k)
        except Exception:
            self.to_pickle(path)

    def read_parquet(path, *a, **k):
        try:
            return real_read_parquet(path, *a, **k)
        except Exception:
            return pd.read_pickle(path)
    pd.DataFrame.to_parquet = to_parquet
    pd.read_parquet = read_parquet
    # pandas 3 (copy-on-write) hands out READ-ONLY arrays from to_numpy(); the reference was written for pandas 1/2
    # and writes into the record in place (data_extract_vitaldb.py:78)
    real_to_numpy = pd.DataFrame.to_numpy
    pd.DataFrame.to_numpy = lambda self, *a, **k: np.array(real_to_numpy(self, *a, **k))

    # (the reference root is on sys.path -- a directory in the build container, the bytecode archive on the GPU box)
    runpy.run_module("data_extract_vitaldb", run_name="__main__")
    import data_load_preprocess.filters as f
    import data_load_preprocess.load_record as lr
    import neurokit2 as nk
    if not pure:
        assert "ae_waveforms_b200" in f.zscore.__module__, f.zscore.__module__
        assert "ae_waveforms_b200" in nk.ecg_delineate.__module__, nk.ecg_delineate.__module__
    assert lr.__file__.startswith(ref_root), lr.__file__          # the reference's own load_record ran
    runpy.run_module("data_create_arrays", run_name="__main__")


if __name__ == "__main__":
    main()
