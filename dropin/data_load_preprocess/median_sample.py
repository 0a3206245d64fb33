"""Reference-path alias: `data_load_preprocess.median_sample` -> `ae_waveforms_b200.data_load_preprocess.median_sample` (put dropin/ first on PYTHONPATH)."""
from ae_waveforms_b200.data_load_preprocess.median_sample import *  # noqa: F401,F403
