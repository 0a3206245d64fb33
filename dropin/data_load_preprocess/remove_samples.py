"""Reference-path alias: `data_load_preprocess.remove_samples` -> `ae_waveforms_b200.data_load_preprocess.remove_samples` (put dropin/ first on PYTHONPATH)."""
from ae_waveforms_b200.data_load_preprocess.remove_samples import *  # noqa: F401,F403
