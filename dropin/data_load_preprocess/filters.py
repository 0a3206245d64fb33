"""Reference-path alias: `data_load_preprocess.filters` -> `ae_waveforms_b200.data_load_preprocess.filters` (put dropin/ first on PYTHONPATH)."""
from ae_waveforms_b200.data_load_preprocess.filters import *  # noqa: F401,F403
