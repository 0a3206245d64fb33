"""Reference-path alias: `data_load_preprocess.sample_from_record` -> `ae_waveforms_b200.data_load_preprocess.sample_from_record` (put dropin/ first on PYTHONPATH)."""
from ae_waveforms_b200.data_load_preprocess.sample_from_record import *  # noqa: F401,F403
