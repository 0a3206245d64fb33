"""Reference-path alias: `data_load_preprocess.feature_extraction.features_ecg` -> `ae_waveforms_b200.data_load_preprocess.feature_extraction.features_ecg` (put dropin/ first on PYTHONPATH)."""
from ae_waveforms_b200.data_load_preprocess.feature_extraction.features_ecg import *  # noqa: F401,F403
