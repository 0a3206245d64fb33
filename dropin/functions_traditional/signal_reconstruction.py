"""Reference-path alias: `functions_traditional.signal_reconstruction` -> `ae_waveforms_b200.functions_traditional.signal_reconstruction` (put dropin/ first on PYTHONPATH)."""
from ae_waveforms_b200.functions_traditional.signal_reconstruction import *  # noqa: F401,F403
