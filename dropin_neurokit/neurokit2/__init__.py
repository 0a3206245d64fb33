"""OPTIONAL stand-in for the two neurokit2 calls on the hot path, for boxes where the reference's
scripts should run their delineation and R-peak detection on the GPU (or where neurokit2 is not
installed).  Put ``dropin_neurokit/`` on PYTHONPATH *in addition to* ``dropin/`` to activate it; leave it
off to keep the installed neurokit2 for everything else.

    nk.ecg_findpeaks(x, sampling_rate=fs)["ECG_R_Peaks"]     remove_samples.py:50-51
    nk.ecg_delineate(x, r_peaks, sampling_rate=fs)[1]        data_extract_vitaldb.py:131
"""
from ae_waveforms_b200.neurokit_shim import ecg_delineate, ecg_findpeaks  # noqa: F401
