"""ae_waveforms_b200 -- the ae_waveforms ECG preprocessing hot path on B200 (sm_100a).

``batched``        device-resident API (throughput path)
``sigma_search``   batched Gaussian-reconstruction objective
``dist``           case sharding + the two small collectives
``data_load_preprocess`` / ``functions_traditional``
                   drop-in functions with the reference's names and signatures
                   (also re-exported under ``dropin/`` with the reference's dotted
                   module paths)
Everything calls hand-written CUDA kernels through the C ABI of
``csrc/libaewave.so`` (include/aewave.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
