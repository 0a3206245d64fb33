"""CPU oracle for the ae_waveforms preprocessing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as a fallback for the CUDA path.

Contents
--------
``ref_port``    NumPy/SciPy restatement of the reference functions on the path
                (a15 via ``nk_delineate``), each citing the
                reference file:line it follows.
``nk_peaks``    restatement of ``neurokit2.ecg_findpeaks(method="neurokit")``
                (third-party, un-vendored, version-unpinned in the reference's
                requirements.txt:10, absent from this image) -- PARITY UNPINNED
                at that boundary: no copy of the library exists offline, the
                reference holds no test or golden vector for it.
``nk_delineate`` restatement of ``neurokit2.ecg_delineate`` (default method "dwt": HR-sized
                epochs for Q/S, 2 kHz cubic-spline resampling, a-trous wavelet rows, per-beat
                searches) -- PARITY UNPINNED for the same reason; the functions that
                consume its output (a16-a18) are pinned by ``tests/golden/features.npz``.
``np_pairwise`` restatement of NumPy's pairwise float64 summation order, the
                order the CUDA reductions reproduce bit-for-bit.
``ecg_synth``   seeded sum-of-Gaussians ECG generator.
``ref_import``  importer for the *real* reference modules under
                ``/root/reference`` (exists only in the build container); used by
                ``tests/golden/make_golden.py`` to freeze golden vectors and by
                the CPU tests to cross-check ``ref_port`` when the tree is there.

Pinning status: every function in ``ref_port`` except the two neurokit2 calls
(R-peak detector, wave delineator) is pinned against outputs of the reference's own code (fixtures under
``tests/golden/`` made by ``tests/golden/make_golden.py``).  The R-peak
detector is pinned only against SciPy primitives it is built from.
"""
