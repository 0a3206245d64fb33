"""Host-side logic: window arithmetic, sharding, sigma-search input preparation, the
drop-in import surface, and the two collectives on world_size-2 gloo (CPU tensors)."""
import os
import sys
import warnings

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from ae_waveforms_b200 import dist as awd
from ae_waveforms_b200.batched import window_count
from ae_waveforms_b200.sigma_search import beat_time_axis, prepare_inputs
from oracle import ecg_synth, ref_port as rp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_window_count_matches_reference(golden):
    g = golden("windows")
    for T, cnt in zip(g["lens"], g["counts"]):
        assert window_count(int(T), 10000) == cnt
    for T in (0, 1, 9999, 10000, 10001, 29999, 30000, 30001, 1800000, 7200000):
        assert window_count(T, 10000) == len(np.arange(10000, T, 10000))


def test_time_axis_and_sigma_inputs_match_oracle(golden):
    g = golden("recon")
    assert np.array_equal(beat_time_axis(), g["x_t"])
    rng = np.random.default_rng(3)
    waves = rng.standard_normal((7, 320)).astype(np.float32)
    feats = rng.standard_normal((7, 35))
    feats[2, 3] = np.nan
    for a, b in zip(prepare_inputs(waves, feats), rp.prepare_sigma_inputs(waves, feats)):
        assert np.array_equal(a, b, equal_nan=True)


def test_chunk_schedule_partitions_the_cases():
    """The feeder's tapered chunks: contiguous, complete, never larger than asked, small at both ends."""
    from ae_waveforms_b200.batched import chunk_schedule
    for n, c in ((1000, 100), (400, 40), (5, 2), (0, 10), (37, 5), (1000, 1), (64, 8), (17, 100), (250, 100), (1000, 7)):
        for taper in (True, False):
            ch = chunk_schedule(n, c, taper)
            if n == 0:
                assert ch == []
                continue
            assert ch[0][0] == 0 and ch[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            assert all(0 < b - a <= c for a, b in ch)
    sizes = [b - a for a, b in chunk_schedule(1000, 100)]
    assert sizes[:3] == [12, 25, 50] and sizes[-3:] == [50, 25, 12] and max(sizes) <= 100
    assert [b - a for a, b in chunk_schedule(1000, 100, taper=False)] == [100] * 10


def test_shard_cases_balances_and_partitions():
    lens = [7200000, 100, 1800000, 1800000, 5, 3600000, 900000, 900000]
    shards = awd.shard_cases(lens, 4)
    assert sorted(sum(shards, [])) == list(range(8))
    loads = [sum(lens[i] for i in s) for s in shards]
    assert max(loads) == 7200000 and min(loads) >= 2700000
    assert awd.shard_cases(lens, 4, rank=2) == shards[2]


def test_dropin_surface_has_the_reference_names():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for k in [k for k in sys.modules if k.split(".")[0] in ("data_load_preprocess", "functions_traditional")]:
            del sys.modules[k]
        from data_load_preprocess.filters import zscore, butterworth  # noqa: F401
        from data_load_preprocess.sample_from_record import sample_from_record  # noqa: F401
        from data_load_preprocess.remove_samples import remove_samples, estimate_hr, rmssd  # noqa: F401
        from data_load_preprocess.median_sample import (compute_median_sample_ecg, compute_median_features_ecg,  # noqa: F401
                                                        plot_median_ecg)
        from data_load_preprocess.feature_extraction.features_ecg import (compute_ecg_morphology_features,  # noqa: F401
                                                                           ecg_features_relative_r)
        from functions_traditional.signal_reconstruction import (reconstruct_gaussian, extract_coordinates,  # noqa: F401
                                                                  gaussian_line)
        import inspect
        assert list(inspect.signature(butterworth).parameters) == [
            "data", "btype", "cutoff_frequency", "sampling_frequency", "filter_order"]
        assert list(inspect.signature(remove_samples).parameters) == [
            "samples_wave_tracks", "indices_samples", "track_names", "sampling_frequency", "pvc_threshold",
            "sd_threshold_ecg"]
        assert inspect.signature(remove_samples).parameters["sd_threshold_ecg"].default == 0.05
        assert list(inspect.signature(compute_median_sample_ecg).parameters) == [
            "sample_ecg", "r_peaks", "sampling_frequency", "resample_ecg", "return_aligned"]
        import torch
        from ae_waveforms_b200._cabi import AewError
        if not torch.cuda.is_available():                  # every drop-in runs on the device or fails loudly
            for call in (lambda: estimate_hr([], [100, 600, 1100], 500, mod="max"),
                         lambda: rmssd([100, 600, 1101], 500),
                         lambda: compute_ecg_morphology_features({}, 0),
                         lambda: butterworth(np.zeros(100), "high", 1)):
                with pytest.raises(AewError):
                    call()
        with pytest.raises(NotImplementedError):
            plot_median_ecg()
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for k in [k for k in sys.modules if k.split(".")[0] in ("data_load_preprocess", "functions_traditional")]:
            del sys.modules[k]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = 3 + 2 * rank
        waves = torch.from_numpy(rng.standard_normal((n, 320)).astype(np.float32))
        info = torch.from_numpy(np.stack([np.full(n, float(10 * (world - rank))), (np.arange(n) + 1) * 10000.0], 1))
        parts = awd.allgather_rows(waves)
        gw, gi = awd.gather_beat_arrays(waves, info)
        feats = torch.from_numpy(np.repeat(info.numpy()[:, 1:2] + rank, 35, axis=1))       # rows tagged by idx_end + rank
        gw2, gi2, gf = awd.gather_beat_arrays(waves, info, np_features=feats)
        assert torch.equal(gw2, gw) and torch.equal(gi2, gi) and gf.shape == (gi.shape[0], 35)
        src_rank = world - (gi[:, 0] / 10).to(torch.int64)                                 # case id encodes the rank
        assert torch.equal(gf[:, 0], gi[:, 1] + src_rank.to(torch.float64))                # features follow their rows
        dw, di, df = awd.gather_beat_arrays(waves, info, np_features=feats, dst=1)         # to one rank only
        if rank == 1:
            assert torch.equal(dw, gw) and torch.equal(di, gi) and torch.equal(df, gf)
        else:
            assert dw is None and di is None and df is None
        ew, ei = awd.gather_beat_arrays(waves[:0] if rank == 0 else waves, info[:0] if rank == 0 else info, dst=0)
        if rank == 0:                                                                       # a rank without rows
            assert ew.shape[0] == ei.shape[0] == 3 + 2 * (world - 1)
        sse = torch.tensor([1.0 + rank, 2.0, float("nan") if rank == 1 else 0.5], dtype=torch.float64)
        awd.allreduce_sse(sse)
        q.put((rank, [p.shape[0] for p in parts], gw.numpy(), gi.numpy(), sse.numpy(), waves.numpy(), info.numpy()))
    finally:
        dist.destroy_process_group()


def test_collectives_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    assert res[0][1] == res[1][1] == [3, 5]
    for r in res:
        gw, gi, sse = r[2], r[3], r[4]
        assert gw.shape == (8, 320) and gi.shape == (8, 2)
        # case-major order: case 10 (rank 1) before case 20 (rank 0), idx_end ascending
        assert np.array_equal(gi[:, 0], [10] * 5 + [20] * 3)
        assert np.array_equal(gi[:5, 1], (np.arange(5) + 1) * 10000.0)
        assert np.array_equal(gw[:5], res[1][5]) and np.array_equal(gw[5:], res[0][5])
        assert sse[0] == 3.0 and sse[1] == 4.0 and np.isnan(sse[2])
    assert np.array_equal(res[0][2], res[1][2])


def test_synthetic_generator_is_seeded():
    a, b = ecg_synth.synth_case(5, 20000), ecg_synth.synth_case(5, 20000)
    assert np.array_equal(a, b) and not np.array_equal(a, ecg_synth.synth_case(6, 20000))
    c = ecg_synth.sigma_candidates(64, "integer")
    assert c.shape == (66, 12) and set(np.unique(c)) <= {1.0, 2.0, 3.0, 4.0, 5.0}
    assert c[-2].tolist() == [1, 1, 1, 2, 2, 3, 3, 3, 4, 2, 2, 1]


def test_array_store_resume_and_final_arrays(tmp_path):
    """per-chunk shards + manifest: resume skips finished cases, a half-written shard is ignored, the final
    arrays are case-major with the reference's dtypes and file names (data_create_arrays.py:17-58)."""
    from ae_waveforms_b200.store import ArrayStore
    rng = np.random.default_rng(1)

    def rows(case, n):
        info = np.stack([np.full(n, float(case)), (np.arange(n) + 1) * 10000.0], 1)
        return rng.standard_normal((n, 320)).astype(np.float32), info, rng.standard_normal((n, 35))

    root = str(tmp_path / "out")
    st = ArrayStore(root)
    a = [rows(7, 3), rows(9, 0), rows(8, 2)]
    st.add_chunk([7, 9, 8], np.concatenate([x[0] for x in a]), np.concatenate([x[1] for x in a]),
                 np.concatenate([x[2] for x in a]))
    b = [rows(2, 4)]
    st.add_chunk([2], b[0][0], b[0][1])                                       # no features: NaN rows
    with pytest.raises(ValueError):
        st.add_chunk([8], *rows(8, 1))                                         # already stored
    open(os.path.join(root, "shards", "50_60.npz.tmp"), "wb").write(b"partial")  # a killed run's leftover
    st2 = ArrayStore(root)                                                     # a new process resumes
    assert st2.pending([2, 7, 8, 9, 11, 12]) == [11, 12]
    st2.add_chunk([11, 12], *rows(12, 2))
    out = st2.finalize()
    assert out["np_waves"].dtype == np.float32 and out["np_features"].dtype == np.float64
    assert out["np_info"][:, 0].tolist() == [2] * 4 + [7] * 3 + [8] * 2 + [12] * 2
    assert out["case_ids"] == [2, 7, 8, 12] and out["samples_total"] == 11     # case 9 and 11: no rows
    assert np.array_equal(out["np_waves"][4:7], a[0][0]) and np.array_equal(out["np_features"][7:9], a[2][2])
    assert np.isnan(out["np_features"][:4]).all()
    for name, shape in (("np_waves", (11, 320)), ("np_features", (11, 35)), ("np_info", (11, 2))):
        assert np.load(os.path.join(root, name + ".npy")).shape == shape


def test_degenerate_beats_raise_what_the_reference_raises():
    """reconstruct_gaussian (signal_reconstruction.py:17-42) raises IndexError / ValueError on degenerate beats; the
    drop-in retraces the index walk that decides it (``degenerate_beat_error``, host integers only).  Against the
    reference's own function on beats with missing fiducials, missing baselines, everything outside the axis, maps
    that are too short -- same exception type, or none where the reference returns."""
    from oracle import ref_import, ref_port as rp
    if not ref_import.available():
        pytest.skip("reference modules not present")
    from ae_waveforms_b200.functions_traditional.signal_reconstruction import degenerate_beat_error
    ref_fn = ref_import.load().signal_reconstruction.reconstruct_gaussian
    x_t = rp.beat_time_axis()
    rng = np.random.default_rng(5)
    base_x = np.array([-180.0, -150, -120, -40, -25, 0, 25, 45, 150, 250, 330])
    base_y = np.array([0.0, 0.1, 0.0, -0.1, 0.0, 1.0, -0.2, 0.0, 0.0, 0.3, 0.0])
    full = [1, 2, 2, 3, 2, 3, 3, 2, 3, 2, 2, 1]
    cases = [(base_x, base_y, 0.0, full)]
    for k in range(40):
        x, y = base_x.copy(), base_y.copy()
        x[rng.random(11) < rng.choice([0.2, 0.6, 0.95])] = np.nan
        y[rng.random(11) < 0.1] = np.nan
        bl = np.nan if k % 5 == 0 else 0.05
        smap = full[:rng.integers(1, 13)] if k % 4 == 0 else full
        cases.append((x, y, bl, smap))
    cases.append((np.full(11, np.nan), base_y, 0.0, full))
    cases.append((np.full(11, np.nan), base_y, np.nan, full))
    cases.append((base_x + 5000.0, base_y, 0.0, full))                    # everything right of the axis
    cases.append((np.full(11, 10.0), base_y, 0.0, full))                  # eleven duplicates, bumped apart by 2 ms
    seen = set()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for x, y, bl, smap in cases:
            try:
                ref_fn(x, y, x_t, bl, smap)
                want = None
            except (IndexError, ValueError) as exc:
                want = type(exc)
            xc, _, mask = rp.extract_coordinates(x, y, x_t, bl)
            got = degenerate_beat_error(xc, mask, x_t, len(smap))
            assert (None if got is None else type(got)) == want, (x, y, bl, smap, got, want)
            seen.add(want)
    assert seen == {None, IndexError, ValueError}
