"""Two ranks over NCCL (multi-GPU tier): the case-sharded result after the gather equals the
single-GPU rows bit for bit, and the all-reduced sigma-search error sums pick the same candidate.  Needs two GPUs
(skipped otherwise; run with ``gpurun --gpus 2``).  The gather / sharding LOGIC is also covered on CPU tensors
under gloo in tests/test_host_logic.py."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, lens, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ae_waveforms_b200 import batched, dist as awd
    from ae_waveforms_b200.sigma_search import SigmaSearch
    from oracle import ecg_synth
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cases = [ecg_synth.synth_case(100 + i, int(n), nan_gaps=i % 3, hr_bpm=60.0 if i % 4 == 0 else None)
                 for i, n in enumerate(lens)]
        ids = np.arange(100, 100 + len(lens))
        mine = awd.shard_cases(lens, world, rank)
        pre = batched.Preprocessor(device=dev)
        res = pre.run(batched.RecordBatch.from_host([cases[i] for i in mine], ids[mine], device=dev), features=True)
        got = awd.gather_beat_arrays(res["np_waves"], res["np_info"], np_features=res["np_features"], dst=0)
        allg = awd.gather_beat_arrays(res["np_waves"], res["np_info"])            # the all-gather form, every rank
        # sigma search: beats sharded, per-candidate sums all-reduced
        tt, fx, fy, bl = ecg_synth.synth_fiducials(4096, seed=5)
        cands = ecg_synth.sigma_candidates(96)
        sl = slice(rank * 2048, (rank + 1) * 2048)
        sse = awd.allreduce_sse(SigmaSearch(fx[sl], fy[sl], bl[sl], tt[sl], device=dev).sse(cands))
        if rank == 0:
            one = batched.Preprocessor(device=dev).run(batched.RecordBatch.from_host(cases, ids, device=dev),
                                                       features=True)
            full = SigmaSearch(fx, fy, bl, tt, device=dev).sse(cands)
            np.savez(os.path.join(tmpdir, "out.npz"),
                     waves=got[0].cpu().numpy(), info=got[1].cpu().numpy(), feats=got[2].cpu().numpy(),
                     waves1=one["np_waves"].cpu().numpy(), info1=one["np_info"].cpu().numpy(),
                     feats1=one["np_features"].cpu().numpy(), allg_waves=allg[0].cpu().numpy(),
                     allg_info=allg[1].cpu().numpy(), sse=sse.cpu().numpy(), sse1=full.cpu().numpy())
        else:
            assert got[0] is None
            np.savez(os.path.join(tmpdir, "rank1.npz"), allg_info=allg[1].cpu().numpy(), sse=sse.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    lens = np.array([60000, 45000, 90001, 20001, 75000, 30000, 120000], dtype=np.int64)
    mp.spawn(_worker, args=(2, 29531, lens, str(tmp_path)), nprocs=2, join=True)
    r = np.load(tmp_path / "out.npz")
    assert r["waves"].shape[0] > 20
    assert np.array_equal(r["waves"], r["waves1"]) and np.array_equal(r["info"], r["info1"])
    assert np.array_equal(r["feats"], r["feats1"], equal_nan=True)
    assert np.array_equal(r["allg_waves"], r["waves1"]) and np.array_equal(r["allg_info"], r["info1"])
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r1["allg_info"], r["info1"])                       # every rank holds the gathered rows
    assert np.array_equal(r1["sse"], r["sse"])                                # and the same reduced sums
    ok = np.isfinite(r["sse1"])
    assert np.allclose(r["sse"][ok], r["sse1"][ok], rtol=1e-12, atol=0)
    assert np.nanargmin(r["sse"]) == np.nanargmin(r["sse1"])


def test_object_on_a_device_that_is_not_current(tmp_path):
    """Preprocessor / SigmaSearch / run_host_pipelined bound to cuda:1 while cuda:0 is the current device: every
    launch goes to the object's device and stream (the library itself never calls cudaSetDevice)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, ROOT)
    from ae_waveforms_b200 import batched
    from ae_waveforms_b200.sigma_search import SigmaSearch
    from oracle import ecg_synth
    torch.cuda.set_device(0)
    cases = [ecg_synth.synth_case(300 + i, 50000 + 1000 * i, hr_bpm=60.0 if i == 1 else None) for i in range(3)]
    want = batched.Preprocessor(device="cuda:0").run(batched.RecordBatch.from_host(cases, device="cuda:0"), features=True)
    pre1 = batched.Preprocessor(device="cuda:1")
    got = pre1.run(batched.RecordBatch.from_host(cases, device="cuda:1"), features=True)
    assert torch.cuda.current_device() == 0 and got["np_waves"].device.index == 1
    for k in ("np_waves", "np_info", "np_features"):
        assert np.array_equal(got[k].cpu().numpy(), want[k].cpu().numpy(), equal_nan=True), k
    lens = np.array([len(c) for c in cases])
    host = torch.from_numpy(np.concatenate(cases)).pin_memory()
    piped = batched.run_host_pipelined(pre1, host, np.concatenate([[0], np.cumsum(lens)]), chunk_cases=1)
    assert np.array_equal(piped["np_waves"].numpy(), want["np_waves"].cpu().numpy())
    tt, fx, fy, bl = ecg_synth.synth_fiducials(256, seed=2)
    cands = ecg_synth.sigma_candidates(8)
    a = SigmaSearch(fx, fy, bl, tt, device="cuda:1").sse(cands)
    b = SigmaSearch(fx, fy, bl, tt, device="cuda:0").sse(cands)
    assert a.device.index == 1 and np.array_equal(a.cpu().numpy(), b.cpu().numpy(), equal_nan=True)
