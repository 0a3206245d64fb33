"""Batched sigma-map search for the Gaussian fiducial reconstruction (K5).

The reference drives ``objective_function`` (analysis_performance_traditional_vs_ae.py:63-90)
through 2,000 sequential ``skopt.gp_minimize`` calls, each a Python loop over every test
beat.  Here the objective of MANY candidate sigma maps is evaluated in one launch;
``best`` is the argmin over the same candidate list.  With several ranks the beats are
sharded and the per-candidate error sums are all-reduced (``dist.allreduce_sse``).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi

BEAT_LEN, BEAT_PRE, BEAT_POST = 320, 106, 214


def beat_time_axis(n: int = BEAT_LEN, fs: float = 500) -> np.ndarray:
    """analysis_performance_traditional_vs_ae.py:48."""
    return np.linspace(-BEAT_PRE, BEAT_POST - 1, n) / fs * 1000


def prepare_inputs(np_waves, np_features):
    """analysis_...:33-46 -- float32 targets (median beat minus its row maximum), the
    ten fiducials reordered by time of occurrence with the R peak inserted at slot 5 as
    (0, -1e-10), and the baseline = nanmean of the six on/offset amplitudes.
    Pure index shuffling on the host; the arithmetic heavy part runs on the device."""
    import warnings
    waves = np.asarray(np_waves)
    feats = np.asarray(np_features, dtype=np.float64)
    target = (waves - waves.max(axis=1, keepdims=True)).astype(np.float32)
    fx = np.insert(feats[:, [2, 0, 4, 8, 6, 12, 10, 16, 14, 18]], 5, 0.0, axis=1)
    fy = np.insert(feats[:, [3, 1, 5, 9, 7, 13, 11, 17, 15, 19]], 5, -1e-10, axis=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        baselines = np.nanmean(feats[:, [3, 5, 9, 11, 17, 19]], axis=1)
    return target, fx, fy, baselines


class SigmaSearch:
    def __init__(self, features_x, features_y, baselines, target=None, x_time=None, device=None):
        self.torch, self.lib = _cabi.require_cuda()
        torch = self.torch
        self.device = torch.device(device or "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())

        def dev(a, dt):
            if isinstance(a, torch.Tensor):
                return a.to(self.device, dt).contiguous()
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device, dt).contiguous()

        self.fx, self.fy = dev(features_x, torch.float64), dev(features_y, torch.float64)
        self.baselines = dev(baselines, torch.float64)
        self.target = None if target is None else dev(target, torch.float32)
        self.xt = dev(beat_time_axis() if x_time is None else x_time, torch.float64)
        self.n_beats, self.n_t = self.fx.shape[0], self.xt.numel()
        if self.fx.shape[1] != 11 or self.fy.shape != self.fx.shape:
            raise _cabi.AewError("features_x / features_y must be (n_beats, 11)")
        self.launches = 0

    @_cabi.on_device
    def _call(self, cand, want_pred):
        torch, lib = self.torch, self.lib
        cand = cand if isinstance(cand, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cand, dtype=np.float64))
        cand = cand.to(self.device, torch.float64).reshape(-1, 12).contiguous()
        n_cand = cand.shape[0]
        sse = torch.empty(n_cand, dtype=torch.float64, device=self.device)
        pred = torch.empty((self.n_beats, self.n_t), dtype=torch.float64, device=self.device) if want_pred else None
        ws_bytes = int(lib.aew_recon_workspace_bytes(self.n_beats, n_cand))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        _cabi.check(lib.aew_recon_sse_f64(p(self.fx), p(self.fy), p(self.baselines), p(self.target), self.n_beats,
                                          p(self.xt), self.n_t, p(cand), n_cand, p(sse), p(pred), p(ws), ws_bytes,
                                          _cabi.stream_ptr(torch)), "recon_sse")
        self.launches += 2
        return sse, pred

    @_cabi.on_device
    def sse_grid(self, values, index):
        """Objective of candidates drawn from a discrete grid: candidate c uses sigma_map[s] = values[index[c, s]].
        One pass over the beats builds the separable error tables, then every candidate is scored from them
        (see aew_recon_sse_grid_f64) -- the cost no longer grows with beats x candidates."""
        torch, lib = self.torch, self.lib
        if self.target is None:
            raise _cabi.AewError("targets are needed for the objective")
        dev = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(self.device, dt)
        vals = dev(values, torch.float64).reshape(-1).contiguous()
        idx = dev(index, torch.uint8).reshape(-1, 12).contiguous()
        if vals.numel() < 1 or vals.numel() > int(lib.aew_recon_max_grid_values()):
            raise _cabi.AewError(f"the table formulation takes 1..{lib.aew_recon_max_grid_values()} grid values")
        n_cand = idx.shape[0]
        sse = torch.empty(n_cand, dtype=torch.float64, device=self.device)
        ws_bytes = int(lib.aew_recon_tables_workspace_bytes(self.n_beats))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _cabi.check(lib.aew_recon_sse_grid_f64(p(self.fx), p(self.fy), p(self.baselines), p(self.target), self.n_beats,
                                               p(self.xt), self.n_t, p(vals), vals.numel(), p(idx), n_cand, p(sse),
                                               p(ws), ws_bytes, _cabi.stream_ptr(torch)), "recon_sse_grid")
        self.launches += 3
        return sse

    @_cabi.on_device
    def sse(self, candidates, method="auto"):
        """Sum over beats and samples of (float32(pred) - target)^2 per candidate (CUDA tensor).
        method "direct" evaluates every (beat, candidate) pair; "tables" requires the candidates to take at most
        aew_recon_max_grid_values() distinct values (the reference's Integer(1, 5) space does) and goes through
        sse_grid; "auto" picks tables whenever they apply and there are enough candidates to pay for them."""
        torch = self.torch
        if self.target is None:
            raise _cabi.AewError("targets are needed for the objective")
        if method not in ("auto", "direct", "tables"):
            raise ValueError(method)
        if method != "direct":
            c = candidates if isinstance(candidates, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(candidates, dtype=np.float64))
            c = c.to(self.device, torch.float64).reshape(-1, 12)
            if method == "tables" or c.shape[0] >= 32:
                vals = torch.unique(c)                 # sorted; NaN entries never compare equal, so they fail the test below
                if vals.numel() <= int(self.lib.aew_recon_max_grid_values()) and not bool(torch.isnan(vals).any()):
                    return self.sse_grid(vals, torch.searchsorted(vals, c.contiguous()).to(torch.uint8))
                if method == "tables":
                    raise _cabi.AewError("candidates take too many distinct values for the table formulation")
            candidates = c
        return self._call(candidates, False)[0]

    def mse(self, candidates):
        """The reference's loss: torch.nn.MSELoss over all beats and samples."""
        return self.sse(candidates) / float(self.n_beats * self.n_t)

    @_cabi.on_device
    def reconstruct(self, sigma_map):
        """(n_beats, 320) float64 reconstructions for one sigma map (analysis_...:121-132)."""
        return self._call(np.asarray(sigma_map, dtype=np.float64).reshape(1, 12), True)[1]

    @_cabi.on_device
    def metrics(self, sigma_map, medians, compression_ratio: float = 320 / 20):
        """Per-beat RMSE, PRD and quality score of the reconstruction for one sigma map against the float64
        median beats (K9; analysis_...:178-197, functions_ae/loss_functions.py:8-20) plus the summary the
        reference prints: median / quartiles / mean / sd of RMSE and QS (np.percentile = linear interpolation)."""
        torch = self.torch
        med = torch.as_tensor(np.ascontiguousarray(medians, dtype=np.float64) if not torch.is_tensor(medians)
                              else medians, dtype=torch.float64, device=self.device).contiguous()
        if tuple(med.shape) != (self.n_beats, self.n_t):
            raise _cabi.AewError("medians must be (n_beats, n_t)")
        recon = self.reconstruct(sigma_map)
        out = [torch.empty(self.n_beats, dtype=torch.float32, device=self.device) for _ in range(3)]
        _cabi.check(self.lib.aew_beat_metrics_f32(
            ctypes.c_void_p(recon.data_ptr()), ctypes.c_void_p(med.data_ptr()), self.n_beats, self.n_t,
            float(compression_ratio), *[ctypes.c_void_p(t.data_ptr()) for t in out],
            _cabi.stream_ptr(torch)), "beat_metrics")
        rmse, prd, qs = out
        q = torch.tensor([0.25, 0.5, 0.75], dtype=torch.float64, device=self.device)

        def summary(v):
            v = v.to(torch.float64)
            ok = v[~torch.isnan(v)]
            qq = torch.quantile(ok, q) if ok.numel() else torch.full((3,), float("nan"), device=self.device)
            return {"q1": float(qq[0]), "median": float(qq[1]), "q3": float(qq[2]),
                    "mean": float(v.mean()), "sd": float(v.std(unbiased=False))}
        return {"rmse": rmse, "prd": prd, "qs": qs, "recon": recon,
                "summary": {"rmse": summary(rmse), "qs": summary(qs), "prd_mean": float(prd.mean())}}

    @_cabi.on_device
    def best(self, candidates, sse=None):
        """argmin over the candidate list (NaN losses never win)."""
        torch = self.torch
        s = self.sse(candidates) if sse is None else sse
        s = torch.where(torch.isnan(s), torch.full_like(s, float("inf")), s)
        i = int(torch.argmin(s).item())
        return i, float(s[i].item()) / float(self.n_beats * self.n_t)
