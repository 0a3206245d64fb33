"""Device-resident batched API of the preprocessing hot path.

This is the throughput path: many cases per launch, every buffer a CUDA tensor,
kernels called through the C ABI (``_cabi``).  The per-call drop-in functions in
``ae_waveforms_b200.data_load_preprocess`` are thin wrappers over the same calls.

Pipeline of ``Preprocessor.run`` = data_extract_vitaldb.py:53-106 for a batch:

    trim all-NaN edges (:53-57)            aew_trim_nan_edges_f64
    whole-record z-score statistics (:64)  aew_zscore_stats_f64        (K1)
    NaN-delimited segments                 aew_segments_find_f64
    high-pass, low-pass filtfilt (:66-76)  aew_filtfilt_f64 x 2        (K2; z-score fused into the load)
    20 s windows (:85-86)                  window table built on the host (index arithmetic only)
    per-window z-score, R-peaks, screen,
    median beat (:93-106)                  aew_windows_beats_f64       (K3 + K4)
    wave delineation (:131)                aew_delineate_f64           (K7; nk.ecg_delineate, "dwt")
    35 features per window (:146-162)      aew_fiducial_features_f64   (K6)
    np_waves / np_info rows                compaction of the kept windows (data_create_arrays.py:41-46)
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _cabi

BEAT_LEN, BEAT_PRE, BEAT_POST = 320, 106, 214        # at 500 Hz


def beat_geometry(fs):
    """median_sample.py:20-25 -- (section, pre, post) = int(0.8 * (60 / 75) * fs) samples, a third (rounded down)
    before the R peak: 320 / 106 / 214 at 500 Hz, 160 / 53 / 107 at 250 Hz."""
    section = int(0.8 * (60 / 75) * fs)
    pre = int(1 / 3 * section)
    return section, pre, section - pre


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


@dataclass
class FilterDesign:
    """Host-side design (SciPy, so b, a, zi carry SciPy's bits) + device table."""
    btype: str
    cutoff_hz: float
    fs: float
    order: int
    b: np.ndarray = field(repr=False, default=None)
    a: np.ndarray = field(repr=False, default=None)
    zi: np.ndarray = field(repr=False, default=None)
    table: object = field(repr=False, default=None)     # torch tensor on the device
    table_host: np.ndarray = field(repr=False, default=None)   # the same table on the host (kernel parameters)
    fast_ok: bool = True

    @property
    def padlen(self) -> int:
        return 3 * max(len(self.a), len(self.b))


_design_cache: dict = {}


def design_filter(btype: str, cutoff_hz: float, fs: float = 500, order: int = 4, device=None) -> FilterDesign:
    """filters.py:29-31 -- butter(N, Wn, btype) in (b, a) form, plus lfilter_zi."""
    torch, lib = _cabi.require_cuda()
    device = torch.device(device or "cuda")
    key = (btype, float(cutoff_hz), float(fs), int(order), str(device))
    if key in _design_cache:
        return _design_cache[key]
    from scipy.signal import butter, lfilter_zi
    b, a = butter(N=order, Wn=cutoff_hz / (0.5 * fs), btype=btype, analog=False)
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    zi = np.ascontiguousarray(lfilter_zi(b, a), dtype=np.float64)
    if len(b) != len(a) or len(a) - 1 > 4:
        raise _cabi.AewError("only low/high-pass designs up to order 4 are supported")
    tab = np.zeros(int(lib.aew_filt_tables_len()), dtype=np.float64)
    rc = lib.aew_filt_tables(_np_ptr(b), _np_ptr(a), _np_ptr(zi), len(a) - 1, _np_ptr(tab))
    if rc not in (0, -3):
        _cabi.check(rc, "aew_filt_tables")
    d = FilterDesign(btype, cutoff_hz, fs, order, b, a, zi, torch.from_numpy(tab).to(device), tab, fast_ok=(rc == 0))
    _design_cache[key] = d
    return d


def window_count(n_rows: int, L: int) -> int:
    """len(np.arange(L, n_rows, L)) -- sample_from_record.py:21 (exclusive stop)."""
    return max(0, (int(n_rows) - 1) // L) if n_rows > L else 0


class RecordBatch:
    """A ragged batch: one flat float64 CUDA tensor + host offsets."""

    def __init__(self, flat, rec_off: np.ndarray, case_ids=None):
        self.flat = flat
        self.rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
        n = len(self.rec_off) - 1
        self.case_ids = np.arange(n, dtype=np.int64) if case_ids is None else np.asarray(case_ids, dtype=np.int64)

    @classmethod
    def from_host(cls, records, case_ids=None, device=None, pinned=None):
        """records: list of 1-D float arrays (one lead each).  ``pinned`` may be a
        pre-filled pinned flat tensor (then ``records`` is only used for lengths)."""
        torch, _ = _cabi.require_cuda()
        device = torch.device(device or "cuda")
        lens = np.array([len(r) for r in records], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        if pinned is None:
            host = torch.empty(int(off[-1]), dtype=torch.float64, pin_memory=True)
            hv = host.numpy()
            for r, s in zip(records, off[:-1]):
                hv[s:s + len(r)] = np.asarray(r, dtype=np.float64)
        else:
            host = pinned
        return cls(host.to(device, non_blocking=True), off, case_ids)

    @property
    def n_rec(self) -> int:
        return len(self.rec_off) - 1

    @property
    def total(self) -> int:
        return int(self.rec_off[-1])


class Preprocessor:
    """filter -> windows -> R-peaks -> screen -> median beat for a RecordBatch."""

    def __init__(self, fs: int = 500, sample_length: int = 20, hp: float = 1, lp: float = 30, order: int = 4,
                 pvc_threshold: float = 150, sd_threshold: float = 0.05, filter_mode: str = "fast",
                 screen_mode: str = "auto", max_segments: int | None = None, device=None,
                 exact_masks: bool = True):
        self.torch, self.lib = _cabi.require_cuda()
        self.device = self.torch.device(device or "cuda")
        if self.device.index is None:                      # bind to the device that is current NOW
            self.device = self.torch.device("cuda", self.torch.cuda.current_device())
        self.fs, self.L = int(fs), int(fs * sample_length)
        self.beat_len, self.beat_pre, self.beat_post = beat_geometry(self.fs)
        self.hp, self.lp, self.order = hp, lp, order
        self.pvc_threshold, self.sd_threshold = float(pvc_threshold), float(sd_threshold)
        self.filter_mode = {"fast": 0, "exact": 1}[filter_mode]
        self.screen_mode = {"auto": 0, "exact": 1, "fast": 2, "forced_fallback": 3}[screen_mode]
        self.max_segments = max_segments
        # With the tiled (fast) filter the windows are ~1e-9 from SciPy's, which flips the last bit of np.std in a
        # z-scored window and with it the rule sd / n_peaks > 0.05 exactly when n_peaks == 20.
        # exact_masks: the records that hold such windows are filtered AGAIN in SciPy's literal order and those
        # windows re-screened from that output, so every keep / remove decision is the reference's bit for bit.
        self.exact_masks = bool(exact_masks) and self.filter_mode == 0
        self.defer_chunks = 6               # run_host_pipelined: chunks finished this many chunks late (exact_masks)
        # side streams of the exact-mask passes (round robin): at least as many as calls a caller keeps in flight,
        # or two passes queue behind each other on one stream
        self.side_stream_count = int(__import__("os").environ.get("AEW_SIDE_STREAMS", "12"))
        # short-memory filters (the low-pass) CAN run forward + backward in one launch (k2_fused, AEW_K2_FUSE=1): half
        # the HBM traffic, but measured SLOWER (3.14 against 2.59 ms per 200 cases): the filter kernels are bound by
        # their FP64 instruction stream (~35 % of the measured DFMA peak), not by bytes, and the halo adds 25 % of it
        self.fuse_passes = bool(int(__import__("os").environ.get("AEW_K2_FUSE", "0")))
        self.lazy_sd = True                 # run(): 'ecg_sd' is not an output; evaluate it only where it decides
        # the screen kernel parks the z-scored windows in the filter scratch for the median kernel (measured: 30.1 ms
        # per 179,000 windows against 32.5 when the median kernel re-derives them; profiles/k3_variants.py)
        self.park_windows = bool(int(__import__("os").environ.get("AEW_K3_PARK", "1")))
        # Behind the tiled filter (~1e-9 from SciPy's output) NumPy's summation order in the per-window z-score buys
        # nothing: the screen kernel then takes mean / sd from one pass of shifted sums, parks nothing, and the median
        # kernel z-scores the middle values of its columns (K3Params.fast_stats).  The windows whose decision hangs
        # on np.std's last bit go through the bit-exact chain anyway (exact_masks); so do filter_mode="exact" and the
        # feature path (the delineation kernel reads the parked z-scored windows).
        self.fast_stats = self.filter_mode == 0 and bool(int(__import__("os").environ.get("AEW_K3_FAST_STATS", "1")))
        self.launches = 0
        if self.L > self.lib.aew_windows_max_len():
            raise _cabi.AewError(f"window of {self.L} samples exceeds the kernel limit")

    # ------------------------------------------------------------------ stages
    UPLOAD_RING_BYTES = 32 << 20

    @_cabi.on_device
    def _upload(self, arr: np.ndarray):
        """Host table -> device tensor WITHOUT the copy engine: the bytes go into a page-locked ring and a small
        kernel (``aew_fetch_host``) reads them over PCIe.  A ``.to(device)`` here would queue behind whatever bulk
        host->device transfer is in flight (``run_host_pipelined`` keeps one in flight all the time) and stall
        the kernels waiting for the table until that transfer has finished."""
        torch = self.torch
        arr = np.ascontiguousarray(arr)
        out = torch.empty(arr.shape, dtype=torch.from_numpy(arr[:0].reshape(-1)).dtype, device=self.device)
        nb = arr.nbytes
        if nb == 0:
            return out
        if nb > self.UPLOAD_RING_BYTES // 2 or out.data_ptr() % 16:
            out.copy_(torch.from_numpy(arr))
            return out
        ring = getattr(self, "_up_ring", None)
        if ring is None:
            ring = self._up_ring = torch.empty(self.UPLOAD_RING_BYTES, dtype=torch.uint8, pin_memory=True)
            self._up_ring_np = ring.numpy()
            self._up_pos = 0
            self._up_events = {}
        if self._up_pos + nb > self.UPLOAD_RING_BYTES:        # wrap: every fetch issued so far must have read its bytes
            # (the fetch kernels only -- one event per stream that ever fetched -- not the whole device: a device-wide
            #  wait here also waited for the exact-mask passes in flight on the side streams, ~300 ms every few calls)
            for ev in self._up_events.values():
                ev.synchronize()
            self._up_events.clear()
            self._up_pos = 0
        pos = self._up_pos
        self._up_ring_np[pos:pos + nb] = arr.reshape(-1).view(np.uint8)
        _cabi.check(self.lib.aew_fetch_host(ring.data_ptr() + pos, nb, out.data_ptr(), _cabi.stream_ptr(torch)),
                    "fetch_host")
        self.launches += 1
        st = torch.cuda.current_stream(self.device)
        ev = self._up_events.get(st.cuda_stream)
        if ev is None:
            ev = self._up_events[st.cuda_stream] = torch.cuda.Event()
        ev.record(st)
        self._up_pos = (pos + nb + 15) & ~15
        return out

    @_cabi.on_device
    def _ranges(self, batch: RecordBatch, trim: bool):
        torch, lib = self.torch, self.lib
        off = self._upload(batch.rec_off)
        beg_h, end_h = batch.rec_off[:-1].copy(), batch.rec_off[1:].copy()
        if trim and batch.n_rec:
            trimmed = torch.empty(2 * batch.n_rec, dtype=torch.int64, device=self.device)
            _cabi.check(lib.aew_trim_nan_edges_f64(_ptr(batch.flat), _ptr(off), _ptr(off[1:]), batch.n_rec,
                                                   _ptr(trimmed), _cabi.stream_ptr(torch)), "trim")
            self.launches += 1
            t = trimmed.cpu().numpy().reshape(-1, 2)           # the one small D2H sync of the pipeline
            beg_h, end_h = np.ascontiguousarray(t[:, 0]), np.ascontiguousarray(t[:, 1])
            tr = trimmed.view(-1, 2)                           # already on the device: no upload of the bounds
            return beg_h, end_h, tr[:, 0].contiguous(), tr[:, 1].contiguous()
        return beg_h, end_h, off[:-1].contiguous(), off[1:].contiguous()

    @_cabi.on_device
    def zscore_stats(self, flat, beg_h, end_h, d_beg, d_end, fast=False):
        """K1: per-record mean, sd (NumPy bits) and non-NaN count.  ``fast``: from one sweep over the samples
        instead of two, within a few ulp of NumPy's but not bit-identical (``aew_zscore_stats_fast_f64``)."""
        torch, lib = self.torch, self.lib
        n_rec = len(beg_h)
        blk = np.zeros(n_rec + 1, dtype=np.int64)
        n_blocks = int(lib.aew_zscore_plan(_np_ptr(beg_h), _np_ptr(end_h), n_rec, _np_ptr(blk)))
        d_blk = self._upload(blk)
        mean = torch.empty(n_rec, dtype=torch.float64, device=self.device)
        sd = torch.empty_like(mean)
        cnt = torch.empty(n_rec, dtype=torch.int64, device=self.device)
        ws_bytes = int(lib.aew_zscore_workspace_bytes(n_rec, n_blocks))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        entry = lib.aew_zscore_stats_fast_f64 if fast else lib.aew_zscore_stats_f64
        _cabi.check(entry(_ptr(flat), _ptr(d_beg), _ptr(d_end), _ptr(d_blk), n_rec, n_blocks, _ptr(mean), _ptr(sd),
                          _ptr(cnt), _ptr(ws), ws_bytes, _cabi.stream_ptr(torch)), "zscore_stats")
        self.launches += 2 if fast else 4
        return mean, sd, cnt

    @_cabi.on_device
    def filter_records(self, flat, beg_h, end_h, d_beg, d_end, mean=None, sd=None, cnt=None, out=None, tmp=None,
                       mode=None, max_seg=None):
        """K2: z-score (fused) + high-pass + low-pass per NaN-delimited segment.
        Returns (filtered, device header tensor).  ``mode`` overrides the object's filter mode (0 tiled, 1 exact with
        one lane per segment, 3 exact with one warp per segment).  The filtered signal comes back in ``out`` OR in
        ``tmp``: a stage whose filter has a short memory (the 30 Hz low-pass) runs forward and backward in ONE launch
        from one buffer into the other, so the two buffers may change roles; the other one is the scratch
        (``self._filter_tmp`` when ``mode`` is None)."""
        torch, lib = self.torch, self.lib
        n_rec, total = len(beg_h), flat.numel()
        max_seg = max_seg or self.max_segments or (4 * n_rec + 4096)
        ws_bytes = int(lib.aew_filt_workspace_bytes(n_rec, total, max_seg))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        y = torch.empty_like(flat) if out is None else out
        tmp = torch.empty_like(flat) if tmp is None else tmp
        filter_mode = self.filter_mode if mode is None else mode
        st = _cabi.stream_ptr(torch)
        stages = [d for d in (self.hp and design_filter("high", self.hp, self.fs, self.order, self.device),
                              self.lp and design_filter("low", self.lp, self.fs, self.order, self.device)) if d]
        # plan: which stage takes the one-launch form, and in which buffer the result ends up (that one gets the NaN
        # fill at NaN inputs; the kernels write segment samples only)
        fused = [filter_mode == 0 and self.fuse_passes and d.fast_ok and bool(lib.aew_filt_fused_ok(_np_ptr(d.table_host)))
                 for d in stages]
        bufs = [y, tmp]
        where, cur = 0, -1                                   # cur: -1 = the input, 0 / 1 = bufs[.]
        for f in fused:
            if f:
                where = 1 - cur if cur >= 0 else 0
            else:
                where = cur if cur >= 0 else 0
            cur = where
        final = bufs[where]
        _cabi.check(lib.aew_segments_find_f64(_ptr(flat), _ptr(d_beg), _ptr(d_end), n_rec, total, _ptr(cnt),
                                              _ptr(final), _ptr(ws), ws_bytes, max_seg, st), "segments_find")
        self.launches += 1
        src, m, s, cur = flat, mean, sd, -1
        for d, f in zip(stages, fused):
            if filter_mode == 0 and not d.fast_ok:
                raise _cabi.AewError(f"{d}: no modal form; use filter_mode='exact'")
            if f:                                            # one launch, src -> the buffer src is not in
                dst = 1 - cur if cur >= 0 else 0
                _cabi.check(lib.aew_filtfilt_f64(_ptr(src), _ptr(m), _ptr(s), _ptr(bufs[dst]), _ptr(bufs[1 - dst]),
                                                 _ptr(d.table), _np_ptr(d.table_host), n_rec, total, _ptr(ws),
                                                 ws_bytes, max_seg, 2, st), "filtfilt (one launch)")
                self.launches += 1
            else:                                            # forward into the other buffer, backward into dst
                dst = cur if cur >= 0 else 0
                _cabi.check(lib.aew_filtfilt_f64(_ptr(src), _ptr(m), _ptr(s), _ptr(bufs[dst]), _ptr(bufs[1 - dst]),
                                                 _ptr(d.table), _np_ptr(d.table_host), n_rec, total, _ptr(ws),
                                                 ws_bytes, max_seg, filter_mode, st), "filtfilt")
                self.launches += 1 if filter_mode else 2
            cur, src, m, s = dst, bufs[dst], None, None
        if not stages:                                       # both cut-offs falsy: z-score only
            if mean is not None:
                _cabi.check(lib.aew_zscore_apply_f64(_ptr(flat), _ptr(d_beg), _ptr(d_end), n_rec, _ptr(mean),
                                                     _ptr(sd), _ptr(y), 0, total, st), "zscore_apply")
                self.launches += 1
            else:
                y.copy_(flat)
        y, tmp = final, bufs[1 - where]
        if mode is None:
            self._filter_tmp = tmp              # free again once the filter is done: scratch for the window kernels
        off = (-ws.data_ptr()) % 128
        header = ws[off:off + 64].view(torch.int64)
        return y, header

    def window_table(self, beg_h, end_h, case_ids):
        """sample_from_record.py:19-30 as index arithmetic: flat start, case and end
        index of every window."""
        L = self.L
        counts = np.array([window_count(e - b, L) for b, e in zip(beg_h, end_h)], dtype=np.int64)
        n_win = int(counts.sum())
        rec_of = np.repeat(np.arange(len(counts)), counts)
        first = np.concatenate([[0], np.cumsum(counts)])[:-1]
        k = np.arange(n_win) - np.repeat(first, counts)
        start = beg_h[rec_of] + k * L if n_win else np.zeros(0, np.int64)
        return (np.ascontiguousarray(start, dtype=np.int64), np.asarray(case_ids)[rec_of] if n_win else np.zeros(0, np.int64),
                ((k + 1) * L).astype(np.int64))

    @_cabi.on_device
    def screen_windows(self, y, win_start_h, want_median=True, want_zwin=False, zscore_windows=True,
                       pvc_threshold=None, sd_threshold=None, L=None, zwin_scratch=None, lazy_sd=False,
                       d_start=None, fast_stats=False):
        """K3 (window kernel + beat kernel) + K4 on the windows whose flat offsets are given.  ``zwin_scratch``:
        optional float64 tensor of at least n_win * L elements; the window kernel parks the z-scored windows there
        for the beat and median kernels (a buffer of this call otherwise).  ``fast_stats`` (with ``lazy_sd``, windows
        z-scored here, not with ``want_zwin`` or the sequential screen mode): the per-window z-score without NumPy's
        summation order and without the parked copy (``K3Params.fast_stats`` in include/aewave.h) -- for the output
        of the tiled filter, which is itself ~1e-9 from SciPy's."""
        torch, lib = self.torch, self.lib
        L = L or self.L
        n_win = len(win_start_h)
        maxp = lib.aew_windows_max_peaks()
        dev = self.device
        out = dict(
            keep=torch.zeros(n_win, dtype=torch.uint8, device=dev),
            has_nan=torch.zeros(n_win, dtype=torch.uint8, device=dev),
            n_peaks=torch.zeros(n_win, dtype=torch.int32, device=dev),
            peaks=torch.empty((n_win, maxp), dtype=torch.int32, device=dev),
            stats=torch.empty((n_win, 4), dtype=torch.float64, device=dev),
            median=torch.empty((n_win, self.beat_len), dtype=torch.float64, device=dev) if want_median else None,
            waves=torch.empty((n_win, self.beat_len), dtype=torch.float32, device=dev) if want_median else None,
            zwin=torch.empty((n_win, L), dtype=torch.float64, device=dev) if want_zwin else None,
            zparams=torch.empty((n_win, 2), dtype=torch.float64, device=dev) if want_median else None,
            status=torch.zeros(1, dtype=torch.int32, device=dev),
            hazards=torch.zeros(1, dtype=torch.int64, device=dev))
        if n_win == 0:
            return out
        if d_start is None:
            d_start = self._upload(np.ascontiguousarray(win_start_h, dtype=np.int64))
        # the z-scored windows are parked for the beat and median kernels: in the caller's buffer ('ecg_sample'),
        # in the scratch the caller lends (the filter's second buffer), else in a buffer of this call
        fast_stats = bool(fast_stats and lazy_sd and zscore_windows and not want_zwin and self.screen_mode != 1)
        zwin_ptr = out["zwin"].data_ptr() if want_zwin else None
        if zwin_ptr is None and zscore_windows and not fast_stats:
            if zwin_scratch is not None and zwin_scratch.dtype == torch.float64 \
                    and zwin_scratch.numel() >= n_win * L and zwin_scratch.data_ptr() != y.data_ptr():
                zwin_ptr = zwin_scratch.data_ptr()
            else:
                out["_zwin_own"] = torch.empty((n_win, L), dtype=torch.float64, device=dev)
                zwin_ptr = out["_zwin_own"].data_ptr()
        ws_bytes = int(lib.aew_windows_workspace_bytes(n_win))
        out["_ws"] = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        fs = self.fs
        p = _cabi.K3Params(
            y=y.data_ptr(), win_start=d_start.data_ptr(), n_win=n_win, L=L,
            smooth_k=int(np.rint(0.1 * fs)), avg_k=int(np.rint(0.75 * fs)), refractory=int(np.rint(0.3 * fs)),
            pre=self.beat_pre, post=self.beat_post, hr_num=float(60 * fs), rr_div=float(fs * 1000), hr_lo=30.0, hr_hi=180.0,
            pvc_threshold=float(self.pvc_threshold if pvc_threshold is None else pvc_threshold),
            sd_threshold=float(self.sd_threshold if sd_threshold is None else sd_threshold),
            mode=self.screen_mode, zscore_windows=1 if zscore_windows else 0,
            keep=out["keep"].data_ptr(), has_nan=out["has_nan"].data_ptr(), n_peaks=out["n_peaks"].data_ptr(),
            peaks=out["peaks"].data_ptr(), stats=out["stats"].data_ptr(),
            median=out["median"].data_ptr() if want_median else None,
            waves=out["waves"].data_ptr() if want_median else None,
            zwin=zwin_ptr,
            zparams=out["zparams"].data_ptr() if want_median else None,
            status=out["status"].data_ptr(), hazard_count=out["hazards"].data_ptr(),
            lazy_sd=1 if (lazy_sd and zscore_windows) else 0,
            workspace=out["_ws"].data_ptr(), workspace_bytes=ws_bytes, fast_stats=1 if fast_stats else 0)
        _cabi.check(lib.aew_windows_beats_f64(ctypes.byref(p), _cabi.stream_ptr(torch)), "windows_beats")
        self.launches += 3 if want_median else 2
        out["_win_start"] = d_start
        return out

    @_cabi.on_device
    def delineate_windows(self, sig, sig_off, peaks, n_peaks, select=None, L=None):
        """K7: nk.ecg_delineate(window, r_peaks, fs)[1] (data_extract_vitaldb.py:131) for every selected window.
        ``sig`` flat float64 CUDA tensor, ``sig_off`` int64 CUDA tensor of window offsets into it, ``peaks``
        (n_win, max_peaks) int32, ``n_peaks`` (n_win,) int32, ``select`` optional uint8 mask.
        Returns ``fid`` (n_win, 10, max_peaks) int32 (FID_NAN = np.nan) and ``flags`` (n_win,) uint8:
        0 ok, 1 NaN features / window kept, 2 window dropped, 3 not selected."""
        torch, lib = self.torch, self.lib
        L = L or self.L
        n_win, maxp = int(peaks.shape[0]), int(peaks.shape[1])
        fid = torch.empty((n_win, 10, maxp), dtype=torch.int32, device=self.device)
        flags = torch.full((n_win,), 3, dtype=torch.uint8, device=self.device)
        if n_win == 0:
            return fid, flags
        ws_bytes = int(lib.aew_delineate_workspace_bytes(n_win, L, float(self.fs)))
        ws = getattr(self, "_k7_ws", None)
        if ws is None or ws.numel() < ws_bytes:
            ws = self._k7_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        _cabi.check(lib.aew_delineate_f64(_ptr(sig), _ptr(sig_off), n_win, L, float(self.fs), _ptr(peaks),
                                          _ptr(n_peaks), maxp, _ptr(select), _ptr(fid), _ptr(flags), _ptr(ws),
                                          ws.numel(), _cabi.stream_ptr(torch)), "delineate")
        self.launches += 1
        return fid, flags

    @_cabi.on_device
    def fiducial_features(self, sig, sig_off, peaks, n_peaks, fid, valid=None, L=None):
        """K6: ecg_features_relative_r + compute_ecg_morphology_features + compute_median_features_ecg
        (data_extract_vitaldb.py:146-162) -> (n_win, 35) float64 in np_features column order."""
        torch, lib = self.torch, self.lib
        L = L or self.L
        n_win, maxp = int(peaks.shape[0]), int(peaks.shape[1])
        out = torch.empty((n_win, 35), dtype=torch.float64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        if n_win:
            _cabi.check(lib.aew_fiducial_features_f64(_ptr(sig), _ptr(sig_off), n_win, L, _ptr(peaks), _ptr(n_peaks),
                                                      maxp, _ptr(fid), _ptr(valid), float(self.fs), _ptr(out),
                                                      _ptr(status), _cabi.stream_ptr(torch)), "fiducial_features")
            self.launches += 1
        return out, status

    @_cabi.on_device
    def window_features(self, zwin_flat, n_win, scr, L=None):
        """K7 + K6 on the kept windows of a screen result; ``zwin_flat`` holds the z-scored windows back to back.
        Returns (features (n_win, 35), flags (n_win,))."""
        torch = self.torch
        L = L or self.L
        off = torch.arange(n_win, dtype=torch.int64, device=self.device) * L
        fid, flags = self.delineate_windows(zwin_flat, off, scr["peaks"], scr["n_peaks"], select=scr["keep"], L=L)
        valid = (flags == 0).to(torch.uint8)
        feats, status = self.fiducial_features(zwin_flat, off, scr["peaks"], scr["n_peaks"], fid, valid, L=L)
        return feats, flags, fid, status

    def hazard_peak_count(self):
        """The beat count at which ``sd / n_peaks > sd_threshold`` hangs on the last bit of np.std of a z-scored
        window (sd = 1 +- a few ulp): n with n * threshold == 1 (20 for the reference's 0.05); None if there is none."""
        if not self.sd_threshold > 0:
            return None
        n = int(round(1.0 / self.sd_threshold))
        return n if n >= 1 and abs(n * self.sd_threshold - 1.0) < 1e-9 else None

    @_cabi.on_device
    def resolve_hazards(self, flat, beg_h, end_h, mean, sd, cnt, win_start, win_rec, scr, zwin_main=None,
                        low_latency=False, after=None):
        """exact_masks: windows whose keep / remove decision depends on the last bit of np.std (n_peaks ==
        hazard_peak_count(), every other rule passed) are decided from SciPy-order filter output.  The records that
        hold such windows are gathered into a compact batch, filtered by the sequential kernel (bit-identical to
        scipy.signal.filtfilt), the hazard windows screened again from that output and their rows of ``scr``
        replaced -- keep, peaks, stats, median beat: for these windows everything is the reference's bit for bit.

        The sequential filter is a dependent chain of ~25 cycles per sample and pass: ~100 ms for a 1 h record
        whatever the batch.  It needs a handful of warps, so it runs on a SIDE stream next to the kernels of the
        following batches; this call only launches it.  Returns None (no such window) or a dict with ``event``
        (recorded on the side stream after the rows of ``scr`` have been replaced), ``n`` and ``changed`` (device).

        ``after``: an event recorded behind the batch's screen kernels.  With it nothing here touches the calling
        stream: the hazard windows are selected on a small stream of their own that waits for the event (so the host
        does not wait for whatever the caller has queued since -- the NEXT batch's kernels), and the records are
        gathered on the side stream.  The caller keeps ``flat`` untouched until the batch is finished."""
        torch = self.torch
        nh = self.hazard_peak_count()
        if nh is None or not len(win_start):
            return None
        main = torch.cuda.current_stream(self.device)
        if after is not None:
            sel = getattr(self, "_resolver", None)
            if sel is None:
                sel = self._resolver = torch.cuda.Stream(device=self.device, priority=-1)
        else:
            sel = main
        with torch.cuda.stream(sel):
            if after is not None:
                sel.wait_event(after)
            st = scr["stats"]
            hr, rms = st[:, 0], st[:, 1]
            hz = (scr["n_peaks"] == nh) & (scr["has_nan"] == 0) & (hr >= 30.0) & (hr <= 180.0) \
                & ~(rms > self.pvc_threshold)
            hz_idx = torch.nonzero(hz).squeeze(1)
            if hz_idx.numel() == 0:                              # (sync: the list of hazard windows)
                return None
            hz_h = hz_idx.cpu().numpy()
            recs = np.unique(win_rec[hz_h])
            lens = (end_h[recs] - beg_h[recs]).astype(np.int64)
            # records without NaN are one segment each: the sequential kernel then gets exactly the blocks it needs
            # (every block keeps an SM to itself; see K2E_HOG in k2_filtfilt.cu)
            clean = bool((cnt.cpu().numpy()[recs] == lens).all())
        coff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        pos = np.searchsorted(recs, win_rec[hz_h])
        start_c = np.ascontiguousarray(win_start[hz_h] - beg_h[win_rec[hz_h]] + coff[pos], dtype=np.int64)
        L = self.L

        def side_buffer(numel, stream):
            # The three big buffers of a pass (1.4 GB each at 100 one-hour records) come from ONE pool, whichever
            # side stream uses them: the caching allocator keeps a pool per stream, so buffers allocated under the
            # side streams themselves were never shared between them -- 16 streams x 4.3 GB cached for good, grown by
            # cudaMalloc in the middle of a run (stalls of 100-250 ms per call while it grows,
            # profiles/loop_stalls.py).  Allocated under a stream that runs nothing and marked as used by the side
            # stream, a freed buffer goes to the next pass as soon as this one is through.
            alloc = getattr(self, "_alloc_stream", None)
            if alloc is None:
                alloc = self._alloc_stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(alloc):
                buf = torch.empty(int(numel), dtype=torch.float64, device=self.device)
            buf.record_stream(stream)
            return buf

        def gather():
            # the hazard records' samples and statistics are copied out, so that the sequential pass touches nothing
            # the next batch may overwrite (input buffers, filter scratch)
            xc = side_buffer(int(lens.sum()), torch.cuda.current_stream(self.device))
            torch.cat([flat[int(b):int(e)] for b, e in zip(beg_h[recs], end_h[recs])], out=xc)
            d_rec = self._upload(recs.astype(np.int64))
            d_off = self._upload(coff)
            if self.fast_stats:
                # the batch's statistics came from the one-sweep kernel: these records get NumPy's bits here
                m_c, s_c, n_c = self.zscore_stats(xc, coff[:-1].copy(), coff[1:].copy(), d_off[:-1].contiguous(),
                                                  d_off[1:].contiguous())
            else:
                m_c, s_c, n_c = mean.index_select(0, d_rec), sd.index_select(0, d_rec), cnt.index_select(0, d_rec)
            return (xc, d_rec, d_off, m_c, s_c, n_c, self._upload(start_c))

        # a POOL of side streams, round robin: one pass takes ~150 ms however few records it holds, so the passes of
        # consecutive batches / chunks must run next to each other, not queue behind each other on one stream
        pool = getattr(self, "_side_streams", None)
        if pool is None:
            pool = self._side_streams = [torch.cuda.Stream(device=self.device, priority=-1)
                                         for _ in range(self.side_stream_count)]
            self._side_next = 0
        side = pool[self._side_next % len(pool)]
        self._side_next += 1
        if after is None:
            xc, d_rec, d_off, m_c, s_c, n_c, d_start_c = gather()   # on the calling stream, before it moves on
            xc.record_stream(side)
            ready = torch.cuda.Event()
            ready.record(main)
        else:
            ready = after
            hz_idx.record_stream(side)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            if after is not None:
                xc, d_rec, d_off, m_c, s_c, n_c, d_start_c = gather()
            yc, tc = side_buffer(xc.numel(), side), side_buffer(xc.numel(), side)
            self.filter_records(xc, coff[:-1].copy(), coff[1:].copy(), d_off[:-1].contiguous(), d_off[1:].contiguous(),
                                m_c, s_c, n_c, out=yc, tmp=tc, mode=3 if low_latency else 1,
                                max_seg=len(recs) if clean else None)
            scr2 = self.screen_windows(yc, start_c, zwin_scratch=tc if tc.numel() >= len(hz_h) * L else None,
                                       d_start=d_start_c)
            changed = (scr2["keep"] != scr["keep"].index_select(0, hz_idx)).sum()
            for k in ("keep", "has_nan", "n_peaks", "peaks", "stats", "median", "waves", "zparams"):
                if scr.get(k) is not None and scr2.get(k) is not None:
                    scr[k].index_copy_(0, hz_idx, scr2[k])
            if zwin_main is not None and tc.numel() >= len(hz_h) * L:   # the parked z-scored windows (features path)
                zwin_main[:scr["keep"].shape[0] * L].view(-1, L).index_copy_(0, hz_idx, tc[:len(hz_h) * L].view(-1, L))
            scr["status"] |= scr2["status"]
            scr["hazards"] += scr2["hazards"]
            done = torch.cuda.Event()
            done.record(side)
        return dict(event=done, n=len(hz_h), changed=changed, hold=(xc, yc, tc, scr2, m_c, s_c, n_c, d_rec, d_off,
                                                                  d_start_c, hz_idx))

    # ------------------------------------------------------------------ whole path
    @_cabi.on_device
    def run(self, batch: RecordBatch, trim: bool = True, keep_intermediates: bool = False,
            scratch: dict | None = None, features: bool = False, strips: bool = False, defer: bool = False,
            low_latency: bool = False, lazy_hazards: bool = False):
        """Returns a dict of CUDA tensors: ``np_waves`` (n_keep, 320) float32,
        ``np_info`` (n_keep, 2) float64 = [case_id, idx_end], ``medians`` float64,
        plus per-window ``keep`` / ``has_nan`` / ``n_peaks`` / ``peaks`` / ``stats`` and
        counters.  ``scratch`` may hold reusable 'y' / 'tmp' tensors shaped like the input.
        ``low_latency``: the exact-mask pass uses one warp per segment (half the wait, 32 x the issue slots: for callers
        whose device is otherwise idle, such as the PCIe-bound feeder).
        ``defer``: return a pending handle right after the kernels have been launched; ``finish(handle)`` gives the
        result.  With exact_masks the sequential re-filtering of the 20-peak records (side stream, ~100 ms of
        latency whatever the batch size) then overlaps the NEXT batches' kernels instead of idling the device.
        ``lazy_hazards`` (with ``defer``): the 20-peak windows of THIS batch are looked for when the next batch's
        kernels have been queued (or when a batch is finished), on streams of their own, so that the host never
        waits on the calling stream with nothing queued behind it; ``batch.flat`` must stay untouched until
        ``finish``."""
        torch = self.torch
        flat = batch.flat
        beg_h, end_h, d_beg, d_end = self._ranges(batch, trim)
        # (fast_stats: one sweep; the records of the exact-mask pass get NumPy's bits in resolve_hazards)
        mean, sd, cnt = self.zscore_stats(flat, beg_h, end_h, d_beg, d_end, fast=self.fast_stats and not features)
        y, header = self.filter_records(flat, beg_h, end_h, d_beg, d_end, mean, sd, cnt,
                                        out=(scratch or {}).get("y"), tmp=(scratch or {}).get("tmp"))
        if strips:
            # BASELINE configs[2] (10 s 12-lead strips): sample_from_record's exclusive stop would cut no window
            # from a record of exactly L samples, so every (record, lead) strip IS one window
            if np.any(end_h - beg_h != self.L):
                raise _cabi.AewError(f"strips=True needs records of exactly {self.L} samples")
            win_start = np.ascontiguousarray(beg_h, dtype=np.int64)
            win_case = np.asarray(batch.case_ids)
            win_end = np.full(len(beg_h), self.L, dtype=np.int64)
            win_rec = np.arange(len(beg_h))
        else:
            win_start, win_case, win_end = self.window_table(beg_h, end_h, batch.case_ids)
            win_rec = np.repeat(np.arange(len(beg_h)), [window_count(e - b, self.L) for b, e in zip(beg_h, end_h)])
        info_h = np.stack([win_case.astype(np.float64), win_end.astype(np.float64)], axis=1) if len(win_start) \
            else np.zeros((0, 2))
        info = self._upload(info_h)
        tmp = getattr(self, "_filter_tmp", None)
        # the z-scored windows are parked (filter scratch) only for the delineation kernel; the median kernel alone
        # re-derives its samples from y and the per-window mean / sd.  np.std is evaluated where it decides the mask.
        scr = self.screen_windows(y, win_start, zwin_scratch=tmp if (features or self.park_windows) else None,
                                  lazy_sd=self.lazy_sd, fast_stats=self.fast_stats and not features)
        hz = None
        lazy = bool(lazy_hazards and defer and not features and self.exact_masks)
        if self.exact_masks and not lazy:
            hz = self.resolve_hazards(flat, beg_h, end_h, mean, sd, cnt, win_start, win_rec, scr,
                                      zwin_main=tmp if features else None, low_latency=low_latency)
        pend = dict(scr=scr, hz=hz, info=info, header=header, features=features, tmp=tmp, y=y, n_win=len(win_start),
                    mean=mean, sd=sd, keep_intermediates=keep_intermediates, beg_h=beg_h, end_h=end_h,
                    win_start=win_start, flat=flat, cnt=cnt)
        if lazy:
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            pend.update(k3_done=done, hz_args=dict(win_rec=win_rec, low_latency=low_latency))
            prev, self._lazy_prev = getattr(self, "_lazy_prev", None), pend
            if prev is not None:                               # its screen kernels finished a batch ago
                self._resolve_lazy(prev)
        if defer and not features:                             # (the delineation reads the shared filter scratch)
            return pend
        return self.finish(pend)

    def flush_hazards(self):
        """lazy_hazards: look for the newest batch's 20-peak windows now (end of a run: its sequential pass should
        start before the older batches are finished one by one)."""
        newest = getattr(self, "_lazy_prev", None)
        if newest is not None:
            self._resolve_lazy(newest)

    def _resolve_lazy(self, pend):
        a = pend.pop("hz_args", None)
        if a is not None:
            pend["hz"] = self.resolve_hazards(pend["flat"], pend["beg_h"], pend["end_h"], pend["mean"], pend["sd"],
                                              pend["cnt"], pend["win_start"], a["win_rec"], pend["scr"],
                                              low_latency=a["low_latency"], after=pend["k3_done"])
        if getattr(self, "_lazy_prev", None) is pend:
            self._lazy_prev = None

    @_cabi.on_device
    def finish(self, pend) -> dict:
        """Second half of ``run``: waits (on the stream, not the host) for the batch's exact-mask pass, compacts the
        kept windows and reads the counters back."""
        torch = self.torch
        if "hz_args" in pend:
            self._resolve_lazy(pend)
        if "k3_done" in pend and "_fin" not in pend:
            # a lazily resolved batch is compacted and read back on a stream of its own, which waits for THAT batch's
            # kernels only: the host's reads below do not wait for what has been queued on the calling stream since
            fin = getattr(self, "_finisher", None)
            if fin is None:
                fin = self._finisher = torch.cuda.Stream(device=self.device, priority=-1)
            main = torch.cuda.current_stream(self.device)
            pend["_fin"] = True
            with torch.cuda.stream(fin):
                fin.wait_event(pend["k3_done"])
                res = self.finish(pend)
            for v in res.values():
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(main)
            main.wait_stream(fin)
            return res
        scr, hz, info, header, features, tmp = (pend[k] for k in ("scr", "hz", "info", "header", "features", "tmp"))
        n_win, y = pend["n_win"], pend["y"]
        if hz is not None:
            torch.cuda.current_stream(self.device).wait_event(hz["event"])
        keep_mask = scr["keep"]
        feats = None
        if features and n_win and not (self.hp and self.lp):
            # data_extract_vitaldb.py:118,166-170: without both filters no delineation, NaN features, every screened
            # window kept -- and data_create_arrays.py:37-40 then writes zeros for such a case (a NaN cell where a
            # {'x', 'y'} dict is expected raises its TypeError branch)
            feats = torch.zeros((n_win, 35), dtype=torch.float64, device=self.device)
            fl = torch.ones(n_win, dtype=torch.uint8, device=self.device)
            fid, fstat = None, torch.zeros(1, dtype=torch.int32, device=self.device)
        elif features and n_win:
            # the screen kernel parked the z-scored windows ('ecg_sample') back to back in the filter scratch
            if tmp is None or tmp.numel() < n_win * self.L or tmp.data_ptr() == y.data_ptr():
                raise _cabi.AewError("features=True needs the filter scratch buffer for the z-scored windows")
            feats, fl, fid, fstat = self.window_features(tmp, n_win, scr)
            keep_mask = scr["keep"] & (fl != 2).to(torch.uint8)      # list-length rule, data_extract_vitaldb.py:140-144
            # a kept window whose delineation RAISED (flag 1) has np.nan cells where data_create_arrays.py:34-35
            # expects {'x', 'y'} dicts: its TypeError branch (:37-40) zeroes all 35 features of EVERY row of that case
            raised = (fl == 1) & scr["keep"].bool()
            if bool(raised.any()):
                d_case = info[:, 0]
                hit = torch.isin(d_case, torch.unique(d_case[raised]))
                feats = torch.where(hit[:, None], torch.zeros_like(feats), feats)
        keep_idx = torch.nonzero(keep_mask).squeeze(1)              # sync: number of kept windows
        # every counter and status word of the batch in ONE device->host read (each .item() is a round trip with
        # the device idle, which the end-to-end feeder pays once per chunk)
        i64 = torch.int64
        words = [header[0:1], header[1:2], header[4:5], scr["status"].to(i64), scr["hazards"].to(i64),
                 scr["n_peaks"].index_select(0, keep_idx).sum(dtype=i64).reshape(1),
                 (hz["changed"] if hz is not None else torch.zeros((), dtype=i64, device=self.device)).reshape(1).to(i64)]
        if feats is not None:
            words.append(fstat.to(i64))
        res = dict(np_waves=scr["waves"].index_select(0, keep_idx), medians=scr["median"].index_select(0, keep_idx),
                   np_info=info.index_select(0, keep_idx), keep=scr["keep"], has_nan=scr["has_nan"],
                   n_peaks=scr["n_peaks"], peaks=scr["peaks"], stats=scr["stats"], keep_idx=keep_idx,
                   n_windows=n_win, mean=pend["mean"], sd=pend["sd"])
        if feats is not None:
            res.update(np_features=feats.index_select(0, keep_idx), delineation_flags=fl, fiducials=fid)
        elif features:
            res.update(np_features=torch.zeros((0, 35), dtype=torch.float64, device=self.device))
        w = torch.cat(words).cpu().numpy()
        _cabi.check_device_status(int(w[2]), "filter")
        _cabi.check_device_status(int(w[3]), "window screen")
        if feats is not None:
            _cabi.check_device_status(int(w[7]), "fiducial features")
        res.update(n_segments=int(w[0]), n_tiles=int(w[1]), hazard_windows=int(w[4]), n_beats=int(w[5]),
                   mask_hazard_windows=0 if hz is None else hz["n"], mask_flips_resolved=int(w[6]))
        if pend["keep_intermediates"]:
            res.update(filtered=y, rec_beg=pend["beg_h"], rec_end=pend["end_h"], win_start=pend["win_start"])
        return res


COPY_PIECE = 1 << 40                # samples per host->device copy piece (whole chunks: the tables the kernels need
                                    # no longer use the copy engine, see Preprocessor._upload)
_pinned_pool: dict = {}             # reusable page-locked result buffers (cudaHostAlloc is far too slow per call)


def _pinned(key: str, shape, dtype, torch):
    """A page-locked host tensor of at least `shape[0]` rows, kept and grown across calls."""
    t = _pinned_pool.get(key)
    if t is None or t.shape[0] < shape[0] or t.dtype != dtype or tuple(t.shape[1:]) != tuple(shape[1:]):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        _pinned_pool[key] = t
    return t


def widen_to_f64(src, out=None, gain: float = 1.0, offset: float = 0.0, sentinel: int | None = None):
    """K8: a float32 or int16 CUDA tensor -> float64 (``x.astype(np.float64)``, resp. ``v * gain + offset`` with
    ``sentinel`` -> NaN): the device half of the input feeder (load_record.py:7-37 delivers float64; the source
    stores narrower)."""
    torch, lib = _cabi.require_cuda()
    code = {torch.float32: 0, torch.int16: 1}.get(src.dtype)
    if code is None:
        raise _cabi.AewError(f"cannot widen {src.dtype}")
    out = torch.empty(src.numel(), dtype=torch.float64, device=src.device) if out is None else out
    with torch.cuda.device(src.device):
        _cabi.check(lib.aew_widen_f64(_ptr(src), code, src.numel(), float(gain), float(offset),
                                      0 if sentinel is None else 1, 0 if sentinel is None else int(sentinel),
                                      _ptr(out), _cabi.stream_ptr(torch)), "widen")
    return out


def chunk_schedule(n_rec: int, chunk_cases: int, taper: bool = True):
    """[(first, last+1)] case ranges of the end-to-end feeder.  The call takes (copy of the first chunk) +
    max(all copies, all kernels) + (kernels of the last chunk): with ``taper`` the chunks grow from an eighth of
    ``chunk_cases`` at the start and shrink back to it at the end, so that both unoverlapped pieces are small while
    the bulk of the cases still moves in full-size chunks.  The result does not depend on the schedule."""
    chunk_cases = max(1, int(chunk_cases))
    ramp = [c for c in (chunk_cases // 8, chunk_cases // 4, chunk_cases // 2) if c >= 1]
    if not taper or not ramp or n_rec < 2 * sum(ramp) + 2 * chunk_cases:
        return [(c0, min(n_rec, c0 + chunk_cases)) for c0 in range(0, n_rec, chunk_cases)]
    rest = n_rec - 2 * sum(ramp)
    n_mid = -(-rest // chunk_cases)
    mid = [rest // n_mid + (1 if i < rest % n_mid else 0) for i in range(n_mid)]
    out, c0 = [], 0
    for c in ramp + mid + ramp[::-1]:
        out.append((c0, c0 + c))
        c0 += c
    return out


def run_host_pipelined(pre: Preprocessor, host_flat, rec_off, case_ids=None, chunk_cases: int = 100,
                       to_host: bool = True, pinned_results: bool = False, gain: float = 1.0, offset: float = 0.0,
                       sentinel: int | None = None, features: bool = False, taper: bool = True) -> dict:
    """End-to-end path for HOST data: ``host_flat`` is a (preferably pinned) CPU tensor holding all records back
    to back -- float64, or the narrower source type (float32; int16 with ``gain`` / ``offset`` and an optional
    ``sentinel`` for missing samples), which then crosses PCIe as it is and is widened on the device (K8).  Cases are fed in chunks through two device
    input buffers: the host->device copy of chunk k+1 (copy stream) overlaps the kernels of
    chunk k (current stream), and the device->host copy of chunk k's rows (third stream, into
    page-locked memory) overlaps both, so the wall time is max(PCIe, kernels) instead of their sum.
    With ``features`` the wave delineation and the 35 features per window run as well (``np_features`` (n,35) f64).
    Returns ``np_waves`` (n,320) f32 and ``np_info`` (n,2) f64 in case order, as host tensors
    (``to_host``) or left on the device (e.g. for a following all-gather).  With ``pinned_results`` the host
    tensors are views of a reusable page-locked pool (valid until the next call) instead of fresh copies."""
    torch = pre.torch
    if pre.device.index is not None and pre.device.index != torch.cuda.current_device():
        with torch.cuda.device(pre.device):            # streams, events and launches below belong to this device
            return run_host_pipelined(pre, host_flat, rec_off, case_ids, chunk_cases, to_host, pinned_results, gain,
                                      offset, sentinel, features, taper)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
    n_rec = len(rec_off) - 1
    case_ids = np.arange(n_rec, dtype=np.int64) if case_ids is None else np.asarray(case_ids, dtype=np.int64)
    chunks = chunk_schedule(n_rec, chunk_cases, taper)
    lens_rec = rec_off[1:] - rec_off[:-1]
    # exact_masks: the sequential pass of a chunk's 20-peak records takes ~150 ms however small the chunk; if such
    # a record sits in one of the LAST chunks the call ends with that wait.  So records that look like holding
    # 20-peak windows (a preview of their first windows, below) are fed FIRST, right after the first two chunks;
    # the rows are put back into case order on the device at the end.  A wrong guess costs time, never results.
    hazard_first = bool(pre.exact_masks and not features and len(chunks) > 3 and pre.hazard_peak_count()
                        and len(np.unique(case_ids)) == n_rec)
    order = np.arange(n_rec)
    sizes = [int(rec_off[c1] - rec_off[c0]) for c0, c1 in chunks]
    if hazard_first:                               # any chunk may end up holding the longest records
        srt = np.sort(lens_rec)[::-1]
        sizes_cap = [int(srt[:c1 - c0].sum()) for c0, c1 in chunks]
    else:
        sizes_cap = sizes
    cap = max(sizes_cap) if sizes_cap else 0
    narrow = host_flat.dtype != torch.float64
    cap_al = (cap + 7) & ~7                                    # keeps every buffer 16-byte aligned
    bufs = [torch.empty(cap_al, dtype=host_flat.dtype, device=pre.device) for _ in range(2)]
    wide = torch.empty(cap_al, dtype=torch.float64, device=pre.device) if narrow else None
    scratch = {"y": torch.empty(cap, dtype=torch.float64, device=pre.device),
               "tmp": torch.empty(cap, dtype=torch.float64, device=pre.device)}
    copy_stream = torch.cuda.Stream(device=pre.device)
    main = torch.cuda.current_stream(pre.device)
    copied = [torch.cuda.Event() for _ in chunks]
    freed = [torch.cuda.Event() for _ in chunks]
    if to_host:
        # every window can be kept at most: an upper bound on the rows, known from the record lengths alone
        max_rows = int(sum(window_count(int(rec_off[r + 1] - rec_off[r]), pre.L) for r in range(n_rec)))
        h_waves = _pinned("waves", (max(max_rows, 1), pre.beat_len), torch.float32, torch)
        h_info = _pinned("info", (max(max_rows, 1), 2), torch.float64, torch)
        h_feat = _pinned("feat", (max(max_rows, 1), 35), torch.float64, torch) if features else None
        back_stream = torch.cuda.Stream(device=pre.device)

    def start_copy(k):
        c0, c1 = chunks[k]
        recs = order[c0:c1]
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(freed[k - 2])          # the buffer's previous chunk has been consumed
            # the copy engine serves its queue in order whatever the stream (measured: pieces do not let a small copy
            # of another stream overtake), so nothing the kernels need may go through it: see Preprocessor._upload
            if len(recs) and np.array_equal(recs, np.arange(recs[0], recs[0] + len(recs))):
                lo, hi = int(rec_off[recs[0]]), int(rec_off[recs[-1] + 1])        # records in a row: one copy
                bufs[k % 2][:hi - lo].copy_(host_flat[lo:hi], non_blocking=True)
            else:
                at = 0
                for r in recs:                                                   # one copy per record
                    lo, hi = int(rec_off[r]), int(rec_off[r + 1])
                    bufs[k % 2][at:at + hi - lo].copy_(host_flat[lo:hi], non_blocking=True)
                    at += hi - lo
            copied[k].record(copy_stream)

    def preview_order():
        """Records 0 .. first-1 stay where they are (their chunks are already on their way); of the others, those
        whose first windows hold hazard_peak_count() +- 2 beats come first.  The preview (the first two windows of
        every remaining record) crosses PCIe behind the first two chunks and goes through K1 / K2 / K3 once."""
        first = chunks[1][1]
        rest = np.arange(first, n_rec)
        take = np.minimum(lens_rec[rest], 2 * pre.L + 1)       # two windows
        poff = np.concatenate([[0], np.cumsum(take)]).astype(np.int64)
        pbuf = torch.empty(int(poff[-1]), dtype=host_flat.dtype, device=pre.device)
        with torch.cuda.stream(copy_stream):
            stride = lens_rec[rest]
            if len(rest) and (stride == stride[0]).all() and (take == take[0]).all() and host_flat.is_pinned():
                # equally long records: the heads are equally spaced, ONE strided copy moves them all
                isz = host_flat.element_size()
                _cabi.check(pre.lib.aew_copy_heads_h2d(
                    ctypes.c_void_p(pbuf.data_ptr()), int(take[0]) * isz,
                    ctypes.c_void_p(host_flat.data_ptr() + int(rec_off[first]) * isz), int(stride[0]) * isz,
                    int(take[0]) * isz, len(rest), copy_stream.cuda_stream), "copy_heads")
            else:
                for j, r in enumerate(rest):
                    lo = int(rec_off[r])
                    pbuf[int(poff[j]):int(poff[j + 1])].copy_(host_flat[lo:lo + int(take[j])], non_blocking=True)
            got = torch.cuda.Event()
            got.record(copy_stream)
        return rest, poff, pbuf, got

    def rank_preview(rest, poff, pbuf, got):
        main.wait_event(got)
        pw = widen_to_f64(pbuf, None, gain, offset, sentinel) if narrow else pbuf
        sub = RecordBatch(pw, poff, rest)
        beg_h, end_h, d_beg, d_end = pre._ranges(sub, trim=False)
        mean, sd, cnt = pre.zscore_stats(pw, beg_h, end_h, d_beg, d_end, fast=pre.fast_stats)
        y, _ = pre.filter_records(pw, beg_h, end_h, d_beg, d_end, mean, sd, cnt, mode=0)
        wstart, wcase, _ = pre.window_table(beg_h, end_h, rest)
        if not len(wstart):
            return rest
        scr = pre.screen_windows(y, wstart, want_median=False, lazy_sd=True, fast_stats=pre.fast_stats)
        npk = scr["n_peaks"].cpu().numpy()
        # closest first: records whose first windows hold exactly hazard_peak_count() beats, then +-1, then +-2 (their
        # sequential pass takes ~200 ms whatever the chunk: the sooner they cross PCIe the better); the others after
        dist = np.full(n_rec, 3, dtype=np.int64)
        np.minimum.at(dist, wcase, np.minimum(np.abs(npk.astype(np.int64) - pre.hazard_peak_count()), 3))
        return rest[np.argsort(dist[rest], kind="stable")]

    waves, infos, feats = [], [], []
    tot = dict(hazards=0, n_windows=0, n_beats=0, rows=0, mask_hazard=0, mask_flips=0)
    # exact_masks: the sequential re-filtering of a chunk's 20-peak records takes ~100 ms on a side stream whatever
    # the chunk size; chunks are therefore finished `depth` chunks late, so that it overlaps the following chunks
    depth = pre.defer_chunks if (pre.exact_masks and not features) else 0
    pending = []

    def finish_one(item):
        res = pre.finish(item) if depth else item
        waves.append(res["np_waves"])
        infos.append(res["np_info"])
        if features:
            feats.append(res["np_features"])
        if to_host and not hazard_first:
            n = res["np_waves"].shape[0]
            rows = tot["rows"]
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(back_stream):
                back_stream.wait_event(done)
                h_waves[rows:rows + n].copy_(res["np_waves"], non_blocking=True)
                h_info[rows:rows + n].copy_(res["np_info"], non_blocking=True)
                if features:
                    h_feat[rows:rows + n].copy_(res["np_features"], non_blocking=True)
            tot["rows"] += n
        tot["hazards"] += res["hazard_windows"]
        tot["n_windows"] += res["n_windows"]
        tot["n_beats"] += res["n_beats"]
        tot["mask_hazard"] += res["mask_hazard_windows"]
        tot["mask_flips"] += res["mask_flips_resolved"]
        tot.setdefault("per_chunk", []).append(res["mask_hazard_windows"])

    if chunks:
        start_copy(0)
    preview = None
    for k, (c0, c1) in enumerate(chunks):
        if hazard_first and k == 0:
            start_copy(1)
            preview = preview_order()                          # queued behind chunks 0 and 1 on the copy engine
        elif hazard_first and k == 1:
            order = np.concatenate([order[:c1], rank_preview(*preview)])
            sizes = [int(lens_rec[order[a:b]].sum()) for a, b in chunks]
            start_copy(2)
        elif k + 1 < len(chunks):
            start_copy(k + 1)
        main.wait_event(copied[k])
        if narrow:
            widen_to_f64(bufs[k % 2][:sizes[k]], wide, gain, offset, sentinel)
            pre.launches += 1
            freed[k].record(main)                              # the narrow buffer is free as soon as it is widened
            chunk_in = wide[:sizes[k]]
        else:
            chunk_in = bufs[k % 2][:sizes[k]]
        recs_k = order[c0:c1]
        sub = RecordBatch(chunk_in, np.concatenate([[0], np.cumsum(lens_rec[recs_k])]).astype(np.int64),
                          case_ids[recs_k])
        sc = {"y": scratch["y"][:sizes[k]], "tmp": scratch["tmp"][:sizes[k]]}
        pending.append(pre.run(sub, scratch=sc, features=features, defer=depth > 0, low_latency=True))
        if not narrow:
            freed[k].record(main)                              # (the hazard records have been copied out by now)
        if hazard_first and depth and pending[-1]["hz"] is None:
            finish_one(pending.pop())                          # nothing to wait for (rows are re-ordered at the end)
        while len(pending) > depth:
            finish_one(pending.pop(0))
    while pending:
        finish_one(pending.pop(0))
    hazards, n_windows, n_beats, rows = tot["hazards"], tot["n_windows"], tot["n_beats"], tot["rows"]
    sorted_dev = None
    if hazard_first:
        # back into case order (rows of one case are already in window order): position of every row's case in
        # the caller's list, stable sort, one gather per array -- all on the device
        w_all = torch.cat(waves) if waves else torch.zeros((0, pre.beat_len), dtype=torch.float32, device=pre.device)
        i_all = torch.cat(infos) if infos else torch.zeros((0, 2), dtype=torch.float64, device=pre.device)
        ids_sorted, inv = np.unique(case_ids, return_inverse=False), np.argsort(case_ids, kind="stable")
        d_ids = pre._upload(ids_sorted.astype(np.int64))
        d_pos = pre._upload(inv.astype(np.int64))
        pos = d_pos.index_select(0, torch.searchsorted(d_ids, i_all[:, 0].to(torch.int64).contiguous()))
        perm = torch.argsort(pos, stable=True)
        sorted_dev = (w_all.index_select(0, perm), i_all.index_select(0, perm))
        rows = int(sorted_dev[0].shape[0])
        if to_host:
            h_waves[:rows].copy_(sorted_dev[0], non_blocking=True)
            h_info[:rows].copy_(sorted_dev[1], non_blocking=True)
            main.synchronize()
    if to_host:
        back_stream.synchronize()
        np_waves, np_info = h_waves[:rows], h_info[:rows]
        np_feat = h_feat[:rows] if features else None
        if not pinned_results:
            np_waves, np_info = np_waves.clone(), np_info.clone()
            np_feat = np_feat.clone() if features else None
    elif sorted_dev is not None:
        np_waves, np_info, np_feat = sorted_dev[0], sorted_dev[1], None
    else:
        np_feat = (torch.cat(feats) if feats else torch.zeros((0, 35), dtype=torch.float64, device=pre.device)) \
            if features else None
        np_waves = torch.cat(waves) if waves else torch.zeros((0, pre.beat_len), dtype=torch.float32,
                                                              device=pre.device)
        np_info = torch.cat(infos) if infos else torch.zeros((0, 2), dtype=torch.float64, device=pre.device)
    out = dict(np_waves=np_waves, np_info=np_info, hazard_windows=hazards, n_windows=n_windows, n_beats=n_beats,
               mask_hazard_windows=tot["mask_hazard"], mask_flips_resolved=tot["mask_flips"],
               mask_hazard_windows_per_chunk=tot.get("per_chunk", []))
    if features:
        out["np_features"] = np_feat
    return out


def process_cases(records, case_ids=None, **kwargs) -> dict:
    """Host convenience: list of 1-D float arrays -> NumPy ``np_waves`` (n,320) f32,
    ``np_features`` (n,35) f64, ``np_info`` (n,2) f64, in case-major order like
    data_create_arrays.py:27-48."""
    pre = Preprocessor(**kwargs)
    batch = RecordBatch.from_host(records, case_ids, device=pre.device)
    res = pre.run(batch, features=True)
    waves = res["np_waves"].cpu().numpy()
    info = res["np_info"].cpu().numpy()
    return dict(np_waves=waves, np_features=res["np_features"].cpu().numpy(), np_info=info,
                n_windows=res["n_windows"], hazard_windows=res["hazard_windows"])
