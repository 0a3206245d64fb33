/* aewave.h -- C ABI of libaewave.so, the B200 (sm_100a) implementation of the
 * ae_waveforms preprocessing hot path.
 *
 * The reference (Computational-Biology-TUe/ae_waveforms) is pure Python and has no
 * FFI; its boundary for this path is the Python import surface
 * (data_load_preprocess.*, functions_traditional.signal_reconstruction).  Each
 * entry point below names the reference function(s) it replaces; the Python
 * mirror in ae_waveforms_b200/ (and dropin/) binds them with ctypes, see
 * INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only.  Pointers marked `d_` / "device"
 * are CUDA device pointers owned by the caller (PyTorch tensors on the Python
 * side); `h_` are host pointers.  Calls are stateless, stream-ordered on
 * `stream` (a cudaStream_t passed as void*), never allocate, free or
 * synchronise, and return 0 on success, a negative AEW_ERR_* code for argument
 * errors or a positive cudaError_t.  Device-side conditions (workspace overflow)
 * are reported through a status word in the workspace / params, read back by the
 * caller after the stream drains.  Ragged batches are one flat float64 array of
 * all records back to back plus int64 offsets rec_off[n_rec + 1].
 */
#ifndef AEWAVE_H
#define AEWAVE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AEW_VERSION 113

/* return codes */
#define AEW_OK 0
#define AEW_ERR_ARG (-1)
#define AEW_ERR_WORKSPACE (-2)
#define AEW_ERR_UNSUPPORTED (-3)
#define AEW_ERR_NO_DEVICE (-4)
/* bits of the device status word */
#define AEW_DEV_SEG_OVERFLOW 1
#define AEW_DEV_TILE_OVERFLOW 2
#define AEW_DEV_LOOKBACK_TIMEOUT 4 /* filter: a tile's bulk copy never completed (the name dates from the look-back over published tile states) */
#define AEW_DEV_RUN_OVERFLOW 8
#define AEW_DEV_FID_RANGE 16 /* a fiducial index outside [0, L) was treated as missing */

int aew_version(void);
/* 0 if a CUDA device of compute capability 10.x is current, AEW_ERR_NO_DEVICE otherwise */
int aew_device_check(void);

/* ---- K1: whole-record z-score statistics --------------------------------------
 * replaces zscore() of data_load_preprocess/filters.py:5-12 as called on the whole
 * track at data_extract_vitaldb.py:64: mean = np.nanmean, sd = np.nanstd, both
 * bit-identical to NumPy (pairwise summation order reproduced). */
long long aew_zscore_plan(const long long* h_rec_beg, const long long* h_rec_end, long long n_rec,
                          long long* h_blk_start /* [n_rec+1] out */);
size_t aew_zscore_workspace_bytes(long long n_rec, long long n_blocks);
/* records are [rec_beg[r], rec_end[r]) in the flat array; a contiguous batch passes (off, off + 1) */
int aew_zscore_stats_f64(const double* d_x, const long long* d_rec_beg, const long long* d_rec_end,
                         const long long* d_blk_start, long long n_rec, long long n_blocks, double* d_mean,
                         double* d_sd, long long* d_count /* non-NaN per record, may be NULL */, void* d_ws,
                         size_t ws_bytes, void* stream);
/* The same statistics from ONE sweep over the samples (8 B/sample instead of 16): per-block (n, mean, M2) around a
 * local shift, merged per record in a fixed order -- deterministic and within a few ulp of np.nanmean / np.nanstd
 * (filters.py:5-12), NOT bit-identical.  For the tiled filter, whose output is ~1e-9 from SciPy's anyway; the
 * bit-exact chain (exact filter mode, the exact-mask pass, the zscore() drop-in) uses aew_zscore_stats_f64. */
int aew_zscore_stats_fast_f64(const double* d_x, const long long* d_rec_beg, const long long* d_rec_end,
                              const long long* d_blk_start, long long n_rec, long long n_blocks, double* d_mean,
                              double* d_sd, long long* d_count /* may be NULL */, void* d_ws, size_t ws_bytes,
                              void* stream);
/* out[i] = (x[i] - mean[r]) / sd[r] over flat positions [lo, hi); NaN propagates, NaN between records */
int aew_zscore_apply_f64(const double* d_x, const long long* d_rec_beg, const long long* d_rec_end, long long n_rec,
                         const double* d_mean, const double* d_sd, double* d_out, long long lo, long long hi,
                         void* stream);
/* leading/trailing all-NaN trim of data_extract_vitaldb.py:53-57:
 * d_trimmed[2r] = first non-NaN position, d_trimmed[2r+1] = one past the last (equal if all NaN) */
int aew_trim_nan_edges_f64(const double* d_x, const long long* d_rec_beg, const long long* d_rec_end, long long n_rec,
                           long long* d_trimmed, void* stream);

/* ---- K2: zero-phase IIR filtering per non-NaN segment ----------------------------
 * replaces butterworth() of data_load_preprocess/filters.py:15-55
 * (= scipy.signal.filtfilt(b, a, seg) per maximal non-NaN segment longer than
 * 3*max(len(a),len(b)); shorter segments copied; NaN stays NaN).  Filter design
 * (scipy.signal.butter, lfilter_zi) stays on the host so b, a, zi carry SciPy's bits. */
long long aew_filt_tables_len(void);
int aew_filt_tables(const double* h_b, const double* h_a, const double* h_zi, int order, double* h_tab /* out */);
long long aew_filt_max_tiles(long long n_rec, long long total_samples, long long max_seg);
size_t aew_filt_workspace_bytes(long long n_rec, long long total_samples, long long max_seg);
int aew_segments_find_f64(const double* d_x, const long long* d_rec_beg, const long long* d_rec_end,
                          long long n_rec, long long total_samples,
                          const long long* d_nonnan_count /* may be NULL */, double* d_nan_fill /* may be NULL */,
                          void* d_ws, size_t ws_bytes, long long max_seg, void* stream);
/* mode 0: tiled block scan, forward launch (src -> tmp) + backward launch (tmp -> out; out may alias src); within 1e-6
 *         relative of SciPy, measured ~1e-9;
 * mode 1: SciPy's sequential DF2T order, bit-identical to scipy.signal.filtfilt; one lane per segment (32 segments
 *         per block in lockstep: least of the machine);
 * mode 3: the same bits, one warp per segment (half the latency per segment, 32 x the issue slots);
 * mode 2: tiled, forward AND backward in one launch over tiles with a 416-sample halo (src -> out, src != out, tmp
 *         unused): only for filters whose memory is that short, aew_filt_fused_ok(h_tab) (the 30 Hz low-pass is,
 *         the 1 Hz high-pass is not). */
int aew_filt_fused_ok(const double* h_tab);
int aew_filtfilt_f64(const double* d_src, const double* d_mean /* may be NULL */, const double* d_sd,
                     double* d_out, double* d_tmp, const double* d_tab, const double* h_tab /* host copy of d_tab */,
                     long long n_rec, long long total_samples, void* d_ws, size_t ws_bytes, long long max_seg, int mode,
                     void* stream);
/* debugging aid: with AEW_K2_TIMING=1 in the environment, k2_pass accumulates the cycles thread 0 of every
 * block spends per phase; this copies and clears the 8 counters (synchronises the device) */
int aew_filt_timing_dump(unsigned long long* h_out8);
/* device address of the 8 x int64 workspace header: [0] segments found, [1] tiles, [4] status bits */
int aew_filt_header(void* d_ws, long long** d_header_out);

/* ---- K3 + K4: window screen and median beat ----------------------------------------
 * replaces, per 20 s window: the per-window zscore (data_extract_vitaldb.py:93),
 * remove_samples() (data_load_preprocess/remove_samples.py:8-82, including
 * nk.ecg_findpeaks, estimate_hr :85-107 and rmssd :110-121) and
 * compute_median_sample_ecg() (data_load_preprocess/median_sample.py:6-54). */
typedef struct K3Params {
  const double* y;            /* device: filtered samples, flat */
  const long long* win_start; /* device: [n_win] flat offset of each window (sample_from_record.py:19-30) */
  long long n_win;
  int L;                      /* samples per window (fs * sample_length) */
  int smooth_k, avg_k;        /* rint(0.1 fs), rint(0.75 fs) */
  int refractory;             /* rint(0.3 fs) */
  int pre, post;              /* 106, 214 (median_sample.py:20-25) */
  double hr_num;              /* 60 * fs */
  double rr_div;              /* fs * 1000 (remove_samples.py:119) */
  double hr_lo, hr_hi, pvc_threshold, sd_threshold;   /* 30, 180, 150, 0.05 */
  int mode;                   /* 0 fast + exact fallback on hazard windows, 1 exact everywhere, 2 fast only,
                                 3 fast then forced fallback (test hook for the hazard transition) */
  int zscore_windows;         /* 1: z-score each window first (data_extract_vitaldb.py:93); 0: windows are
                                 already z-scored by the caller (the remove_samples() drop-in) */
  unsigned char* keep;        /* device outputs, one entry per window */
  unsigned char* has_nan;
  int* n_peaks;
  int* peaks;                 /* [n_win][aew_windows_max_peaks()], -1 padded */
  double* stats;              /* [n_win][4]: ecg_hr, ecg_hrv_rmssd, ecg_sd, ecg_sd/hr */
  double* median;             /* [n_win][pre + post] float64 median beat of kept windows (may be NULL); the section
                                 is int(0.64 fs) samples, 320 at 500 Hz (median_sample.py:20-25) */
  float* waves;               /* [n_win][pre + post] float32 copy = np_waves rows (may be NULL) */
  double* zwin;               /* [n_win][L] z-scored windows = 'ecg_sample' column.  REQUIRED with zscore_windows = 1
                                 (unless fast_stats, where it must be NULL):
                                 the beat and median kernels read the samples they need from it (may be NULL with
                                 zscore_windows = 0, where y itself holds them) */
  double* zparams;            /* [n_win][2] optional output: mean, sd of each window's z-score (may be NULL) */
  int* status;                /* device status word */
  unsigned long long* hazard_count; /* device counter of windows that took the exact fallback (may be NULL) */
  int lazy_sd;                /* 1: np.std of the (z-scored) window is evaluated only where it decides the mask
                                 (n_peaks * sd_threshold == 1); stats[2], stats[3] are NaN elsewhere.
                                 0: always (the 'ecg_sd' / 'ecg_sd/hr' columns of remove_samples()) */
  void* workspace;            /* device scratch of aew_windows_workspace_bytes(n_win) bytes: what the window kernel
                                 (z-score, QRS indicator) hands to the beat kernel (runs, R peaks, screen) */
  unsigned long long workspace_bytes;
  int fast_stats;             /* 1 (with zscore_windows = 1, lazy_sd = 1, mode != 1, zwin = NULL): the per-window z-score
                                 (data_extract_vitaldb.py:93) without NumPy's summation order -- mean and sd from one
                                 pass of shifted sums (~1e-15 from np.mean / np.std), R peaks taken from the unscaled
                                 samples, the median beat z-scored after the column medians; nothing is parked.  For
                                 filter output that is itself ~1e-9 from SciPy's (the tiled filter); the bit-exact
                                 chain is fast_stats = 0 */
} K3Params;
int aew_windows_max_peaks(void);
int aew_windows_max_len(void);
size_t aew_windows_workspace_bytes(long long n_win);
int aew_windows_beats_f64(const K3Params* h_params, void* stream);
/* debugging aid: with AEW_K3_TIMING=1 in the environment the window kernel accumulates the cycles thread 0 of
 * every block spends per phase; this copies and clears the 16 counters (synchronises the device) */
int aew_windows_timing_dump(unsigned long long* h_out16);
/* compute_median_sample_ecg for given peaks: windows [n_win][L], peaks CSR; aligned may be NULL */
int aew_median_beats_f64(const double* d_windows, long long n_win, int L, const int* d_peaks,
                         const long long* d_peak_off, int max_peaks_per_window, int pre, int post,
                         double* d_median, double* d_aligned, void* stream);

/* ---- K9: per-beat error metrics of a reconstruction ------------------
 * replaces sample_rmse_loss / sample_prd_loss (functions_ae/loss_functions.py:8-20) and the quality score
 * as used at analysis_performance_traditional_vs_ae.py:178-197: true = float32(median beat),
 * pred = float32(y_gauss + rowmax(median beat)); rmse = sqrt(mean((pred-true)^2)),
 * prd = 100 sqrt(sum((true-pred)^2) / sum(true^2)), qs = compression_ratio / prd (320 / 20 in the reference).
 *   d_recon, d_median [n_beats][n_t] float64; outputs [n_beats] float32 */
int aew_beat_metrics_f32(const double* d_recon, const double* d_median, long long n_beats, int n_t,
                         double compression_ratio, float* d_rmse, float* d_prd, float* d_qs, void* stream);

/* ---- K8: narrow source samples -> float64 (the input side) ----------
 * the float64 frame that data_load_preprocess/load_record.py:7-37 hands to process_case_id
 * (data_extract_vitaldb.py:45-51) is stored narrower at the source; the end-to-end feeder moves the narrow
 * type over PCIe and widens on the device.  AEW_SRC_F32: out = (double)x, exactly x.astype(np.float64);
 * AEW_SRC_I16: out = v * gain + offset (one multiplication, one addition), v == sentinel -> NaN when
 * use_sentinel.  d_src and d_out must be 16-byte aligned. */
#define AEW_SRC_F32 0
#define AEW_SRC_I16 1
int aew_widen_f64(const void* d_src, int src_type, long long n, double gain, double offset, int use_sentinel,
                  int sentinel, double* d_out, void* stream);

/* The path's small host tables (record bounds as in data_extract_vitaldb.py:53-57, window starts of
 * sample_from_record.py:19-30, the np_info rows of data_create_arrays.py:37-38) read by the SMs directly from
 * PAGE-LOCKED host memory (unified addressing: the host pointer is the device pointer) instead of a copy-engine
 * transfer, which would wait behind the next chunk's bulk host->device copy.  Both pointers 16-byte aligned;
 * the host buffer must stay untouched until the stream has passed this call. */
int aew_fetch_host(const void* h_pinned, long long n_bytes, void* d_out, void* stream);
/* The first `width` bytes of `rows` host records `src_pitch` bytes apart -> device rows `dst_pitch` apart, one
 * strided copy (the feeder's preview of every record's first windows; page-locked source). */
int aew_copy_heads_h2d(void* d_dst, size_t dst_pitch, const void* h_src, size_t src_pitch, size_t width, size_t rows,
                       void* stream);

/* ---- K7: wavelet delineation of the beats of a window --------------------------------
 * replaces nk.ecg_delineate(ecg_sample, r_peaks, sampling_rate=fs)[1] (neurokit2, default method
 * "dwt") as called at data_extract_vitaldb.py:131, including the wrapper's sanitising and the
 * caller's validity rules (data_extract_vitaldb.py:130-144).  neurokit2 is not vendored in the
 * reference and version-unpinned: the kernel follows the restatement in oracle/nk_delineate.py.
 *   d_sig, d_sig_off  window w's samples are d_sig[d_sig_off[w] .. + L) (the z-scored 'ecg_sample')
 *   d_peaks           [n_win][max_peaks] R-peak indices, d_n_peaks [n_win]
 *   d_select          optional [n_win]: only windows with a non-zero entry are processed
 *   d_fid             [n_win][10][max_peaks] int32 out, order P_Peaks, P_Onsets, P_Offsets, Q_Peaks,
 *                     R_Onsets, R_Offsets, S_Peaks, T_Peaks, T_Onsets, T_Offsets; AEW_FID_NAN = np.nan
 *   limits            64 <= L <= 10,240 samples, fs <= 1,000 Hz (the analysis runs at 2,000 Hz), max_peaks <= 72
 *   d_flags           [n_win] out: 0 ok; 1 the library would raise (NaN features, window kept);
 *                     2 a list would lose an entry (lengths differ -> the caller drops the window) */
int aew_delineate_blocks(long long n_win);
size_t aew_delineate_workspace_bytes(long long n_win, int L, double fs);
int aew_delineate_f64(const double* d_sig, const long long* d_sig_off, long long n_win, int L, double fs,
                      const int* d_peaks, const int* d_n_peaks, int max_peaks, const unsigned char* d_select,
                      int* d_fid, unsigned char* d_flags, void* d_ws, size_t ws_bytes, void* stream);

/* debugging aid: with AEW_K7_TIMING=1 in the environment k7_delineate accumulates the cycles thread 0 of every block
 * spends per phase; this copies and clears the 8 counters (synchronises the device) */
int aew_delineate_timing_dump(unsigned long long* h_out8);

/* ---- K6: the 35 np_features of a window from per-beat fiducial indices ----------------
 * replaces ecg_features_relative_r + compute_ecg_morphology_features
 * (data_load_preprocess/feature_extraction/features_ecg.py:56-136) and
 * compute_median_features_ecg (data_load_preprocess/median_sample.py:57-82) as called at
 * data_extract_vitaldb.py:146-162, flattened like data_create_arrays.py:34-36:
 * [x, y of P_Peaks, P_Onsets, P_Offsets, Q_Peaks, R_Onsets, R_Offsets, S_Peaks, T_Peaks, T_Onsets,
 *  T_Offsets | 10 durations | 5 prominences], bit-identical to the reference's arithmetic.
 *   d_sig, d_sig_off  window w's samples are d_sig[d_sig_off[w] .. + L) (the z-scored 'ecg_sample')
 *   d_peaks           [n_win][max_peaks] R-peak indices, d_n_peaks [n_win]
 *   d_fid             [n_win][10][max_peaks] int32 fiducial indices in the order above, AEW_FID_NAN = np.nan
 *   d_valid           optional [n_win]: 0 = delineation failed -> NaN row (data_extract_vitaldb.py:132-135)
 *   d_features        [n_win][35] out */
#define AEW_FID_NAN (-2147483647 - 1)
int aew_fiducial_features_f64(const double* d_sig, const long long* d_sig_off, long long n_win, int L,
                              const int* d_peaks, const int* d_n_peaks, int max_peaks, const int* d_fid,
                              const unsigned char* d_valid, double fs, double* d_features, int* d_status,
                              void* stream);

/* The three stages of K6 as calls of their own, for the per-call drop-ins (any number of beats):
 *   aew_fiducial_relative_f64  ecg_features_relative_r (features_ecg.py:56-86): d_fid [n_fid][n_beats] int32 indices
 *                              (AEW_FID_NAN = np.nan) -> d_xy [n_fid][2][n_beats] = x (ms) then y per fiducial
 *   aew_morphology_f64         compute_ecg_morphology_features (features_ecg.py:89-136): d_x, d_y [11][n_beats]
 *                              (the ten fiducials in np_features order, then the inserted R peak) -> d_out [15][n_beats]
 *   aew_nanmedian_rows_f64     compute_median_features_ecg (median_sample.py:57-82): np.nanmedian of every row of
 *                              d_vals [rows][n]; d_scratch [rows][n]
 *   aew_peak_stats_f64         estimate_hr / rmssd (remove_samples.py:85-121): d_out[0] = hr_num / median(intervals)
 *                              (mode 0 'min', 1 'mean', 2 'max' picks the lists), d_out[1] = RMSSD of d_max_peaks with
 *                              the reference's divisor rr_div = fs * 1000; d_scratch [2 * (n_min + n_max) + 4] */
int aew_fiducial_relative_f64(const double* d_sig, int L, const int* d_peaks, const int* d_fid, int n_fid,
                              int n_beats, double fs, double* d_xy, int* d_status, void* stream);
int aew_morphology_f64(const double* d_x, const double* d_y, int n_beats, double* d_out, void* stream);
int aew_nanmedian_rows_f64(const double* d_vals, int rows, int n, double* d_scratch, double* d_out, void* stream);
int aew_peak_stats_f64(const long long* d_min_peaks, int n_min, const long long* d_max_peaks, int n_max, int mode,
                       double hr_num, double rr_div, double* d_scratch, double* d_out, void* stream);

/* ---- K5: Gaussian fiducial reconstruction + sigma-map search objective --------------
 * replaces gaussian_line / extract_coordinates / reconstruct_gaussian
 * (functions_traditional/signal_reconstruction.py:5-81) and objective_function
 * (analysis_performance_traditional_vs_ae.py:63-90) for n_cand sigma maps at once.
 *   d_fx, d_fy   [n_beats][11] float64 fiducials (NaN = missing), order of analysis_...:40-44
 *   d_target     [n_beats][n_t] float32 targets (may be NULL if only d_pred is wanted)
 *   d_xt         [n_t] beat time axis in ms (analysis_...:48), n_t <= 320
 *   d_cand       [n_cand][12] float64 sigma maps
 *   d_sse        [n_cand] out: sum over beats and samples of (float32(pred) - target)^2
 *   d_pred       optional [n_beats][n_t] float64 reconstruction for candidate 0 */
size_t aew_recon_workspace_bytes(long long n_beats, long long n_cand);
int aew_recon_sse_f64(const double* d_fx, const double* d_fy, const double* d_baseline, const float* d_target,
                      long long n_beats, const double* d_xt, int n_t, const double* d_cand, long long n_cand,
                      double* d_sse, double* d_pred, void* d_ws, size_t ws_bytes, void* stream);
/* The same objective for a DISCRETE sigma grid (the reference searches Integer(1, 5)^12,
 * analysis_performance_traditional_vs_ae.py:58): every output sample depends on at most two entries of
 * the sigma map, so per-(slot, value) and per-(slot pair, value pair) error tables are accumulated once
 * over all beats and each candidate is then scored by 157 table look-ups.  Same per-sample float32
 * squared errors as aew_recon_sse_f64; only the (float64) summation order differs.
 *   d_vals [n_vals] the grid values, n_vals <= aew_recon_max_grid_values()
 *   d_idx  [n_cand][12] uint8: index into d_vals of every candidate entry */
int aew_recon_max_grid_values(void);
size_t aew_recon_tables_workspace_bytes(long long n_beats);
int aew_recon_sse_grid_f64(const double* d_fx, const double* d_fy, const double* d_baseline, const float* d_target,
                           long long n_beats, const double* d_xt, int n_t, const double* d_vals, int n_vals,
                           const unsigned char* d_idx, long long n_cand, double* d_sse, void* d_ws, size_t ws_bytes,
                           void* stream);
/* gaussian_line(x1, y1, x2, y2, s) (signal_reconstruction.py:5-14): n = int(x2) - int(x1) values */
int aew_gaussian_line_f64(double x1, double y1, double x2, double y2, double s, double* d_out, int n, void* stream);
/* extract_coordinates (signal_reconstruction.py:45-81) for one beat: d_fx, d_fy [11];
 * d_coords [26] out = x[13] then y[13]; d_meta [14] out = count, nan_mask[13] */
int aew_extract_coordinates_f64(const double* d_fx, const double* d_fy, double baseline, const double* d_xt, int n_t,
                                double* d_coords, int* d_meta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AEWAVE_H */
