/*
 * ssr_b200.h -- C ABI of the B200-native ssr_eval DSP hot path (libssr_b200.so).
 *
 * The reference (haoheliu/ssr_eval) is pure Python and has no FFI: its boundary for this path is
 * a Python class API.  These entry points are what a binding for that path
 * binds; each cites the reference interface it replaces (paths relative to the reference repo).
 * Plain pointers and sizes only: no torch / C++ types.  All `*_dev` pointers are CUDA device
 * pointers owned by the caller; `stream` is a `cudaStream_t` passed as `void*` (NULL = default
 * stream).  Every call is asynchronous on `stream` unless stated otherwise, allocates no device
 * memory (workspace is caller-provided; size queries below) and returns an `int` status
 * (0 = SSR_OK).  Nothing throws across the ABI; `ssr_last_error()` gives the message of the last
 * failure on the calling thread.
 *
 * Ragged batches: utterance i of a batch lives at `x_dev[offsets[i] .. offsets[i+1])` (float32).
 * Offsets are given twice -- `offsets_host` (used on the host to size grids / workspaces) and
 * `offsets_dev` (the same n+1 int64 values in device memory, read by the kernels).
 */
#ifndef SSR_B200_H_
#define SSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSR_OK 0
#define SSR_ERR_INVALID 1   /* bad argument */
#define SSR_ERR_CUDA 2      /* a CUDA runtime call / launch failed */
#define SSR_ERR_WORKSPACE 3 /* workspace too small */

/* metric selection flags; out layout is out[pair*4 + {0:lsd, 1:log_sispec, 2:sispec, 3:ssim}] */
#define SSR_METRIC_LSD 1u
#define SSR_METRIC_LOG_SISPEC 2u
#define SSR_METRIC_SISPEC 4u
#define SSR_METRIC_SSIM 8u
#define SSR_METRIC_ALL 15u

int ssr_version(void);
const char* ssr_last_error(void);
/* number of kernels this library has launched in the calling process (for bench accounting) */
uint64_t ssr_launch_count(void);

/* Optional device-side timing of the dominant kernel (k_stft_metrics): when enabled, every launch is
 * bracketed by CUDA events ON THE STREAM IT IS LAUNCHED ON.  ssr_timing_collect synchronises those
 * events, returns the summed kernel milliseconds and the number of launches since the last
 * collect, and resets.  Used by bench.py for the roofline line; off by default. */
int ssr_timing_enable(int on);
int ssr_timing_collect(double* total_ms, int* n_launches);

/* ------------------------------------------------------------------------------------------
 * K1 + K2: batched STFT -> magnitude -> {LSD, log-sispec, sispec, SSIM}
 * Replaces AudioMetrics.__init__/wav_to_spectrogram/lsd/sispec/ssim and the metric half of
 * AudioMetrics.evaluation (ssr_eval/metrics.py:16-19, 26-30, 92-132; ssr_eval/utils.py:43-92).
 * STFT semantics = librosa.stft(y, n_fft, hop) 0.9.x: centre, reflect pad n_fft//2, periodic Hann
 * (float64), float64 transform, complex64 store, float32 magnitude.  Any n_fft in [65, 8192]
 * (power of two: direct FFT; otherwise Bluestein), any hop >= 1.
 * ------------------------------------------------------------------------------------------ */
typedef struct ssr_stft_plan ssr_stft_plan;

/* window_host: n_fft float64 values (the analysis window), or NULL for the periodic Hann.
 * Synchronous (uploads twiddle / chirp tables to the current device). */
int ssr_stft_plan_create(ssr_stft_plan** plan, int n_fft, int hop, const double* window_host);
int ssr_stft_plan_destroy(ssr_stft_plan* plan);
/* frames of a centred STFT of `length` samples: 1 + (length + 2*(n_fft/2) - n_fft) / hop */
int64_t ssr_stft_num_frames(const ssr_stft_plan* plan, int64_t length);

/* workspace bytes needed by ssr_stft_metrics_batched for this batch and flag set; the workspace pointer must be
 * 16-byte aligned (any cudaMalloc / torch allocation is) */
size_t ssr_stft_metrics_workspace_bytes(const ssr_stft_plan* plan, const int64_t* offsets_host,
                                        int n_pairs, unsigned flags);

/* est_dev / tgt_dev: the two ragged float32 batches (same offsets: the reference truncates both
 * waveforms of a pair to the shorter, metrics.py:89-90 -- the caller does that when packing).
 * Every utterance needs length > n_fft/2 (reflect padding).  out_dev: n_pairs*4 float64;
 * metrics not requested in `flags` are written as NaN. */
int ssr_stft_metrics_batched(const ssr_stft_plan* plan, const float* est_dev, const float* tgt_dev,
                             const int64_t* offsets_host, const int64_t* offsets_dev, int n_pairs,
                             unsigned flags, double* out_dev, void* workspace_dev,
                             size_t workspace_bytes, void* stream);

/* Same, for a float64 ESTIMATE batch (target stays float32).  The reference's zero-phase IIR low-pass filters
 * (ssr_eval/lowpass.py:94-131, scipy sosfiltfilt) return float64, an identity-like testee hands that to
 * AudioMetrics.evaluation (ssr_eval/eval.py:138-151), and librosa / torch then keep the estimate's spectrum,
 * magnitude and every formula that touches it in float64 (metrics.py:26-30, 109-121): only the target is
 * rounded to complex64 / float32.  Same workspace as ssr_stft_metrics_batched. */
int ssr_stft_metrics_batched_f64est(const ssr_stft_plan* plan, const double* est_dev, const float* tgt_dev,
                                    const int64_t* offsets_host, const int64_t* offsets_dev, int n_pairs,
                                    unsigned flags, double* out_dev, void* workspace_dev,
                                    size_t workspace_bytes, void* stream);

/* Same, for a float64 estimate AND a float64 target (e.g. soundfile.read's default dtype handed straight to
 * AudioMetrics.evaluation, metrics.py:51-107): librosa keeps both spectra in complex128 (metrics.py:26-30) and every torch
 * formula runs in float64.  A float64 target with a float32 estimate is scored through this entry with the estimate
 * widened (exact); the reference would round |E| to float32 first, a ~6e-8 relative difference per bin. */
int ssr_stft_metrics_batched_f64(const ssr_stft_plan* plan, const double* est_dev, const double* tgt_dev,
                                 const int64_t* offsets_host, const int64_t* offsets_dev, int n_pairs,
                                 unsigned flags, double* out_dev, void* workspace_dev, size_t workspace_bytes,
                                 void* stream);

/* Magnitude spectrogram only (AudioMetrics.wav_to_spectrogram, metrics.py:26-30) of one ragged
 * batch: spec_dev receives, utterance after utterance, T_i x F float32 row-major (F = n_fft/2+1).
 * Needs the same workspace as ssr_stft_metrics_batched with flags = 0. */
int ssr_stft_magnitude_batched(const ssr_stft_plan* plan, const float* x_dev,
                               const int64_t* offsets_host, const int64_t* offsets_dev, int n,
                               float* spec_dev, void* workspace_dev, size_t workspace_bytes,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * K3: polyphase FIR resampler = scipy.signal.resample_poly(x, up, down) (upfirdn, zero
 * extension) for a given FIR.  Replaces the resample_poly calls of subsampling()
 * (ssr_eval/lowpass.py:134-144) and of librosa.resample(res_type="polyphase")
 * (ssr_eval/eval.py:144-150).  `taps_host` is scipy's `h` AFTER the `h *= up` scaling and BEFORE
 * its zero padding: n_taps = 2*half_len+1 float32 values (firwin(..., ("kaiser", 5.0))).
 * up/down must already be gcd-reduced.
 * ------------------------------------------------------------------------------------------ */
typedef struct ssr_resample_plan ssr_resample_plan;
int ssr_resample_plan_create(ssr_resample_plan** plan, int up, int down, const float* taps_host,
                             int n_taps);
int ssr_resample_plan_destroy(ssr_resample_plan* plan);
/* ceil(n_in * up / down) */
int64_t ssr_resample_out_len(const ssr_resample_plan* plan, int64_t n_in);
int ssr_resample_poly_batched(const ssr_resample_plan* plan, const float* x_dev,
                              const int64_t* in_offsets_host, const int64_t* in_offsets_dev,
                              float* y_dev, const int64_t* out_offsets_host,
                              const int64_t* out_offsets_dev, int n, void* stream);

/* float64 form (taps, input and output float64): what librosa.resample(res_type="polyphase") does when the
 * testee hands back a float64 waveform (scipy.signal.resample_poly keeps the input dtype, eval.py:144-150).
 * A plan is either float32 or float64; mixing them is SSR_ERR_INVALID. */
int ssr_resample_plan_create_f64(ssr_resample_plan** plan, int up, int down, const double* taps_host,
                                 int n_taps);
int ssr_resample_poly_batched_f64(const ssr_resample_plan* plan, const double* x_dev,
                                  const int64_t* in_offsets_host, const int64_t* in_offsets_dev,
                                  double* y_dev, const int64_t* out_offsets_host,
                                  const int64_t* out_offsets_dev, int n, void* stream);

/* Explicit polyphase bank instead of a prototype FIR, for resamplers whose per-phase weights are not samples of one
 * prototype -- resampy's ``kaiser_best`` (what librosa 0.9's librosa.load(sr=...) runs, ssr_eval/eval.py:242,
 * ssr_eval/metrics.py:22-23): a 512-per-zero-crossing table read with linear interpolation and a truncated table step.
 * Output j sits at time j * down / up (input samples): n = floor, phase = (j * down) % up; tap k of that phase
 * (bank_host[phase * K + k]) multiplies x[n + lead - k], zero extension outside the utterance.  Such a plan produces
 * floor(n_in * up / down) outputs (resampy's length; ssr_resample_out_len knows).  float32 only. */
int ssr_resample_plan_create_bank(ssr_resample_plan** plan, int up, int down, const float* bank_host, int K, int lead);

/* ------------------------------------------------------------------------------------------
 * K4: STFT hard low-pass = stft_hard_lowpass_v0 (ssr_eval/lowpass.py:17-28) through
 * FDomainHelper.wav_to_spectrogram_phase / spectrogram_phase_to_wav (ssr_eval/dsp.py:76-119):
 * STFT (n_fft, hop, periodic Hann, centre, reflect) -> zero bins >= cut_bin (kept bins pass through; the reference
 * rebuilds them as mag*cos, mag*sin with eps 1e-8: the bin itself up to two roundings -- the dense mode below and the
 * n_fft != 2048 kernel do that) -> ISTFT (x window / n_fft, overlap-add, / clamp(overlap-added window^2, 1e-11))
 * -> drop n_fft/2 -> exactly `length` samples.  float32 arithmetic.  n_fft must be a power of
 * two in [256, 4096]; the reference always uses n_fft 2048 / hop 441 (dsp.py:9-10).
 * cut_bins_dev: one int32 per utterance = int((n_fft/2+1) * lowpass_ratio).
 * ------------------------------------------------------------------------------------------ */
typedef struct ssr_lowpass_plan ssr_lowpass_plan;
int ssr_lowpass_plan_create(ssr_lowpass_plan** plan, int n_fft, int hop);
int ssr_lowpass_plan_destroy(ssr_lowpass_plan* plan);
int ssr_stft_hard_lowpass_batched(const ssr_lowpass_plan* plan, const float* x_dev,
                                  const int64_t* offsets_host, const int64_t* offsets_dev, int n,
                                  const int32_t* cut_bins_dev, float* y_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4d: reference-faithful ("dense") mode of the STFT hard low-pass.  torchlibrosa's STFT / ISTFT
 * (ssr_eval/dsp.py:21-39, used by ssr_eval/lowpass.py:17-28) are not FFTs but dense float32 convolutions with
 * (DFT matrix x Hann) kernels; the bins above the cutoff of the result ARE the rounding noise of that arithmetic and
 * LSD / log-sispec of such an estimate measure it.  This mode multiplies by the same float32 matrices (built on the
 * host with torchlibrosa's formula: stft_w_* = (n_fft/2+1) x n_fft, istft_w_* = n_fft x n_fft (out, in),
 * ola_window = window^2) in the accumulation order of the reference's convolutions (see stft_lowpass_dense.cu), with
 * IEEE sqrt / division: bit-identical to the CPU reference for almost every sample, ~100x the arithmetic of K4.
 * Utterances must be longer than n_fft/2 (torchlibrosa's reflect padding).  Any workspace that holds the longest
 * utterance works (the batch is chunked); ssr_stft_hard_lowpass_dense_workspace_bytes = one chunk for everything.
 * ------------------------------------------------------------------------------------------ */
typedef struct ssr_lowpass_dense_plan ssr_lowpass_dense_plan;
int ssr_lowpass_dense_plan_create(ssr_lowpass_dense_plan** plan, int n_fft, int hop, const float* stft_w_real_host,
                                  const float* stft_w_imag_host, const float* istft_w_real_host,
                                  const float* istft_w_imag_host, const float* ola_window_host);
int ssr_lowpass_dense_plan_destroy(ssr_lowpass_dense_plan* plan);
size_t ssr_stft_hard_lowpass_dense_workspace_bytes(const ssr_lowpass_dense_plan* plan, const int64_t* offsets_host, int n);
int ssr_stft_hard_lowpass_dense_batched(const ssr_lowpass_dense_plan* plan, const float* x_dev, const int64_t* offsets_host,
                                        const int64_t* offsets_dev, int n, const int32_t* cut_bins_dev, float* y_dev,
                                        void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6 ("next" row): the STFT splice of BasicTestee.postprocessing
 * (ssr_eval/eval.py:33-41): librosa.stft (n_fft 2048, hop 512) of the model input x and of the model
 * output, bins below cut_bin taken from x, librosa.istft(length = len(out)).  x and out of a pair
 * must have the same length (the reference's array assignment requires equal frame counts).
 * cut_bins_dev: one int32 per utterance = BasicTestee._get_cutoff_index(x) (eval.py:28-31).
 * ------------------------------------------------------------------------------------------ */
typedef struct ssr_splice_plan ssr_splice_plan;
int ssr_splice_plan_create(ssr_splice_plan** plan, int n_fft, int hop);
int ssr_splice_plan_destroy(ssr_splice_plan* plan);
int ssr_stft_splice_istft_batched(const ssr_splice_plan* plan, const float* x_dev, const float* out_dev,
                                  const int64_t* offsets_host, const int64_t* offsets_dev, int n,
                                  const int32_t* cut_bins_dev, float* y_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7 ("next" row): zero-phase IIR filtering = scipy.signal.sosfiltfilt(sos, x)
 * with the default odd padding, the recursion behind lowpass_filter / bandpass_filter
 * (ssr_eval/lowpass.py:54-131).  The filter design stays on the host (scipy, as in the reference):
 * sos_host = n_sections x 6 float64 (a0 == 1), zi_host = scipy.signal.sosfilt_zi(sos) (n_sections x 2),
 * edge = scipy's default padlen = 3 * ntaps.  Input float32, output float64 (scipy's result type).
 * Every utterance must be longer than `edge`.  Workspace: ssr_sosfiltfilt_workspace_bytes.
 * ------------------------------------------------------------------------------------------ */
size_t ssr_sosfiltfilt_workspace_bytes(const int64_t* offsets_host, int n, int edge);
int ssr_sosfiltfilt_batched(const double* sos_host, int n_sections, const double* zi_host, int edge,
                            const float* x_dev, const int64_t* offsets_host, const int64_t* offsets_dev,
                            int n, double* y_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K0: 16-bit PCM -> float32, dst[i] = (float)src[i] / 32768 -- what librosa.load / soundfile.read hand the
 * reference for a 16-bit wav (ssr_eval/metrics.py:22-23, ssr_eval/eval.py:133-134, 242).  Host-buffer callers
 * upload the 2-byte samples and convert on the device: half the PCIe traffic, bit-identical values.
 * ------------------------------------------------------------------------------------------ */
int ssr_pcm16_to_float(const int16_t* src_dev, float* dst_dev, int64_t n, void* stream);

/* Measurement support: the measured FP64 instruction rate of the current device (thread-instructions / s, DFMA chains),
 * the denominator of bench.py's secondary (FP64-pipe) roofline for K1.  Synchronises the stream. */
int ssr_probe_fp64_rate(double* thread_instr_per_s, void* stream);

/* ------------------------------------------------------------------------------------------
 * K8 ("next" row): the alignment step of the mp3 degradation,
 *   np.argmax(scipy.signal.correlate(decoded, x))            (ssr_eval/eval.py:319; the caller subtracts len(x), :319)
 * for a batch of equal-length (decoded, x) float32 pairs (eval.py:318 unifies the lengths): FFT cross-correlation
 * (zero-padded to 2^m >= 2L-1, float32 complex like scipy's own path) + first-maximum argmax over scipy's 'full'
 * index k = 0 .. 2L-2.  L <= 524288 samples.  The codec itself (the sox binary, eval.py:308-316) is out of scope.
 * argmax_dev: one int64 per pair.  Any workspace that holds the longest pair works; ssr_xcorr_workspace_bytes = one pass.
 * ------------------------------------------------------------------------------------------ */
size_t ssr_xcorr_workspace_bytes(const int64_t* offsets_host, int n);
int ssr_xcorr_argmax_batched(const float* a_dev, const float* x_dev, const int64_t* offsets_host,
                             const int64_t* offsets_dev, int n, int64_t* argmax_dev, void* workspace_dev,
                             size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSR_B200_H_ */
