"""ctypes binding of the C ABI in include/ssr_b200.h (libssr_b200.so, hand-written sm_100a CUDA).

There is NO CPU fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

# SSR_B200_LIB: another build of the same library (A/B timing of kernel variants, tools/); default = the in-tree build
_LIB_PATH = os.environ.get("SSR_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                          "libssr_b200.so")
_lib = None

SSR_OK = 0
METRIC_LSD, METRIC_LOG_SISPEC, METRIC_SISPEC, METRIC_SSIM, METRIC_ALL = 1, 2, 4, 8, 15
METRIC_NAMES = ("lsd", "log_sispec", "sispec", "ssim")


class NativeError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def lib():
    """Load (once) and return the CUDA extension; fail loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeError(
            "ssr_eval_b200: %s not found -- build it with `make` (or __graft_entry__.build()); "
            "there is no CPU fallback" % _LIB_PATH)
    L = ctypes.CDLL(_LIB_PATH)
    c_int, c_uint, c_sz, c_i64, vp = ctypes.c_int, ctypes.c_uint, ctypes.c_size_t, ctypes.c_int64, ctypes.c_void_p
    L.ssr_version.restype = c_int
    L.ssr_last_error.restype = ctypes.c_char_p
    L.ssr_launch_count.restype = ctypes.c_uint64
    L.ssr_timing_enable.argtypes = [c_int]
    L.ssr_timing_collect.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int)]
    L.ssr_stft_plan_create.argtypes = [ctypes.POINTER(vp), c_int, c_int, vp]
    L.ssr_stft_plan_destroy.argtypes = [vp]
    L.ssr_stft_num_frames.argtypes = [vp, c_i64]
    L.ssr_stft_num_frames.restype = c_i64
    L.ssr_stft_metrics_workspace_bytes.argtypes = [vp, vp, c_int, c_uint]
    L.ssr_stft_metrics_workspace_bytes.restype = c_sz
    L.ssr_stft_metrics_batched.argtypes = [vp, vp, vp, vp, vp, c_int, c_uint, vp, vp, c_sz, vp]
    L.ssr_stft_metrics_batched_f64est.argtypes = [vp, vp, vp, vp, vp, c_int, c_uint, vp, vp, c_sz, vp]
    L.ssr_stft_metrics_batched_f64.argtypes = [vp, vp, vp, vp, vp, c_int, c_uint, vp, vp, c_sz, vp]
    L.ssr_stft_magnitude_batched.argtypes = [vp, vp, vp, vp, c_int, vp, vp, c_sz, vp]
    L.ssr_resample_plan_create.argtypes = [ctypes.POINTER(vp), c_int, c_int, vp, c_int]
    L.ssr_resample_plan_create_f64.argtypes = [ctypes.POINTER(vp), c_int, c_int, vp, c_int]
    L.ssr_resample_poly_batched_f64.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_int, vp]
    L.ssr_resample_plan_create_bank.argtypes = [ctypes.POINTER(vp), c_int, c_int, vp, c_int, c_int]
    L.ssr_resample_plan_destroy.argtypes = [vp]
    L.ssr_resample_out_len.argtypes = [vp, c_i64]
    L.ssr_resample_out_len.restype = c_i64
    L.ssr_resample_poly_batched.argtypes = [vp, vp, vp, vp, vp, vp, vp, c_int, vp]
    L.ssr_lowpass_plan_create.argtypes = [ctypes.POINTER(vp), c_int, c_int]
    L.ssr_lowpass_plan_destroy.argtypes = [vp]
    L.ssr_stft_hard_lowpass_batched.argtypes = [vp, vp, vp, vp, c_int, vp, vp, vp]
    L.ssr_splice_plan_create.argtypes = [ctypes.POINTER(vp), c_int, c_int]
    L.ssr_splice_plan_destroy.argtypes = [vp]
    L.ssr_stft_splice_istft_batched.argtypes = [vp, vp, vp, vp, vp, c_int, vp, vp, vp]
    L.ssr_sosfiltfilt_workspace_bytes.argtypes = [vp, c_int, c_int]
    L.ssr_sosfiltfilt_workspace_bytes.restype = c_sz
    L.ssr_sosfiltfilt_batched.argtypes = [vp, c_int, vp, c_int, vp, vp, vp, c_int, vp, vp, c_sz, vp]
    L.ssr_pcm16_to_float.argtypes = [vp, vp, c_i64, vp]
    L.ssr_probe_fp64_rate.argtypes = [ctypes.POINTER(ctypes.c_double), vp]
    L.ssr_lowpass_dense_plan_create.argtypes = [ctypes.POINTER(vp), c_int, c_int, vp, vp, vp, vp, vp]
    L.ssr_lowpass_dense_plan_destroy.argtypes = [vp]
    L.ssr_stft_hard_lowpass_dense_workspace_bytes.argtypes = [vp, vp, c_int]
    L.ssr_stft_hard_lowpass_dense_workspace_bytes.restype = c_sz
    L.ssr_stft_hard_lowpass_dense_batched.argtypes = [vp, vp, vp, vp, c_int, vp, vp, vp, c_sz, vp]
    L.ssr_xcorr_workspace_bytes.argtypes = [vp, c_int]
    L.ssr_xcorr_workspace_bytes.restype = c_sz
    L.ssr_xcorr_argmax_batched.argtypes = [vp, vp, vp, vp, c_int, vp, vp, c_sz, vp]
    _lib = L
    return L


def check(status, what):
    if status != SSR_OK:
        msg = lib().ssr_last_error()
        raise NativeError("%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))


def timing_enable(on=True):
    check(lib().ssr_timing_enable(1 if on else 0), "ssr_timing_enable")


def timing_collect():
    """(summed k_stft_metrics milliseconds, launches) since the last collect."""
    ms, n = ctypes.c_double(0.0), ctypes.c_int(0)
    check(lib().ssr_timing_collect(ctypes.byref(ms), ctypes.byref(n)), "ssr_timing_collect")
    return ms.value, n.value


def probe_fp64_rate(stream=None):
    """Measured FP64 thread-instructions per second of the current device (see include/ssr_b200.h)."""
    v = ctypes.c_double(0.0)
    check(lib().ssr_probe_fp64_rate(ctypes.byref(v), stream), "ssr_probe_fp64_rate")
    return v.value


def launch_count():
    return int(lib().ssr_launch_count())


EXPORTED_SYMBOLS = (
    "ssr_version", "ssr_last_error", "ssr_launch_count", "ssr_timing_enable", "ssr_timing_collect",
    "ssr_stft_plan_create", "ssr_stft_plan_destroy", "ssr_stft_num_frames",
    "ssr_stft_metrics_workspace_bytes", "ssr_stft_metrics_batched", "ssr_stft_metrics_batched_f64est", "ssr_stft_metrics_batched_f64",
    "ssr_stft_magnitude_batched",
    "ssr_resample_plan_create", "ssr_resample_plan_destroy", "ssr_resample_out_len",
    "ssr_resample_poly_batched", "ssr_resample_plan_create_f64", "ssr_resample_poly_batched_f64",
    "ssr_resample_plan_create_bank",
    "ssr_lowpass_plan_create", "ssr_lowpass_plan_destroy", "ssr_stft_hard_lowpass_batched",
    "ssr_splice_plan_create", "ssr_splice_plan_destroy", "ssr_stft_splice_istft_batched",
    "ssr_sosfiltfilt_workspace_bytes", "ssr_sosfiltfilt_batched",
    "ssr_pcm16_to_float", "ssr_probe_fp64_rate",
    "ssr_lowpass_dense_plan_create", "ssr_lowpass_dense_plan_destroy",
    "ssr_stft_hard_lowpass_dense_workspace_bytes", "ssr_stft_hard_lowpass_dense_batched",
    "ssr_xcorr_workspace_bytes", "ssr_xcorr_argmax_batched",
)
