"""ctypes binding of libsdaopt_b200.so (include/sdaopt_b200.h).

The library is the product: there is no CPU fallback.  Loading fails loudly if the shared object
has not been built, and every compute entry point fails with ``SdaError`` if CUDA is unusable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDA_LIB_PATH") or os.path.join(_HERE, "libsdaopt_b200.so")  # override: A/B builds

SDA_OK = 0
SDA_E_INVALID, SDA_E_BOUNDS, SDA_E_CUDA, SDA_E_WORKSPACE, SDA_E_UNSUPPORTED = -1, -2, -3, -4, -5
RNG_PHILOX, RNG_TAPE = 0, 1
CHAIN_REINIT_FAILED, CHAIN_TAPE_EXHAUSTED = 1, 2

# every symbol include/sdaopt_b200.h declares
EXPORTS = ["sda_abi_version", "sda_has_user_functor", "sda_error_string", "sda_last_cuda_error", "sda_functor_id",
           "sda_functor_name", "sda_device_count", "sda_workspace_bytes", "sda_batch_run", "sda_last_launch_count",
           "sda_batch_run_host", "sda_eval", "sda_eval_host", "sda_visiting_host",
           "sda_visit_fn_host", "sda_local_search_host", "sda_fp64_peak", "sda_kernel_timing",
           "sda_kernel_times", "sda_math_probe_host", "sda_functor_arity", "sda_energy_reset_host"]


class SdaProblem(C.Structure):
    _fields_ = [("functor", C.c_int32), ("dim", C.c_int32), ("lower", C.c_void_p),
                ("upper", C.c_void_p), ("params", C.c_void_p), ("n_params", C.c_int32)]


class SdaConfig(C.Structure):
    _fields_ = [("maxiter", C.c_int64), ("initial_temp", C.c_double), ("visit", C.c_double),
                ("accept", C.c_double), ("maxfun", C.c_double), ("pure_sa", C.c_int32),
                ("temperature_restart", C.c_double), ("rng_mode", C.c_int32),
                ("max_calls", C.c_int64), ("fglob", C.c_double), ("tol", C.c_double),
                ("trace_len", C.c_int64), ("temp_table", C.c_void_p),
                ("sigmax_table", C.c_void_p), ("table_len", C.c_int64)]


RESULT_FIELDS = ("x", "fun", "nit", "ncall", "ncall_ls", "n_ls", "status", "hit", "ncall_hit",
                 "f_hit", "trace_e", "trace_d", "tape_used")


class SdaResult(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in RESULT_FIELDS]


class SdaError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = "libsdaopt_b200: error %d" % code
        try:
            msg = "libsdaopt_b200: %s" % lib().sda_error_string(code).decode()
            if code == SDA_E_CUDA:
                msg += " [%s]" % lib().sda_last_cuda_error().decode()
        except Exception:
            pass
        super().__init__(msg + (" " + detail if detail else ""))


_LIBS = {}
SDA_FN_USER = 255


def lib(path=None):
    """The loaded CUDA library (default build, or the build that carries a user functor).
    Raises if it was not built (python -m sdaopt_b200.build)."""
    path = path or LIB_PATH
    if path not in _LIBS:
        if not os.path.exists(path):
            raise ImportError(
                "%s is missing: build it with `python -m sdaopt_b200.build` "
                "(needs nvcc); sdaopt_b200 has no CPU fallback" % path)
        L = C.CDLL(path)
        L.sda_abi_version.restype = C.c_int
        L.sda_error_string.restype = C.c_char_p
        L.sda_error_string.argtypes = [C.c_int]
        L.sda_last_cuda_error.restype = C.c_char_p
        L.sda_functor_id.restype = C.c_int
        L.sda_functor_id.argtypes = [C.c_char_p]
        L.sda_functor_name.restype = C.c_char_p
        L.sda_functor_name.argtypes = [C.c_int]
        L.sda_device_count.restype = C.c_int
        L.sda_workspace_bytes.restype = C.c_size_t
        L.sda_workspace_bytes.argtypes = [C.POINTER(SdaProblem), C.POINTER(SdaConfig), C.c_int64]
        L.sda_batch_run.restype = C.c_int
        L.sda_batch_run.argtypes = [C.POINTER(SdaProblem), C.POINTER(SdaConfig), C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.POINTER(SdaResult), C.c_void_p, C.c_size_t, C.c_void_p]
        L.sda_batch_run_host.restype = C.c_int
        L.sda_batch_run_host.argtypes = [C.POINTER(SdaProblem), C.POINTER(SdaConfig), C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.POINTER(SdaResult), C.c_int]
        L.sda_eval.restype = C.c_int
        L.sda_eval.argtypes = [C.POINTER(SdaProblem), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.sda_eval_host.restype = C.c_int
        L.sda_eval_host.argtypes = [C.POINTER(SdaProblem), C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.sda_visiting_host.restype = C.c_int
        L.sda_visiting_host.argtypes = [C.POINTER(SdaProblem), C.c_double, C.c_void_p, C.c_int64,
                                        C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.POINTER(C.c_int64), C.c_int]
        L.sda_visit_fn_host.restype = C.c_int
        L.sda_visit_fn_host.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_int64, C.POINTER(C.c_int64), C.c_int]
        L.sda_local_search_host.restype = C.c_int
        L.sda_local_search_host.argtypes = [C.POINTER(SdaProblem), C.c_void_p, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.sda_fp64_peak.restype = C.c_int
        L.sda_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.sda_has_user_functor.restype = C.c_int
        L.sda_last_launch_count.restype = C.c_longlong
        L.sda_kernel_timing.restype = C.c_int
        L.sda_kernel_timing.argtypes = [C.c_int]
        L.sda_math_probe_host.restype = C.c_int
        L.sda_math_probe_host.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        if hasattr(L, "sda_functor_arity"):      # (absent from round-1 builds loaded through SDA_LIB_PATH for A/B runs)
            L.sda_functor_arity.restype = C.c_int
            L.sda_functor_arity.argtypes = [C.c_int]
            L.sda_energy_reset_host.restype = C.c_int
            L.sda_energy_reset_host.argtypes = [C.POINTER(SdaProblem), C.c_void_p, C.c_void_p, C.c_int64,
                                                C.c_int32, C.c_void_p, C.POINTER(C.c_double), C.c_void_p,
                                                C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                                C.POINTER(C.c_int64), C.c_int]
        L.sda_kernel_times.restype = C.c_int
        L.sda_kernel_times.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        if L.sda_abi_version() != 1:
            raise ImportError("libsdaopt_b200.so ABI mismatch")
        _LIBS[path] = L
    return _LIBS[path]


def check(rc, detail=""):
    if rc == SDA_E_BOUNDS:
        raise ValueError('Bounds are note consistent min < max')    # _sda.py:366-367
    if rc != SDA_OK:
        raise SdaError(rc, detail)


def require_cuda():
    n = lib().sda_device_count()
    if n <= 0:
        raise SdaError(SDA_E_CUDA, "no CUDA device: sdaopt_b200 has no CPU fallback")
    return n
