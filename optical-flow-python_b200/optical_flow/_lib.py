"""ctypes binding of libb200flow.so (include/b200flow.h).

This is the only place the Python drop-in touches native code.  There is no CPU implementation behind it:
if the shared library is missing, or no sm_100 (B200) GPU is usable, every operator raises.
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200FLOW_LIB", os.path.join(os.path.dirname(_HERE), "libb200flow.so"))

EINVAL, ECUDA, ENOCONV = -1, -2, -3
# most page-locked host memory one context hands out as result buffers (Context.pinned_empty)
PINNED_CAP_BYTES = int(os.environ.get("B200FLOW_PINNED_CAP_MB", "2048")) << 20


class Penalty(C.Structure):
    _fields_ = [("kind", C.c_int), ("p0", C.c_double), ("p1", C.c_double)]


class Params(C.Structure):
    _fields_ = [
        ("method", C.c_int), ("interp", C.c_int), ("texture", C.c_int),
        ("gnc_iters", C.c_int), ("max_iters", C.c_int), ("max_linear", C.c_int),
        ("max_warping_iters", C.c_int), ("limit_update", C.c_int),
        ("pyramid_levels", C.c_int), ("auto_level", C.c_int), ("gnc_pyramid_levels", C.c_int),
        ("pyramid_spacing", C.c_double), ("gnc_pyramid_spacing", C.c_double),
        ("lambda_", C.c_double), ("lambda_q", C.c_double), ("alpha0", C.c_double), ("alp", C.c_double),
        ("blend", C.c_double), ("deriv_filter", C.c_double * 5),
        ("sigmaD2", C.c_double), ("sigmaS2", C.c_double),
        ("rho_su", Penalty * 2), ("rho_sv", Penalty * 2), ("rho_d", Penalty),
        ("qua_su", Penalty * 2), ("qua_sv", Penalty * 2), ("qua_d", Penalty),
        ("median_h", C.c_int), ("median_w", C.c_int), ("mf_iter", C.c_int), ("area_hsz", C.c_int),
        ("sigma_i", C.c_double), ("occ_sigma_d", C.c_double), ("occ_sigma_i", C.c_double),
        ("solver", C.c_int), ("tol", C.c_double), ("maxit", C.c_int), ("rof_iters", C.c_int),
        ("rof_theta", C.c_double), ("final_median", C.c_int),
    ]


class Stats(C.Structure):
    _fields_ = [("solves", C.c_int), ("pcg_iters", C.c_longlong), ("pcg_pixel_iters", C.c_longlong),
                ("kernel_launches", C.c_int), ("not_converged", C.c_int),
                ("solver_ms", C.c_double), ("warp_ms", C.c_double), ("filter_ms", C.c_double),
                ("pre_ms", C.c_double), ("total_ms", C.c_double),
                ("kernel_ms", C.c_double * 11), ("kernel_bytes", C.c_double * 11), ("kernel_calls", C.c_int * 11)]

    KERNEL_GROUPS = ("rof", "pyramid", "resample_flow", "level_prep", "warp_assemble", "solver", "clip_add", "occlusion",
                     "weighted_median", "median", "misc")

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("kernel_ms", "kernel_bytes", "kernel_calls")}
        d["kernels"] = {n: {"ms": self.kernel_ms[i], "bytes": self.kernel_bytes[i], "calls": self.kernel_calls[i]}
                        for i, n in enumerate(self.KERNEL_GROUPS)}
        return d


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> argtypes after the leading ctx pointer (restype is always int)
_SIGS = {
    "b200flow_estimate": [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.POINTER(Stats)],
    "b200flow_estimate_dev": [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.POINTER(Stats)],
    "b200flow_estimate_rgb8": [C.POINTER(Params), C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, C.POINTER(Stats)],
    "b200flow_estimate_rgb8_dev": [C.POINTER(Params), C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, C.POINTER(Stats)],
    "b200flow_rgb2gray": [_vp, C.c_int, C.c_int, _vp],
    "b200flow_rgb2lab": [_vp, C.c_int, C.c_int, C.c_int, _vp],
    "b200flow_scale_image": [_vp, C.c_longlong, C.c_double, C.c_double, _vp],
    "b200flow_rof_texture": [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, _vp],
    "b200flow_pyramid": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_double, C.POINTER(_vp), _ip, _ip],
    "b200flow_resample_flow": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "b200flow_partial_deriv": [_vp, _vp, C.c_int, C.c_int, C.c_int, _dp, C.c_double, _vp, _vp, _vp],
    "b200flow_robust_eval": [Penalty, C.c_int, _vp, C.c_longlong, _vp],
    "b200flow_operator_apply": [C.POINTER(Params), C.c_double, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp],
    "b200flow_solve_increment": [C.POINTER(Params), C.c_double, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _ip, _dp],
    "b200flow_median_filter": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "b200flow_detect_occlusion": [_vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double, _vp],
    "b200flow_weighted_median": [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _vp],
    "b200flow_estimate_mc": [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp,
                             C.POINTER(Stats)],
    "b200flow_partial_deriv_mc": [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double, _vp, _vp, _vp],
    "b200flow_operator_apply_mc": [C.POINTER(Params), C.c_double, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                   _vp, _vp, _vp, _vp],
    "b200flow_solve_increment_mc": [C.POINTER(Params), C.c_double, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                    _vp, _ip, _dp],
    "b200flow_solve_rhs_mc": [C.POINTER(Params), C.c_double, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp,
                              _vp, _ip, _dp],
    "b200flow_detect_occlusion_mc": [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp],
    "b200flow_flow_error": [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "b200flow_flow_error_dev": [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "b200flow_flow_to_color": [_vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp],
    "b200flow_flow_to_flo": [_vp, C.c_int, C.c_int, C.c_int, _vp],
    "b200flow_debug_pcg_bench": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp],
}
_SIGS["b200flow_host_alloc"] = [C.c_ulonglong, C.POINTER(_vp)]
_SIGS["b200flow_host_free"] = [_vp]
_SIGS["b200flow_ctx_set_split"] = [C.c_int, C.c_int]
_SIGS["b200flow_band_init"] = [C.c_int, C.c_int, C.c_ulonglong, C.c_int]
_SIGS["b200flow_band_export"] = [_vp, C.POINTER(_vp)]
_SIGS["b200flow_band_connect"] = [C.c_int, _vp, _vp]
_SIGS["b200flow_band_close"] = []
_SIGS["b200flow_ctx_set_log"] = [C.c_int]
_SIGS["b200flow_ctx_get_log"] = [_vp, C.c_int, C.POINTER(C.c_int)]

EXPORTS = sorted(list(_SIGS) + ["b200flow_abi_version", "b200flow_ctx_create", "b200flow_ctx_destroy",
                                "b200flow_last_error", "b200flow_ctx_set_timing", "b200flow_ctx_sync",
                                "b200flow_ctx_stream", "b200flow_ctx_num_sms"])

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libb200flow.so and declare every prototype.  Needs no GPU (used by the symbol-export test)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb200flow.so not found at %s -- run `python optical-flow-python_b200/build.py` "
                               "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.b200flow_abi_version.restype = C.c_int
        lib.b200flow_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
        lib.b200flow_ctx_create.restype = C.c_int
        lib.b200flow_ctx_destroy.argtypes = [_vp]
        lib.b200flow_ctx_destroy.restype = None
        lib.b200flow_last_error.argtypes = [_vp]
        lib.b200flow_last_error.restype = C.c_char_p
        lib.b200flow_ctx_set_timing.argtypes = [_vp, C.c_int]
        lib.b200flow_ctx_sync.argtypes = [_vp]
        lib.b200flow_ctx_stream.argtypes = [_vp]
        lib.b200flow_ctx_stream.restype = _vp
        lib.b200flow_ctx_num_sms.argtypes = [_vp]
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = [_vp] + args
            fn.restype = C.c_int
        _lib = lib
        return lib


class Context:
    """One (device, stream, arena).  Not thread-safe; use one per thread / per GPU."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.b200flow_ctx_create(int(device), C.byref(h))
        if rc != 0:
            msg = self.lib.b200flow_last_error(None).decode()
            raise RuntimeError("b200flow: cannot create a context on cuda:%d: %s" % (device, msg))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            for free in self.__dict__.get("_pinned_free", {}).values():     # buffers still owned by live arrays stay valid
                while free:
                    self.lib.b200flow_host_free(self.handle, _vp(free.pop()))
            self.lib.b200flow_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self):
        return self.lib.b200flow_last_error(self.handle).decode()

    def call(self, name, *args, allow_noconv=False):
        rc = getattr(self.lib, name)(self.handle, *args)
        if rc == 0 or (rc == ENOCONV and allow_noconv):
            return rc
        msg = self.error()
        if rc == EINVAL:
            raise ValueError(msg)
        raise RuntimeError("b200flow %s failed (%d): %s" % (name, rc, msg))

    def set_timing(self, on):
        self.lib.b200flow_ctx_set_timing(self.handle, int(bool(on)))

    def sync(self):
        self.lib.b200flow_ctx_sync(self.handle)

    def set_log(self, on):
        """display=True: record one row per linear solve of single-pair runs (include/b200flow.h, b200flow_ctx_set_log)."""
        self.call("b200flow_ctx_set_log", int(bool(on)))

    def get_log(self):
        """Rows (gnc, level, warp, linearisation, norm) of the last single-pair run, 0-based indices."""
        n = C.c_int(0)
        self.call("b200flow_ctx_get_log", None, 0, C.byref(n))
        rows = np.zeros((max(n.value, 1), 5))
        self.call("b200flow_ctx_get_log", ptr(rows), n.value, C.byref(n))
        return rows[:n.value]

    def set_split(self, groups, solver_ctas_per_sm=0):
        """Concurrent sub-batches: batched calls run `groups` groups of pairs on their own streams (include/b200flow.h)."""
        self.call("b200flow_ctx_set_split", int(groups), int(solver_ctas_per_sm))

    @property
    def stream(self):
        return self.lib.b200flow_ctx_stream(self.handle)

    @property
    def num_sms(self):
        return self.lib.b200flow_ctx_num_sms(self.handle)

    # ---- page-locked result buffers ---------------------------------------------------------------------------
    def pinned_empty(self, shape, dtype=np.float64):
        """Uninitialised C-contiguous NumPy array in page-locked host memory (cudaHostAlloc), so that the device->host
        copy of a result is one DMA instead of a staged copy into freshly mapped pageable pages.  The memory goes back to
        a per-context free list when the array (and every view of it) is garbage collected and is handed out again for
        the next request of the same size: a loop of `uv = estimate_flow_batch(...)` allocates twice and then recycles.
        A caller that keeps many results alive is not allowed to pin the host: beyond PINNED_CAP_BYTES of page-locked
        memory owned by this context the array is an ordinary pageable np.empty."""
        shape = tuple(int(v) for v in np.atleast_1d(shape))
        count = int(np.prod(shape))
        nbytes = count * np.dtype(dtype).itemsize
        pool = self.__dict__.setdefault("_pinned_free", {})
        free = pool.setdefault(nbytes, [])
        if free:
            addr = free.pop()
        else:
            owned = self.__dict__.get("_pinned_owned", 0)
            if owned + nbytes > PINNED_CAP_BYTES:
                return np.empty(shape, dtype=dtype)
            p = _vp()
            self.call("b200flow_host_alloc", C.c_ulonglong(max(nbytes, 1)), C.byref(p))
            addr = p.value
            self._pinned_owned = owned + nbytes
        buf = (C.c_char * max(nbytes, 1)).from_address(addr)
        weakref.finalize(buf, free.append, addr)          # recycled, not freed: pinning 80 MB costs ~10 ms
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


_tls = threading.local()


def default_context(device=None):
    """Per-thread default context (device from B200FLOW_DEVICE or LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("B200FLOW_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    if device not in cache:
        cache[device] = Context(device)
    return cache[device]


def f64(a):
    """C-contiguous float64 view/copy of a."""
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)
