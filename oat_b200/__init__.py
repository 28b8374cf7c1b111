"""oat_b200 -- B200-native implementation of Oat's per-frame tracking hot path.

This package is a thin ctypes binding of ``liboatgpu.so`` (C ABI in ``include/oatgpu.h``;
kernels in ``oat_b200/csrc``) that mirrors the reference's operator interface for the path:

* :class:`BackgroundSubtractorMOG`  -- ``framefilt mog``  (src/framefilter/BackgroundSubtractorMOG.cpp:114-127)
* :func:`color_convert_hsv`        -- ``framefilt col -C HSV`` (src/framefilter/ColorConvert.cpp:101-107)
* :class:`BackgroundSubtractor`     -- ``framefilt bsub`` (src/framefilter/BackgroundSubtractor.cpp:87-100)
* :class:`HSVDetector`              -- ``posidet hsv``    (src/positiondetector/HSVDetector.cpp:142-173)
* :class:`Tracker`                  -- the three chained, fused on the device

There is NO CPU fallback: importing works without a GPU (so the build can be checked), but
every compute call raises :class:`OatError` when the CUDA library or a device is missing.
Nothing here imports ``oracle/`` (test infrastructure).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OAT_B200_LIB") or os.path.join(_HERE, "liboatgpu.so")  # env override: kernel experiments
DBL_MAX = float(np.finfo(np.float64).max)

OAT_OK = 0


class OatError(RuntimeError):
    """Raised for any non-zero oat_status (message = oat_last_error())."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"liboatgpu error {code}: {msg}")
        self.code = code


class MogParams(C.Structure):
    _fields_ = [
        ("history", C.c_int),
        ("nmixtures", C.c_int),
        ("var_threshold", C.c_float),
        ("var_threshold_gen", C.c_float),
        ("background_ratio", C.c_float),
        ("var_init", C.c_float),
        ("var_min", C.c_float),
        ("var_max", C.c_float),
        ("complexity_reduction_threshold", C.c_float),
        ("detect_shadows", C.c_int),
        ("shadow_value", C.c_int),
        ("shadow_threshold", C.c_float),
    ]


class HsvParams(C.Structure):
    """HSVDetector options (src/positiondetector/HSVDetector.cpp:54-69)."""

    _fields_ = [
        ("h_min", C.c_int), ("h_max", C.c_int),
        ("s_min", C.c_int), ("s_max", C.c_int),
        ("v_min", C.c_int), ("v_max", C.c_int),
        ("erode_px", C.c_int),
        ("dilate_px", C.c_int),
        ("min_area", C.c_double),
        ("max_area", C.c_double),
    ]

    @classmethod
    def make(cls, h=(0, 256), s=(0, 256), v=(0, 256), erode=0, dilate=10, area=(0.0, DBL_MAX)):
        return cls(h[0], h[1], s[0], s[1], v[0], v[1], erode, dilate, area[0], area[1])


class Position(C.Structure):
    """oat_position: the Position2D fields the position filters/combiners touch (Position2D.h:112-155)."""
    _fields_ = [
        ("position_valid", C.c_int32),
        ("velocity_valid", C.c_int32),
        ("heading_valid", C.c_int32),
        ("reserved", C.c_int32),
        ("x", C.c_double),
        ("y", C.c_double),
        ("vx", C.c_double),
        ("vy", C.c_double),
        ("hx", C.c_double),
        ("hy", C.c_double),
    ]


class KalmanParams(C.Structure):
    """oat_kalman_params: oat posifilt kalman --dt/--timeout/--sigma-accel/--sigma-noise (KalmanFilter2D.cpp:40-92)."""
    _fields_ = [("dt", C.c_double), ("timeout", C.c_double), ("sigma_accel", C.c_double), ("sigma_noise", C.c_double)]


class Detection(C.Structure):
    _fields_ = [
        ("position_valid", C.c_int32),
        ("n_components", C.c_int32),
        ("x", C.c_double),
        ("y", C.c_double),
        ("area", C.c_double),
    ]

    def as_tuple(self):
        return (bool(self.position_valid), self.x, self.y, self.area)


_lib = None

_SIGS = {
    # name: (restype, argtypes)
    "oat_abi_version": (C.c_int, []),
    "oat_last_error": (C.c_char_p, []),
    "oat_device_count": (C.c_int, []),
    "oat_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "oat_ctx_destroy": (C.c_int, [C.c_void_p]),
    "oat_ctx_sync": (C.c_int, [C.c_void_p]),
    "oat_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "oat_ctx_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "oat_ctx_clip_host_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "oat_ctx_profile_resident": (C.c_int, [C.c_void_p, C.c_int]),
    "oat_ctx_profile_resident_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64),
                                                C.POINTER(C.c_uint64)]),
    "oat_mog_default_params": (None, [C.POINTER(MogParams)]),
    "oat_mog_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(MogParams), C.POINTER(C.c_void_p)]),
    "oat_mog_destroy": (C.c_int, [C.c_void_p]),
    "oat_mog_reset": (C.c_int, [C.c_void_p]),
    "oat_mog_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                C.c_double]),
    "oat_mog_apply_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.c_double]),
    "oat_mog_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oat_mog_live_modes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "oat_bgr2hsv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "oat_bsub_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_void_p)]),
    "oat_bsub_destroy": (C.c_int, [C.c_void_p]),
    "oat_bsub_set_background": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "oat_bsub_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "oat_hsv_default_params": (None, [C.POINTER(HsvParams)]),
    "oat_hsvdet_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "oat_hsvdet_destroy": (C.c_int, [C.c_void_p]),
    "oat_hsvdet_detect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(HsvParams), C.POINTER(Detection),
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "oat_sift_contours": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(HsvParams), C.POINTER(Detection),
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "oat_thresh_detect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(HsvParams),
                                    C.POINTER(Detection), C.c_void_p, C.c_size_t, C.c_void_p]),
    "oat_diffdet_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "oat_diffdet_destroy": (C.c_int, [C.c_void_p]),
    "oat_diffdet_reset": (C.c_int, [C.c_void_p]),
    "oat_diffdet_detect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.POINTER(Detection), C.c_void_p, C.c_size_t]),
    "oat_keep_where": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "oat_tracker_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(MogParams), C.c_int,
                                     C.POINTER(C.c_void_p)]),
    "oat_tracker_destroy": (C.c_int, [C.c_void_p]),
    "oat_tracker_reset": (C.c_int, [C.c_void_p]),
    "oat_tracker_track": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.POINTER(HsvParams),
                                    C.POINTER(Detection), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                    C.c_size_t, C.c_void_p, C.c_size_t]),
    "oat_tracker_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.POINTER(HsvParams), C.c_void_p,
                                     C.c_size_t]),
    "oat_tracker_collect": (C.c_int, [C.c_void_p, C.POINTER(Detection)]),
    "oat_tracker_wait_ingest": (C.c_int, [C.c_void_p]),
    "oat_tracker_run_clip": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t, C.c_double,
                                       C.POINTER(HsvParams), C.c_int, C.POINTER(Detection), C.POINTER(Position)]),
    "oat_tracker_live_modes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "oat_tracker_stream_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.POINTER(HsvParams), C.c_uint]),
    "oat_tracker_stream_wait_ingest": (C.c_int, [C.c_void_p]),
    "oat_tracker_stream_flush": (C.c_int, [C.c_void_p, C.c_int]),
    "oat_tracker_stream_poll": (C.c_int, [C.c_void_p, C.POINTER(Detection), C.POINTER(Position), C.c_size_t, C.c_int,
                                          C.POINTER(C.c_size_t)]),
    "oat_tracker_stream_pending": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "oat_kalman_default_params": (None, [C.POINTER(KalmanParams)]),
    "oat_posfilt_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(KalmanParams), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "oat_posfilt_destroy": (C.c_int, [C.c_void_p]),
    "oat_posfilt_reset": (C.c_int, [C.c_void_p]),
    "oat_posfilt_apply": (C.c_int, [C.c_void_p, C.POINTER(Position), C.POINTER(Position)]),
    "oat_tracker_attach_posfilt": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oat_tracker_collect_position": (C.c_int, [C.c_void_p, C.POINTER(Detection), C.POINTER(Position)]),
    "oat_tracker_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oat_tracker_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "oat_tracker_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "oat_tracker_run_clips": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t,
                                        C.c_double, C.POINTER(HsvParams), C.c_int, C.POINTER(Detection)]),
    "oat_tracker_submit_fused_only": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.POINTER(HsvParams)]),
    "oat_tracker_tail_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "oat_synth_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32]),
    "oat_alloc_device": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "oat_free_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oat_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "oat_free_pinned": (C.c_int, [C.c_void_p]),
    "oat_register_host": (C.c_int, [C.c_void_p, C.c_size_t]),
    "oat_unregister_host": (C.c_int, [C.c_void_p]),
    "oat_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p]),
    "oat_ipc_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "oat_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oat_memcpy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oat_ctx_idle": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "oat_memcpy_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oat_memcpy_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "oat_memcpy_done": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "oat_flush_l2": (C.c_int, [C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def build(force: bool = False) -> str:
    """Compile liboatgpu.so in-tree with nvcc for sm_100a (oat_b200/csrc/Makefile)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", "Makefile"))]
    srcs.append(os.path.join(_HERE, "..", "include", "oatgpu.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-s", "-C", csrc] + (["-B"] if force else []), check=True)
    return LIB_PATH


def lib():
    """Load liboatgpu.so; raises OatError (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OatError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ck(code: int):
    if code != OAT_OK:
        raise OatError(code, lib().oat_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return lib().oat_device_count()


def _ptr(a):
    """numpy array -> host pointer; int -> raw (device or pinned) pointer; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, DeviceBuffer):
        return C.c_void_p(a.ptr)
    return C.c_void_p(int(a))


def frame_pointers(frames):
    """A clip as the C ABI takes it: an array of frame pointers (device buffers, pinned pointers or numpy arrays --
    the caller keeps them alive).  Lets a harness build the array outside its timed region."""
    return (C.c_void_p * len(frames))(*[_ptr(f).value for f in frames])


class Context:
    """One per (thread, device); owns the CUDA streams (BackgroundSubtractorMOG::configureGPU,
    src/framefilter/BackgroundSubtractorMOG.cpp:92-111)."""

    def __init__(self, device_index: int = 0):
        self._h = C.c_void_p()
        _ck(lib().oat_ctx_create(device_index, C.byref(self._h)))
        self.device_index = device_index

    def close(self):
        if getattr(self, "_h", None):
            lib().oat_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _ck(lib().oat_ctx_sync(self._h))

    @property
    def stream(self) -> int:
        return int(lib().oat_ctx_stream(self._h) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(lib().oat_ctx_kernel_launches(self._h))

    def clip_host_stats(self):
        """(busy us, waiting us, frames) of the calling thread inside the resident clip engine since the last call."""
        b, w, f = C.c_double(), C.c_double(), C.c_uint64()
        _ck(lib().oat_ctx_clip_host_stats(self._h, C.byref(b), C.byref(w), C.byref(f)))
        return b.value, w.value, f.value

    def profile_resident(self, enable: bool):
        """Bracket every launch of the resident fused kernel with CUDA events (oat_ctx_profile_resident)."""
        _ck(lib().oat_ctx_profile_resident(self._h, 1 if enable else 0))

    def profile_resident_read(self):
        """(total ms, launches, frames) of the resident fused kernel since the last read."""
        ms, n, f = C.c_double(), C.c_uint64(), C.c_uint64()
        _ck(lib().oat_ctx_profile_resident_read(self._h, C.byref(ms), C.byref(n), C.byref(f)))
        return ms.value, n.value, f.value

    def flush_l2(self):
        _ck(lib().oat_flush_l2(self._h))

    def alloc(self, nbytes: int) -> "DeviceBuffer":
        return DeviceBuffer(self, nbytes)

    def memcpy(self, dst, src, nbytes: int):
        _ck(lib().oat_memcpy(self._h, _ptr(dst), _ptr(src), nbytes))

    def synth_frame(self, rows: int, cols: int, seed: int, t: int, out=None, pitch: int | None = None):
        """Synthetic BGR frame; out = numpy array, DeviceBuffer or None (-> numpy);
        pitch: row pitch in bytes of a device buffer (default tight)."""
        ret = None
        if out is None:
            out = ret = np.empty((rows, cols, 3), np.uint8)
        _ck(lib().oat_synth_frame(self._h, _ptr(out), pitch or cols * 3, rows, cols, seed, t))
        return ret if ret is not None else out


class DeviceBuffer:
    """Device-resident Frame variant: raw HBM allocation handed to the C ABI by pointer."""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        self.nbytes = nbytes
        p = C.c_void_p()
        _ck(lib().oat_alloc_device(ctx._h, nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if getattr(self, "ptr", None):
            lib().oat_free_device(self.ctx._h, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        self.ctx.memcpy(self, a, a.nbytes)
        return self

    def download(self, shape, dtype=np.uint8) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        self.ctx.memcpy(out, self, out.nbytes)
        return out


class PinnedArray:
    """Pinned-host Frame variant: page-locked numpy view (async DMA source/target)."""

    def __init__(self, shape, dtype=np.uint8):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _ck(lib().oat_alloc_pinned(n, C.byref(p)))
        self.ptr = p.value
        buf = (C.c_uint8 * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib().oat_free_pinned(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def default_mog_params() -> MogParams:
    p = MogParams()
    lib().oat_mog_default_params(C.byref(p))
    return p


def _img(a, rows, cols, ch):
    if isinstance(a, np.ndarray):
        want = (rows, cols, ch) if ch > 1 else (rows, cols)
        if a.dtype != np.uint8 or a.shape != want or not a.flags.c_contiguous:
            raise ValueError(f"expected contiguous uint8 array of shape {want}, got {a.dtype} {a.shape}")
    return a


def _state_arrays(rows, cols, K):
    return (np.empty((rows, cols), np.uint8), np.empty((rows, cols, K), np.float32),
            np.empty((rows, cols, K), np.float32), np.empty((rows, cols, K, 3), np.float32))


class BackgroundSubtractorMOG:
    """``framefilt mog``: ``filter(frame)`` = MOG2 apply + zero the background
    (src/framefilter/BackgroundSubtractorMOG.cpp:114-127).  ``adaptation_coeff`` is ``-a``."""

    def __init__(self, ctx: Context, rows: int, cols: int, adaptation_coeff: float = 0.0,
                 params: MogParams | None = None):
        if not (0.0 <= adaptation_coeff <= 1.0):  # getNumericValue range check, ...MOG.cpp:87-88
            raise ValueError("adaptation-coeff must be in [0, 1]")
        self.ctx, self.rows, self.cols = ctx, rows, cols
        self.learning_coeff = adaptation_coeff
        self.params = params or default_mog_params()
        self._h = C.c_void_p()
        _ck(lib().oat_mog_create(ctx._h, rows, cols, C.byref(self.params), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().oat_mog_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, frame, learning_rate=None, want_mask=True, want_frame=True):
        """-> (filtered frame or None, mask or None) as numpy arrays (host in, host out)."""
        _img(frame, self.rows, self.cols, 3)
        out = np.empty((self.rows, self.cols, 3), np.uint8) if want_frame else None
        mask = np.empty((self.rows, self.cols), np.uint8) if want_mask else None
        lr = self.learning_coeff if learning_rate is None else learning_rate
        _ck(lib().oat_mog_apply(self._h, _ptr(frame), self.cols * 3, _ptr(out), self.cols * 3, _ptr(mask), self.cols,
                                lr))
        return out, mask

    def apply_async(self, frame, out=None, mask=None, learning_rate=None, pitch=None):
        """Device-resident frame in, device-resident filtered frame / mask out; returns at once (Context.sync waits)."""
        lr = self.learning_coeff if learning_rate is None else learning_rate
        _ck(lib().oat_mog_apply_async(self._h, _ptr(frame), pitch or self.cols * 3, _ptr(out), pitch or self.cols * 3,
                                      _ptr(mask), self.cols, lr))

    def filter(self, frame: np.ndarray) -> np.ndarray:
        """In place on a host frame, like FrameFilter::filter(cv::Mat&)."""
        _img(frame, self.rows, self.cols, 3)
        _ck(lib().oat_mog_apply(self._h, _ptr(frame), self.cols * 3, _ptr(frame), self.cols * 3, None, 0,
                                self.learning_coeff))
        return frame

    def reset(self):
        _ck(lib().oat_mog_reset(self._h))

    def state(self):
        K = self.params.nmixtures
        m, w, v, mu = _state_arrays(self.rows, self.cols, K)
        _ck(lib().oat_mog_get_state(self._h, _ptr(m), _ptr(w), _ptr(v), _ptr(mu)))
        return m, w, v, mu

    def live_modes(self) -> int:
        s = C.c_uint64()
        _ck(lib().oat_mog_live_modes(self._h, C.byref(s)))
        return s.value


def color_convert_hsv(ctx: Context, bgr: np.ndarray) -> np.ndarray:
    """``framefilt col -C HSV`` on a host BGR frame."""
    rows, cols = bgr.shape[:2]
    _img(bgr, rows, cols, 3)
    out = np.empty_like(bgr)
    _ck(lib().oat_bgr2hsv(ctx._h, _ptr(bgr), cols * 3, _ptr(out), cols * 3, rows, cols))
    return out


def threshold_filter(ctx: Context, frame: np.ndarray, i_min: int, i_max: int) -> np.ndarray:
    """``framefilt thresh`` (src/framefilter/Threshold.cpp:67-81) on a host GREY or BGR frame."""
    rows, cols = frame.shape[:2]
    ch = 1 if frame.ndim == 2 else frame.shape[2]
    _img(frame, rows, cols, ch)
    out = np.empty_like(frame)
    _ck(lib().oat_keep_where(ctx._h, _ptr(frame), cols * ch, _ptr(out), cols * ch, rows, cols, ch, None, 0, i_min, i_max))
    return out


def mask_filter(ctx: Context, frame: np.ndarray, roi: np.ndarray) -> np.ndarray:
    """``framefilt mask`` (src/framefilter/FrameMasker.cpp:71-75): frame.setTo(0, roi == 0)."""
    rows, cols = frame.shape[:2]
    ch = 1 if frame.ndim == 2 else frame.shape[2]
    _img(frame, rows, cols, ch)
    _img(roi, rows, cols, 1)
    out = np.empty_like(frame)
    _ck(lib().oat_keep_where(ctx._h, _ptr(frame), cols * ch, _ptr(out), cols * ch, rows, cols, ch, _ptr(roi), cols, 0, 0))
    return out


class BackgroundSubtractor:
    """``framefilt bsub`` (src/framefilter/BackgroundSubtractor.cpp:87-100)."""

    def __init__(self, ctx: Context, rows: int, cols: int, channels: int = 3, adaptation_coeff: float = 0.0):
        self.ctx, self.rows, self.cols, self.ch = ctx, rows, cols, channels
        self._h = C.c_void_p()
        _ck(lib().oat_bsub_create(ctx._h, rows, cols, channels, adaptation_coeff, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().oat_bsub_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_background(self, img: np.ndarray):
        _img(img, self.rows, self.cols, self.ch)
        _ck(lib().oat_bsub_set_background(self._h, _ptr(img), self.cols * self.ch))

    def filter(self, frame: np.ndarray) -> np.ndarray:
        _img(frame, self.rows, self.cols, self.ch)
        out = np.empty_like(frame)
        pitch = self.cols * self.ch
        _ck(lib().oat_bsub_apply(self._h, _ptr(frame), pitch, _ptr(out), pitch))
        return out


class HSVDetector:
    """``posidet hsv``: ``detect(hsv)`` = inRange -> erode -> dilate -> siftContours
    (src/positiondetector/HSVDetector.cpp:142-173, DetectorFunc.cpp:31-66)."""

    def __init__(self, ctx: Context, rows: int, cols: int, params: HsvParams | None = None):
        self.ctx, self.rows, self.cols = ctx, rows, cols
        self.params = params or HsvParams.make()
        self._h = C.c_void_p()
        _ck(lib().oat_hsvdet_create(ctx._h, rows, cols, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().oat_hsvdet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, fn, img, ch, want_thresh, want_labels):
        _img(img, self.rows, self.cols, ch)
        d = Detection()
        thr = np.empty((self.rows, self.cols), np.uint8) if want_thresh else None
        lab = np.empty((self.rows, self.cols), np.int32) if want_labels else None
        _ck(fn(self._h, _ptr(img), self.cols * ch, C.byref(self.params), C.byref(d), _ptr(thr), self.cols, _ptr(lab)))
        return d, thr, lab

    def detect(self, hsv, want_thresh=False, want_labels=False):
        """-> (Detection, thresh mask or None, labels or None)."""
        return self._run(lib().oat_hsvdet_detect, hsv, 3, want_thresh, want_labels)

    def thresh_detect(self, grey, t_min, t_max, want_thresh=False, want_labels=False):
        """``posidet thresh`` on a GREY frame (src/positiondetector/SimpleThreshold.cpp:169-182)."""
        _img(grey, self.rows, self.cols, 1)
        d = Detection()
        thr = np.empty((self.rows, self.cols), np.uint8) if want_thresh else None
        lab = np.empty((self.rows, self.cols), np.int32) if want_labels else None
        _ck(lib().oat_thresh_detect(self._h, _ptr(grey), self.cols, t_min, t_max, C.byref(self.params), C.byref(d), _ptr(thr),
                                    self.cols, _ptr(lab)))
        return d, thr, lab

    def sift_contours(self, mask, want_thresh=False, want_labels=False):
        """siftContours on a binary mask (morphology per self.params applied first)."""
        return self._run(lib().oat_sift_contours, mask, 1, want_thresh, want_labels)


class DifferenceDetector:
    """``posidet diff`` (src/positiondetector/DifferenceDetector.cpp:118-173): motion by frame differencing."""

    def __init__(self, ctx: Context, rows: int, cols: int, diff_threshold: int = 10, blur: int = 2,
                 area=(0.0, DBL_MAX)):
        self.ctx, self.rows, self.cols = ctx, rows, cols
        self.diff_threshold, self.blur, self.area = diff_threshold, blur, tuple(area)
        self._h = C.c_void_p()
        _ck(lib().oat_diffdet_create(ctx._h, rows, cols, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().oat_diffdet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def detect(self, grey, want_thresh=False):
        _img(grey, self.rows, self.cols, 1)
        d = Detection()
        thr = np.empty((self.rows, self.cols), np.uint8) if want_thresh else None
        _ck(lib().oat_diffdet_detect(self._h, _ptr(grey), self.cols, self.diff_threshold, self.blur, self.area[0], self.area[1],
                                     C.byref(d), _ptr(thr), self.cols))
        return d, thr


class PositionFilter:
    """`oat posifilt kalman` (KalmanFilter2D, src/positionfilter/KalmanFilter2D.cpp:95-200) per source and/or
    `oat posicom mean` (MeanPosition::combine, src/positioncombiner/MeanPosition.cpp:60-118) over the sources,
    one kernel launch per sample; the filter state lives on the device."""

    def __init__(self, ctx: Context, n_sources: int = 1, kalman: KalmanParams | dict | None = None,
                 combine_mean: bool = False, heading_anchor: int = -1):
        self.ctx, self.n, self.combine = ctx, n_sources, combine_mean
        kp = None
        if kalman is not None:
            kp = KalmanParams()
            lib().oat_kalman_default_params(C.byref(kp))
            if isinstance(kalman, dict):
                for k, v in kalman.items():
                    setattr(kp, k, v)
            else:
                kp = kalman
        self._h = C.c_void_p()
        _ck(lib().oat_posfilt_create(ctx._h, n_sources, C.byref(kp) if kp is not None else None, int(combine_mean),
                                     heading_anchor, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().oat_posfilt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _ck(lib().oat_posfilt_reset(self._h))

    def apply(self, sources):
        """sources: n Position structs -> one Position (combine_mean) or a list of n."""
        arr = (Position * self.n)(*sources)
        out = (Position * (1 if self.combine else self.n))()
        _ck(lib().oat_posfilt_apply(self._h, arr, out))
        return out[0] if self.combine else list(out)


class Tracker:
    """mog -> col HSV -> hsv fused on the device; one instance = one video stream."""

    def __init__(self, ctx: Context, rows: int, cols: int, adaptation_coeff: float = 0.0,
                 hsv: HsvParams | None = None, mog_params: MogParams | None = None, ring_depth: int = 0):
        if not (0.0 <= adaptation_coeff <= 1.0):
            raise ValueError("adaptation-coeff must be in [0, 1]")
        self.ctx, self.rows, self.cols = ctx, rows, cols
        self.learning_coeff = adaptation_coeff
        self.hsv = hsv or HsvParams.make()
        self.mog_params = mog_params or default_mog_params()
        self._h = C.c_void_p()
        _ck(lib().oat_tracker_create(ctx._h, rows, cols, C.byref(self.mog_params), ring_depth, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib().oat_tracker_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _ck(lib().oat_tracker_reset(self._h))

    def track(self, bgr, egress=(), learning_rate=None):
        """One frame in (numpy host array, DeviceBuffer or raw pointer), one Detection out.
        egress: any of 'bgr', 'fgmask', 'hsv', 'thresh' -> returned as numpy arrays in a dict."""
        _img(bgr, self.rows, self.cols, 3)
        r, c = self.rows, self.cols
        outs = {
            "bgr": np.empty((r, c, 3), np.uint8) if "bgr" in egress else None,
            "fgmask": np.empty((r, c), np.uint8) if "fgmask" in egress else None,
            "hsv": np.empty((r, c, 3), np.uint8) if "hsv" in egress else None,
            "thresh": np.empty((r, c), np.uint8) if "thresh" in egress else None,
        }
        d = Detection()
        lr = self.learning_coeff if learning_rate is None else learning_rate
        _ck(lib().oat_tracker_track(self._h, _ptr(bgr), c * 3, lr, C.byref(self.hsv), C.byref(d),
                                    _ptr(outs["bgr"]), c * 3, _ptr(outs["fgmask"]), c, _ptr(outs["hsv"]), c * 3,
                                    _ptr(outs["thresh"]), c))
        return d, {k: v for k, v in outs.items() if v is not None}

    def submit(self, bgr, bgr_out=None, learning_rate=None, pitch=None):
        lr = self.learning_coeff if learning_rate is None else learning_rate
        _ck(lib().oat_tracker_submit(self._h, _ptr(bgr), pitch or self.cols * 3, lr, C.byref(self.hsv), _ptr(bgr_out),
                                     self.cols * 3))

    def submit_fused_only(self, bgr, learning_rate=None, pitch=None):
        """Diagnostic: only the fused kernel of the next frame (see oat_tracker_submit_fused_only)."""
        lr = self.learning_coeff if learning_rate is None else learning_rate
        _ck(lib().oat_tracker_submit_fused_only(self._h, _ptr(bgr), pitch or self.cols * 3, lr, C.byref(self.hsv)))

    def wait_ingest(self):
        """Blocks until the last submitted frame's input buffer may be reused (oat_tracker_wait_ingest)."""
        _ck(lib().oat_tracker_wait_ingest(self._h))

    def collect(self) -> Detection:
        d = Detection()
        _ck(lib().oat_tracker_collect(self._h, C.byref(d)))
        return d

    def run_clip(self, frames, depth=4, learning_rate=None, positions=False, pitch=None, out=None):
        """frames: device buffers / arrays of one clip -> list of Detection (and of Position with a filter attached).
        Device-resident frames go through the resident engine (one launch per chunk of frames), anything else
        through the natively looped submit/collect pipeline (oat_tracker_run_clip).
        out: a caller-owned (Detection * n)() the detections are written to and that is returned as it is -- nothing is
        allocated or converted per frame on the Python side (what a measurement loop wants)."""
        lr = self.learning_coeff if learning_rate is None else learning_rate
        ptrs = frames if isinstance(frames, C.Array) else frame_pointers(frames)
        n = len(ptrs)
        if out is not None:
            assert len(out) >= n and not positions
            _ck(lib().oat_tracker_run_clip(self._h, ptrs, n, pitch or self.cols * 3, lr, C.byref(self.hsv), depth, out, None))
            return out
        out = (Detection * n)()
        pos = (Position * n)() if positions else None
        _ck(lib().oat_tracker_run_clip(self._h, ptrs, n, pitch or self.cols * 3, lr, C.byref(self.hsv), depth, out, pos))
        return (list(out), list(pos)) if positions else list(out)

    # ---- streaming use of the resident engine (oat_tracker_stream_*) ----
    def stream_push(self, bgr, copy=False, learning_rate=None, pitch=None):
        """One frame into the stream; copy=True: the frame's memory is the caller's again after stream_wait_ingest()."""
        lr = self.learning_coeff if learning_rate is None else learning_rate
        _ck(lib().oat_tracker_stream_push(self._h, _ptr(bgr), pitch or self.cols * 3, lr, C.byref(self.hsv), 1 if copy else 0))

    def stream_wait_ingest(self):
        _ck(lib().oat_tracker_stream_wait_ingest(self._h))

    def stream_flush(self, block=False):
        _ck(lib().oat_tracker_stream_flush(self._h, 1 if block else 0))

    def stream_poll(self, cap=64, block=False, positions=False):
        """Finished detections in push order (at most cap) -> list of Detection (, list of Position)."""
        out = (Detection * cap)()
        pos = (Position * cap)() if positions else None
        got = C.c_size_t()
        _ck(lib().oat_tracker_stream_poll(self._h, out, pos, cap, 1 if block else 0, C.byref(got)))
        return (list(out[:got.value]), list(pos[:got.value])) if positions else list(out[:got.value])

    def stream_pending(self):
        """(frames gathered, frames in flight on the GPU, detections waiting for stream_poll)."""
        a, b, c_ = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _ck(lib().oat_tracker_stream_pending(self._h, C.byref(a), C.byref(b), C.byref(c_)))
        return a.value, b.value, c_.value

    def attach_posfilt(self, f: "PositionFilter | None"):
        """Fuse a single-source position filter behind this tracker (device-side epilogue, frame order)."""
        _ck(lib().oat_tracker_attach_posfilt(self._h, f._h if f is not None else None))
        self._pf = f

    def collect_position(self):
        d, p = Detection(), Position()
        _ck(lib().oat_tracker_collect_position(self._h, C.byref(d), C.byref(p)))
        return d, p

    def live_modes(self) -> int:
        s = C.c_uint64()
        _ck(lib().oat_tracker_live_modes(self._h, C.byref(s)))
        return s.value

    def state(self):
        K = self.mog_params.nmixtures
        m, w, v, mu = _state_arrays(self.rows, self.cols, K)
        _ck(lib().oat_tracker_get_state(self._h, _ptr(m), _ptr(w), _ptr(v), _ptr(mu)))
        return m, w, v, mu

    def tail_stats(self):
        """Diagnostics of the detect tail of the last collected frame (see oat_tracker_tail_stats)."""
        a = (C.c_uint32 * 15)()
        _ck(lib().oat_tracker_tail_stats(self._h, a))
        v = list(a)
        return {"status": v[0], "nodes": v[1], "replays": v[2], "fast": v[3] & 1, "prelabelled": (v[3] >> 1) & 1, "cyc": v[4:12],
                "generic_frames": v[12], "slow_groups": v[13], "clip_frames": v[14]}

    @staticmethod
    def run_clips(trackers, frames, learning_rate=None, fused_only=False, pitch=None):
        """Several independent streams of one GPU through ONE queue of the resident engine (oat_tracker_run_clips).
        frames: list over time of lists over trackers (device buffers), or a prebuilt pointer array of
        len(trackers) * n_frames entries, frame-major.  Returns detections[i][s] (None with fused_only)."""
        S = len(trackers)
        t0 = trackers[0]
        lr = t0.learning_coeff if learning_rate is None else learning_rate
        ptrs = frames if isinstance(frames, C.Array) else frame_pointers([f for row in frames for f in row])
        n = len(ptrs) // S
        handles = (C.c_void_p * S)(*[t._h.value for t in trackers])
        out = None if fused_only else (Detection * (n * S))()
        _ck(lib().oat_tracker_run_clips(handles, S, ptrs, n, pitch or t0.cols * 3, lr, C.byref(t0.hsv), 1 if fused_only else 0, out))
        if fused_only:
            return None
        return [[out[i * S + s] for s in range(S)] for i in range(n)]

    def profile(self, enable: bool):
        _ck(lib().oat_tracker_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        ms, n = C.c_double(), C.c_uint64()
        _ck(lib().oat_tracker_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value
