"""ctypes binding of libsspyr.so -- the C ABI declared in include/sspyr.h.

There is no Python or CPU implementation behind this module: if the CUDA library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C sift-parallel-optimization_b200/csrc``)
importing it raises, and on a machine without a GPU ``sspyr_create`` fails with SSPYR_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsspyr.so")

# ---- constants (include/sspyr.h) -------------------------------------------------------------------
OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
MODE_REF, MODE_CONV = 0, 1
OUT_GAUSS, OUT_DOG, OUT_GAUSS_TOP, OUT_EXTREMA, OUT_KEYPOINTS = 1, 2, 4, 8, 16
OUT_INPLACE = OUT_DOG | OUT_GAUSS_TOP
OUT_ALL = OUT_GAUSS | OUT_DOG
PIXEL_I32, PIXEL_F32, PIXEL_U8 = 0, 1, 2
KIND_GAUSS, KIND_DOG, KIND_INPLACE, KIND_EXTREMA, KIND_KEYPOINTS = 0, 1, 2, 3, 4
STAGE_INIT, STAGE_FILTER, STAGE_DOG = 0, 1, 2
MAX_OCTAVES, MAX_LEVELS = 16, 16

# every symbol include/sspyr.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "sspyr_version", "sspyr_default_config", "sspyr_create", "sspyr_destroy", "sspyr_last_error",
    "sspyr_num_octaves", "sspyr_num_levels", "sspyr_num_dogs", "sspyr_level_dims", "sspyr_algorithmic_bytes",
    "sspyr_set_stream", "sspyr_upload", "sspyr_set_input_device", "sspyr_build", "sspyr_build_stage",
    "sspyr_build_batch", "sspyr_sync", "sspyr_elapsed_ms", "sspyr_last_launches", "sspyr_download",
    "sspyr_download_inplace", "sspyr_download_gauss", "sspyr_download_keypoints", "sspyr_device_ptr", "sspyr_window_table",
    "sspyr_host_alloc", "sspyr_host_free", "sspyr_conv_taps", "sspyr_set_tuning", "sspyr_halo_rows", "sspyr_halo_ptrs", "sspyr_conv_step",
    "sspyr_ipc_export", "sspyr_ipc_attach", "sspyr_peer_attach_local",
)
IPC_BLOB_BYTES = 1024
SIDE_ABOVE, SIDE_BELOW = 0, 1


class Config(C.Structure):
    """struct sspyr_config, field for field."""
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32), ("octaves", C.c_int32), ("S", C.c_int32),
        ("sigma0", C.c_float), ("mode", C.c_int32), ("outputs", C.c_int32), ("pixel_type", C.c_int32),
        ("frames", C.c_int32), ("device", C.c_int32), ("band_row0", C.c_int32), ("full_height", C.c_int32),
        ("sigma_in", C.c_float), ("radius_sigmas", C.c_float), ("extrema_thresh", C.c_float),
        ("max_keypoints", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class SspyrError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"sspyr error {code}: {text}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libsspyr.so (once).  Raises if it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library is not built (run __graft_entry__.build()); "
            "this package has no CPU implementation")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    pi = C.POINTER(C.c_int)
    sig = {
        "sspyr_version": ([], i),
        "sspyr_default_config": ([C.POINTER(Config)], i),
        "sspyr_create": ([C.POINTER(Config), C.POINTER(H)], i),
        "sspyr_destroy": ([H], i),
        "sspyr_last_error": ([H], C.c_char_p),
        "sspyr_num_octaves": ([H], i),
        "sspyr_num_levels": ([H], i),
        "sspyr_num_dogs": ([H], i),
        "sspyr_level_dims": ([H, i, pi, pi, C.POINTER(sz)], i),
        "sspyr_algorithmic_bytes": ([H, C.POINTER(C.c_uint64)], i),
        "sspyr_set_stream": ([H, vp], i),
        "sspyr_upload": ([H, i, vp, sz], i),
        "sspyr_set_input_device": ([H, i, vp, sz], i),
        "sspyr_build": ([H, i], i),
        "sspyr_build_stage": ([H, i, i], i),
        "sspyr_build_batch": ([H, i, i], i),
        "sspyr_sync": ([H], i),
        "sspyr_elapsed_ms": ([H, C.POINTER(C.c_float)], i),
        "sspyr_last_launches": ([H], i),
        "sspyr_download": ([H, i, i, i, i, vp, sz], i),
        "sspyr_download_inplace": ([H, i, vp], i),
        "sspyr_download_gauss": ([H, i, vp], i),
        "sspyr_download_keypoints": ([H, i, vp, i, vp], i),
        "sspyr_device_ptr": ([H, i, i, i, i, C.POINTER(vp)], i),
        "sspyr_host_alloc": ([sz, C.POINTER(vp)], i),
        "sspyr_host_free": ([vp], i),
        "sspyr_window_table": ([H, i, i, i, vp, i], i),
        "sspyr_conv_taps": ([H, i, vp, i, pi], i),
        "sspyr_set_tuning": ([H, C.c_char_p, i], i),
        "sspyr_halo_rows": ([H, i, i, pi], i),
        "sspyr_halo_ptrs": ([H, i, i, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                             C.POINTER(sz)], i),
        "sspyr_conv_step": ([H, i, i, i], i),
        "sspyr_ipc_export": ([H, vp, sz, C.POINTER(sz)], i),
        "sspyr_ipc_attach": ([H, i, vp, sz], i),
        "sspyr_peer_attach_local": ([H, i, H], i),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (argtypes, restype) in sig.items():
        f = getattr(L, name)      # AttributeError here = the library does not export a declared symbol
        f.argtypes, f.restype = argtypes, restype
    _lib = L
    return L


def check(handle, rc: int) -> int:
    if rc < 0:
        text = load().sspyr_last_error(handle)
        raise SspyrError(rc, text.decode() if text else "")
    return rc
