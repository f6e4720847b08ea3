"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers:

* ``liborc.so``              our plain-C restatement (oracle/sspyr_oracle.c), any H x W / octaves / S / sigma0
* ``_ref/libsiftref*.so``    the UNMODIFIED reference headers compiled where they lie (oracle/ref_wrap*.cpp):
                             square images, all octaves, sigma = 2.0 only

Only tests/, oracle/make_golden.py, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of bench.py import this module.  The product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


# --------------------------------------------------------------------------------------------------
# dense layout bookkeeping (mirrors the comment at the top of sspyr_oracle.c)
# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Geometry:
    h: int
    w: int
    octaves: int
    S: int

    @property
    def levels(self) -> int:
        return self.S + 3

    def dims(self, o: int) -> tuple[int, int]:
        return self.h >> o, self.w >> o

    def plane_pixels(self) -> int:
        return sum((self.h >> o) * (self.w >> o) for o in range(self.octaves))

    def split(self, flat: np.ndarray, planes_per_octave: int) -> list[np.ndarray]:
        """flat dense array -> list over octaves of [planes, H_o, W_o] views."""
        out, off = [], 0
        for o in range(self.octaves):
            ho, wo = self.dims(o)
            n = planes_per_octave * ho * wo
            out.append(flat[off:off + n].reshape(planes_per_octave, ho, wo))
            off += n
        assert off == flat.size, (off, flat.size)
        return out


def octaves_all(h: int, w: int) -> int:
    """floor(log2(min(h, w))) + 1 -- GuassDePyramid.h:48-53 applied to the short side."""
    return int(min(h, w)).bit_length()


# --------------------------------------------------------------------------------------------------
# loaders
# --------------------------------------------------------------------------------------------------
_port = None
_ref = None
_ref512 = None


def port_path() -> str:
    return os.path.join(_HERE, "liborc.so")


def ref_path() -> str:
    return os.path.join(_HERE, "_ref", "libsiftref.so")


def ref512_path() -> str:
    return os.path.join(_HERE, "_ref", "libsiftref_avx512.so")


def load_port():
    global _port
    if _port is None:
        if not os.path.exists(port_path()):
            raise RuntimeError("oracle/liborc.so missing: run `make -C oracle port` (or __graft_entry__.build())")
        L = C.CDLL(port_path())
        L.orc_sigma_ref.restype = C.c_float
        L.orc_pi_ref.restype = C.c_float
        L.orc_octaves_all.argtypes = [C.c_int, C.c_int]
        L.orc_window.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, _f32p]
        L.orc_ref_mirror.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_ref_build.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_ref_sweeps_mt.argtypes = [_i32p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                        _f32p, C.c_int, C.c_int]
        L.orc_conv_sigma_inc.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_conv_sigma_inc.restype = C.c_double
        L.orc_conv_radius.argtypes = [C.c_double, C.c_float]
        L.orc_conv_taps.argtypes = [C.c_double, C.c_float, _f32p]
        L.orc_conv_build.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.c_float, C.c_float, _f32p, C.c_void_p, C.c_int]
        L.orc_extrema_octave.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_float, _u8p]
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_fnv1a64.restype = C.c_uint64
        _port = L
    return _port


def have_ref() -> bool:
    return os.path.exists(ref_path())


def load_ref():
    """The real reference header, compiled (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None and have_ref():
        L = C.CDLL(ref_path())
        L.sref_total_floats.argtypes = [C.c_int, C.c_int]
        L.sref_total_floats.restype = C.c_longlong
        for name in ("sref_serial_dog", "sref_serial_gauss", "sref_serial_init"):
            f = getattr(L, name)
            f.argtypes = [_i32p, C.c_int, C.c_int, _f32p]
            f.restype = C.c_longlong
        for name in ("sref_pthread_i_dog", "sref_omp_dog"):
            f = getattr(L, name)
            f.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, _f32p]
            f.restype = C.c_longlong
        L.sref_time_serial.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        L.sref_time_pthread_i.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        L.sref_time_omp.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        _ref = L
    return _ref


def cpu_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False


def load_ref_avx512():
    global _ref512
    if _ref512 is None and os.path.exists(ref512_path()) and cpu_has_avx512():
        L = C.CDLL(ref512_path())
        L.sref_a512xp_dog.argtypes = [_i32p, C.c_int, C.c_int, _f32p]
        L.sref_a512xp_dog.restype = C.c_longlong
        L.sref_time_a512xp.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        L.sref_time_a512omp.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        _ref512 = L
    return _ref512


# --------------------------------------------------------------------------------------------------
# the real header (square, all octaves, sigma = 2)
# --------------------------------------------------------------------------------------------------
def _sq(img: np.ndarray) -> tuple[np.ndarray, int]:
    img = np.ascontiguousarray(img, dtype=np.int32)
    assert img.ndim == 2 and img.shape[0] == img.shape[1], "the reference header is square-only"
    return img, img.shape[0]


def header_run(img: np.ndarray, S: int, what: str = "dog", threads: int = 0) -> list[np.ndarray]:
    """Run the unmodified header.  what: 'dog' (GenerateDoG, in-place layout), 'gauss' (GaussFilter on
    every octave only), 'init' (GaussPyInit only), 'pthread_i', 'omp', 'a512xp' (bit-exact parallel
    variants, in-place layout).  Returns a list over octaves of [S+3, len_o, len_o] arrays."""
    img, n = _sq(img)
    R = load_ref()
    if R is None:
        raise RuntimeError("oracle/_ref/libsiftref.so missing (built only where /root/reference exists)")
    geo = Geometry(n, n, octaves_all(n, n), S)
    out = np.empty(R.sref_total_floats(n, S), dtype=np.float32)
    if what in ("dog", "gauss", "init"):
        got = getattr(R, f"sref_serial_{what}")(img, n, S, out)
    elif what == "pthread_i":
        got = R.sref_pthread_i_dog(img, n, S, threads, out)
    elif what == "omp":
        got = R.sref_omp_dog(img, n, S, threads, out)
    elif what == "a512xp":
        R5 = load_ref_avx512()
        if R5 is None:
            raise RuntimeError("AVX-512 reference build unavailable on this host")
        got = R5.sref_a512xp_dog(img, n, S, out)
    else:
        raise ValueError(what)
    assert got == out.size
    return geo.split(out, S + 3)


# --------------------------------------------------------------------------------------------------
# our restatement
# --------------------------------------------------------------------------------------------------
def _img_args(img: np.ndarray):
    img = np.ascontiguousarray(img)
    if img.dtype == np.float32:
        return img, None, img.ctypes.data
    img = np.ascontiguousarray(img, dtype=np.int32)
    return img, img.ctypes.data, None


def window(axis_len: int, o: int, s: int, sigma0: float = 2.0) -> np.ndarray:
    f = np.empty(max(axis_len >> o, 1), dtype=np.float32)
    n = load_port().orc_window(axis_len, o, s, sigma0, f)
    return f[:n]


def ref_mirror(img: np.ndarray, octaves: int | None = None, S: int = 3, sigma0: float = 2.0,
               do_dog: bool = True, row0: int = 0, full_h: int = 0) -> list[np.ndarray]:
    """Line-by-line mirror of the serial header -> in-place layout (or gauss layout if not do_dog)."""
    keep, pi, pf = _img_args(img)
    h, w = keep.shape
    octaves = octaves or octaves_all(full_h or h, w)
    geo = Geometry(h, w, octaves, S)
    out = np.empty(geo.plane_pixels() * (S + 3), dtype=np.float32)
    rc = load_port().orc_ref_mirror(pi, pf, w, h, w, octaves, S, sigma0, row0, full_h, int(do_dog), out)
    assert rc == 0, rc
    return geo.split(out, S + 3)


def ref_build(img: np.ndarray, octaves: int | None = None, S: int = 3, sigma0: float = 2.0,
              row0: int = 0, full_h: int = 0, threads: int = 0, want=("gauss", "dog", "inplace")) -> dict:
    """Closed-form restatement -> {'gauss': [...], 'dog': [...], 'inplace': [...]} lists over octaves."""
    keep, pi, pf = _img_args(img)
    h, w = keep.shape
    octaves = octaves or octaves_all(full_h or h, w)
    geo = Geometry(h, w, octaves, S)
    px = geo.plane_pixels()
    bufs = {
        "gauss": np.empty(px * (S + 3), dtype=np.float32) if "gauss" in want else None,
        "dog": np.empty(px * (S + 2), dtype=np.float32) if "dog" in want else None,
        "inplace": np.empty(px * (S + 3), dtype=np.float32) if "inplace" in want else None,
    }
    ptr = lambda a: None if a is None else a.ctypes.data
    rc = load_port().orc_ref_build(pi, pf, w, h, w, octaves, S, sigma0, row0, full_h,
                                   ptr(bufs["gauss"]), ptr(bufs["dog"]), ptr(bufs["inplace"]), threads)
    assert rc == 0, rc
    planes = {"gauss": S + 3, "dog": S + 2, "inplace": S + 3}
    return {k: geo.split(v, planes[k]) for k, v in bufs.items() if v is not None}


def conv_taps(s: int, S: int, sigma0: float, sigma_in: float, radius_sigmas: float) -> np.ndarray:
    L = load_port()
    si = L.orc_conv_sigma_inc(s, S, sigma0, sigma_in)
    R = L.orc_conv_radius(si, radius_sigmas)
    t = np.empty(2 * R + 1, dtype=np.float32)
    assert L.orc_conv_taps(si, radius_sigmas, t) == R
    return t


def conv_build(img: np.ndarray, octaves: int, S: int = 3, sigma0: float = 1.6, sigma_in: float = 0.5,
               radius_sigmas: float = 3.0, threads: int = 0, want_dog: bool = True) -> dict:
    keep, pi, pf = _img_args(img)
    h, w = keep.shape
    geo = Geometry(h, w, octaves, S)
    px = geo.plane_pixels()
    g = np.empty(px * (S + 3), dtype=np.float32)
    d = np.empty(px * (S + 2), dtype=np.float32) if want_dog else None
    rc = load_port().orc_conv_build(pi, pf, w, h, w, octaves, S, sigma0, sigma_in, radius_sigmas, g,
                                    None if d is None else d.ctypes.data, threads)
    assert rc == 0, rc
    out = {"gauss": geo.split(g, S + 3)}
    if d is not None:
        out["dog"] = geo.split(d, S + 2)
    return out


def extrema_octave(dog: np.ndarray, thresh: float) -> np.ndarray:
    """dog: [S+2, H, W] float32 -> flags [S, H, W] uint8."""
    dog = np.ascontiguousarray(dog, dtype=np.float32)
    S = dog.shape[0] - 2
    flags = np.empty((S,) + dog.shape[1:], dtype=np.uint8)
    load_port().orc_extrema_octave(dog.reshape(-1), S, dog.shape[1], dog.shape[2], thresh, flags.reshape(-1))
    return flags


def fnv1a64(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return f"{load_port().orc_fnv1a64(a.ctypes.data, a.nbytes):016x}"


# --------------------------------------------------------------------------------------------------
# CPU timing legs (bench.py cpu_baseline / --impl reference)
# --------------------------------------------------------------------------------------------------
def time_header_serial(img: np.ndarray, S: int, warm: int, reps: int, include_init: bool = False) -> np.ndarray:
    img, n = _sq(img)
    ms = np.zeros(reps, dtype=np.float64)
    load_ref().sref_time_serial(img, n, S, warm, reps, int(include_init), ms)
    return ms


def time_header_variant(img: np.ndarray, S: int, variant: str, threads: int, warm: int, reps: int) -> np.ndarray:
    img, n = _sq(img)
    ms = np.zeros(reps, dtype=np.float64)
    if variant == "pthread_i":
        load_ref().sref_time_pthread_i(img, n, S, threads, warm, reps, 0, ms)
    elif variant == "omp":
        load_ref().sref_time_omp(img, n, S, threads, warm, reps, 0, ms)
    elif variant == "a512xp":
        load_ref_avx512().sref_time_a512xp(img, n, S, warm, reps, ms)
    elif variant == "a512omp":
        load_ref_avx512().sref_time_a512omp(img, n, S, threads, warm, reps, ms)
    else:
        raise ValueError(variant)
    return ms
