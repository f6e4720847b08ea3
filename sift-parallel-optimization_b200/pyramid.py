"""Host-side mirror of the reference's operator interface over the C ABI.

Two layers:

* :class:`ScaleSpace` -- the superset handle (rectangular frames, chosen octave count, sigma0, REF/CONV
  mode, frame slots, row bands, device-resident results).  A thin, explicit wrapper of include/sspyr.h.
* :class:`GaussPyramid` -- the reference's class surface, member for member
  (``GaussPyramid(int** img, int len, int S)``, ``GaussPyInit()``, ``GaussFilter(theLayer)``,
  ``GenerateDoG()``, ``output()``, public ``data`` / ``GaussPy`` / ``initialized``; GuassDePyramid.h:11-29),
  so the parity tests read like code written against the reference.  The C++ twin for ``main.cpp`` is
  include/GaussDePyramid-CUDA.h.

All arithmetic happens in libsspyr.so's CUDA kernels; numpy is used here only to hold host buffers.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib as L

_PIX = {np.dtype(np.int32): L.PIXEL_I32, np.dtype(np.float32): L.PIXEL_F32, np.dtype(np.uint8): L.PIXEL_U8}


class ScaleSpace:
    """One sspyr handle: one GPU, `frames` resident frame slots of an H x W (band of a) frame."""

    def __init__(self, height: int, width: int, octaves: int = 0, S: int = 3, sigma0: float = 0.0,
                 mode: int = L.MODE_REF, outputs: int = L.OUT_ALL, pixel_type: int = L.PIXEL_I32,
                 frames: int = 1, device: int = -1, band_row0: int = 0, full_height: int = 0,
                 sigma_in: float = 0.5, radius_sigmas: float = 3.0, extrema_thresh: float = 0.0, max_keypoints: int = 0):
        self._lib = L.load()
        cfg = L.Config()
        L.check(None, self._lib.sspyr_default_config(C.byref(cfg)))
        cfg.height, cfg.width, cfg.octaves, cfg.S = height, width, octaves, S
        cfg.sigma0, cfg.mode, cfg.outputs, cfg.pixel_type = sigma0, mode, outputs, pixel_type
        cfg.frames, cfg.device, cfg.band_row0, cfg.full_height = frames, device, band_row0, full_height
        cfg.sigma_in, cfg.radius_sigmas, cfg.extrema_thresh = sigma_in, radius_sigmas, extrema_thresh
        cfg.max_keypoints = max_keypoints
        self._h = C.c_void_p()
        L.check(None, self._lib.sspyr_create(C.byref(cfg), C.byref(self._h)))
        self.height, self.width, self.S, self.frames = height, width, S, max(frames, 1)
        self.mode, self.outputs, self.pixel_type = mode, outputs or L.OUT_ALL, pixel_type
        self.max_keypoints = max_keypoints or (1 << 20)
        self.octaves = self._lib.sspyr_num_octaves(self._h)
        self.levels = self._lib.sspyr_num_levels(self._h)
        self.dogs = self._lib.sspyr_num_dogs(self._h)
        self._keep = {}            # host/device arrays that must outlive async copies

    # ---- life cycle -----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.sspyr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc: int) -> int:
        return L.check(self._h, rc)

    # ---- geometry ---------------------------------------------------------------------------------
    def level_dims(self, octave: int) -> tuple[int, int, int]:
        r, c, p = C.c_int(), C.c_int(), C.c_size_t()
        self._ck(self._lib.sspyr_level_dims(self._h, octave, C.byref(r), C.byref(c), C.byref(p)))
        return r.value, c.value, p.value

    def algorithmic_bytes(self) -> int:
        b = C.c_uint64()
        self._ck(self._lib.sspyr_algorithmic_bytes(self._h, C.byref(b)))
        return b.value

    def plane_pixels(self) -> int:
        return sum(self.level_dims(o)[0] * self.level_dims(o)[1] for o in range(self.octaves))

    # ---- streams / tuning ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int) -> None:
        self._ck(self._lib.sspyr_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_tuning(self, **kw: int) -> None:
        for k, v in kw.items():
            self._ck(self._lib.sspyr_set_tuning(self._h, k.encode(), int(v)))

    # ---- input ----------------------------------------------------------------------------------
    def upload(self, img: np.ndarray, frame: int = 0) -> None:
        """Host frame -> slot (async if `img` is pinned; the array is kept alive until the next sync)."""
        if _PIX.get(img.dtype) != self.pixel_type or img.shape != (self.height, self.width):
            raise ValueError(f"expected a {self.height}x{self.width} frame of pixel type {self.pixel_type}")
        if img.strides[1] != img.itemsize:
            img = np.ascontiguousarray(img)
        self._keep[("in", frame)] = img
        self._ck(self._lib.sspyr_upload(self._h, frame, C.c_void_p(img.ctypes.data), img.strides[0]))

    def upload_ptr(self, host_ptr: int, pitch_bytes: int = 0, frame: int = 0) -> None:
        self._ck(self._lib.sspyr_upload(self._h, frame, C.c_void_p(host_ptr), pitch_bytes))

    def set_input_device(self, dev_ptr: int, pitch_bytes: int = 0, frame: int = 0) -> None:
        self._ck(self._lib.sspyr_set_input_device(self._h, frame, C.c_void_p(dev_ptr), pitch_bytes))

    # ---- the hot path -------------------------------------------------------------------------------
    def build(self, frame: int = 0, stage: int = L.STAGE_DOG) -> None:
        if stage == L.STAGE_DOG:
            self._ck(self._lib.sspyr_build(self._h, frame))
        else:
            self._ck(self._lib.sspyr_build_stage(self._h, frame, stage))

    def build_batch(self, first: int, count: int) -> None:
        self._ck(self._lib.sspyr_build_batch(self._h, first, count))

    def sync(self) -> None:
        self._ck(self._lib.sspyr_sync(self._h))
        self._keep = {k: v for k, v in self._keep.items() if k[0] == "dev"}

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        self._ck(self._lib.sspyr_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def last_launches(self) -> int:
        return self._lib.sspyr_last_launches(self._h)

    # ---- results ----------------------------------------------------------------------------------
    def download(self, octave: int, level: int, kind: int = L.KIND_GAUSS, frame: int = 0) -> np.ndarray:
        r, c, _ = self.level_dims(octave)
        out = np.empty((r, c), dtype=np.uint8 if kind == L.KIND_EXTREMA else np.float32)
        self._ck(self._lib.sspyr_download(self._h, frame, octave, level, kind, C.c_void_p(out.ctypes.data), 0))
        return out

    def _split(self, flat: np.ndarray, planes: int) -> list[np.ndarray]:
        out, off = [], 0
        for o in range(self.octaves):
            r, c, _ = self.level_dims(o)
            out.append(flat[off:off + planes * r * c].reshape(planes, r, c))
            off += planes * r * c
        return out

    def download_inplace(self, frame: int = 0, out: np.ndarray | None = None) -> list[np.ndarray]:
        """The reference's in-place result: per octave [S+3, H_o, W_o] = DoG_0..DoG_{S+1}, G_{S+2}."""
        if out is None:
            out = np.empty(self.plane_pixels() * self.levels, dtype=np.float32)
        self._ck(self._lib.sspyr_download_inplace(self._h, frame, C.c_void_p(out.ctypes.data)))
        self.sync()
        return self._split(out, self.levels)

    def download_gauss(self, frame: int = 0, out: np.ndarray | None = None) -> list[np.ndarray]:
        if out is None:
            out = np.empty(self.plane_pixels() * self.levels, dtype=np.float32)
        self._ck(self._lib.sspyr_download_gauss(self._h, frame, C.c_void_p(out.ctypes.data)))
        self.sync()
        return self._split(out, self.levels)

    def download_dog(self, frame: int = 0) -> list[np.ndarray]:
        return [np.stack([self.download(o, s, L.KIND_DOG, frame) for s in range(self.dogs)])
                for o in range(self.octaves)]

    def download_keypoints(self, frame: int = 0, capacity: int = 1 << 20) -> tuple[np.ndarray, int]:
        """(records [n, 4] int32: x, y, octave << 16 | level, value bits; number of extrema found).  Synchronises."""
        rec = np.empty((capacity, 4), dtype=np.int32)
        cnt = np.zeros(1, dtype=np.int32)
        self._ck(self._lib.sspyr_download_keypoints(self._h, frame, C.c_void_p(rec.ctypes.data), capacity, C.c_void_p(cnt.ctypes.data)))
        self.sync()
        n = int(cnt[0])
        return rec[:min(n, capacity, self.max_keypoints)], n

    def download_keypoints_ptr(self, rec_ptr: int, capacity: int, count_ptr: int, frame: int = 0) -> None:
        """Async variant into caller-owned (pinned) memory; call sync() before reading."""
        self._ck(self._lib.sspyr_download_keypoints(self._h, frame, C.c_void_p(rec_ptr), capacity, C.c_void_p(count_ptr)))

    def download_inplace_ptr(self, host_ptr: int, frame: int = 0) -> None:
        """Async variant into caller-owned (pinned) memory; call sync() before reading."""
        self._ck(self._lib.sspyr_download_inplace(self._h, frame, C.c_void_p(host_ptr)))

    def device_ptr(self, octave: int, level: int, kind: int = L.KIND_GAUSS, frame: int = 0) -> int:
        p = C.c_void_p()
        self._ck(self._lib.sspyr_device_ptr(self._h, frame, octave, level, kind, C.byref(p)))
        return p.value

    def window_table(self, octave: int, level: int, axis: int) -> np.ndarray:
        r, c, _ = self.level_dims(octave)
        out = np.empty(r if axis == 0 else c, dtype=np.float32)
        self._ck(self._lib.sspyr_window_table(self._h, octave, level, axis, C.c_void_p(out.ctypes.data), out.size))
        return out

    # ---- CONV mode, level by level (row bands: exchange halos between the steps) ----------------------
    def halo_rows(self, octave: int, level: int) -> int:
        r = C.c_int()
        self._ck(self._lib.sspyr_halo_rows(self._h, octave, level, C.byref(r)))
        return r.value

    def halo_ptrs(self, octave: int, level: int, frame: int = 0) -> tuple[int, int, int, int, int]:
        """(send_up, send_down, recv_up, recv_down, nbytes) device pointers for the halo of (octave, level)."""
        p = [C.c_void_p() for _ in range(4)]
        n = C.c_size_t()
        self._ck(self._lib.sspyr_halo_ptrs(self._h, frame, octave, level, *[C.byref(x) for x in p], C.byref(n)))
        return p[0].value, p[1].value, p[2].value, p[3].value, n.value

    def conv_step(self, octave: int, level: int, frame: int = 0) -> None:
        self._ck(self._lib.sspyr_conv_step(self._h, frame, octave, level))

    # ---- CONV row bands over NVLink peer memory (halo rows read inside the blur kernel) -----------------
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(L.IPC_BLOB_BYTES)
        n = C.c_size_t()
        self._ck(self._lib.sspyr_ipc_export(self._h, buf, L.IPC_BLOB_BYTES, C.byref(n)))
        return buf.raw[:n.value]

    def ipc_attach(self, side: int, blob: bytes) -> None:
        self._ck(self._lib.sspyr_ipc_attach(self._h, side, blob, len(blob)))

    def peer_attach_local(self, side: int, neighbour: "ScaleSpace") -> None:
        self._ck(self._lib.sspyr_peer_attach_local(self._h, side, neighbour._h))

    def conv_taps(self, level: int) -> np.ndarray:
        buf = np.empty(2 * 64 + 1, dtype=np.float32)
        R = C.c_int()
        self._ck(self._lib.sspyr_conv_taps(self._h, level, C.c_void_p(buf.ctypes.data), buf.size, C.byref(R)))
        return buf[:2 * R.value + 1].copy()


class GaussPyramid:
    """The reference's ``class GaussPyramid`` (GuassDePyramid.h:11-29) backed by the B200 kernels.

    Same constructor, methods and public members; ``GaussPy[o][s][r][c]`` is a host mirror refreshed by
    each call.  One documented difference: every call recomputes from ``data`` (the reference multiplies
    the stored levels again when ``GaussFilter``/``GenerateDoG`` are called twice without
    ``GaussPyInit()``, an artefact of its in-place storage -- main.cpp:66-73 -- that is not reproduced).
    """

    def __init__(self, img=None, len: int = 0, S: int = 0, device: int = -1):
        # GaussPyramid::GaussPyramid()  GuassDePyramid.h:31-34
        self.data = None
        self.initialized = False
        self.GaussPy = None
        self._ss = None
        if img is None:
            return
        # GaussPyramid::GaussPyramid(int** img, int len, int S)  GuassDePyramid.h:36-58
        img = np.asarray(img)
        if img.ndim != 2 or img.shape[0] < len or img.shape[1] < len or len < 1:
            raise ValueError("img must hold at least len x len pixels")
        self.length = int(len)
        self.S = int(S)
        self.data = np.array(img[:len, :len], dtype=np.int32, order="C")          # deep copy, :38-46
        self._ss = ScaleSpace(self.length, self.length, 0, self.S, outputs=L.OUT_ALL, device=device)
        self._ss.set_tuning(timing=1)
        self.layer = self._ss.octaves                                              # :48-53
        self._ss.upload(self.data)
        self.GaussPyInit()                                                         # :57

    def GaussPyInit(self) -> None:
        """GuassDePyramid.h:60-87 -- every level of every octave := decimated original (K0, on the GPU)."""
        self._ss.upload(self.data)      # `data` is public and may have been edited by the caller
        self._ss.build(stage=L.STAGE_INIT)
        self._filtered = False
        self.GaussPy = [list(a) for a in self._ss.download_gauss()]
        self.initialized = True

    def GaussFilter(self, theLayer: int) -> None:
        """GuassDePyramid.h:106-134 -- window-multiply all S+3 levels of octave `theLayer`."""
        if not 0 <= theLayer < self.layer:
            raise IndexError("theLayer out of range")
        if not self._filtered:          # one fused pass filters every octave; `for o: GaussFilter(o)` reuses it
            self._ss.build(stage=L.STAGE_FILTER)
            self._filtered = True
        self.GaussPy[theLayer] = [self._ss.download(theLayer, s, L.KIND_GAUSS) for s in range(self.S + 3)]

    def GenerateDoG(self) -> None:
        """GuassDePyramid.h:136-149 -- slots 0..S+1 become DoG_s = G_s - G_{s+1}; slot S+2 keeps G_{S+2}."""
        self._ss.build(stage=L.STAGE_DOG)
        self._filtered = True           # (a full build leaves every Gaussian level in place too)
        self.GaussPy = [list(a) for a in self._ss.download_inplace()]

    def output(self, file=None) -> None:
        """GuassDePyramid.h:89-104 -- print level 0 of every octave, '==' rulers between octaves."""
        file = file or sys.stdout
        ln = self.length
        for i in range(self.layer):
            for j in range(ln):
                file.write(" ".join(f"{v:g}" for v in self.GaussPy[i][0][j][:ln]) + " \n")
            file.write("==" * ln + "\n")
            ln //= 2

    def elapsed_ms(self) -> float:
        return self._ss.elapsed_ms()

    def close(self) -> None:
        if self._ss is not None:
            self._ss.close()
            self._ss = None
