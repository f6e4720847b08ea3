"""sift-parallel-optimization_b200 -- B200-native SIFT scale-space (Gaussian + DoG pyramid) builder.

The directory name carries a hyphen (it mirrors the upstream repository name), so import it through
``__graft_entry__.load_package()`` (which registers it as ``sift_parallel_optimization_b200``) or put the
repo root on ``sys.path`` and use ``importlib``.  The compute lives in ``libsspyr.so`` (csrc/, sm_100a
CUDA, C ABI in include/sspyr.h); there is no CPU implementation in this package.
"""
from . import _lib, synth                                    # noqa: F401
from ._lib import (KIND_DOG, KIND_EXTREMA, KIND_GAUSS, KIND_INPLACE, KIND_KEYPOINTS, MODE_CONV, MODE_REF,    # noqa: F401
                   OUT_ALL, OUT_DOG, OUT_EXTREMA, OUT_GAUSS, OUT_GAUSS_TOP, OUT_INPLACE, OUT_KEYPOINTS, PIXEL_F32,
                   PIXEL_I32, PIXEL_U8, STAGE_DOG, STAGE_FILTER, STAGE_INIT, SspyrError)
from .exchange import DistExchanger, LocalExchanger, LocalPeerLink, PeerExchanger   # noqa: F401
from .partition import band_rows, shard_frames               # noqa: F401
from .pyramid import GaussPyramid, ScaleSpace                # noqa: F401

__all__ = ["GaussPyramid", "ScaleSpace", "SspyrError", "synth", "band_rows", "shard_frames", "LocalExchanger",
           "DistExchanger", "LocalPeerLink", "PeerExchanger"]
