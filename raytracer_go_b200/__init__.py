"""raytracer_go_b200 — host-side Python view of librt_b200.so, the B200 (sm_100a) CUDA
implementation of raytracer-go's per-pixel / per-sample path-tracing loop.

The product is the C-ABI shared library (include/rt_b200.h, csrc/); this package only loads it
(`lib`), mirrors the reference's scene-description API on top of it (`api`) and synthesises the
deterministic benchmark scenes (`scenes`).  There is no CPU rendering path here: every compute
entry point fails loudly when the library or a B200 is missing.
"""
from . import abi  # noqa: F401

__all__ = ["abi", "api", "lib", "scenes", "philox"]
