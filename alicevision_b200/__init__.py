"""alicevision_b200 — Blackwell (sm_100a) engine for AliceVision's descriptor-matching hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of include/b200match.h),
``adaptor/`` (C++ classes deriving from the reference's ArrayMatcher / IImageCollectionMatcher),
``matching`` (the same surface bound from Python for tests and bench), ``synth`` (deterministic inputs).
"""
from .matching import (ArrayMatcherB200, B200MatchError, Context, EMatcherType, ImageCollectionMatcherB200,  # noqa: F401
                       createImageCollectionMatcher, default_context, load_library)

__all__ = ["ArrayMatcherB200", "ImageCollectionMatcherB200", "EMatcherType", "Context", "B200MatchError",
           "createImageCollectionMatcher", "default_context", "load_library"]
