"""rendering_learning_b200 — B200-native per-pixel ray loop behind the reference's scene API.

`rtc` mirrors ray-tracer-challenge, `ow` mirrors ray-tracing-one-weekend; both render through
librl_b200.so (hand-written sm_100a CUDA, C ABI in include/rl_b200.h).  No CPU fallback.
(The directory is spelled with an underscore because `rendering-learning_b200` is not an
importable Python identifier.)
"""
from . import _abi  # noqa: F401
from ._abi import RlError, load_library  # noqa: F401
from .context import Context, default_context  # noqa: F401
