"""pythoncrt_b200 — B200-native per-frame CRT effect chain.

Drop-in for the hot path of jaylikesbunda/PythonCRT (`crt_filter.py`):
`apply_crt_effect`, `apply_static_effects`, `make_triad_mask`, `make_vignette`
keep the reference's signatures; the work is done by hand-written sm_100a CUDA
kernels behind the C ABI in include/crt_b200.h.  No CPU fallback.
"""
from .params import CrtParams  # noqa: F401
from .tables import make_triad_mask, make_vignette  # noqa: F401

__all__ = ["CrtParams", "CrtEngine", "apply_crt_effect", "apply_static_effects", "make_triad_mask", "make_vignette",
           "process_clip", "DeviceState"]


def __getattr__(name):
    # compute-path objects are imported lazily so that the parameter/table helpers
    # stay importable on a machine without the CUDA library.
    if name == "CrtEngine":
        from .engine import CrtEngine
        return CrtEngine
    if name in ("apply_crt_effect", "apply_static_effects", "DeviceState"):
        from . import effects
        return getattr(effects, name)
    if name == "process_clip":
        from .clip import process_clip
        return process_clip
    raise AttributeError(name)
