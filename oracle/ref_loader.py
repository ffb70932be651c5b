"""Import the UNMODIFIED reference module in the build container (test infra).

`/root/reference/crt_filter.py` runs `ensure_deps()` at import time
(crt_filter.py:17-47), which shells out to `pip install` when moviepy /
PySide6 / imageio_ffmpeg are missing.  They are missing here and there is no
network, so before executing the module we register empty stand-ins for those
packages (with a real `__spec__`, otherwise `importlib.util.find_spec` raises
and `ensure_deps` takes the pip branch, crt_filter.py:22-25) and make
`subprocess.run` refuse to run while the module body executes.

Only the pure numpy/cv2 functions of the reference are used afterwards.
/root/reference does not exist on the GPU box; the copy staged under
baseline/_ref by oracle/install_ref.py (git-ignored, shipped with the gpurun
snapshot) is used there by `bench.py --impl reference`.  Tests never rely on
it: they skip when `available()` is False and use the golden fixtures.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import subprocess
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the read-only checkout in the build container, else the copy staged by oracle/install_ref.py (travels to the GPU box)
_CANDIDATES = [os.environ.get("CRT_REFERENCE_PATH", ""), "/root/reference/crt_filter.py",
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref", "crt_filter.py")]
REFERENCE_PATH = next((c for c in _CANDIDATES if c and os.path.isfile(c)), _CANDIDATES[1])

_STUBS = {
    "moviepy": {},
    "moviepy.editor": {"VideoFileClip": object},
    "moviepy.video": {},
    "moviepy.video.io": {},
    "moviepy.video.io.ffmpeg_writer": {"FFMPEG_VideoWriter": object},
    "imageio_ffmpeg": {"get_ffmpeg_exe": lambda: "ffmpeg"},
    "PySide6": {},
}

_cached = None


def available() -> bool:
    return os.path.isfile(REFERENCE_PATH)


def _install_stubs():
    added = []
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is not None:
                continue
        except (ImportError, ValueError):
            pass
        mod = types.ModuleType(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        mod.__path__ = []  # behave like a package so submodule imports resolve
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        added.append(name)
    return added


def load():
    """Return the reference module object (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_PATH}")
    _install_stubs()
    real_run = subprocess.run

    def _no_subprocess(*a, **k):
        raise RuntimeError("reference tried to spawn a subprocess during import")

    subprocess.run = _no_subprocess
    try:
        spec = importlib.util.spec_from_file_location("crt_filter_reference", REFERENCE_PATH)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["crt_filter_reference"] = mod
        spec.loader.exec_module(mod)
    finally:
        subprocess.run = real_run
    _cached = mod
    return mod
