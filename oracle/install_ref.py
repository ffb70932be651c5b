"""Stage the UNMODIFIED reference under baseline/_ref (TEST / BENCH INFRASTRUCTURE).

    python -m oracle.install_ref

`bench.py --impl reference` times the reference's own CPU implementation of the hot path.  The
reference is a single Python module without packaging metadata, so the contract's
`pip install --no-index --target baseline/_ref /root/reference` cannot work ("neither setup.py nor
pyproject.toml found" — the outcome is recorded in baseline/_ref/INSTALL_NOTE.txt and DESIGN.md);
the module file is staged as it is instead.  baseline/_ref is git-ignored (reference sources never
enter the repository history) but travels to the GPU box with the gpurun snapshot, where
/root/reference does not exist.  oracle/ref_loader.py imports the staged file with the same
stand-in modules it uses for /root/reference.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose: bool = True) -> bool:
    """Returns True when baseline/_ref/crt_filter.py is in place afterwards."""
    staged = os.path.join(DST, "crt_filter.py")
    if not os.path.isfile(os.path.join(SRC, "crt_filter.py")):
        return os.path.isfile(staged)
    os.makedirs(DST, exist_ok=True)
    tmp = "/tmp/_crt_ref_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp)
    res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
                          "/opt/wheelhouse", "--target", DST, tmp], capture_output=True, text=True)
    pip_note = (res.stdout + res.stderr).strip().splitlines()[-1:] or ["(no output)"]
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copyfile(os.path.join(SRC, "crt_filter.py"), staged)
    with open(os.path.join(DST, "INSTALL_NOTE.txt"), "w") as f:
        f.write(f"pip install --target baseline/_ref /root/reference: rc={res.returncode}: {pip_note[0]}\n"
                "staged instead: crt_filter.py, byte-identical to /root/reference/crt_filter.py\n")
    if verbose:
        print(f"reference staged at {staged} (pip: rc={res.returncode}, {pip_note[0]})")
    return True


if __name__ == "__main__":
    raise SystemExit(0 if install() else 1)
