"""Oracle vs the real reference, where it is importable (build container)."""
import numpy as np
import pytest

from oracle import harness, ref_loader
from oracle.cases import CASES

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present on this machine")]


@pytest.mark.parametrize("case", [c for c in CASES if "large" not in c.tags], ids=lambda c: c.name)
@pytest.mark.parametrize("variant", ["gui", "export"])
def test_bit_exact_against_reference(case, variant):
    ref_out, ref_state = harness.run_reference(case, variant)
    for backend in ("cv2", "numpy"):
        out, state = harness.run_oracle(case, variant, backend=backend)
        for j, (a, b) in enumerate(zip(ref_out, out)):
            assert np.array_equal(a, b), (backend, j, harness.diff_stats(a, b))
        if backend == "cv2":
            assert state.dtype == ref_state.dtype and np.array_equal(state, ref_state)


def test_gui_and_export_agree_without_glitch():
    """both entry points give identical uint8 output when glitch is off."""
    from oracle.cases import CASES_BY_NAME
    case = CASES_BY_NAME["cfg3_warp"]
    a, _ = harness.run_reference(case, "gui")
    b, _ = harness.run_reference(case, "export")
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
