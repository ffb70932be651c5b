"""Parity of the CUDA path (through the C ABI) against the oracle, on the B200.

Bar (north_star): within +-1 LSB per uint8 channel and PSNR >= 50 dB, with the
RNG-driven stages fed the reference's own draws.  Asserted as max |delta| <= 1
for EVERY case, colour gamma included: numpy's float32 `power` (SVML, within
1 ulp, not correctly rounded) cannot be matched bit for bit, and a 1-ulp
difference could in principle flip a triad LUT bin on a dark sample, but no
committed seed does (profiles/r01/parity_*.jsonl, profiles/r02)."""
import numpy as np
import pytest

from gpu_util import log_report, run_case_gpu
from oracle import harness
from oracle.cases import CASES, CASES_BY_NAME

pytestmark = pytest.mark.gpu


def _check(case, want, got, what):
    worst = {"max": 0, "frac_gt1": 0.0, "frac_ne": 0.0, "psnr": float("inf")}
    for a, b in zip(want, got):
        st = harness.diff_stats(a, b)
        worst = {"max": max(worst["max"], st["max"]), "frac_gt1": max(worst["frac_gt1"], st["frac_gt1"]),
                 "frac_ne": max(worst["frac_ne"], st["frac_ne"]), "psnr": min(worst["psnr"], st["psnr"])}
    log_report(case=case.name, what=what, **worst)
    assert worst["psnr"] >= 50.0, worst
    assert worst["max"] <= 1, worst
    assert worst["frac_ne"] <= 5e-3, worst


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
@pytest.mark.parametrize("variant", ["gui", "export"])
@pytest.mark.parametrize("policy", ["staged", "auto"])
def test_cuda_matches_oracle(case, variant, policy, golden):
    got, state, fused = run_case_gpu(case, variant, policy)
    if policy == "auto" and not fused:
        pytest.skip("fused kernel not selected for this parameter set (staged path already covered)")
    want, want_state = harness.run_oracle(case, variant, backend="cv2")
    _check(case, want, got, f"{variant}/{policy}/fused={fused}")
    d = np.abs(want_state.astype(np.float64) - state)
    assert d.max() < 1.0 / 255 and (d > 4e-6).mean() < 2e-3, (d.max(), (d > 4e-6).mean())
    # and against the fixture made from the unmodified reference
    st = harness.diff_stats(golden["last_frame"](case.name, variant), got[-1])
    assert st["psnr"] >= 50.0 and st["max"] <= 1, st


def test_identity_chain_is_exact():
    case = CASES_BY_NAME["identity"]
    from oracle.cases import case_frames
    got, _, _ = run_case_gpu(case, "export")
    assert all(np.array_equal(a, b) for a, b in zip(got, case_frames(case)))


@pytest.mark.parametrize("name", ["cfg1_cli_default", "cfg4_full", "high_persistence"])
def test_frame_by_frame_equals_batched(name):
    """Determinism of the state hand-off: one crt_process call per frame gives the same bytes."""
    case = CASES_BY_NAME[name]
    a, sa, _ = run_case_gpu(case, "export")
    b, sb, _ = run_case_gpu(case, "export", batch=1)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(sa, sb)
