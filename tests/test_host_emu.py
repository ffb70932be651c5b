"""The kernels' own per-pixel arithmetic (csrc/crt_math.cuh, crt_stages.cuh),
compiled for the host by tests/host_emu, against the oracle on every parity
case.  CPU only.  Tolerance: +-1 LSB (north_star); in practice all but a
handful of rounding ties are identical."""
import numpy as np
import pytest

import host_emu
from oracle import harness
from oracle.cases import CASES


@pytest.mark.parametrize("case", [c for c in CASES if "large" not in c.tags], ids=lambda c: c.name)
@pytest.mark.parametrize("variant", ["gui", "export"])
def test_kernel_arithmetic_matches_oracle(case, variant):
    want, want_state = harness.run_oracle(case, variant, backend="cv2")
    got, state = host_emu.run_case(case, variant)
    for a, b in zip(want, got):
        st = harness.diff_stats(a, b)
        assert st["max"] <= 1 and st["frac_ne"] <= 2e-4 and st["psnr"] >= 50.0, st
    # float32 state vs the reference's float64 state: rounding noise only, except where a
    # 1-ulp difference in float32 pow (numpy's SVML vs libm / CUDA) flips a triad LUT bin
    d = np.abs(want_state.astype(np.float64) - state)
    assert d.max() < 1.0 / 255 and (d > 2e-6).mean() < 1e-3, (d.max(), (d > 2e-6).mean())


def test_static_image_matches_apply_static_effects():
    from oracle import crt_oracle as O
    from oracle.cases import CASES_BY_NAME, case_frames
    case = CASES_BY_NAME["cfg3_warp"]
    img, _ = host_emu.run_case(case, "export", static=True)
    ref = O.static_chain(case_frames(case)[0], case.params, phase_px=harness.frame_scalars(case, 0)[0], variant="export")
    assert np.max(np.abs(ref - img[0])) < 2e-6


def test_pow_unit_is_correctly_rounded_almost_everywhere():
    """csrc/crt_math.cuh pow_unit (colour gamma): against float64 pow rounded once.  numpy's own
    float32 power differs from that reference in ~20 % of inputs (SVML, <= 1 ulp); pow_unit in < 1 %
    over the GUI's gamma range (0.1 .. 5), never by more than 1 ulp.  The error grows with 1 / gamma:
    at the CLI's floor (gamma = 0.001, y = 1000) a few ulp are allowed."""
    import ctypes as C
    L = host_emu.lib()
    rng = np.random.default_rng(0)
    x = rng.random(400_000, dtype=np.float32)
    x[:256] = np.arange(256, dtype=np.float32) / 255
    x[256:259] = [1.0, 1e-30, 0.0]
    x[300:20300] = np.float32(0.5) + rng.random(20000, dtype=np.float32) * np.float32(0.0117)     # widest |r| of the log table
    out = np.empty_like(x)
    for gamma, max_ulp, max_miss in ((1.1, 1, 1e-2), (2.2, 1, 1e-2), (0.8, 1, 1e-2), (3.0, 1, 1e-2), (5.0, 1, 1e-2), (0.5, 1, 1e-2),
                                     (0.2, 1, 2e-2), (0.1, 1, 4e-2), (0.05, 1, 8e-2), (0.001, 16, 1.0)):
        y = np.float32(1.0 / gamma)
        L.emu_pow_unit(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_int(x.size), C.c_double(float(y)))
        ref = np.power(x.astype(np.float64), np.float64(y)).astype(np.float32)
        big = ref > 1e-37                                   # results below 2^-125 flush to zero
        ulp = np.abs(out.view(np.int32).astype(np.int64) - ref.view(np.int32))[big]
        assert ulp.max() <= max_ulp and (ulp > 0).mean() < max_miss, (gamma, ulp.max(), (ulp > 0).mean())
        assert np.all(out[~big] < 1e-37)
        assert out[256] == 1.0 and out[258] == 0.0


def test_constant_divisor_division_is_the_ieee_quotient():
    """csrc/crt_math.cuh div_const (bloom threshold, crt_filter.py:603): reciprocal, exact residual, one correction —
    must be the correctly rounded quotient for every operand the stage can see (v - thr with v in [0, 1])."""
    import ctypes as C
    L = host_emu.lib()
    rng = np.random.default_rng(3)
    for thr in (0.7, 0.5, 0.3, 0.99, 0.01, 0.123456, 0.9, 1.0 - 1e-6, 0.6180339):
        thr32 = np.float32(min(0.99, max(0.0, thr)))
        d = np.float32(max(1e-6, 1.0 - float(min(0.99, max(0.0, thr)))))
        v = np.concatenate([rng.random(1_000_000, dtype=np.float32), np.arange(256, dtype=np.float32) / np.float32(255.0),
                            np.float32(thr32) + d * rng.random(200_000, dtype=np.float32)])
        n = (v - thr32).astype(np.float32)
        out, want = np.empty_like(n), np.empty_like(n)
        L.emu_div_const(n.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), C.c_int(n.size), C.c_float(float(d)))
        assert np.array_equal(want, n / d)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (thr, int((out != want).sum()))
