"""include/mz_math.h accuracy (host build) against float64 libm, and the pins lifted from the reference's own
function bodies (tests/golden/make_reference_pins.py: muax/utils.py:65-102, muax/nn.py:37-44)."""
import os

import numpy as np

from helpers import GOLDEN_DIR


def _ulp_err(y, ref):
    ulp = np.spacing(np.abs(ref.astype(np.float32))).astype(np.float64)
    return np.abs(y.astype(np.float64) - ref) / ulp


def test_expf_logf_expm1f_within_2ulp(c_oracle):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-87, 88, 400000), rng.uniform(-1, 1, 200000)]).astype(np.float32)
    assert _ulp_err(c_oracle.expf(x), np.exp(x.astype(np.float64))).max() < 1.5
    x = np.concatenate([rng.uniform(1e-38, 10, 200000), rng.uniform(0.5, 2, 200000),
                        np.exp(rng.uniform(-87, 88, 200000))]).astype(np.float32)
    assert _ulp_err(c_oracle.logf(x), np.log(x.astype(np.float64))).max() < 1.5
    x = np.concatenate([-np.exp(rng.uniform(-30, 3, 400000)), rng.uniform(-1, 0, 200000)]).astype(np.float32)
    assert _ulp_err(c_oracle.expm1f(x), np.expm1(x.astype(np.float64))).max() < 2.0


def test_special_values(c_oracle):
    e = c_oracle.expf([0.0, -np.inf, np.inf, -200.0, 100.0])
    assert e.tolist() == [1.0, 0.0, np.inf, 0.0, np.inf]
    assert np.isnan(c_oracle.expf([np.nan])[0])
    l = c_oracle.logf([0.0, 1.0, np.inf])
    assert l.tolist() == [-np.inf, 0.0, np.inf]
    assert np.isnan(c_oracle.logf([-1.0])[0])
    tiny = np.finfo(np.float32).tiny
    assert abs(float(c_oracle.logf([tiny])[0]) - np.log(float(tiny))) < 1e-5


def test_reference_function_pins(c_oracle):
    g = np.load(os.path.join(GOLDEN_DIR, "reference_pins.npz"))
    assert np.array_equal(c_oracle.inv_scaling(g["inv_x"]), g["inv_y"])
    assert np.array_equal(c_oracle.support_from_probs(g["sts_probs"], 10), g["sts_y"])
    assert np.array_equal(c_oracle.min_max_normalize(g["mm_x"]), g["mm_y"])
    assert np.abs(c_oracle.pb_c(g["pbc_visits"]).astype(np.float64) - g["pbc_y"]).max() < 5e-7


def test_branch_free_forms_are_exhaustively_identical(c_oracle):
    """include/mz_math.h evaluates exp/expm1 with selects instead of early returns (GPU warps would serialise
    on the branches).  The two formulations must agree on every float: all 2^32 bit patterns were checked when
    the change was made (31 s on 8 cores); the suite re-checks every threshold neighbourhood completely and a
    strided sweep of the whole space."""
    import struct

    def bits(f):
        return struct.unpack("<I", struct.pack("<f", f))[0]

    for centre in (0.34657359, -0.34657359, 88.72283905, -103.972084, -104.0, 89.0, 0.0, -0.0, -87.33655, 1.0):
        b = bits(centre)
        lo, hi = max(b - 200000, 0), min(b + 200000, 2**32 - 1)
        assert c_oracle.check_branch_free(lo, hi)[0] == 0, centre
    for start in range(0, 2**32, 2**32 // 64):  # 64 windows of 2^18 consecutive patterns across the space
        assert c_oracle.check_branch_free(start, start + 2**18)[0] == 0
    # NaN / inf patterns
    assert c_oracle.check_branch_free(0x7F800000 - 1000, 0x7F800000 + 100000)[0] == 0
    assert c_oracle.check_branch_free(0xFF800000 - 1000, 0xFF800000 + 100000)[0] == 0
