"""CPU model of the fixed-point all-reduce of the single-reduction CG kernels (csrc/jf_cg_allreduce.cuh:fused_fx_allreduce):
the constants are read from the source, the encode / accumulate / decode steps are restated with Python integers and
checked for the invariants the kernel relies on -- limbs never spill into the arrival count, the count is exact for
every grid size up to 1023, the decoded total is the exact sum of the partials up to one rounding (plus the bits
below the low unit), partials beyond the headroom are flagged."""
import math
import os
import re
from fractions import Fraction

import numpy as np
import pytest

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "jointforces_b200", "csrc", "jf_cg_allreduce.cuh")


def constants():
    text = open(SRC).read()
    c = {}
    for name in ("FX_COUNT_SHIFT", "FX_BITS", "FX_HEADROOM"):
        c[name] = int(re.search(r"constexpr int %s = (\d+);" % name, text).group(1))
    assert "FX_HI_SHIFT = FX_BITS - 2 - FX_HEADROOM, FX_LO_SHIFT = FX_HI_SHIFT + FX_BITS" in text
    c["FX_HI_SHIFT"] = c["FX_BITS"] - 2 - c["FX_HEADROOM"]
    c["FX_LO_SHIFT"] = c["FX_HI_SHIFT"] + c["FX_BITS"]
    return c


C = constants()


def exponent_of(total):
    """fx_exponent_ok: ilogb + 1."""
    return math.frexp(total)[1]


def encode(x, E):
    """One partial -> (word_hi, word_lo, ok), as lane 0 does."""
    ok = abs(x) < 2.0 ** (E + C["FX_HEADROOM"])
    if not ok:
        hi = lo = 0
    else:
        hi = math.floor(x * 2.0 ** (C["FX_HI_SHIFT"] - E))
        unit_hi = 2.0 ** (E - C["FX_HI_SHIFT"])
        rem = x - hi * unit_hi                                  # one rounding (the kernel's fma): exact unless x is a tiny
        exact_rem = Fraction(x) - hi * Fraction(unit_hi)        # negative number next to a unit boundary; then the error
        assert abs(Fraction(rem) - exact_rem) <= Fraction(unit_hi) * Fraction(2) ** -53    # is far below the low unit
        assert 0.0 <= rem <= unit_hi                            # (== unit_hi after rounding up: the clamp below)
        lo = math.floor(rem * 2.0 ** (C["FX_LO_SHIFT"] - E))
        lo = min(max(lo, 0), (1 << C["FX_BITS"]) - 1)
    one = 1 << C["FX_COUNT_SHIFT"]
    return one + hi + (1 << (C["FX_BITS"] - 1)), one + lo, ok


def decode(w_hi, w_lo, grid, E):
    mask = (1 << C["FX_COUNT_SHIFT"]) - 1
    assert w_hi >> C["FX_COUNT_SHIFT"] == grid and w_lo >> C["FX_COUNT_SHIFT"] == grid      # arrival counts intact
    H = (w_hi & mask) - grid * (1 << (C["FX_BITS"] - 1))
    L = w_lo & mask
    return float(H) * 2.0 ** (E - C["FX_HI_SHIFT"]) + float(L) * 2.0 ** (E - C["FX_LO_SHIFT"])


def test_layout_constants():
    assert C["FX_COUNT_SHIFT"] + 10 == 64                                   # ten count bits
    assert 1023 * ((1 << C["FX_BITS"]) - 1) < 1 << C["FX_COUNT_SHIFT"]      # 1023 full limbs stay below the count
    assert C["FX_HEADROOM"] + C["FX_HI_SHIFT"] + 2 == C["FX_BITS"]          # |hi| < 2^(BITS-2): biased limb in (0, 2^BITS)
    assert C["FX_LO_SHIFT"] >= 64                                           # finer than a double's ulp at the scale


@pytest.mark.parametrize("grid", [1, 2, 11, 134, 296, 1023])
@pytest.mark.parametrize("scale", [1e-30, 3.7e-9, 1.0, 8.1e12])
def test_fixed_point_sum_is_exact(grid, scale):
    rng = np.random.default_rng(grid)
    for trial in range(20):
        # partials of mixed sign and widely different magnitude, some far above the previous total (cancellation)
        x = rng.normal(size=grid) * scale * 10.0 ** rng.uniform(-12, 3, size=grid)
        prev_total = abs(float(np.sum(x))) * 10.0 ** rng.uniform(-1, 1) + scale * 1e-3
        E = exponent_of(prev_total)
        w_hi = w_lo = 0
        all_ok = True
        for v in x:                                    # integer addition: any order gives the same words
            a, b, ok = encode(float(v), E)
            w_hi += a
            w_lo += b
            all_ok &= ok
        assert w_hi < 1 << 64 and w_lo < 1 << 64
        if not all_ok:
            assert np.max(np.abs(x)) >= 2.0 ** (E + C["FX_HEADROOM"])
            continue
        got = decode(w_hi, w_lo, grid, E)
        exact = sum(Fraction(float(v)) for v in x)
        unit_lo = Fraction(2) ** (E - C["FX_LO_SHIFT"])
        # each partial loses less than one low unit; the decode rounds twice at most
        err = abs(Fraction(got) - exact)
        slack = grid * (unit_lo + Fraction(2) ** (E - C["FX_HI_SHIFT"] - 53)) + 2 * abs(exact) * Fraction(2) ** -52 \
            + Fraction(2) ** (E - C["FX_HI_SHIFT"]) * grid * Fraction(2) ** -52
        assert err <= slack
        # and it beats (or equals) the sequential double sum the slot tables give
        seq = 0.0
        for v in x:
            seq += float(v)
        assert err <= abs(Fraction(seq) - exact) + slack


def test_partials_beyond_the_headroom_are_flagged():
    E = exponent_of(1.0)
    assert encode(2.0 ** (E + C["FX_HEADROOM"]) * 0.999, E)[2]
    assert not encode(2.0 ** (E + C["FX_HEADROOM"]), E)[2]
    assert not encode(float("nan"), E)[2] and not encode(float("inf"), E)[2]
    # negative partials: floor keeps the remainder non-negative, the bias keeps the limb non-negative
    a, b, ok = encode(-3.25e-7, E)
    assert ok and (a & ((1 << C["FX_COUNT_SHIFT"]) - 1)) > 0
    assert abs(decode(a, b, 1, E) - -3.25e-7) <= 2.0 ** (E - C["FX_LO_SHIFT"])     # bits below the low unit are dropped
    a, b, ok = encode(-0.40625, E)
    assert ok and decode(a, b, 1, E) == -0.40625
