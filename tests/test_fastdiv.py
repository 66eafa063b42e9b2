"""CPU: the division of the fused histogram kernel (csrc/icpf_histfused.cu, vote_bin_fast) restated with exact rational
arithmetic.  q = RN(q0 + RN(x - q0 b) y) with y = RN(1/b), q0 = RN(x y) must equal the correctly rounded quotient RN(x / b)
-- the value hist_cuda_core.cuh:54-56 computes with an IEEE division -- for the divisors the reference's settings produce
(bins.max() - bins.min()).  tools/check_fastdiv.cu runs the same comparison on the GPU for EVERY dividend in [0, b);
here: the edge cases plus a random sample, every rounding done exactly (fractions, round-half-even to 24 bits)."""
from fractions import Fraction

import numpy as np
import pytest


def rn32(v: Fraction) -> Fraction:
    """Round a rational to the nearest binary32 value (ties to even); normal range only."""
    if v == 0:
        return Fraction(0)
    sign = -1 if v < 0 else 1
    v = abs(v)
    e = v.numerator.bit_length() - v.denominator.bit_length()
    if Fraction(2) ** e > v:
        e -= 1
    assert -126 <= e <= 127, "outside the normal range"
    ulp = Fraction(2) ** (e - 23)
    q, r = divmod(v, ulp)
    if r * 2 > ulp or (r * 2 == ulp and q % 2 == 1):
        q += 1
    return sign * q * ulp


def fast_div(x: Fraction, b: Fraction) -> Fraction:
    y = rn32(1 / b)                     # __frcp_rn
    q0 = rn32(x * y)                    # __fmul_rn
    r = rn32(x - q0 * b)                # __fmaf_rn(-q0, b, x): one rounding of the exact value
    return rn32(q0 + r * y)             # __fmaf_rn(r, y, q0)


def _divisors():
    out = []
    tau = np.float32(0.1)
    for F in (2.0, 3.333, 3.34, 6.666, 10.0, 13.332):
        n = int((F + 0.1 - 1e-8 + F) / 0.1) + 1
        for first in (np.float32(-F),):
            last = np.float32(-F + (n - 1) * 0.1)
            d = np.float32(last - first)
            out += [np.nextafter(d, np.float32(0)), d, np.nextafter(d, np.float32(100))]
    out += [np.float32(tau + tau), np.float32(0.2), np.float32(0.30000001)]
    return [Fraction(float(v)) for v in out]


@pytest.mark.parametrize("b", _divisors())
def test_markstein_quotient_is_correctly_rounded(b):
    rng = np.random.default_rng(int(b * 1000) % 2**31)
    bf = np.float32(float(b))
    xs = [np.float32(0), np.nextafter(bf, np.float32(0)), np.float32(1e-20), np.float32(float(b) / 3)]
    xs += list(rng.uniform(0, float(b), 1500).astype(np.float32))
    xs += list((rng.uniform(0, 1, 300).astype(np.float32) ** 8 * bf).astype(np.float32))       # many small dividends
    # dividends that land on or next to bin boundaries (the values that decide a vote's bin)
    for k in range(0, 136, 5):
        edge = np.float32(float(b) * k / 135)
        xs += [edge, np.nextafter(edge, np.float32(0)), np.nextafter(edge, np.float32(100))]
    for x in xs:
        x = float(x)
        if not (0.0 <= x < float(b)):
            continue
        xf = Fraction(x)
        want = rn32(xf / b) if xf > Fraction(1, 10**30) else None
        if want is None:
            continue
        assert fast_div(xf, b) == want, (x, float(b))


def test_guard_conditions_cover_the_exceptional_divisors():
    """An all-ones significand is the one divisor class for which the correction step can miss; launch_hist_fused() falls
    back to the IEEE division there (and outside a +-2^40 exponent window)."""
    b = Fraction(float(np.nextafter(np.float32(2.0), np.float32(0))))        # 1.99999988 = 0x3fffffff
    bits = np.float32(float(b)).view(np.uint32)
    assert (int(bits) & 0x7FFFFF) == 0x7FFFFF                                  # the host check rejects exactly this
