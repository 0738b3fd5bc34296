"""The device computes AS183's three quotients B/30269, B/30307, B/30323 with two FMAs instead of a
division (erlamsa_b200/csrc/eb_rng.cuh, Rng::div_exact). Exhaustive proof, in exact rational arithmetic,
that the result is the correctly rounded IEEE quotient for every operand the generator can produce."""
from fractions import Fraction as F


def test_fma_division_is_correctly_rounded_for_every_as183_operand():
    bad = []
    for c in (30269, 30307, 30323):
        cf = float(c)
        y = 1.0 / cf
        for a in range(c):
            af = float(a)
            q0 = af * y                                  # __dmul_rn
            r = float(F(af) - F(q0) * F(cf))             # __fma_rn(-q0, c, a): one rounding of the exact value
            q = float(F(r) * F(y) + F(q0))               # __fma_rn(r, y, q0)
            if q != af / cf:
                bad.append((c, a))
    assert not bad, bad[:5]
