#!/usr/bin/env python3
"""Replays exact_div() of dzahui_b200/csrc/tridiag_kernels.cuh (Markstein's division by a precomputed correctly rounded
reciprocal, two corrections) in exact rational arithmetic and compares it with the correctly rounded quotient, on random and
adversarial operands (significands near all ones / near a power of two / short).  Also checks that the second residual
is exactly representable, which is the hypothesis of the theorem.
usage: tools/check_exact_div.py [seed] [cases]   ->  exit status 0 iff no mismatch"""
import random
import struct
import sys
from fractions import Fraction as F


def rnd_double(rng, special):
    if special == 0:
        return rng.uniform(-1, 1) * 2.0 ** rng.randint(-60, 60)
    m = rng.getrandbits(52)
    if special == 1:
        m |= ((1 << 52) - 1) ^ rng.getrandbits(6)
    elif special == 2:
        m &= rng.getrandbits(8)
    elif special == 3:
        m = rng.getrandbits(52) & ~((1 << rng.randint(0, 40)) - 1)
    elif special == 4:
        m = ((1 << 52) - 1) ^ rng.getrandbits(3)
    e = rng.randint(1023 - 50, 1023 + 50)
    return struct.unpack("<d", struct.pack("<Q", (rng.getrandbits(1) << 63) | (e << 52) | m))[0]


def fma(x, y, z):
    return float(F(x) * F(y) + F(z))  # float(Fraction) rounds to nearest even: one rounding, like the hardware fma


def exact_div(a, den):
    r = float(F(1) / F(den))  # RN(1 / den) = __drcp_rn
    q0 = a * r
    e0 = fma(-den, q0, a)
    q1 = fma(e0, r, q0)
    e1 = fma(-den, q1, a)
    return fma(e1, r, q1), F(e1) == F(a) - F(den) * F(q1)


def run(seed, cases):
    rng = random.Random(seed)
    bad = inexact = 0
    for _ in range(cases):
        a, den = rnd_double(rng, rng.randint(0, 4)), rnd_double(rng, rng.randint(0, 4))
        if a == 0.0 or den == 0.0:
            continue
        q, exact = exact_div(a, den)
        inexact += not exact
        if q != a / den:
            bad += 1
            if bad < 5:
                print("MISMATCH", a.hex(), den.hex(), q.hex(), (a / den).hex())
    return bad, inexact


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    bad, inexact = run(seed, cases)
    print(f"cases {cases} mismatches {bad} inexact second residuals {inexact}")
    sys.exit(1 if bad or inexact else 0)
