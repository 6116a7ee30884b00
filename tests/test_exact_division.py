"""The cost-volume kernel replaces divisions by known divisors with `div_rc` (csrc/costvol.cu): a
reciprocal multiply plus two exact FMA remainder corrections.  Replay that sequence in rational
arithmetic (each FMA rounded once, as the hardware does) and check it returns the correctly rounded
quotient -- the bits numpy's `/` produces in the reference -- including adversarial divisors."""
import math
import random
from fractions import Fraction as Fr


def fma(a, b, c):
    return float(Fr(a) * Fr(b) + Fr(c))     # float(Fraction) rounds to nearest-even once


def div_rc(a, d, rd):
    q = a * rd
    r = fma(-d, q, a)
    q = fma(r, rd, q)
    r = fma(-d, q, a)
    return fma(r, rd, q)


def _rand(rng, emin=-60, emax=60):
    m = rng.getrandbits(52) | (1 << 52)
    return rng.choice((-1, 1)) * math.ldexp(m, rng.randint(emin, emax) - 52)


def test_div_rc_is_correctly_rounded():
    rng = random.Random(1)
    for i in range(30000):
        a, d = _rand(rng), _rand(rng)
        if i % 5 == 0:
            d = 3.0                                                    # the channel mean
        elif i % 7 == 0:
            d = math.ldexp((1 << 53) - 1, rng.randint(-60, 8))          # all-ones significand
        elif i % 11 == 0:
            d = math.ldexp((1 << 52) + rng.randint(1, 3), rng.randint(-60, 8))   # just above a power of two
        if i % 13 == 0:
            a = float(Fr(d) * Fr(_rand(rng, -5, 5)))                    # quotient (nearly) representable
        assert div_rc(a, d, 1.0 / d) == a / d, (a.hex(), d.hex())


def test_div_rc_zero_numerator():
    for d in (3.0, -0.75, 1.0000000000000002):
        assert div_rc(0.0, d, 1.0 / d) == 0.0
