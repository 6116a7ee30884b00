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


def div3(a):
    c = 1.0 / 3.0
    q = a * c
    r = fma(-3.0, q, a)
    return fma(r, c, q)


def test_div3_one_correction_is_correctly_rounded():
    """The channel mean (sum / 3.0) takes ONE remainder correction (csrc/costvol.cu::div3; the argument why one is
    enough for the divisor 3 is in the source).  Random operands, every residue of the significand mod 3 next to
    powers of two and to all-ones, multiples of 3, and the values the kernel actually sees (small sums)."""
    rng = random.Random(3)
    for i in range(60000):
        a = _rand(rng)
        if i % 4 == 1:                                                  # significands round 2^52, 2^53 - 1, and k * 2^52 / 3
            base = rng.choice(((1 << 52), (1 << 53) - 1, (1 << 54) // 3, (1 << 53) // 3 * 2))
            m = min(max(base + rng.randint(-4, 4), 1 << 52), (1 << 53) - 1)
            a = math.ldexp(m, rng.randint(-40, 40) - 52)
        elif i % 4 == 2:
            a = 3.0 * _rand(rng, -20, 20)                               # quotient (nearly) representable
        elif i % 4 == 3:
            a = (rng.random() - 0.5) * 10.0 ** rng.randint(-6, 3)       # what the residual sums look like
        assert div3(a) == a / 3.0, a.hex()
    assert div3(0.0) == 0.0
