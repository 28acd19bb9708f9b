"""CPU tests of the input side: the field conversion the device parser runs (numparse.cuh, reached through the
host-only hook tp_test_parse_field) must be the correctly rounded decimal -> double conversion the reference's
read.big.matrix performs (R/TADpole.R:17); Python's float() is the checker.  No GPU is touched."""
import math
import random
import struct
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np
import pytest

from oracle import tadpole_oracle as O


@pytest.fixture(scope="module")
def parse():
    import __graft_entry__ as g
    g.build()
    from tadpole_b200 import _lib
    return _lib.parse_field


def same_bits(a, b):
    return struct.pack("d", a) == struct.pack("d", b) or (a != a and b != b)


def test_special_fields(parse):
    for s in ["NA", "NaN", "nan", "", "  ", "NAN"]:
        st, v = parse(s)
        assert st == 0 and math.isnan(v), s
    assert parse("Inf") == (0, math.inf) and parse("-Inf") == (0, -math.inf) and parse("+inf") == (0, math.inf)
    for s in ["abc", "1e", "1e+", "--1", "1.2.3", "0x10", "1,5", "-", "."]:
        assert parse(s)[0] == 1, s                 # left to the host, which rejects them (or accepts hex like strtod)
    st, v = parse("-0")
    assert st == 0 and v == 0 and math.copysign(1, v) == -1
    assert parse(" 12 ") == (0, 12.0) and parse("5.") == (0, 5.0) and parse(".5") == (0, 0.5) and parse("7\r") == (0, 7.0)


def test_known_hard_cases(parse):
    hard = ["9007199254740993", "9007199254740992", "9007199254740991", "2.2250738585072011e-308", "2.2250738585072014e-308",
            "4.9e-324", "2.4703282292062328e-324", "1.7976931348623157e308", "1.7976931348623159e308", "1e400", "-1e400",
            "1e-400", "0.1", "0.30000000000000004", "123456789012345678", "1234567890123456789", "12345678901234567890",
            "1e22", "1e23", "8.5e22", "1.0000000000000002", "1.00000000000000011102230246251565404236316680908203125",
            "1.00000000000000011102230246251565404236316680908203124", "1.00000000000000011102230246251565404236316680908203126",
            "6.02214076e23", "1E+05", "1e-5", "100000000000000000000000", "0.000000000000000000000000000001"]
    host = 0
    for s in hard:
        st, v = parse(s)
        if st == 1:
            host += 1
            continue
        assert same_bits(v, float(s)), (s, v, float(s))
    assert host <= 3          # only >19-digit fields sitting on a rounding boundary may go to the host


def test_random_fields_match_python_float(parse):
    rnd = random.Random(7)
    getcontext().prec = 1200
    n = host = 0
    for _ in range(60000):
        k = rnd.random()
        if k < 0.25:
            s = str(rnd.randrange(0, 10 ** rnd.randrange(1, 9)))
        elif k < 0.45:
            s = repr(rnd.random() * 10 ** rnd.randrange(-5, 6))
        elif k < 0.65:
            x = struct.unpack("d", struct.pack("Q", rnd.getrandbits(64)))[0]
            if x != x or math.isinf(x):
                continue
            s = repr(x)
        elif k < 0.8:
            nd = rnd.randrange(1, 30)
            d = "".join(rnd.choice("0123456789") for _ in range(nd))
            p = rnd.randrange(0, nd + 1)
            s = (d[:p] or "0") + "." + d[p:] + ("e%d" % rnd.randrange(-330, 310) if rnd.random() < 0.5 else "")
        elif k < 0.9:
            s = "%.*e" % (rnd.randrange(0, 25), rnd.random() * 10.0 ** rnd.randrange(-320, 308))
        else:                                   # exact midpoint between two adjacent doubles, printed in full
            x = abs(struct.unpack("d", struct.pack("Q", rnd.getrandbits(62)))[0])
            if not (1e-5 < x < 1e20):
                continue
            mid = (Fraction(x) + Fraction(float(np.nextafter(x, np.inf)))) / 2
            s = format(Decimal(mid.numerator) / Decimal(mid.denominator), "f")
        st, v = parse(s)
        n += 1
        if st == 1:
            host += 1
            continue
        assert same_bits(v, float(s)), (s, v, float(s))
    assert n > 50000 and host < 0.02 * n


def test_oracle_text_roundtrip():
    rng = np.random.default_rng(3)
    m = rng.poisson(3.0, (40, 40)).astype(float)
    m[2, 5] = np.nan
    m[7, 7] = 0.125
    back = O.read_matrix_text(O.matrix_to_text(m))
    assert np.array_equal(np.isnan(back), np.isnan(m)) and np.array_equal(np.nan_to_num(back), np.nan_to_num(m))
    with pytest.raises(ValueError):
        O.read_matrix_text("1\t2\n3\n")
