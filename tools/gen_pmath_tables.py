#!/usr/bin/env python
"""Generates the constant tables of wave_tracer_b200/csrc/pmath.h (the portable, bit-reproducible f32 math library shared by the
device code and the CPU oracle): pi/2 split in two doubles, and 2^e mod pi/2 (+ quadrant) for the huge-argument path of sin/cos.
Pure integer / decimal arithmetic; run once, output pasted into pmath.h (tests/test_pmath.py re-derives and checks the table)."""
from decimal import Decimal, getcontext
from fractions import Fraction
import struct

getcontext().prec = 120
PI = Decimal("3.14159265358979323846264338327950288419716939937510582097494459230781640628620899862803482534211706798214808651328230664709384460955058223172535940812848111745")


def to_double(d):
    return float(d)      # Decimal -> nearest double (correctly rounded)


def hexd(x):
    return "0x%016x" % struct.unpack("<Q", struct.pack("<d", x))[0]


def pio2_split():
    p = PI / 2
    hi = to_double(p)
    lo = to_double(p - Decimal(hi))
    return hi, lo


def residue_table(e0=17, e1=104):
    p = PI / 2
    rows = []
    for e in range(e0, e1 + 1):
        x = Decimal(2) ** e
        q = int(x / p)
        r = x - Decimal(q) * p
        assert 0 <= r < p
        rows.append((e, q & 3, to_double(r)))
    return rows


if __name__ == "__main__":
    hi, lo = pio2_split()
    print("PIO2_HI", repr(hi), hexd(hi)); print("PIO2_LO", repr(lo), hexd(lo))
    print("TWO_OVER_PI", repr(to_double(2 / PI)))
    print("LN2_HI/LO", repr(to_double(Decimal(2).ln())), repr(to_double(Decimal(2).ln() - Decimal(to_double(Decimal(2).ln())))))
    print("LOG2E", repr(to_double(1 / Decimal(2).ln())))
    rows = residue_table()
    print("static const double kR[%d] = {" % len(rows))
    print(",\n".join("    " + ", ".join("%r" % r for _, _, r in rows[i:i + 4]) for i in range(0, len(rows), 4)))
    print("};\nstatic const unsigned char kQ[%d] = {" % len(rows))
    print("    " + ", ".join(str(q) for _, q, _ in rows))
    print("};")
