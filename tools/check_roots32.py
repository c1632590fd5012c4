"""re-derives offt_b200/csrc/roots32.h (exp(-2*pi*i*k/32), k < 32) with 40-digit decimal arithmetic"""
import re
import sys
from decimal import Decimal, getcontext
from pathlib import Path

getcontext().prec = 40
PI = Decimal("3.14159265358979323846264338327950288419716939937510")


def series(x, first, start):
    s = t = first
    n = start
    while abs(t) > Decimal(10) ** -38:
        n += 2
        t = -t * x * x / ((n - 1) * n)
        s += t
    return s


txt = (Path(__file__).resolve().parents[1] / "offt_b200" / "csrc" / "roots32.h").read_text()
bad = 0
for m in re.finditer(r"Root32<(\d+)> \{ static constexpr double re = ([-\d.]+); static constexpr double im = ([-\d.]+);", txt):
    a = PI * 2 * int(m.group(1)) / 32
    c, s_ = series(a, Decimal(1), 0), series(a, a, 1)
    if abs(Decimal(m.group(2)) - c) > Decimal("1e-24") or abs(Decimal(m.group(3)) + s_) > Decimal("1e-24"):
        bad += 1
        print("mismatch at k =", m.group(1))
print("roots32.h:", "ok" if not bad else f"{bad} bad entries")
sys.exit(1 if bad else 0)
