#!/usr/bin/env python3
"""Generate the polynomial coefficient tables used by include/tg_detmath.h.

The library's numeric contract needs transcendental functions whose results are
bit-identical on the host (gcc, x86-64, no FMA) and on the device (nvcc
-fmad=false).  glibc libm and the CUDA math library disagree in the last ulp,
and that one ulp is amplified by the ill-conditioned reduced system
(SURVEY.md H1), so the library carries its own +,-,*,/ only implementations.
This script derives near-minimax (Chebyshev-fit) coefficients with mpmath at
60 digits and prints them as C hex-float literals.  Run:  python tools/gen_detmath_coeffs.py
"""
import mpmath as mp

mp.mp.dps = 60


def hexf(x):
    return float(x).hex()


def fit(f, a, b, n):
    # chebyfit returns coefficients highest power first
    c, err = mp.chebyfit(f, [a, b], n, error=True)
    return [c[len(c) - 1 - i] for i in range(len(c))], err


def emit(name, coeffs, err):
    print(f"// {name}: max abs fit error {mp.nstr(err, 5)}")
    print(f"TG_DM_CONST double {name}[{len(coeffs)}] = {{")
    for c in coeffs:
        print(f"    {hexf(c)},  // {mp.nstr(c, 20)}")
    print("};")


# log: log(1+f) = 2*atanh(s), s = f/(2+f); R(z) = (2*atanh(s)/s - 2)/z with z = s^2
# on |s| <= (sqrt2-1)/(sqrt2+1) -> z in [0, 0.02944]
def log_R(z):
    if z == 0:
        return mp.mpf(2) / 3
    s = mp.sqrt(z)
    return (2 * mp.atanh(s) / s - 2) / z


c, e = fit(log_R, 0, mp.mpf("0.0295"), 8)
emit("kLogR", c, e)


# exp: c(r) = r*(e^r+1)/(e^r-1) = 2 + z*P(z), z = r^2, |r| <= ln2/2 -> z in [0, 0.1202]
def exp_P(z):
    if z == 0:
        return mp.mpf(1) / 6
    r = mp.sqrt(z)
    return (r * (mp.exp(r) + 1) / (mp.exp(r) - 1) - 2) / z


c, e = fit(exp_P, 0, mp.mpf("0.1202"), 6)
emit("kExpP", c, e)


# sin: sin(r) = r + r^3*S(z), z=r^2, |r|<=pi/4
def sin_S(z):
    if z == 0:
        return -mp.mpf(1) / 6
    r = mp.sqrt(z)
    return (mp.sin(r) - r) / (r * z)


c, e = fit(sin_S, 0, (mp.pi / 4) ** 2 * mp.mpf("1.001"), 7)
emit("kSinS", c, e)


# cos: cos(r) = 1 - z/2 + z^2*C(z)
def cos_C(z):
    if z == 0:
        return mp.mpf(1) / 24
    r = mp.sqrt(z)
    return (mp.cos(r) - 1 + z / 2) / (z * z)


c, e = fit(cos_C, 0, (mp.pi / 4) ** 2 * mp.mpf("1.001"), 7)
emit("kCosC", c, e)


# atan: atan(t) = t - t^3*T(z), z=t^2, |t| <= 7/16
def atan_T(z):
    if z == 0:
        return mp.mpf(1) / 3
    t = mp.sqrt(z)
    return (t - mp.atan(t)) / (t * z)


c, e = fit(atan_T, 0, (mp.mpf(7) / 16) ** 2 * mp.mpf("1.001"), 13)
emit("kAtanT", c, e)

# cbrt seed: cbrt(m) on [1, 8) (after exponent reduction mod 3) -- quadratic seed
c, e = fit(lambda m: mp.cbrt(m), 1, 8, 4)
emit("kCbrtSeed", c, e)

print("// constants")
for name, v in [
    ("kLn2Hi", None),
]:
    pass
ln2 = mp.log(2)
# ln2_hi: top 32 bits of ln2 so that k*ln2_hi is exact for |k| < 2^20
import struct


def trunc_bits(x, keep):
    f = float(x)
    b = struct.unpack("<Q", struct.pack("<d", f))[0]
    b &= ~((1 << (52 - keep)) - 1)
    return struct.unpack("<d", struct.pack("<Q", b))[0]


ln2_hi = trunc_bits(ln2, 31)
ln2_lo = float(ln2 - mp.mpf(ln2_hi))
print("ln2_hi", ln2_hi.hex(), "ln2_lo", ln2_lo.hex(), "inv_ln2", float(1 / ln2).hex())
pio2 = mp.pi / 2
p1 = trunc_bits(pio2, 32)
p2 = trunc_bits(pio2 - mp.mpf(p1), 32)
p3 = float(pio2 - mp.mpf(p1) - mp.mpf(p2))
print("pio2_1", p1.hex(), "pio2_2", p2.hex(), "pio2_3", p3.hex(), "two_over_pi", float(2 / mp.pi).hex())
for nm, v in [("atan(0.5)", mp.atan(mp.mpf(1) / 2)), ("atan(1)", mp.atan(1)), ("atan(1.5)", mp.atan(mp.mpf(3) / 2)), ("pi/2", mp.pi / 2), ("pi", mp.pi)]:
    hi = float(v)
    lo = float(v - mp.mpf(hi))
    print(nm, "hi", hi.hex(), "lo", lo.hex())
