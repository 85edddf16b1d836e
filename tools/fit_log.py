"""Fit the polynomial of log_fast() (cloudsc_column.cuh): log(m) = 2s + 2 s z g(z), s=(m-1)/(m+1), z=s^2,
g(z) = 1/3 + z/5 + z^2/7 + ...  on z in [0, ((sqrt2-1)/(sqrt2+1))^2].  Chebyshev interpolation in 60-digit
arithmetic (near-minimax); prints coefficients rounded to binary64 and the resulting max relative error of log(m)."""
import mpmath as mp
mp.mp.dps = 60
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
zmax = smax ** 2 * mp.mpf('1.0001')
def g(z):
    if z == 0: return mp.mpf(1) / 3
    s = mp.sqrt(z)
    return (mp.atanh(s) / s - 1) / z
for deg in (5, 6, 7):
    n = deg + 1
    nodes = [zmax / 2 * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) for k in range(n)]
    A = mp.matrix(n, n); b = mp.matrix(n, 1)
    for i, z in enumerate(nodes):
        for j in range(n): A[i, j] = z ** j
        b[i] = g(z)
    c = mp.lu_solve(A, b)
    cf = [float(x) for x in c]
    # error of log(m) relative
    worst = 0
    for k in range(1, 2000):
        s = smax * k / 2000
        z = s * s
        p = sum(mp.mpf(cf[j]) * z ** j for j in range(n))
        approx = 2 * s + 2 * s * z * p
        exact = 2 * mp.atanh(s)
        worst = max(worst, abs(approx / exact - 1))
    print(deg, mp.nstr(worst, 3))
    print("  ", ", ".join(repr(x) for x in cf))
