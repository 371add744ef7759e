"""Generate the coefficient tables of ccgp_math.h's K0/K1 (modified Bessel functions of the second
kind) with 50-digit mpmath: exact power-series coefficients for 0 < t <= 2 and Chebyshev
coefficients of sqrt(2t/pi) e^t K_nu(t) in s = 4/t - 1 for t > 2.  Prints C initialisers."""
import mpmath as mp

mp.mp.dps = 50
NS = 16     # series terms (t <= 2: (t^2/4)^k / (k!)^2 <= 1/(k!)^2)
print("// I0 series: sum a0[k] z^k, z = t^2/4")
print("static const double BES_I0[%d] = {%s};" % (NS, ", ".join(repr(float(1 / mp.factorial(k) ** 2)) for k in range(NS))))
print("static const double BES_K0P[%d] = {%s};" % (NS, ", ".join(repr(float(mp.digamma(k + 1) / mp.factorial(k) ** 2)) for k in range(NS))))
print("static const double BES_I1[%d] = {%s};" % (NS, ", ".join(repr(float(1 / (mp.factorial(k) * mp.factorial(k + 1)))) for k in range(NS))))
print("static const double BES_K1P[%d] = {%s};" % (NS, ", ".join(repr(float((mp.digamma(k + 1) + mp.digamma(k + 2)) / (mp.factorial(k) * mp.factorial(k + 1)))) for k in range(NS))))


def cheb(nu, N, M=200):
    f = lambda s: mp.sqrt(2 * (4 / (s + 1)) / mp.pi) * mp.e ** (4 / (s + 1)) * mp.besselk(nu, 4 / (s + 1))
    th = [mp.pi * (k + mp.mpf(1) / 2) / M for k in range(M)]
    fv = [f(mp.cos(t)) for t in th]
    return [2 / mp.mpf(M) * sum(fv[k] * mp.cos(j * th[k]) for k in range(M)) for j in range(N)]


for nu in (0, 1):
    c = cheb(nu, 40)
    # keep terms until they drop below 1e-18
    n = max(j for j in range(40) if abs(c[j]) > mp.mpf("1e-18")) + 1
    print("static const double BES_K%dA[%d] = {%s};  // Chebyshev, c[0] is halved at evaluation" % (nu, n, ", ".join(repr(float(x)) for x in c[:n])))
