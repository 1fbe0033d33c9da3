"""Generates the minimax polynomial coefficients of ilqc_b200/csrc/fastmath.cuh (weighted Remez exchange in
50-digit arithmetic, mpmath) and prints them as C++ literals together with the achieved relative error.

    python tools/gen_fastmath_coeffs.py

  sin(r)  = r + r z S(z),            z = r^2, |r| <= pi/4 (+margin)     S: 6 coefficients
  cos(r)  = 1 - z/2 + z^2 C(z)                                          C: 6 coefficients
  atan(t) = t + t z A(z),            z = t^2, |t| <= tan(pi/8) (+margin) A: 12 coefficients
"""
import mpmath as mp

mp.mp.dps = 60


def remez(g, w, a, b, n, iters=30, grid=4000):
    """Polynomial p of degree n in z minimising max |w(z) (p(z) - g(z))| on [a, b]."""
    # initial reference: Chebyshev extrema
    ref = [a + (b - a) * (1 - mp.cos(mp.pi * i / (n + 1))) / 2 for i in range(n + 2)]
    zs = [a + (b - a) * (1 - mp.cos(mp.pi * i / grid)) / 2 for i in range(grid + 1)]
    gz = [g(z) for z in zs]
    wz = [w(z) for z in zs]
    coef = None
    for _ in range(iters):
        A = mp.matrix(n + 2, n + 2)
        rhs = mp.matrix(n + 2, 1)
        for i, z in enumerate(ref):
            for j in range(n + 1):
                A[i, j] = z ** j
            A[i, n + 1] = (-1) ** i / w(z)
            rhs[i] = g(z)
        sol = mp.lu_solve(A, rhs)
        coef = [sol[j] for j in range(n + 1)]
        err = [wz[k] * (mp.polyval(coef[::-1], zs[k]) - gz[k]) for k in range(len(zs))]
        # new reference: one extremum per sign-alternating stretch
        ext = []
        k = 0
        while k < len(zs):
            s = mp.sign(err[k])
            best = k
            while k < len(zs) and (mp.sign(err[k]) == s or err[k] == 0):
                if abs(err[k]) > abs(err[best]):
                    best = k
                k += 1
            ext.append(best)
        if len(ext) < n + 2:
            break
        # keep the n+2 consecutive extrema with the largest minimum
        while len(ext) > n + 2:
            if abs(err[ext[0]]) < abs(err[ext[-1]]):
                ext.pop(0)
            else:
                ext.pop()
        new_ref = [zs[k] for k in ext]
        emax = max(abs(err[k]) for k in ext)
        emin = min(abs(err[k]) for k in ext)
        ref = new_ref
        if emax - emin < emax * mp.mpf("1e-6"):
            break
    return coef


def as_double(c):
    return float(c)


def report(name, coef, f_exact, f_poly, lo, hi, n=20000):
    cd = [as_double(c) for c in coef]
    worst = mp.mpf(0)
    for i in range(1, n + 1):
        x = lo + (hi - lo) * mp.mpf(i) / n
        e = abs((f_poly(cd, x) - f_exact(x)) / f_exact(x))
        worst = max(worst, e)
    print("// %s: %d coefficients, max rel. error of the real-arithmetic polynomial 2^%.1f" %
          (name, len(cd), float(mp.log(worst, 2))))
    for k, c in enumerate(cd):
        print("//   %s[%d] = %s  (%s)" % (name, k, repr(c), float.hex(c)))
    print("static constexpr double k%s[%d] = {%s};" % (name, len(cd), ", ".join(repr(c) for c in cd)))
    return cd


def main():
    q = mp.pi / 4 * mp.mpf("1.002")
    zmax = q * q
    eps = mp.mpf("1e-30")
    # sin
    g = lambda z: (mp.sin(mp.sqrt(z)) / mp.sqrt(z) - 1) / z if z > eps else -mp.mpf(1) / 6 + z / 120
    w = lambda z: (z if z > eps else eps) / (mp.sin(mp.sqrt(z)) / mp.sqrt(z) if z > eps else 1)
    S = remez(g, w, mp.mpf(0), zmax, 5)
    report("Sin", S, mp.sin, lambda c, x: x + x * (x * x) * mp.polyval([mp.mpf(v) for v in c[::-1]], x * x), mp.mpf(0), q)
    # cos
    g = lambda z: (mp.cos(mp.sqrt(z)) - 1 + z / 2) / (z * z) if z > mp.mpf("1e-20") else mp.mpf(1) / 24 - z / 720
    w = lambda z: (z * z if z > eps else eps) / mp.cos(mp.sqrt(z))
    C = remez(g, w, mp.mpf(0), zmax, 5)
    report("Cos", C, mp.cos, lambda c, x: 1 - (x * x) / 2 + (x * x) ** 2 * mp.polyval([mp.mpf(v) for v in c[::-1]], x * x),
           mp.mpf(0), q)
    # atan
    tmax = mp.tan(mp.pi / 8) * mp.mpf("1.002")
    g = lambda z: (mp.atan(mp.sqrt(z)) / mp.sqrt(z) - 1) / z if z > eps else -mp.mpf(1) / 3 + z / 5
    w = lambda z: (z if z > eps else eps) / (mp.atan(mp.sqrt(z)) / mp.sqrt(z) if z > eps else 1)
    for n in (10, 11, 12):
        A = remez(g, w, mp.mpf(0), tmax * tmax, n)
        report("Atan%d" % (n + 1), A, mp.atan,
               lambda c, x: x + x * (x * x) * mp.polyval([mp.mpf(v) for v in c[::-1]], x * x), mp.mpf(0), tmax)


if __name__ == "__main__":
    main()
