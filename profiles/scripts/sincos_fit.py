"""Coefficients and accuracy of the sin / cos kernels of csrc/update_tc.cu (reduction by pi, polynomials in r^2 on [-pi/2, pi/2]).
Lawson-reweighted least squares on Chebyshev nodes (near-minimax), then the fp32 evaluation error against fp64 over |x| < 60."""
import numpy as np

h = np.pi / 2


def fit(f, deg, n=4000):
    k = np.arange(n)
    u = (np.cos(np.pi * (k + 0.5) / n) + 1) / 2 * h * h
    V = np.vander(u, deg + 1, increasing=True)
    y = f(u)
    w = np.ones(n)
    for _ in range(200):
        c, *_ = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)
        e = np.abs(V @ c - y)
        w = w * (0.5 + e / e.max())
        w /= w.max()
    return c


if __name__ == '__main__':
    sin_c = fit(lambda u: np.where(u > 0, np.sin(np.sqrt(u)) / np.sqrt(np.maximum(u, 1e-300)), 1.0), 4)
    cos_c = fit(lambda u: np.cos(np.sqrt(u)), 5)
    print('sin r * P(r^2):', [float(x) for x in sin_c])
    print('cos Q(r^2):   ', [float(x) for x in cos_c])
    f = np.float32
    x = np.random.default_rng(0).uniform(-60, 60, 2000000).astype(f)
    k = np.rint(x.astype(np.float64) * 0.318309886)
    r = (x.astype(np.float64) - k * 3.140625 - k * 9.67502593994140625e-4 - k * 1.509957990978376432e-7).astype(f)
    r2 = (r * r).astype(f)

    def horner(c):
        acc = np.full_like(r2, f(c[-1]))
        for cc in c[-2::-1]:
            acc = (acc * r2 + f(cc)).astype(f)
        return acc
    sign = np.where(k.astype(np.int64) & 1, -1, 1).astype(f)
    s, c = (r * horner(sin_c)).astype(f) * sign, horner(cos_c) * sign
    print('max abs error: sin %.3e cos %.3e' % (np.abs(s - np.sin(x.astype(np.float64))).max(), np.abs(c - np.cos(x.astype(np.float64))).max()))
