"""Generates the polynomial schedule hard-coded in csrc/fad.cu (kPolarCoef).

The Frechet distance needs c = sum of singular values of M (fad.cu).  With the polar
decomposition M = U H, c = tr(U^T M), and U is the limit of  X <- X p_k(X^T X)  started
from X_0 = M / |M|_F for odd polynomials p_k that push every singular value in
[l_0, 1] to 1.  Each p_k(x) = x (a + b x^2 + c x^4) is the quintic that maximises the
image of the current lower bound, l_{k+1} = min p_k over [l_k, 1 + eta], subject to
p_k <= 1 on [l_k, 1 + eta] (a linear program on a grid).  The margin eta above 1 keeps a
singular value that rounding pushed slightly past 1 inside the design interval: without
it the slope p'(1) ~ 13 of the early polynomials amplifies that excess step after step.
With eta > 0 the greedy bound stalls near 0.95, so once l_k >= 0.5 the classical quintic
Newton-Schulz step x (15 - 10 x^2 + 3 x^4) / 8 (cubically convergent, p'(1) = p''(1) = 0)
finishes the job.

    python polar_schedule.py [l0=1e-12] [eta=1/64]
"""
import sys

import numpy as np
import scipy.optimize as so


def best_quintic(lo, eta, ngrid=4000):
    u = 1 + eta
    x = np.unique(np.concatenate([np.geomspace(lo, u, ngrid), np.linspace(lo, u, ngrid), np.linspace(0.5, u, ngrid)]))
    q = np.stack([np.ones_like(x), x**2, x**4], 1)
    # variables (a, b, c, s):  q(x) >= s lo / x   and   x q(x) <= 1;  maximise s
    g = np.concatenate([np.concatenate([-q, (lo / x)[:, None]], 1),
                        np.concatenate([q * x[:, None], np.zeros((len(x), 1))], 1)])
    h = np.concatenate([np.zeros(len(x)), np.ones(len(x))])
    r = so.linprog([0, 0, 0, -1], A_ub=g, b_ub=h, bounds=[(None, None)] * 3 + [(0, 1 / lo)], method="highs")
    a, b, c, s = r.x
    return (a, b, c), s * lo


def main():
    lo = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-12
    eta = float(sys.argv[2]) if len(sys.argv) > 2 else 1 / 64
    steps = []
    while lo < 0.5:
        co, t = best_quintic(lo, eta)
        steps.append(co)
        print("    {%.17g, %.17g, %.17g},   // l: %.3e -> %.3e" % (*co, lo, t))
        lo = t
    e = 1 - lo
    n_tail = 0
    while e > 1e-15:   # classical quintic: 1 - p(1 - e) = (5/2) e^3 + O(e^4) near 1; evaluate exactly
        x = 1 - e
        e = 1 - x * (15 - 10 * x**2 + 3 * x**4) / 8
        n_tail += 1
        print("    {15.0 / 8, -10.0 / 8, 3.0 / 8},   // 1 - l -> %.3e" % e)
    print(len(steps) + n_tail, "steps")


if __name__ == "__main__":
    main()
