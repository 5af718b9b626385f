"""CPU only: the reference's Hertz acceptance criterion evaluated WITHOUT this repository's CUDA path --
the plugin's own stiffness table (oracle/_ref), numpy.fft for the force, the same rigid 12-6 sphere and
FIRE iteration as tests/test_compound.py.  It separates "the restated test is sound" from "the kernels
are right": DESIGN.md section 6b quotes its residuals.

    python tools/hertz_numpy_check.py sc100 | fcc100 | sc100_a0_1.3        (1.5 - 4 minutes each)
"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gfmd_oracle as O  # noqa: E402

CASES = {   # name: (kernel, lattice constant, contact modulus, dmax)   -- tests/TEST_Hertz_*/lammps.in, eval.py
    "sc100": ("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128", 1.0, 8.0 / 3, None),
    "fcc100": ("ft fcc100 1 1.0 pair-potential 1 1.0 height 128", 1.0, 1.39, None),
    "sc100_a0_1.3": ("ft sc100 1.3 1 pair-potential 2 1.0 1.0 height 128", 1.3, 8.0 / 3 / 1.3, 0.1),
}


def main():
    kernel, a0, E, dmax = CASES[sys.argv[1]]
    nx = ny = 128
    k = O.RefKernel(kernel)
    Phi = k.phi(nx, ny).reshape(nx, ny, 3, 3) * (nx * ny)
    linf = k.linf()
    k.close()
    gid = np.array([(ix, iy) for ix in range(nx) for iy in range(ny)])
    xeq = np.stack([gid[:, 0] * a0 + 0.5, gid[:, 1] * a0 + 0.5, np.full(len(gid), 0.5)], 1)
    xprd, yprd = nx * a0, ny * a0
    cz, R = 99.5, 100.0
    eps, sig, cut = 1.38888888888889, 0.890898718140339, 1.0
    c1, c2 = 48 * eps * sig ** 12, 24 * eps * sig ** 6

    def force(x):
        u = x - xeq
        u[:, 0] -= xprd * np.round(u[:, 0] / xprd)
        u[:, 1] -= yprd * np.round(u[:, 1] / yprd)
        fq = -np.einsum("xyij,xyj->xyi", Phi, np.fft.fft2(u.reshape(nx, ny, 3), axes=(0, 1)))
        fq[0, 0, 2] += linf[0] * nx * ny
        fg = np.real(np.fft.ifft2(fq, axes=(0, 1))).reshape(-1, 3)
        f = fg.copy()
        rx, ry, rz = x[:, 0].copy(), x[:, 1].copy(), x[:, 2] - cz
        rx -= xprd * np.round(rx / xprd)
        ry -= yprd * np.round(ry / yprd)
        r = np.sqrt(rx * rx + ry * ry + rz * rz)
        rinv = 1 / np.maximum(r - R, 1e-300)
        r6 = rinv ** 6
        df = np.where(r < R + cut, r6 * (c1 * r6 - c2) * rinv, 0.0)
        f += (df / r)[:, None] * np.stack([rx, ry, rz], 1)
        return f, fg

    x = xeq.copy()
    x[:, 2] -= 2.0
    v = np.zeros_like(x)
    dt, dtmax, alpha, npos = 0.05, 0.25, 0.1, 0
    t0 = time.time()
    for it in range(200000):
        f, fg = force(x)
        if it % 50 == 0:
            fn = np.linalg.norm(f)
            if not np.isfinite(fn) or fn <= 1e-6:
                break
        if (f * v).sum() > 0:
            v = (1 - alpha) * v + f * (alpha * np.linalg.norm(v) / np.linalg.norm(f))
            npos += 1
            if npos > 5:
                dt = min(dt * 1.1, dtmax)
                alpha *= 0.99
        else:
            v[:] = 0
            dt *= 0.5
            alpha, npos = 0.1, 0
        v += dt * f
        if dmax is not None:
            m = np.abs(v).max() * dt
            if m > dmax:
                v *= dmax / m
        x += dt * v
    f_xy = fg[:, 2].reshape(nx, ny)
    xs = np.arange(nx) + 0.5
    xs = np.where(xs > nx / 2, xs - nx, xs) * a0
    r_xy = np.sqrt((xs ** 2).reshape(-1, 1) + (xs ** 2).reshape(1, -1))
    N = f_xy.sum()
    a = R * (3. / 4 * (N / (E * R ** 2))) ** (1. / 3)
    p0 = 3 * N / (2 * math.pi * a * a)
    pa = np.where(r_xy < a, p0 * np.sqrt(np.maximum(0, 1 - (r_xy / a) ** 2)), 0.0)
    res = np.sum((f_xy / (a0 * a0) - pa) ** 2)
    print("%s: |f| %.2e after %d iterations (%.0f s); N %.4f a %.3f p0 %.4f; residual %.4f (bound 1e-2) %s"
          % (sys.argv[1], fn, it, time.time() - t0, N, a, p0, res, "ok" if res < 1e-2 else "FAIL"))


if __name__ == "__main__":
    main()
