"""An FFT-free, table-free anchor for the whole path ("sign-convention KAT"):
the explicit harmonic network the reference's `sc100` stiffness kernel stands for.

Simple cubic lattice, unit nearest-neighbour and next-nearest-neighbour (face diagonal)
central-force springs, energy 1/2 (n.(u_i - u_j))^2 per spring, periodic nx x ny; the surface
layer, `height` layers below it (towards -z) and a clamped layer underneath.  Prescribing the
surface displacements, relaxing the interior by a dense linear solve and reading off the force on
the surface atoms must give what `fix gfmd` computes as real(IDFT[-Phi(q) DFT[u]]) with
Phi = U0 + V (U + ...)^-1 (-V^H) (src/main/surface_stiffness.cpp:811-873,
src/stiffness_kernels/sc100_stiffness.cpp:170-228) -- which pins, independently of any FFT
library, the forward sign e^{-i q r} (reference README.md:36-38), the orientation (substrate
below), the meaning of `height`, and the stiffness table itself.  With the other FFT sign the same
comparison is off by 10 %."""
import itertools

import numpy as np


def surface_force(u0, height):
    """u0: [3, nx, ny] surface displacements -> force on the surface atoms [3, nx, ny]."""
    d, nx, ny = u0.shape
    assert d == 3
    nlay = height + 1                                     # free layers 0 .. height; height + 1 is clamped
    n = nx * ny * nlay

    def idx(ix, iy, lay):
        return (lay * nx + ix % nx) * ny + iy % ny

    K = np.zeros((3 * n, 3 * n))
    bonds = [b for b in itertools.product((-1, 0, 1), repeat=3) if 0 < sum(c * c for c in b) <= 2]
    for lay in range(nlay):
        for ix in range(nx):
            for iy in range(ny):
                i = idx(ix, iy, lay)
                for bx, by, bz in bonds:
                    other = lay + bz                      # bz = +1: one layer deeper, i.e. towards -z
                    if other < 0:
                        continue                          # vacuum above the surface
                    nvec = np.array([bx, by, -bz], dtype=float)
                    nvec /= np.linalg.norm(nvec)
                    kk = np.outer(nvec, nvec)
                    K[3 * i:3 * i + 3, 3 * i:3 * i + 3] += kk
                    if other <= height:
                        j = idx(ix + bx, iy + by, other)
                        K[3 * i:3 * i + 3, 3 * j:3 * j + 3] -= kk
    ns = 3 * nx * ny
    us = np.moveaxis(u0, 0, -1).reshape(ns)               # surface dofs ordered (ix, iy, component)
    u_int = np.linalg.solve(K[ns:, ns:], -K[ns:, :ns] @ us)
    f = -(K[:ns, :ns] @ us + K[:ns, ns:] @ u_int)
    return np.moveaxis(f.reshape(nx, ny, 3), -1, 0)
