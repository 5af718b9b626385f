"""Generates tests/golden/*.npz by running the REFERENCE's own sources
(oracle/_ref/libgfmd_ref.so = /root/reference solver + stiffness kernels compiled
unchanged, FFT3d shim backed by a direct long-double DFT) on seeded inputs.

Run in the build container only (needs /root/reference to build oracle/_ref):
    make -C oracle ref && python tests/golden/make_golden.py

Each file holds: kernel (string), nx, ny, ndof, phi [nx*ny, d, d] complex128
(normalised, fill_phi_buffer), linf, and for every case c: u_c, f_c, epot_c, u0_c.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gfmd_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, kernel string, nx, ny, synthetic linf (fed through the reference arithmetic)
    ("small_sc100_16x12", "ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128", 16, 12, None),
    ("small_fcc111_8x7", "ft fcc111 1 1.0 pair-potential 1 1.0 height 128", 8, 7, [0.03, -0.03]),
    ("small_sc100_9x15_h4", "sc100 height 4", 9, 15, [0.25]),
    ("C3_fcc100_two_layers_10x10", "ft fcc100 1.0 2 pair-potential 2 1.0 -0.1 height 10", 10, 10, None),
    ("C1_sc100_128x128", "ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128", 128, 128, None),
    ("C2_fcc111_64x37", "ft fcc111 1 1.0 pair-potential 1 1.0 height 128", 64, 37, None),
]


def hertz_field(nx, ny, ndof, R=100.0, delta=0.5):
    """u_z = -max(0, delta - r^2/(2R)) around (0,0), periodic; u_x = u_y = 0."""
    ix = np.arange(nx)
    iy = np.arange(ny)
    dx = np.minimum(ix, nx - ix)[:, None].astype(float)
    dy = np.minimum(iy, ny - iy)[None, :].astype(float)
    uz = -np.maximum(0.0, delta - (dx * dx + dy * dy) / (2 * R))
    u = np.zeros((ndof, nx, ny))
    for a in range(ndof // 3):
        u[3 * a + 2] = uz
    return u


def main():
    for name, ks, nx, ny, linf_syn in CASES:
        k = O.RefKernel(ks)
        d = k.ndof
        phi = k.phi(nx, ny)
        linf = k.linf() if linf_syn is None else np.asarray(linf_syn, dtype=float)
        s = O.RefSolver(nx, ny, d, fft_backend=0)
        if linf_syn is None:
            s.set_kernel(k)
        else:
            s.set_phi(phi, linf)
        out = dict(kernel=ks, nx=nx, ny=ny, ndof=d, phi=phi.reshape(nx * ny, d, d), linf=linf)
        rng = np.random.default_rng(12472634)
        fields = {"uniform": rng.uniform(-0.1, 0.1, size=(d, nx, ny)),
                  "hertz": hertz_field(nx, ny, d),
                  "shift": np.full((d, nx, ny), -2.0) * (np.arange(d) % 3 == 2)[:, None, None]}
        for c, u in fields.items():
            f, e, u0 = s.post_force(u)
            out["u_" + c] = u
            out["f_" + c] = f
            out["epot_" + c] = e
            out["u0_" + c] = u0
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "ndof", d, "epot(uniform)", out["epot_uniform"])
        s.close()
        k.close()


if __name__ == "__main__":
    main()
