"""Generates tests/golden/aux/*.npz: outputs of the REFERENCE's own sources
(oracle/_ref/libgfmd_ref.so, long-double DFT backend) for the off-path services of the solver,
on the tables and fields of the golden files next door:

  gP          GFMDSolverStatic::prec_gradient(cavg, g)      (gfmd_solver_static.cpp:253-271)
  dump_<f>    every <prefix>.q.<f>.out file GFMDSolverFFT::dump wrote for post_force(u, f,
              prefix), parsed back ([ny, nx]; the files carry 11 significant digits)
                                                             (gfmd_solver_fft.cpp:209-287)

Run in the build container only:  make -C oracle ref && python tests/golden/make_golden_aux.py
"""
import glob
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gfmd_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "aux")

# base golden file, diagonal of cavg (times 1/(nx*ny), the scale of the stored table), dump?
CASES = [("small_sc100_16x12", 0.5, True), ("small_fcc111_8x7", 1.0, True),
         ("C3_fcc100_two_layers_10x10", 1.0, True), ("C2_fcc111_64x37", 2.0, False),
         ("C1_sc100_128x128", 0.25, False)]


def main():
    os.makedirs(OUT, exist_ok=True)
    for base, cdiag, dump in CASES:
        z = np.load(os.path.join(HERE, base + ".npz"))
        nx, ny, d = int(z["nx"]), int(z["ny"]), int(z["ndof"])
        s = O.RefSolver(nx, ny, d, fft_backend=0)
        s.set_phi(z["phi"], z["linf"])
        rng = np.random.default_rng(20261017)
        cavg = (cdiag * np.eye(d) + 0.05 * cdiag * rng.standard_normal((d, d))) / (nx * ny)
        g = z["u_uniform"]
        out = dict(base=base, cavg=cavg, gP=s.prec_gradient(cavg, g))
        if dump:
            with tempfile.TemporaryDirectory() as td:
                s.post_force_dump(g, os.path.join(td, "dump"))
                for p in sorted(glob.glob(os.path.join(td, "dump.q.*.out"))):
                    field = os.path.basename(p)[len("dump.q."):-len(".out")]
                    out["dump_" + field] = np.loadtxt(p, ndmin=2)
        np.savez_compressed(os.path.join(OUT, base + ".npz"), **out)
        print(base, "gP max", np.abs(out["gP"]).max(), "fields", sum(k.startswith("dump_") for k in out))
        s.close()


if __name__ == "__main__":
    main()
