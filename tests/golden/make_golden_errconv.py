"""Makes tests/golden/compound/errconv_fcc100_two_layers_lj.npz (about ten minutes; needs oracle/_ref, i.e. /root/reference):
the stiffness table of the reference plugin for the Lennard-Jones fcc(100) substrate of tests/errconv.py, its
linear forces, and the force changes on the probe atom of the ALL-ATOM twin (plain numpy, no GFMD)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import errconv  # noqa: E402
from oracle import gfmd_oracle as O  # noqa: E402

dsteps = np.array([1e-2, 1e-3])
k = O.RefKernel(errconv.kernel_string())
phi = k.phi(errconv.NX, errconv.NY)
k.close()
dF_full = errconv.full_atom_twin(list(dsteps))
np.savez_compressed(os.path.join(HERE, "compound", "errconv_fcc100_two_layers_lj.npz"), dsteps=dsteps, phi=phi,
                    linf=errconv.linear_forces(), dF_full=dF_full, kernel=errconv.kernel_string(), a=errconv.A_NN)
print(dF_full)
