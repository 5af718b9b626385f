"""Checks of the off-path solver services (q-space dump fields, prec_gradient), shared by the
GPU tests (tests/test_aux_services.py) and the CPU-emulation tests
(tests/test_emulated_kernels.py).  B is the gfmd_b200 module bound to either library."""
import os

import numpy as np

from conftest import GOLDEN_DIR, load_golden, rel_err

TOL = 1e-11          # BASELINE.json north_star: <= 1e-11 relative
TOL_TEXT = 2e-10     # fields parsed back from the reference's " %20.10e " dump files


def aux_names():
    import glob
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "aux", "*.npz")))


def load_aux(name):
    z = np.load(os.path.join(GOLDEN_DIR, "aux", name + ".npz"))
    return {k: z[k] for k in z.files}


def _solver(B, g):
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(g["phi"], g["linf"])
    return s, nx, ny, d


def check_spectrum(B, O, name):
    """u~(q) and Phi(q).u~(q) against the numpy oracle, and -- where the reference's dump files
    were recorded -- every field of GFMDSolverFFT::dump against them."""
    g, a = load_golden(name), load_aux(name)
    s, nx, ny, d = _solver(B, g)
    u = np.ascontiguousarray(g["u_uniform"].reshape(d, nx * ny))
    uq, fq = s.spectrum(u)
    uq_ref, fq_ref = O.spectrum(g["u_uniform"], g["phi"])
    assert rel_err(uq, uq_ref) < TOL
    assert rel_err(fq, fq_ref) < TOL
    uq_only, none = s.spectrum(u, with_force=False)
    assert none is None and np.array_equal(uq_only, uq)
    fields = O.dump_fields(uq, fq, nx, ny)
    recorded = sorted(k[5:] for k in a if k.startswith("dump_"))
    if recorded:
        assert recorded == sorted(fields)
    for k in recorded:
        assert rel_err(fields[k], a["dump_" + k]) < TOL_TEXT, k
    # the force of an ordinary step is unaffected by an interleaved spectrum request
    f = np.full_like(u, np.nan)
    e = s.post_force(u, f)
    assert rel_err(f.reshape(d, nx, ny), g["f_uniform"]) < TOL
    assert abs(e - float(g["epot_uniform"])) <= TOL * abs(float(g["epot_uniform"]))
    s.close()


def check_prec_gradient(B, O, name):
    """gP against the reference's own GFMDSolverStatic::prec_gradient output (recorded) --
    with the reference's ndof > 3 behaviour -- and, all components, against the oracle."""
    g, a = load_golden(name), load_aux(name)
    s, nx, ny, d = _solver(B, g)
    grad = np.ascontiguousarray(g["u_uniform"].reshape(d, nx * ny))
    gP = np.full_like(grad, np.nan)
    s.prec_gradient(a["cavg"], grad, gP)                       # reference_quirk=True
    assert rel_err(gP.reshape(d, nx, ny), a["gP"]) < TOL
    gP2 = np.full_like(grad, np.nan)
    s.prec_gradient(a["cavg"], grad, gP2, reference_quirk=False)
    ref2 = O.prec_gradient(g["u_uniform"], g["phi"], a["cavg"])
    assert rel_err(gP2.reshape(d, nx, ny), ref2) < TOL
    if d == 3:
        assert np.array_equal(gP, gP2)
    else:
        assert rel_err(gP2.reshape(d, nx, ny), a["gP"]) > 1e-3       # the quirk is not a no-op
    s.close()


def check_large_column_sets(B, O, nx, ny, d):
    """Column sets beyond one CTA's shared memory (e.g. the 4096-wide bench surface): the services run
    through the three-phase column stage -- transform phase, per-q kernel on the spectrum in HBM,
    transform phase -- whatever table layout the per-step kernels of the handle use."""
    rng = np.random.default_rng(nx + ny + d)
    Dr = rng.standard_normal((nx, ny, d, d)) * np.exp(-3 * rng.random((nx, ny, 1, 1)))
    Dm = Dr[(-np.arange(nx)) % nx][:, (-np.arange(ny)) % ny]
    Dr = 0.5 * (Dr + np.swapaxes(Dm, 2, 3))
    phi = np.fft.fft2(Dr, axes=(0, 1)).reshape(nx * ny, d, d) / (nx * ny)
    phi += np.eye(d)[None] * (2.0 / (nx * ny))               # well conditioned for the preconditioner
    u = rng.uniform(-0.1, 0.1, size=(d, nx, ny))
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(phi, np.zeros(d // 3))
    uu = np.ascontiguousarray(u.reshape(d, nx * ny))
    uq, fq = s.spectrum(uu)
    uq_ref, fq_ref = O.spectrum(u, phi)
    assert rel_err(uq, uq_ref) < TOL and rel_err(fq, fq_ref) < TOL
    if d in (3, 6, 9, 12):
        cavg = 0.01 * rng.standard_normal((d, d)) / (nx * ny)
        gP = np.full_like(uu, np.nan)
        s.prec_gradient(cavg, uu, gP, reference_quirk=False)
        gP_ref = O.prec_gradient(u, phi, cavg, reference_quirk=False)
        assert rel_err(gP.reshape(d, nx, ny), gP_ref) < 1e-9
    # the per-step path is unaffected
    f = np.full_like(uu, np.nan)
    e = s.post_force(uu, f)
    f_ref, e_ref, _ = O.post_force(u, phi, np.zeros(d // 3))
    assert rel_err(f.reshape(d, nx, ny), f_ref) < TOL and abs(e - e_ref) <= TOL * abs(e_ref)
    s.close()
