"""The reference's error-convergence acceptance test, restated without LAMMPS
(tests/TEST_error_convergence_ft_two_layers_lj_cut: in.structure, in.gf, in.full, eval.py:207-226).

An fcc(100) Lennard-Jones crystal (nearest-neighbour distance a, cutoff 1.75: two neighbour shells), 8 x 8
surface cells, free top surface.  GF system: 12 atomic layers, the two lowest are the GFMD layers (two atoms per
cell, ndof 6) on a 40-layer harmonic substrate -- stiffness kernel
`ft fcc100 a 2 pair-potential 2x k1 kappa1 k2 kappa2 height 20` with k = V''(r), kappa = V'(r) / r of the LJ
potential, i.e. what the reference's `lj/cut` force constants evaluate (src/force_constants/fc_lj_cut.cpp:88-146),
and the linear forces of fc_lj_cut.cpp:149-174; pairs of two GFMD atoms are skipped by the pair style
(pair_lj_cut_gf.cpp:131-132).  All-atom twin: the same crystal with all 54 layers explicit, the two lowest fixed.
In both, everything is relaxed, then ONE atom of the lowest GFMD layer (the reference's `probe1`) is moved by
(-dstep, 0, -dstep) and held while the rest relaxes again; the change of the force on it is compared.  GFMD is the
harmonic expansion of the twin: the difference must shrink like dstep^2 (criterion of eval.py: exponent >= 1.9
between a small and a large step, and < 10 % relative error at the large one).

The twin needs no stiffness kernel and no FFT: its force changes are a golden fixture
(tests/golden/compound/errconv_fcc100_two_layers_lj.npz, made by tests/golden/make_golden_errconv.py from this file's
plain-numpy Lennard-Jones code), together with the kernel's table from the reference plugin.
"""
import numpy as np

A_NN = 1.1126198391757889      # zero-pressure nearest-neighbour distance of the two-shell LJ crystal (in.structure: 1.11262)
NX = NY = 8
RC = 1.75
NFULL, NGF = 54, 12


def dV(r):
    return 4.0 * (-12.0 * r**-13 + 6.0 * r**-7)


def d2V(r):
    return 4.0 * (156.0 * r**-14 - 42.0 * r**-8)


def kernel_string():
    a, r2 = A_NN, A_NN * np.sqrt(2.0)
    return ("ft fcc100 %.16f 2 pair-potential 2x %.16f %.16f %.16f %.16f height 20"
            % (a, d2V(a), dV(a) / a, d2V(r2), dV(r2) / r2))


def build(nlayers):
    """Layer l at height l a / sqrt 2, shifted by (a/2, a/2) on odd layers; atoms of a layer in (ix, iy) order."""
    a, h = A_NN, A_NN / np.sqrt(2.0)
    pos, lay = [], []
    for l in range(nlayers):
        off = 0.5 * a * (l % 2)
        for ix in range(NX):
            for iy in range(NY):
                pos.append((ix * a + off + 0.25 * a, iy * a + off + 0.25 * a, l * h))
                lay.append(l)
    return np.array(pos), np.array(lay)


def neighbor_pairs(x, rlist=2.2):
    L = np.array([NX * A_NN, NY * A_NN])
    n = len(x)
    I, J = [], []
    for i in range(n):
        d = x - x[i]
        d[:, :2] -= L * np.round(d[:, :2] / L)
        r = np.sqrt((d**2).sum(1))
        j = np.nonzero((r < rlist) & (np.arange(n) > i))[0]
        I += [i] * len(j)
        J += list(j)
    return np.array(I), np.array(J)


def lj_forces(x, I, J, skip=None):
    L = np.array([NX * A_NN, NY * A_NN])
    d = x[J] - x[I]
    d[:, :2] -= L * np.round(d[:, :2] / L)
    r = np.sqrt((d**2).sum(1))
    m = r < RC
    if skip is not None:
        m &= ~skip
    rs = np.where(m, r, 1.0)
    fvec = np.where(m, -dV(rs) / rs, 0.0)[:, None] * d         # force on j
    n = len(x)
    return np.stack([np.bincount(J, fvec[:, c], n) - np.bincount(I, fvec[:, c], n) for c in range(3)], axis=1)


def fire(force, x, free, ftol, maxit=100000, dt0=0.02, dtmax=0.2):
    """FIRE relaxation to max |f| < ftol on the free atoms."""
    v = np.zeros_like(x)
    dt, alpha, npos = dt0, 0.1, 0
    for it in range(maxit):
        f = force(x) * free
        if np.abs(f).max() < ftol:
            return x, it
        if (f * v).sum() > 0:
            v = (1 - alpha) * v + alpha * f * np.sqrt((v**2).sum() / max((f**2).sum(), 1e-300))
            npos += 1
            if npos > 5:
                dt = min(dt * 1.1, dtmax)
                alpha *= 0.99
        else:
            v[:] = 0
            dt *= 0.5
            alpha, npos = 0.1, 0
        v += dt * f
        x = x + dt * v
    raise RuntimeError("FIRE did not converge")


def probe_experiment(force, x_start, free, probe, dsteps, ftol):
    """Relax; then for every dstep move the probe by (-dstep, 0, -dstep), hold it, relax the rest: change of the
    force on the probe (eval.py subtracts the force of the relaxed start)."""
    xr, _ = fire(force, x_start.copy(), free, ftol)
    f_init = force(xr)[probe].copy()
    out = []
    for d in dsteps:
        x = xr.copy()
        x[probe] += (-d, 0.0, -d)
        fr = free.copy()
        fr[probe] = 0.0
        x, _ = fire(force, x, fr, ftol)
        out.append(force(x)[probe] - f_init)
    return np.array(out)


def full_atom_twin(dsteps, ftol=1e-10):
    """in.full: all layers explicit, the two lowest fixed."""
    x0, lay = build(NFULL)
    I, J = neighbor_pairs(x0)
    free = np.ones_like(x0)
    free[lay < 2] = 0.0
    probe = np.nonzero(lay == NFULL - NGF)[0][4 * NY + 4]
    return probe_experiment(lambda x: lj_forces(x, I, J), x0, free, probe, dsteps, ftol)


def gf_system(solver_step):
    """in.gf.  solver_step(u [6, NX*NY]) -> f [6, NX*NY]: one GFMD force evaluation (linf included)."""
    x0, lay = build(NFULL)
    sel = lay >= NFULL - NGF
    x0, lay = x0[sel], lay[sel] - (NFULL - NGF)
    I, J = neighbor_pairs(x0)
    isgf = lay < 2
    skip = isgf[I] & isgf[J]
    up, lo = np.nonzero(lay == 1)[0], np.nonzero(lay == 0)[0]        # iu = 0: upper GFMD layer, iu = 1: lower

    def force(x):
        f = lj_forces(x, I, J, skip)
        u = np.empty((6, NX * NY))
        u[0:3] = (x[up] - x0[up]).T
        u[3:6] = (x[lo] - x0[lo]).T
        fg = solver_step(np.ascontiguousarray(u))
        f[up] += fg[0:3].T
        f[lo] += fg[3:6].T
        return f
    return x0, force, lo[4 * NY + 4]


def linear_forces():
    """Minus the z-force of the explicit atoms on the GFMD atoms at the ideal positions (equal, by the symmetry of
    the bulk, to fc_lj_cut.cpp:149-174: the z-force of the neighbours in the same surface cell and below)."""
    x0, lay = build(NFULL)
    sel = lay >= NFULL - NGF
    x0, lay = x0[sel], lay[sel] - (NFULL - NGF)
    I, J = neighbor_pairs(x0)
    isgf = lay < 2
    f0 = lj_forces(x0, I, J, isgf[I] & isgf[J])
    l0, l1 = -f0[lay == 1, 2].mean(), -f0[lay == 0, 2].mean()
    return np.array([0.5 * (l0 - l1), -0.5 * (l0 - l1)])


def check(gfmd_b200, golden, ftol=1e-10):
    """Runs in.gf on the CUDA path (library bound to gfmd_b200) and applies eval.py's criterion against the
    golden all-atom force changes."""
    dsteps = [float(d) for d in golden["dsteps"]]
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(NX, NY, 6)
    s.set_kernel(golden["phi"], golden["linf"])
    fbuf = np.empty((6, NX * NY))

    def step(u):
        s.post_force(u, fbuf)
        return fbuf.copy()
    x0, force, probe = gf_system(step)
    assert np.abs(force(x0)[:2 * NX * NY]).max() < 1e-10          # the GFMD layers of the ideal crystal are in equilibrium
    dF = probe_experiment(force, x0, np.ones_like(x0), probe, dsteps, ftol)
    s.close()
    err = np.sqrt(((dF - golden["dF_full"])**2).sum(1))
    mag = np.sqrt((golden["dF_full"]**2).sum(1))
    i_large = [i for i, d in enumerate(dsteps) if 1e-2 <= d < 4e-2][0]
    i_small = [i for i, d in enumerate(dsteps) if 1e-4 <= d < 2e-3][0]
    assert err[i_large] / mag[i_large] < 0.1, "more than 10 %% error at dstep %g" % dsteps[i_large]
    goeslike = np.log(err[i_small] / err[i_large]) / np.log(dsteps[i_small] / dsteps[i_large])
    assert 2.0 - goeslike <= 0.1, "force error goes like dstep^%.3f, not quadratically" % goeslike
    return goeslike, err / mag
