"""The reference's compound acceptance tests, restated without LAMMPS and run on the
CUDA path.  Criteria are the reference's own:

* tests/TEST_energy_conservation_two_layers (lammps.in, eval.py:34-40): 10 x 10 grid,
  two-layer fcc100 kernel (ndof 6), random displacements 0.1 (seed 12472634),
  v = (0.1, 0, 0), `fix nve`, dt = 0.01, 100 000 steps, thermo every 1000:
  max|E - <E>| / <E> <= 1e-4.
* tests/TEST_Hertz_sc100_128x128 (lammps.in, eval.py:33-39,73-86): rigid sphere R = 100
  (fix contact/sphere: 12-6 wall in r - R, src/extras/fix_contact_sphere.cpp:130-141,
  :207-264) pressed on the 128 x 128 sc100 layer by moving all atoms 2.0 down, minimised
  to |f| <= 1e-6; final z-force map vs. the Hertz profile with E* = 8/3:
  sum (f - f_Hertz)^2 < 1e-2.

The stiffness tables are the reference plugin's own (tests/golden, produced by
`ft fcc100 1.0 2 pair-potential 2 1.0 -0.1 height 10` and
`ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128`).  The integrator / minimiser
(LAMMPS' job in the reference) are a few torch element-wise lines here; every force and
energy comes from gather -> solver -> scatter of libgfmd_b200."""
import math

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def make_layer_atoms(nx, ny, nu):
    gid = np.array([(ix, iy, iu) for ix in range(nx) for iy in range(ny) for iu in range(nu)], dtype=np.int32)
    xeq = np.stack([gid[:, 0] + 0.5, gid[:, 1] + 0.5, 0.5 - gid[:, 2]], axis=1).astype(np.float64)
    return gid, xeq


def test_energy_conservation_two_layers():
    import torch
    import gfmd_b200
    run_energy_conservation_two_layers(gfmd_b200, torch.device("cuda"), 100000, 1000)


def plugin_table(O, kernel, nx, ny):
    """Phi table and linf of a reference stiffness kernel (oracle/_ref, the plugin's own sources)."""
    k = O.RefKernel(kernel)
    t = {"nx": nx, "ny": ny, "ndof": k.ndof, "phi": k.phi(nx, ny), "linf": k.linf()}
    k.close()
    return t


def bind_stream(s, dev):
    """On a GPU the library's work is ordered with torch's current stream; under the CPU
    emulation build (tests/test_emulated_kernels.py) everything is synchronous."""
    import torch
    if dev.type == "cuda":
        s.set_stream(torch.cuda.current_stream().cuda_stream)


def run_energy_conservation_two_layers(gfmd_b200, dev, nsteps, every):
    run_energy_conservation(gfmd_b200, dev, nsteps, every, table=load_golden("C3_fcc100_two_layers_10x10"), vx=0.1,
                            expect_shift=True)


def run_energy_conservation(gfmd_b200, dev, nsteps, every, table, vx, expect_shift):
    import torch
    g = table
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    gid, xeq = make_layer_atoms(nx, ny, d // 3)
    n = gid.shape[0]
    rng = np.random.default_rng(12472634)
    x0 = xeq + rng.uniform(-0.1, 0.1, size=(n, 3))           # displace_atoms all random 0.1 0.1 0.1
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(g["phi"], g["linf"])
    bind_stream(s, dev)                                       # library work is ordered with torch's
    x = torch.tensor(x0, device=dev)
    dxeq = torch.tensor(xeq, device=dev)
    dgid = torch.tensor(gid, device=dev)
    dmask = torch.ones(n, dtype=torch.int32, device=dev)
    v = torch.zeros((n, 3), device=dev, dtype=torch.float64)
    v[:, 0] = vx                                             # velocity all set 0.1 0.0 0.0 (two layers)
    f = torch.zeros((n, 3), device=dev, dtype=torch.float64)
    dt, mass = 0.01, 1.0

    shift = [0, 0]                                           # FixGFMD xshift / yshift

    def force(check_shift=False):
        # fix_gfmd.cpp:695-715: when the layer's centre of mass has moved by a lattice constant
        # the grid indices are shifted (gid rewritten by the gather); physics is unchanged
        dxs = dys = 0
        if check_shift:
            dcm = (x - dxeq).mean(dim=0).cpu().numpy()
            cur = [int(dcm[0] - 0.5) if dcm[0] < 0 else int(dcm[0] + 0.5),
                   int(dcm[1] - 0.5) if dcm[1] < 0 else int(dcm[1] + 0.5)]      # the fix's nearbyint macro
            dxs, dys = shift[0] - cur[0], shift[1] - cur[1]
            shift[0], shift[1] = cur
        f.zero_()
        s.gather(x, dxeq, dgid, dmask, 1, n, float(nx), float(ny), dxs, dys)
        s.post_force_device()
        s.scatter(dgid, dmask, 1, n, n, f)
        return dxs != 0 or dys != 0

    force()
    etot, com = [], []
    nshifts = 0
    for step in range(nsteps + 1):
        if step % every == 0:
            r = s.results()                                   # synchronises
            ke = 0.5 * mass * float((v * v).sum().item())
            etot.append(ke + r["epot"])
            com.append(float((x - dxeq)[:, 0].mean().item()))
            assert r["natoms_gathered"] == n and r["natoms_scattered"] == n
        if step == nsteps:
            break
        v.add_(f, alpha=0.5 * dt / mass)                      # fix nve: velocity Verlet
        x.add_(v, alpha=dt)
        nshifts += force(check_shift=(step % 25 == 0))
        v.add_(f, alpha=0.5 * dt / mass)
    e = np.array(etot)
    de = np.max(np.abs(e - e.mean()))
    # the layer oscillates by more than half a lattice constant: the re-indexing path ran
    # (it first crosses half a lattice constant after about 500 steps)
    if nsteps >= 1000 and expect_shift:
        assert max(abs(c) for c in com) > 0.5 and nshifts > 0
    assert e.mean() > 0
    assert de / e.mean() <= 1e-4, (de, e.mean())
    s.close()


def hertz_minimise(s, x, dxeq, dgid, dmask, n, xprd, yprd, dmax=None):
    """Rigid sphere (fix contact/sphere 0 0 99.5 100.0 1.38888888888889 0.890898718140339 1.0) on
    the GFMD layer, FIRE to |f| <= 1e-6 (LAMMPS: min_style cg; minimize 0.0 1e-6 ...).
    Returns the GFMD force on the atoms at the minimum.  dmax: largest move of any coordinate per
    iteration (LAMMPS' min_modify dmax, default 0.1 there): the layer first meets the 12-6 wall at
    speed, and without the cap an atom can step through the wall's singularity (it does for lattice
    constant 1.3).  None keeps the uncapped iteration the first two Hertz cases were verified with."""
    import torch
    cx, cy, cz, R = 0.0, 0.0, 99.5, 100.0
    eps, sig, cut = 1.38888888888889, 0.890898718140339, 1.0
    c1, c2 = 48.0 * eps * sig ** 12, 24.0 * eps * sig ** 6
    f = torch.zeros_like(x)

    def force():
        f.zero_()
        s.full_step(x, dxeq, dgid, dmask, 1, n, n, xprd, yprd, f)
        rx = x[:, 0] - cx
        ry = x[:, 1] - cy
        rz = x[:, 2] - cz
        rx = rx - xprd * torch.round(rx / xprd)              # domain->minimum_image
        ry = ry - yprd * torch.round(ry / yprd)
        r = torch.sqrt(rx * rx + ry * ry + rz * rz)
        inside = r < R + cut
        rinv = 1.0 / torch.clamp(r - R, min=1e-300)
        r6 = rinv ** 6
        df = torch.where(inside, r6 * (c1 * r6 - c2) * rinv, torch.zeros_like(r))
        f[:, 0] += df * rx / r
        f[:, 1] += df * ry / r
        f[:, 2] += df * rz / r
        return r

    v = torch.zeros_like(x)
    dt, dtmax, alpha, npos = 0.05, 0.25, 0.1, 0
    fnorm = None
    for it in range(200000):
        r = force()
        if it % 50 == 0:
            assert bool((r > R).all()), "fix contact/sphere: atom inside sphere"
            fnorm = float(torch.linalg.vector_norm(f).item())
            if fnorm <= 1e-6:
                break
        p = float((f * v).sum().item())
        if p > 0:
            vn = torch.linalg.vector_norm(v)
            fn = torch.linalg.vector_norm(f)
            v.mul_(1.0 - alpha).add_(f * (alpha * vn / fn))
            npos += 1
            if npos > 5:
                dt = min(dt * 1.1, dtmax)
                alpha *= 0.99
        else:
            v.zero_()
            dt *= 0.5
            alpha = 0.1
            npos = 0
        v.add_(f, alpha=dt)
        if dmax is not None:
            m = float((v.abs().max() * dt).item())
            if m > dmax:
                v.mul_(dmax / m)
        x.add_(v, alpha=dt)
    assert fnorm is not None and fnorm <= 1e-6, fnorm
    fg = torch.zeros_like(x)
    s.full_step(x, dxeq, dgid, dmask, 1, n, n, xprd, yprd, fg)
    s.results()
    return fg


def hertz_profile(r, N, E, R):
    a = R * (3.0 / 4 * (N / (E * R ** 2))) ** (1.0 / 3)
    p0 = 3 * N / (2 * math.pi * a * a)
    return np.where(r < a, p0 * np.sqrt(np.maximum(0.0, 1 - (r / a) ** 2)), np.zeros_like(r)), a, p0


def test_hertz_sc100_128x128():
    import torch
    import gfmd_b200
    run_hertz_sc100_128x128(gfmd_b200, torch.device("cuda"))


def run_hertz_sc100_128x128(gfmd_b200, dev):
    run_hertz_cubic(gfmd_b200, dev, load_golden("C1_sc100_128x128"), 1.0, 8.0 / 3)


def run_hertz_cubic(gfmd_b200, dev, table, a0, E, dmax=None):
    """One atom per cubic surface cell of lattice constant a0 (lattice sc a0; create_atoms;
    displace_atoms all move 0.5 0.5 0.5 units box)."""
    import torch
    g = table
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    gid, xeq = make_layer_atoms(nx, ny, 1)
    xeq[:, 0] = gid[:, 0] * a0 + 0.5
    xeq[:, 1] = gid[:, 1] * a0 + 0.5
    n = gid.shape[0]
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(g["phi"], g["linf"])
    bind_stream(s, dev)
    dxeq = torch.tensor(xeq, device=dev)
    dgid = torch.tensor(gid, device=dev)
    dmask = torch.ones(n, dtype=torch.int32, device=dev)
    x = dxeq.clone()
    x[:, 2] -= 2.0                                           # displace_atoms all move 0 0 -2.0
    fg = hertz_minimise(s, x, dxeq, dgid, dmask, n, nx * a0, ny * a0, dmax)
    # eval.py: z-force of the GFMD layer (gfmd.*.r.f2.out = f_xy[2]) vs Hertz, contact modulus E;
    # for a0 != 1 the comparison is in pressure, f / a0^2, at radii in box units
    f_xy = fg[:, 2].reshape(nx, ny).cpu().numpy()
    xs = np.arange(nx) + 0.5
    xs = np.where(xs > nx / 2, xs - nx, xs) * a0
    ys = np.arange(ny) + 0.5
    ys = np.where(ys > ny / 2, ys - ny, ys) * a0
    r_xy = np.sqrt((xs ** 2).reshape(-1, 1) + (ys ** 2).reshape(1, -1))
    N = np.sum(f_xy)
    fa_xy, a, p0 = hertz_profile(r_xy, N, E, 100.0)
    res = np.sum((f_xy / (a0 * a0) - fa_xy) ** 2)
    assert N > 0 and a > 3
    assert res < 1e-2, (res, N, a, p0)
    s.close()


def test_hertz_fcc111_64x37():
    """tests/TEST_Hertz_fcc111_64x37_2: non-power-of-two grid (Bluestein rows), two atoms
    per rectangular surface cell (ndof 6), E = 1.54, residual of the PRESSURE < 2e-3
    (eval.py:33-34, :94-106)."""
    import torch
    import gfmd_b200
    run_hertz_fcc111_64x37(gfmd_b200, torch.device("cuda"))


def run_hertz_fcc111_64x37(gfmd_b200, dev):
    import torch
    g = load_golden("C2_fcc111_64x37")
    nx, ny, d = int(g["nx"]), int(g["ny"]), int(g["ndof"])
    s3 = math.sqrt(3.0)
    # lattice custom: a1 = (1,0,0), a2 = (0,sqrt 3,0), basis (0,0,0) and (1/2,1/2,0)
    gid = np.array([(ix, iy, iu) for ix in range(nx) for iy in range(ny) for iu in range(2)], dtype=np.int32)
    xeq = np.stack([gid[:, 0] + 0.5 * gid[:, 2], (gid[:, 1] + 0.5 * gid[:, 2]) * s3,
                    np.zeros(len(gid))], axis=1)
    n = gid.shape[0]
    s = gfmd_b200.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    s.set_kernel(g["phi"], g["linf"])
    bind_stream(s, dev)
    dxeq = torch.tensor(xeq, device=dev)
    dgid = torch.tensor(gid, device=dev)
    dmask = torch.ones(n, dtype=torch.int32, device=dev)
    x = dxeq.clone()
    x[:, 2] -= 2.0
    xprd, yprd = float(nx), ny * s3
    fg = hertz_minimise(s, x, dxeq, dgid, dmask, n, xprd, yprd)
    fz = fg[:, 2].cpu().numpy()
    xe = xeq[:, 0].copy()
    ye = xeq[:, 1].copy()
    xe = np.where(xe > xprd / 2, xe - xprd, xe)
    ye = np.where(ye > yprd / 2, ye - yprd, ye)
    r = np.sqrt(xe ** 2 + ye ** 2)
    A0 = s3 / 2                                              # area per atom
    N = np.sum(fz)
    pa, a, p0 = hertz_profile(r, N, 1.54, 100.0)
    res = np.sum((fz / A0 - pa) ** 2)
    assert N > 0 and a > 3
    assert res < 0.002, (res, N, a, p0)
    s.close()
