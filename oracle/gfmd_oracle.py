"""ORACLE / TEST INFRASTRUCTURE ONLY.

numpy restatement of the per-step elastic-force path of ``fix gfmd`` plus thin
ctypes bindings to the two oracle libraries built by ``oracle/Makefile``:

* ``oracle/_ref/libgfmd_oracle.so`` -- the plain-C restatement (gfmd_oracle.c)
* ``oracle/_ref/libgfmd_ref.so``    -- the reference's OWN sources compiled
  unchanged against ``oracle/shim`` (see ref_driver.cpp)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Nothing under
``user-gfmd_b200/`` does.

Reference lines restated here:
  post_force      src/solvers/gfmd_solver_static.cpp:145-249
  fft_forward     src/solvers/gfmd_solver_fft.cpp:96-147  (unnormalised, exp(-i q r))
  fft_reverse     src/solvers/gfmd_solver_fft.cpp:150-195 (unnormalised, exp(+i q r), real part)
  q ordering      src/main/gfmd_misc.cpp:43-48
  grid partition  src/main/gfmd_solver.cpp:95-135
  gather          src/main/fix_gfmd.cpp:734-803
  scatter         src/main/fix_gfmd.cpp:952-1010, :896-902
  spectrum/dump   src/solvers/gfmd_solver_fft.cpp:209-287
  prec_gradient   src/solvers/gfmd_solver_static.cpp:253-271, src/main/gfmd_misc.h:39-133

PINNING: checked against libgfmd_ref.so and tests/golden in tests/test_oracle.py.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


# --------------------------------------------------------------------------
# numpy restatement
# --------------------------------------------------------------------------

def q_vectors(nx, ny):
    """qx[i], qy[j] exactly as fill_phi_buffer builds them (gfmd_misc.cpp:43-48)."""
    i = np.arange(nx)
    j = np.arange(ny)
    qx = np.where(i <= nx // 2, 2.0 * np.pi * i / nx, 2.0 * np.pi * (i - nx) / nx)
    qy = np.where(j <= ny // 2, 2.0 * np.pi * j / ny, 2.0 * np.pi * (j - ny) / ny)
    return qx, qy


def grid_partition(nx, ny, sublo, subhi, boxlo, prd):
    """Local brick of GFMDSolver::set_grid_size (gfmd_solver.cpp:95-135).

    Returns (xlo_loc, xhi_loc, ylo_loc, yhi_loc, nxy_loc, gammai)."""
    def rnd(v):  # C round(): half away from zero
        return int(np.floor(abs(v) + 0.5) * np.sign(v)) if v != 0 else 0
    xlo = rnd(nx * (sublo[0] - boxlo[0]) / prd[0])
    xhi = rnd(nx * (subhi[0] - boxlo[0]) / prd[0]) - 1
    ylo = rnd(ny * (sublo[1] - boxlo[1]) / prd[1])
    yhi = rnd(ny * (subhi[1] - boxlo[1]) / prd[1]) - 1
    nx_loc, ny_loc = xhi - xlo + 1, yhi - ylo + 1
    gammai = -1
    if xlo <= 0 <= xhi and ylo <= 0 <= yhi:
        gammai = -xlo * ny_loc - ylo
    return xlo, xhi, ylo, yhi, nx_loc * ny_loc, gammai


def post_force(u, phi, linf, fft=None):
    """GFMDSolverStatic::post_force on the full grid (single rank).

    u    : [ndof, nx, ny] float64 (the reference's u_xy[idof][ix*ny+iy])
    phi  : [nx*ny, ndof, ndof] complex128, already divided by nx*ny
    linf : [ndof//3]
    Returns f [ndof, nx, ny], epot, u0 [ndof].
    """
    if fft is None:
        fft = np.fft
    ndof, nx, ny = u.shape
    nu = ndof // 3
    linf = np.asarray(linf, dtype=np.float64)
    uq = fft.fft2(u, axes=(1, 2))                       # exp(-i q r), unnormalised
    q = np.ascontiguousarray(np.moveaxis(uq.reshape(ndof, nx * ny), 0, 1))  # [idq][idim]
    u0 = q[0].real.copy()
    epot = 0.0
    for i in range(nu):
        epot -= 2 * linf[i] * q[0, 3 * i + 2].real
    F = np.einsum("qij,qj->qi", phi.reshape(nx * ny, ndof, ndof), q)
    epot += float(np.sum((np.conj(F) * q).real))
    q = -F
    for i in range(nu):
        q[0, 3 * i + 2] += linf[i]
    epot *= 0.5
    fq = np.moveaxis(q, 0, 1).reshape(ndof, nx, ny)
    f = fft.ifft2(fq, axes=(1, 2)).real * (nx * ny)     # unnormalised backward
    return np.ascontiguousarray(f), epot, u0


def spectrum(u, phi, fft=None):
    """The q-space fields GFMDSolverFFT::dump writes (src/solvers/gfmd_solver_fft.cpp:209-287,
    called from gfmd_solver_static.cpp:181-182): u(q) = unnormalised forward DFT in the
    q_buffer layout [idq = ix*ny + iy][idim], and F(q) = Phi(q).u(q) (MatMulVec, no sign flip).
    Returns uq, fq: [nx*ny, ndof] complex128."""
    if fft is None:
        fft = np.fft
    ndof, nx, ny = u.shape
    uq = fft.fft2(u, axes=(1, 2))
    q = np.ascontiguousarray(np.moveaxis(uq.reshape(ndof, nx * ny), 0, 1))
    F = np.einsum("qij,qj->qi", phi.reshape(nx * ny, ndof, ndof), q)
    return q, F


def dump_fields(uq, fq, nx, ny):
    """What each <prefix>.q.*.out file of GFMDSolverFFT::dump holds, as [ny, nx] arrays (file
    rows = iy, columns = ix; gfmd_solver_fft.cpp:247-275): u<i>.real/imag, f<i>.real/imag,
    uP = sum |u|^2, fP = sum |F|^2, e = Re sum u conj(F)."""
    ndof = uq.shape[1]
    U = uq.reshape(nx, ny, ndof)
    F = fq.reshape(nx, ny, ndof)
    out = {}
    for i in range(ndof):
        out["u%d.real" % i] = U[:, :, i].real.T
        out["u%d.imag" % i] = U[:, :, i].imag.T
        out["f%d.real" % i] = F[:, :, i].real.T
        out["f%d.imag" % i] = F[:, :, i].imag.T
    out["uP"] = (np.abs(U) ** 2).sum(axis=2).T
    out["fP"] = (np.abs(F) ** 2).sum(axis=2).T
    out["e"] = (U * np.conj(F)).sum(axis=2).real.T
    return out


def prec_gradient(g, phi, cavg, fft=None, reference_quirk=False):
    """GFMDSolverStatic::prec_gradient (src/solvers/gfmd_solver_static.cpp:253-271) with
    precondition_gradient<DEF_G> (src/main/gfmd_misc.h:39-133):
    gP = IDFT[(Phi(q) + Cavg)^-1 DFT[g]], transforms unnormalised, Phi as stored (divided by
    nx*ny).  g: [ndof, nx, ny]; cavg: [ndof, ndof] real.

    reference_quirk: for ndof > 3 the reference's general branch copies only the first
    three components of the result back (``idim < 3``, gfmd_misc.h:113-115), leaving the
    others un-preconditioned; True reproduces that."""
    if fft is None:
        fft = np.fft
    ndof, nx, ny = g.shape
    gq = fft.fft2(g, axes=(1, 2))
    q = np.ascontiguousarray(np.moveaxis(gq.reshape(ndof, nx * ny), 0, 1))
    M = phi.reshape(nx * ny, ndof, ndof) + np.asarray(cavg, dtype=np.float64).reshape(1, ndof, ndof)
    y = np.linalg.solve(M, q[..., None])[..., 0]
    if reference_quirk and ndof > 3:
        y[:, 3:] = q[:, 3:]
    yq = np.moveaxis(y, 0, 1).reshape(ndof, nx, ny)
    return np.ascontiguousarray(fft.ifft2(yq, axes=(1, 2)).real * (nx * ny))


def gather(x, xeq, gid, mask, groupbit, nx, ny, ndof, xprd, yprd,
           dxshift=0, dyshift=0, xlo_loc=0, ylo_loc=0, nx_loc=None, ny_loc=None,
           u_xy=None):
    """FixGFMD::pre_force list->grid (fix_gfmd.cpp:734-803).  gid is updated in
    place when shifted, as the reference does.  Vectorised; duplicates resolve
    to the LAST atom in list order like the reference's sequential loop."""
    nx_loc = nx if nx_loc is None else nx_loc
    ny_loc = ny if ny_loc is None else ny_loc
    if u_xy is None:
        u_xy = np.zeros((ndof, nx_loc * ny_loc))
    sel = (mask & groupbit) != 0
    ix = gid[:, 0] - dxshift
    iy = gid[:, 1] - dyshift
    if dxshift != 0 or dyshift != 0:
        ix = np.mod(ix, nx)
        iy = np.mod(iy, ny)
        gid[sel, 0] = ix[sel]
        gid[sel, 1] = iy[sel]
    ix = ix - xlo_loc
    iy = iy - ylo_loc
    inside = sel & (ix >= 0) & (ix < nx_loc) & (iy >= 0) & (iy < ny_loc)
    d = x - xeq
    ux, uy, uz = d[:, 0].copy(), d[:, 1].copy(), d[:, 2]
    # while (ux > xprd_half) ux -= xprd; while (ux < -xprd_half) ux += xprd
    for arr, prd in ((ux, xprd), (uy, yprd)):
        half = 0.5 * prd
        while True:
            m = arr > half
            if not m.any():
                break
            arr[m] -= prd
        while True:
            m = arr < -half
            if not m.any():
                break
            arr[m] += prd
    idx = np.nonzero(inside)[0]
    iloc = ix[idx] * ny_loc + iy[idx]
    idof = 3 * gid[idx, 2]
    u_xy[idof, iloc] = ux[idx]
    u_xy[idof + 1, iloc] = uy[idx]
    u_xy[idof + 2, iloc] = uz[idx]
    return u_xy, int(inside.sum())


def scatter(f_xy, gid, mask, groupbit, f, nlocal=None, xlo_loc=0, xhi_loc=None,
            ylo_loc=0, yhi_loc=None, nx=None, ny=None):
    """grid_to_list + f += f_i (fix_gfmd.cpp:952-1010, :896-902)."""
    xhi_loc = nx - 1 if xhi_loc is None else xhi_loc
    yhi_loc = ny - 1 if yhi_loc is None else yhi_loc
    ny_loc = yhi_loc - ylo_loc + 1
    nall = gid.shape[0]
    nlocal = nall if nlocal is None else nlocal
    sel = (mask & groupbit) != 0
    ix, iy = gid[:, 0], gid[:, 1]
    known = sel & (ix >= xlo_loc) & (ix <= xhi_loc) & (iy >= ylo_loc) & (iy <= yhi_loc)
    idx = np.nonzero(known)[0]
    iloc = (ix[idx] - xlo_loc) * ny_loc + (iy[idx] - ylo_loc)
    idof = 3 * gid[idx, 2]
    fi = np.stack([f_xy[idof, iloc], f_xy[idof + 1, iloc], f_xy[idof + 2, iloc]], axis=1)
    f[idx] += fi
    loc = idx < nlocal
    fsum = fi[loc].sum(axis=0)
    return f, fsum, int(known.sum())


# --------------------------------------------------------------------------
# Slab-decomposed restatement (used by the gloo tests of the multi-GPU plan):
# the same arithmetic split at the points where the B200 path exchanges data.
# --------------------------------------------------------------------------

def rows_forward(u_slab):
    """Row (y) real-to-complex transforms of one x-slab: [d, nx_loc, ny] ->
    half spectrum transposed [d, nyh, nx_loc]."""
    return np.ascontiguousarray(np.swapaxes(np.fft.rfft(u_slab, axis=2), 1, 2))


def columns_contract(ut_cols, phi_cols, linf, ky0, weights):
    """Column (x) transforms + contraction on a ky-slab.

    ut_cols  : [d, nky, nx] complex (x complete)
    phi_cols : [nky, nx, d, d] complex, normalised, phi(kx, ky0+k)
    weights  : [nky] multiplicity of each ky in the full spectrum (1 or 2)
    Returns (ft_cols [d, nky, nx], epot contribution (before the 0.5), u0 or None)."""
    d = ut_cols.shape[0]
    uq = np.fft.fft(ut_cols, axis=2)
    F = np.einsum("kxij,jkx->ikx", phi_cols, uq)
    e = float(np.sum(weights[None, :, None] * (np.conj(F) * uq).real))
    u0 = None
    Fq = -F
    if ky0 == 0:
        u0 = uq[:, 0, 0].real.copy()
        for i in range(d // 3):
            e -= 2 * linf[i] * uq[3 * i + 2, 0, 0].real
            Fq[3 * i + 2, 0, 0] += linf[i]
    ft = np.fft.ifft(Fq, axis=2) * uq.shape[2]
    return ft, e, u0


def rows_inverse(ft_slab, ny):
    """[d, nyh, nx_loc] half spectrum -> real forces [d, nx_loc, ny] (unnormalised)."""
    return np.fft.irfft(np.swapaxes(ft_slab, 1, 2), n=ny, axis=2) * ny


# --------------------------------------------------------------------------
# ctypes: plain-C restatement
# --------------------------------------------------------------------------

_clib = None


def clib():
    global _clib
    if _clib is None:
        path = os.path.join(_REF_DIR, "libgfmd_oracle.so")
        lib = ctypes.CDLL(path)
        lib.gfmd_oracle_post_force.restype = ctypes.c_double
        lib.gfmd_oracle_post_force.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p,
            c_double_p, c_double_p, c_double_p, ctypes.c_int]
        lib.gfmd_oracle_gather.restype = ctypes.c_int
        lib.gfmd_oracle_gather.argtypes = [
            ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_int_p, c_int_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
            ctypes.c_int, ctypes.c_int, c_double_p]
        lib.gfmd_oracle_scatter.restype = ctypes.c_int
        lib.gfmd_oracle_scatter.argtypes = [
            ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p,
            c_double_p, c_double_p]
        _clib = lib
    return _clib


def c_post_force(u, phi, linf, fft_backend=1):
    ndof, nx, ny = u.shape
    u = np.ascontiguousarray(u, dtype=np.float64)
    phi = np.ascontiguousarray(phi, dtype=np.complex128)
    linf = np.ascontiguousarray(linf, dtype=np.float64)
    f = np.empty_like(u)
    u0 = np.empty(ndof)
    e = clib().gfmd_oracle_post_force(nx, ny, ndof, _dp(phi.view(np.float64)), _dp(linf),
                                      _dp(u), _dp(f), _dp(u0), fft_backend)
    return f, e, u0


def c_gather(x, xeq, gid, mask, groupbit, nx, ny, ndof, xprd, yprd, dxshift=0, dyshift=0,
             u_xy=None):
    nall = x.shape[0]
    if u_xy is None:
        u_xy = np.zeros((ndof, nx * ny))
    gid = np.ascontiguousarray(gid, dtype=np.int32)
    mask = np.ascontiguousarray(mask, dtype=np.int32)
    n = clib().gfmd_oracle_gather(nall, nall, _dp(np.ascontiguousarray(x)),
                                  _dp(np.ascontiguousarray(xeq)), _ip(gid), _ip(mask),
                                  groupbit, nx, ny, 0, 0, nx, ny, xprd, yprd, dxshift,
                                  dyshift, _dp(u_xy))
    return u_xy, n, gid


def c_scatter(f_xy, gid, mask, groupbit, f, nx, ny):
    nall = gid.shape[0]
    gid = np.ascontiguousarray(gid, dtype=np.int32)
    mask = np.ascontiguousarray(mask, dtype=np.int32)
    fsum = np.zeros(3)
    n = clib().gfmd_oracle_scatter(nall, nall, _ip(gid), _ip(mask), groupbit, 0, nx - 1, 0,
                                   ny - 1, _dp(np.ascontiguousarray(f_xy)), _dp(f), _dp(fsum))
    return f, fsum, n


# --------------------------------------------------------------------------
# ctypes: the reference's own sources (oracle/_ref/libgfmd_ref.so)
# --------------------------------------------------------------------------

_rlib = None


def ref_available():
    return os.path.exists(os.path.join(_REF_DIR, "libgfmd_ref.so"))


def rlib():
    global _rlib
    if _rlib is None:
        lib = ctypes.CDLL(os.path.join(_REF_DIR, "libgfmd_ref.so"))
        lib.ref_kernel_create.restype = ctypes.c_void_p
        lib.ref_kernel_create.argtypes = [ctypes.c_char_p, ctypes.c_int]
        lib.ref_kernel_ndof.argtypes = [ctypes.c_void_p]
        lib.ref_kernel_nu.argtypes = [ctypes.c_void_p]
        lib.ref_fill_phi.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p,
                                     ctypes.c_int]
        lib.ref_phi_at.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                   ctypes.c_double, c_double_p]
        lib.ref_get_linf.argtypes = [ctypes.c_void_p, c_double_p]
        lib.ref_dynamical_matrices_columns.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                       ctypes.c_int, ctypes.c_int, c_double_p]
        lib.ref_kernel_height.argtypes = [ctypes.c_void_p]
        lib.ref_kernel_destroy.argtypes = [ctypes.c_void_p]
        lib.ref_solver_create.restype = ctypes.c_void_p
        lib.ref_solver_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.ref_solver_set_kernel.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_solver_set_phi.argtypes = [ctypes.c_void_p, c_double_p, c_double_p]
        lib.ref_solver_post_force.restype = ctypes.c_double
        lib.ref_solver_post_force.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, c_double_p]
        lib.ref_solver_destroy.argtypes = [ctypes.c_void_p]
        lib.ref_solver_prec_gradient.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, c_double_p]
        lib.ref_solver_post_force_dump.restype = ctypes.c_double
        lib.ref_solver_post_force_dump.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_char_p]
        lib.ref_solver_dump_tables.argtypes = [ctypes.c_void_p]
        lib.ref_set_fft_backend.argtypes = [ctypes.c_int]
        lib.fftp_set_threads.argtypes = [ctypes.c_int]
        lib.fftp_get_threads.restype = ctypes.c_int
        _rlib = lib
    return _rlib


def set_fft_threads(n):
    """Threads of the substitute FFT (oracle/fft_plain.c) in every oracle library that is loaded or will
    be: explicit, so that a launcher's OMP_NUM_THREADS=1 (torchrun) cannot shrink the CPU baseline.
    Returns the thread count in effect."""
    got = 0
    if ref_available():
        rlib().fftp_set_threads(int(n))
        got = rlib().fftp_get_threads()
    try:
        lib = clib()
        lib.fftp_set_threads.argtypes = [ctypes.c_int]
        lib.fftp_get_threads.restype = ctypes.c_int
        lib.fftp_set_threads(int(n))
        got = got or lib.fftp_get_threads()
    except Exception:
        pass
    return got


class RefKernel:
    """A stiffness-kernel plugin of the reference, e.g.
    RefKernel("ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128")."""

    def __init__(self, kernel_string, invariant=False):
        self.h = rlib().ref_kernel_create(kernel_string.encode(), int(invariant))
        if not self.h:
            raise ValueError("unknown stiffness kernel: " + kernel_string)
        self.ndof = rlib().ref_kernel_ndof(self.h)
        self.nu = self.ndof // 3

    def phi(self, nx, ny, normalize=True):
        """fill_phi_buffer (gfmd_misc.cpp:32-101): [nx*ny, ndof, ndof] complex128."""
        out = np.empty((nx * ny, self.ndof, self.ndof), dtype=np.complex128)
        rlib().ref_fill_phi(self.h, nx, ny, _dp(out.view(np.float64)), int(normalize))
        return out

    def phi_at(self, nx, ny, qx, qy):
        out = np.empty((self.ndof, self.ndof), dtype=np.complex128)
        rlib().ref_phi_at(self.h, nx, ny, qx, qy, _dp(out.view(np.float64)))
        return out

    def dynamical_matrices(self, nx, ny, ky_first, nky):
        """(U0, U, V) of get_dynamical_matrices per q: [nx, nky, 3, ndof, ndof] complex128."""
        out = np.empty((nx, nky, 3, self.ndof, self.ndof), dtype=np.complex128)
        rlib().ref_dynamical_matrices_columns(self.h, nx, ny, ky_first, nky, _dp(out.view(np.float64)))
        return out

    def height(self):
        return rlib().ref_kernel_height(self.h)

    def linf(self):
        out = np.zeros(max(self.nu, 1))
        rlib().ref_get_linf(self.h, _dp(out))
        return out

    def close(self):
        if self.h:
            rlib().ref_kernel_destroy(self.h)
            self.h = None


class RefSolver:
    """The reference's GFMDSolverStatic (from gfmd_solver_factory("static")),
    single rank, FFT3d replaced by the shim.  fft_backend 0 = long-double DFT."""

    def __init__(self, nx, ny, ndof, fft_backend=0):
        self.nx, self.ny, self.ndof = nx, ny, ndof
        rlib().ref_set_fft_backend(fft_backend)
        self.backend = fft_backend
        self.h = rlib().ref_solver_create(nx, ny, ndof)

    def set_kernel(self, kernel, normalize=True):
        rlib().ref_solver_set_kernel(self.h, kernel.h, int(normalize))

    def set_phi(self, phi, linf):
        phi = np.ascontiguousarray(phi, dtype=np.complex128)
        linf = np.ascontiguousarray(linf, dtype=np.float64)
        rlib().ref_solver_set_phi(self.h, _dp(phi.view(np.float64)), _dp(linf))

    def post_force(self, u):
        rlib().ref_set_fft_backend(self.backend)
        u = np.ascontiguousarray(u, dtype=np.float64)
        f = np.empty_like(u)
        u0 = np.empty(self.ndof)
        e = rlib().ref_solver_post_force(self.h, _dp(u), _dp(f), _dp(u0))
        return f.reshape(self.ndof, self.nx, self.ny), e, u0

    def prec_gradient(self, cavg, g):
        """GFMDSolverStatic::prec_gradient of the reference sources."""
        rlib().ref_set_fft_backend(self.backend)
        g = np.ascontiguousarray(g, dtype=np.float64)
        cavg = np.ascontiguousarray(cavg, dtype=np.float64)
        gP = np.empty_like(g)
        rlib().ref_solver_prec_gradient(self.h, _dp(cavg), _dp(g), _dp(gP))
        return gP.reshape(self.ndof, self.nx, self.ny)

    def post_force_dump(self, u, prefix):
        """post_force on a `dumpq_every` step: writes <prefix>.q.*.out (GFMDSolverFFT::dump)."""
        rlib().ref_set_fft_backend(self.backend)
        u = np.ascontiguousarray(u, dtype=np.float64)
        f = np.empty_like(u)
        e = rlib().ref_solver_post_force_dump(self.h, _dp(u), _dp(f), prefix.encode())
        return f.reshape(self.ndof, self.nx, self.ny), e

    def dump_tables(self, directory):
        """dump_stiffness + dump_greens_function; they write into the current directory."""
        cwd = os.getcwd()
        os.chdir(directory)
        try:
            rlib().ref_solver_dump_tables(self.h)
        finally:
            os.chdir(cwd)

    def close(self):
        if self.h:
            rlib().ref_solver_destroy(self.h)
            self.h = None
