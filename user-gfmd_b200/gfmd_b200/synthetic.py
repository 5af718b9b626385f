"""Deterministic synthetic inputs for benchmarks and size-independent tests
(configuration C4): a closed-form simple-cubic-like surface
stiffness table and a displacement field made of random-phase plane waves."""
import numpy as np


def phi_columns(nx, ny, ky_first, nky):
    """Closed-form 3x3 Hermitian stiffness, Phi(-q) = conj Phi(q), positive definite,
    UNNORMALISED, for kx in [0, nx), ky in [ky_first, ky_first + nky) -- in the layout of
    fill_phi_buffer(ndof, nx, 0, nx-1, ny, ky_first, ...): [nx, nky, 3, 3] complex128.

    Nearest/next-nearest-neighbour simple-cubic surface layer (the U0 block of the
    reference's sc100 kernel has this structure, src/stiffness_kernels/sc100_stiffness.cpp);
    used where timing, not elasticity, is the point: the solver's work does not depend
    on the table's values."""
    qx = 2.0 * np.pi * np.fft.fftfreq(nx)
    qy = 2.0 * np.pi * np.arange(ky_first, ky_first + nky) / ny
    QX, QY = np.meshgrid(qx, qy, indexing="ij")
    cx, cy, sx, sy = np.cos(QX), np.cos(QY), np.sin(QX), np.sin(QY)
    P = np.zeros((nx, nky, 3, 3), dtype=np.complex128)
    P[..., 0, 0] = 4.0 - 2.0 * cx * (1.0 + cy) + 0.02
    P[..., 1, 1] = 4.0 - 2.0 * cy * (1.0 + cx) + 0.02
    P[..., 2, 2] = 3.0 - cx - cy + 0.02
    P[..., 0, 1] = 2.0 * sx * sy
    P[..., 1, 0] = P[..., 0, 1]
    P[..., 0, 2] = 1j * sx
    P[..., 2, 0] = -1j * sx
    P[..., 1, 2] = 1j * sy
    P[..., 2, 1] = -1j * sy
    return P


def phi_full(nx, ny):
    """Same table for the whole grid in the reference layout [nx*ny, 3, 3], normalised
    by 1/(nx*ny) like fill_phi_buffer(normalize=true)."""
    P = phi_columns(nx, ny, 0, ny)
    return (P / (nx * ny)).reshape(nx * ny, 3, 3)


def displacement_field(nx, ny, x0=0, nx_loc=None, seed=1, nwaves=64, amp=1e-3):
    """Sum of `nwaves` random-phase plane waves of amplitude `amp` plus uniform noise of
    the same size, on rows [x0, x0 + nx_loc): returns [3, nx_loc, ny] float64.  The wave
    parameters depend on the seed only, so slabs of different ranks fit together."""
    nx_loc = nx if nx_loc is None else nx_loc
    rng = np.random.default_rng(seed)
    kx = rng.integers(0, nx, size=(nwaves, 3))
    ky = rng.integers(0, ny, size=(nwaves, 3))
    ph = rng.uniform(0, 2 * np.pi, size=(nwaves, 3))
    ix = np.arange(x0, x0 + nx_loc)[:, None]
    iy = np.arange(ny)[None, :]
    u = np.zeros((3, nx_loc, ny))
    for c in range(3):
        for w in range(nwaves):
            u[c] += np.cos(2 * np.pi * (kx[w, c] * ix / nx + ky[w, c] * iy / ny) + ph[w, c])
    u *= amp
    noise = np.random.default_rng(seed + 1000 + x0).uniform(-amp, amp, size=u.shape)
    return u + noise


def atoms_for_slab(nx, ny, x0, nx_loc, u):
    """One atom per cell at xeq = lattice site + 0.5, displaced by u: returns x, xeq
    [n, 3] float64, gid [n, 3] int32 (ix, iy, 0), mask [n] int32 (group bit 1)."""
    ix, iy = np.meshgrid(np.arange(x0, x0 + nx_loc), np.arange(ny), indexing="ij")
    n = nx_loc * ny
    gid = np.zeros((n, 3), dtype=np.int32)
    gid[:, 0] = ix.ravel()
    gid[:, 1] = iy.ravel()
    xeq = np.empty((n, 3))
    xeq[:, 0] = gid[:, 0] + 0.5
    xeq[:, 1] = gid[:, 1] + 0.5
    xeq[:, 2] = 0.5
    x = xeq + np.moveaxis(u.reshape(3, n), 0, 1)
    mask = np.ones(n, dtype=np.int32)
    return x, xeq, gid, mask


def sc100_dynamical_matrices(nx, ny, ky_first, nky):
    """U0(q), U(q), V(q) of the simple-cubic (100) surface with unit nearest- and
    next-nearest-neighbour springs -- what the reference's `sc100` stiffness kernel feeds
    into the transfer-matrix recursion (SC100StiffnessKernel::get_per_layer_dynamical_matrices,
    src/stiffness_kernels/sc100_stiffness.cpp:170-228).  Layout of
    gfmd_b200_build_phi_columns: [nx, nky, 3, 3, 3] complex128, q as in fill_phi_buffer."""
    i = np.arange(nx)
    j = np.arange(ky_first, ky_first + nky)
    qx = np.where(i <= nx // 2, 2.0 * np.pi * i / nx, 2.0 * np.pi * (i - nx) / nx)
    qy = np.where(j <= ny // 2, 2.0 * np.pi * j / ny, 2.0 * np.pi * (j - ny) / ny)
    QX, QY = np.meshgrid(qx, qy, indexing="ij")
    cx, cy, sx, sy = np.cos(QX), np.cos(QY), np.sin(QX), np.sin(QY)
    M = np.zeros((nx, nky, 3, 3, 3), dtype=np.complex128)
    U0, U, V = M[:, :, 0], M[:, :, 1], M[:, :, 2]
    U[..., 0, 0] = 6 - 2 * cx * (1 + cy)
    U[..., 1, 1] = 6 - 2 * cy * (1 + cx)
    U[..., 2, 2] = 6
    U[..., 0, 1] = U[..., 1, 0] = 2 * sx * sy
    U0[..., 0, 0] = 5 - 2 * cx * (1 + cy)
    U0[..., 1, 1] = 5 - 2 * cy * (1 + cx)
    U0[..., 2, 2] = 3
    U0[..., 0, 1] = U0[..., 1, 0] = 2 * sx * sy
    V[..., 0, 0] = -cx
    V[..., 1, 1] = -cy
    V[..., 2, 2] = -1 - cx - cy
    V[..., 0, 2] = V[..., 2, 0] = 1j * sx
    V[..., 1, 2] = V[..., 2, 1] = 1j * sy
    return M


# ---- device-side generators (torch tensors on the GPU) for grids too large to build on the host ---
# Same physics as above; every value is a pure function of the GLOBAL cell index, so slabs of
# different ranks -- and a single-GPU run of the whole grid -- see the same field.

def _hash_noise_torch(ix, iy, c, seed):
    """Uniform(-1, 1) from an integer hash of (ix, iy, c, seed); ix [n,1], iy [1,m] int64 tensors."""
    h = (ix * 73856093 + iy * 19349663 + c * 83492791 + seed * 2654435761) & 0x7FFFFFFF
    for _ in range(3):
        h = (h * 1103515245 + 12345) & 0x7FFFFFFF
        h = h ^ (h >> 15)
    return h.to(dtype=__import__("torch").float64) * (2.0 / 2147483648.0) - 1.0


def displacement_field_torch(nx, ny, x0, nx_loc, device, seed=1, nwaves=8, amp=1e-3):
    """[3, nx_loc, ny] float64 on `device`: `nwaves` random-phase plane waves of amplitude `amp` per
    component (separable form cos(a + b) = cos a cos b - sin a sin b, i.e. two small matrix products)
    plus hash noise of the same size.  Wave parameters as in displacement_field (seed only)."""
    import torch
    rng = np.random.default_rng(seed)
    kx = rng.integers(0, nx, size=(nwaves, 3))
    ky = rng.integers(0, ny, size=(nwaves, 3))
    ph = rng.uniform(0, 2 * np.pi, size=(nwaves, 3))
    ix = torch.arange(x0, x0 + nx_loc, device=device, dtype=torch.int64)
    iy = torch.arange(ny, device=device, dtype=torch.int64)
    u = torch.empty((3, nx_loc, ny), device=device, dtype=torch.float64)
    for c in range(3):
        kxc = torch.tensor(kx[:, c], device=device, dtype=torch.int64)
        kyc = torch.tensor(ky[:, c], device=device, dtype=torch.int64)
        phc = torch.tensor(ph[:, c], device=device, dtype=torch.float64)
        # exact argument reduction in integers before the conversion to double
        ax = ((kxc[None, :] * ix[:, None]) % nx).to(torch.float64) * (2.0 * np.pi / nx) + phc[None, :]
        ay = ((kyc[:, None] * iy[None, :]) % ny).to(torch.float64) * (2.0 * np.pi / ny)
        u[c] = torch.cos(ax) @ torch.cos(ay) - torch.sin(ax) @ torch.sin(ay)
        u[c] += _hash_noise_torch(ix[:, None], iy[None, :], c, seed)
    u *= amp
    return u


def atoms_for_slab_torch(nx, ny, x0, nx_loc, u):
    """Device version of atoms_for_slab: x, xeq [n, 3] float64, gid [n, 3] int32, mask [n] int32."""
    import torch
    dev = u.device
    n = nx_loc * ny
    ix = torch.arange(x0, x0 + nx_loc, device=dev, dtype=torch.int32)[:, None].expand(nx_loc, ny).reshape(n)
    iy = torch.arange(ny, device=dev, dtype=torch.int32)[None, :].expand(nx_loc, ny).reshape(n)
    gid = torch.zeros((n, 3), device=dev, dtype=torch.int32)
    gid[:, 0] = ix
    gid[:, 1] = iy
    xeq = torch.empty((n, 3), device=dev, dtype=torch.float64)
    xeq[:, 0] = ix.to(torch.float64) + 0.5
    xeq[:, 1] = iy.to(torch.float64) + 0.5
    xeq[:, 2] = 0.5
    x = xeq + u.reshape(3, n).t()
    mask = torch.ones(n, device=dev, dtype=torch.int32)
    return x, xeq, gid, mask


def sc100_dynamical_matrices_torch(nx, ny, ky_first, nky, device):
    """sc100_dynamical_matrices evaluated on the GPU: [nx, nky, 3, 3, 3] complex128 on `device`."""
    import torch
    i = torch.arange(nx, device=device, dtype=torch.float64)
    j = torch.arange(ky_first, ky_first + nky, device=device, dtype=torch.float64)
    qx = torch.where(i <= nx // 2, 2.0 * np.pi * i / nx, 2.0 * np.pi * (i - nx) / nx)[:, None]
    qy = torch.where(j <= ny // 2, 2.0 * np.pi * j / ny, 2.0 * np.pi * (j - ny) / ny)[None, :]
    cx, cy, sx, sy = torch.cos(qx), torch.cos(qy), torch.sin(qx), torch.sin(qy)
    M = torch.zeros((nx, nky, 3, 3, 3, 2), device=device, dtype=torch.float64)     # (re, im) pairs
    U0, U, V = M[:, :, 0], M[:, :, 1], M[:, :, 2]
    one = torch.ones((nx, nky), device=device, dtype=torch.float64)
    U[..., 0, 0, 0] = 6 - 2 * cx * (1 + cy)
    U[..., 1, 1, 0] = 6 - 2 * cy * (1 + cx)
    U[..., 2, 2, 0] = 6 * one
    U[..., 0, 1, 0] = 2 * sx * sy
    U[..., 1, 0, 0] = 2 * sx * sy
    U0[..., 0, 0, 0] = 5 - 2 * cx * (1 + cy)
    U0[..., 1, 1, 0] = 5 - 2 * cy * (1 + cx)
    U0[..., 2, 2, 0] = 3 * one
    U0[..., 0, 1, 0] = 2 * sx * sy
    U0[..., 1, 0, 0] = 2 * sx * sy
    V[..., 0, 0, 0] = -cx * one
    V[..., 1, 1, 0] = -cy * one
    V[..., 2, 2, 0] = -1 - cx - cy
    V[..., 0, 2, 1] = sx * one
    V[..., 2, 0, 1] = sx * one
    V[..., 1, 2, 1] = sy * one
    V[..., 2, 1, 1] = sy * one
    return M
