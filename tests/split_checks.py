"""Size-independent checks of the three-phase column stage, shared by the GPU test
(tests/test_split_columns_gpu.py, full size) and the emulated one (small, forced split)."""
import numpy as np


def energy_identity_and_linearity(B, nx, ny, d, expect=()):
    """With linf = 0: E = -1/2 sum_r f.u (SURVEY 8a restatement) and f is linear in u.  The table is a
    smooth positive-definite Phi(q) = w(q) A + i sin(qx) C + 0.01, A symmetric positive definite,
    C antisymmetric: Hermitian, Phi(-q) = conj Phi(q), complex off-diagonals."""
    rng = np.random.default_rng(5)
    s = B.GFMDSolverB200()
    s.set_grid_size(nx, ny, d)
    for e in expect:
        assert e in s.describe(), s.describe()
    A = rng.standard_normal((d, d))
    A = A @ A.T + d * np.eye(d)
    C = rng.standard_normal((d, d))
    C = C - C.T
    qx = 2 * np.pi * np.fft.fftfreq(nx)
    for k0 in range(0, s.nky, 32):
        nk = min(32, s.nky - k0)
        qy = 2 * np.pi * np.arange(k0, k0 + nk) / ny
        w = (2.0 - np.cos(qx)[:, None] - np.cos(qy)[None, :])[..., None, None]
        sx = np.sin(qx)[:, None, None, None] * np.ones((1, nk, 1, 1))
        cols = w * A + 1j * sx * C + 0.01 * np.eye(d)
        s.set_kernel_columns(np.ascontiguousarray(cols.reshape(nx * nk, d, d)), k0, normalized=False)
    s.set_linf(np.zeros(d // 3))
    u1 = rng.uniform(-1e-3, 1e-3, size=(d, nx * ny))
    u2 = rng.uniform(-1e-3, 1e-3, size=(d, nx * ny))
    f1, f2, f3 = (np.full_like(u1, np.nan) for _ in range(3))
    e1 = s.post_force(u1, f1)
    s.post_force(u2, f2)
    s.post_force(2.0 * u1 - 3.0 * u2, f3)
    assert e1 > 0.0
    assert abs(e1 + 0.5 * np.sum(f1 * u1)) <= 1e-11 * abs(e1)
    assert np.abs(f3 - (2.0 * f1 - 3.0 * f2)).max() <= 1e-11 * np.abs(f3).max()
    s.close()
