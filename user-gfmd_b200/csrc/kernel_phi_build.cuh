// Device-side stiffness-table builder.
//
// The host plugin only evaluates the three per-q matrices U0(q), U(q), V(q)
// (StiffnessKernel::get_dynamical_matrices, reference src/main/surface_stiffness.cpp:127-292
// and the plugins' get_per_layer_dynamical_matrices); the expensive part, the transfer-matrix
// recursion of greens_function_transfer_matrix_stiffness (surface_stiffness.cpp:811-873) with
// iterate_Gnn (:493-548),
//     Vd = -V^H ;  VT = V U^-1 Vd ;  repeat (height-1) times: VT = V (U + VT)^-1 Vd ;
//     Phi = U0 + VT            (height == 0: Phi = U0)
// runs here, one thread per wavevector, and the result is written straight into the
// Hermitian-packed half-spectrum planes the column kernels stream.  The linear solves use
// Gaussian elimination with partial pivoting (the reference inverts with full-pivot
// Gauss-Jordan; the recursion is a contraction, so results agree to rounding).
// height < 0 (semi-infinite substrate): iterate until sum |VT - VT_prev|^2 <= (1e-8)^2, at most
// 100000 times (surface_stiffness.cpp:849-851, :514-539); a wavevector that runs out of iterations
// raises *not_converged, which the host turns into the reference's "Out of iterations while
// evaluating the continued fraction" error.  Near q = 0 the recursion converges like 1/n, so a few
// threads run ~1e4 iterations while the rest of the grid is long done -- init-time cost only.
#pragma once

#include "fft_pow2.cuh"

namespace gfmd {

// Y = M^-1 B for D x D complex matrices (row-major); M and B are destroyed.
template <int D>
__device__ __forceinline__ void csolve(double2 *M, double2 *B)
{
  for (int k = 0; k < D; ++k) {
    // partial pivoting
    int p = k;
    double best = M[k * D + k].x * M[k * D + k].x + M[k * D + k].y * M[k * D + k].y;
    for (int i = k + 1; i < D; ++i) {
      const double a = M[i * D + k].x * M[i * D + k].x + M[i * D + k].y * M[i * D + k].y;
      if (a > best) { best = a; p = i; }
    }
    if (p != k) {
      for (int j = 0; j < D; ++j) {
        double2 t = M[k * D + j]; M[k * D + j] = M[p * D + j]; M[p * D + j] = t;
        t = B[k * D + j]; B[k * D + j] = B[p * D + j]; B[p * D + j] = t;
      }
    }
    const double2 piv = M[k * D + k];
    const double inv = 1.0 / (piv.x * piv.x + piv.y * piv.y);
    const double2 pinv = make_double2(piv.x * inv, -piv.y * inv);
    for (int j = 0; j < D; ++j) {
      M[k * D + j] = cmul(M[k * D + j], pinv);
      B[k * D + j] = cmul(B[k * D + j], pinv);
    }
    for (int i = 0; i < D; ++i) {
      if (i == k) continue;
      const double2 fct = M[i * D + k];
      for (int j = 0; j < D; ++j) {
        const double2 a = cmul(fct, M[k * D + j]);
        M[i * D + j] = make_double2(M[i * D + j].x - a.x, M[i * D + j].y - a.y);
        const double2 b = cmul(fct, B[k * D + j]);
        B[i * D + j] = make_double2(B[i * D + j].x - b.x, B[i * D + j].y - b.y);
      }
    }
  }
}

// uuv: [nx][nky][3][D*D] complex (U0, U, V) for kx in [0,nx), ky = ky_first + [0,nky);
// phi_cols: first plane block of ky_first inside this handle's table.
template <int D>
__global__ void __launch_bounds__(64)
k_build_phi(const double2 *__restrict__ uuv, int nx, int nky, int height, double scale, int fast, int top,
            int lognx, double *__restrict__ phi_cols, int *not_converged)
{
  const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long) nx * nky) return;
  const int kx = (int) (idx / nky), kyl = (int) (idx % nky);
  const double2 *src = uuv + (size_t) idx * 3 * D * D;
  double2 U[D * D], V[D * D], Vd[D * D], VT[D * D], M[D * D], Y[D * D];
  for (int i = 0; i < D * D; ++i) {
    U[i] = src[D * D + i];
    V[i] = src[2 * D * D + i];
  }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) Vd[i * D + j] = make_double2(-V[j * D + i].x, V[j * D + i].y);

  if (height != 0) {
    const int maxit = height > 0 ? height : 100001;      // 1 initial step + up to 100000 iterations
    const double eps2 = 1e-8 * 1e-8;
    double dnorm = 1.0;
    int it = 0;
    for (; it < maxit && (height > 0 || it == 0 || dnorm > eps2); ++it) {
      for (int i = 0; i < D * D; ++i) {
        M[i] = it == 0 ? U[i] : make_double2(U[i].x + VT[i].x, U[i].y + VT[i].y);
        Y[i] = Vd[i];
      }
      csolve<D>(M, Y);
      if (it > 0) dnorm = 0.0;       // the step after the initial one always runs (iterate_Gnn starts at 1.0)
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
          double2 acc = make_double2(0.0, 0.0);
          for (int k = 0; k < D; ++k) {
            const double2 t = cmul(V[i * D + k], Y[k * D + j]);
            acc.x += t.x;
            acc.y += t.y;
          }
          if (it > 0) {
            const double dx = acc.x - VT[i * D + j].x, dy = acc.y - VT[i * D + j].y;
            dnorm += dx * dx + dy * dy;
          }
          M[i * D + j] = acc;        // M is free again: VT is still needed for the differences
        }
      for (int i = 0; i < D * D; ++i) VT[i] = M[i];
    }
    if (height < 0 && dnorm > eps2 && not_converged) *not_converged = 1;   // every writer stores 1
  }
  // Phi = U0 (+ VT), Hermitian part, scaled, packed
  size_t off, cstride;
  phi_slot(fast, top, lognx, nx, (size_t) D * D, kx, off, cstride);
  double *dst = phi_cols + (size_t) kyl * D * D * nx + off;
  auto P = [&](int i, int j) {
    double2 a = src[i * D + j];
    if (height != 0) { a.x += VT[i * D + j].x; a.y += VT[i * D + j].y; }
    return a;
  };
  int c = D;
  for (int i = 0; i < D; ++i) {
    dst[(size_t) i * cstride] = P(i, i).x * scale;
    for (int j = i + 1; j < D; ++j) {
      const double2 a = P(i, j), b = P(j, i);
      dst[(size_t) c * cstride] = 0.5 * (a.x + b.x) * scale;
      dst[(size_t) (c + 1) * cstride] = 0.5 * (a.y - b.y) * scale;
      c += 2;
    }
  }
}

}  // namespace gfmd
