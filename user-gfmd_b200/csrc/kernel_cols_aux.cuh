// Auxiliary column kernels.  NOT on the per-step path: they
// serve the reference solver's diagnostics and its optional preconditioner, for any grid
// the generic kernels cover, on a single rank (the reference's dumps are single-rank too,
// src/solvers/gfmd_solver_fft.cpp:211-212).
//
//   AUX_SPECTRUM  u~(q) = forward DFT of u, and F(q) = Phi(q).u~(q) (no sign flip) -- the two
//                 fields GFMDSolverFFT::dump writes on `dumpq_every` steps
//                 (src/solvers/gfmd_solver_fft.cpp:209-287, called from
//                 src/solvers/gfmd_solver_static.cpp:181-182).  No inverse transform.
//   AUX_PREC      gP(q) = (Phi(q) + Cavg)^-1 g(q), then the inverse transform:
//                 GFMDSolverStatic::prec_gradient (src/solvers/gfmd_solver_static.cpp:253-271)
//                 with precondition_gradient<DEF_G> (src/main/gfmd_misc.h:39-133).  The
//                 reference inverts with invert3x3 / full-pivot Gauss-Jordan; here the linear
//                 system is solved by elimination with partial pivoting (same result to
//                 rounding).  Cavg is a real ndof x ndof matrix, so M(-q) = conj M(q) and the
//                 half spectrum is still exact, although M need not be Hermitian.
//                 ncopy = ndof: all components of gP(q) are replaced; ncopy = 3 reproduces the
//                 reference's general (ndof > 3) branch, which copies back only the first
//                 three (`idim < 3`, gfmd_misc.h:113-115).
//
// Both read the Hermitian-packed table in whichever layout the per-step kernels of this
// handle use (phi_slot: plane-major, or digit-reversed/interleaved for the specialised ones).
#pragma once

#include "fft_pow2.cuh"
#include "kernels_generic.cuh"

namespace gfmd {

enum { AUX_SPECTRUM = 1, AUX_PREC = 2 };

// x = M^-1 b for one D x D complex system (row-major M); M and b are destroyed, b holds x.
template <int D>
__device__ __forceinline__ void csolve_vec(double2 *M, double2 *b)
{
  for (int k = 0; k < D; ++k) {
    int p = k;
    double best = M[k * D + k].x * M[k * D + k].x + M[k * D + k].y * M[k * D + k].y;
    for (int i = k + 1; i < D; ++i) {
      const double a = M[i * D + k].x * M[i * D + k].x + M[i * D + k].y * M[i * D + k].y;
      if (a > best) { best = a; p = i; }
    }
    if (p != k) {
      for (int j = k; j < D; ++j) {
        const double2 t = M[k * D + j]; M[k * D + j] = M[p * D + j]; M[p * D + j] = t;
      }
      const double2 t = b[k]; b[k] = b[p]; b[p] = t;
    }
    const double2 piv = M[k * D + k];
    const double inv = 1.0 / (piv.x * piv.x + piv.y * piv.y);
    const double2 pinv = make_double2(piv.x * inv, -piv.y * inv);
    for (int j = k; j < D; ++j) M[k * D + j] = cmul(M[k * D + j], pinv);
    b[k] = cmul(b[k], pinv);
    for (int i = 0; i < D; ++i) {
      if (i == k) continue;
      const double2 fct = M[i * D + k];
      for (int j = k; j < D; ++j) {
        const double2 a = cmul(fct, M[k * D + j]);
        M[i * D + j] = make_double2(M[i * D + j].x - a.x, M[i * D + j].y - a.y);
      }
      const double2 c = cmul(fct, b[k]);
      b[i] = make_double2(b[i].x - c.x, b[i].y - c.y);
    }
  }
}

// One CTA per ky column (single rank: stage layout [d][nyh][nx]).  DT = compile-time ndof
// (3, 6, 9, 12), or 0 = run-time ndof (AUX_SPECTRUM only).
//   stage    in: row-transformed field; out: u~ (AUX_SPECTRUM) or the column-inverse of gP~
//   stage_f  AUX_SPECTRUM: receives Phi.u~, same layout (may be null)
template <int DT, int MODE>
__global__ void __launch_bounds__(512)
k_cols_aux(double2 *__restrict__ stage, double2 *__restrict__ stage_f, GridDesc g, FftDesc fd,
           const double *__restrict__ phi, const double *__restrict__ cavg, int fast_phi, int lognx, int ld,
           int ncopy)
{
  extern __shared__ double2 smem[];
  const int d = DT > 0 ? DT : g.d;
  const int kl = blockIdx.x;
  const int nx = g.nx;
  const size_t dsq = (size_t) d * d;

  for (int idx = threadIdx.x; idx < d * nx; idx += blockDim.x) {
    const int dof = idx / nx, ix = idx - dof * nx;
    smem[dof * ld + ix] = stage[((size_t) dof * g.kyb + kl) * nx + ix];
  }
  __syncthreads();
  fft_batch<-1>(smem, ld, d, fd);

  const double *ph = phi + (size_t) kl * dsq * nx;
  for (int kx = threadIdx.x; kx < nx; kx += blockDim.x) {
    size_t off, cs;
    phi_slot(fast_phi, 0, lognx, nx, dsq, kx, off, cs);
    if (MODE == AUX_SPECTRUM) {
      if (!stage_f) continue;
      if (DT > 0) {
        double2 uv[DT > 0 ? DT : 1], F[DT > 0 ? DT : 1];
#pragma unroll
        for (int i = 0; i < DT; ++i) uv[i] = smem[i * ld + kx];
        phi_matvec<DT>(uv, F, [&](int c) { return __ldg(ph + off + (size_t) c * cs); });
#pragma unroll
        for (int i = 0; i < DT; ++i) stage_f[((size_t) i * g.kyb + kl) * nx + kx] = F[i];
      } else {
        double2 uv[24], F[24];
        for (int i = 0; i < d; ++i) uv[i] = smem[i * ld + kx];
        for (int i = 0; i < d; ++i) {
          const double a = __ldg(ph + off + (size_t) i * cs);
          F[i] = make_double2(a * uv[i].x, a * uv[i].y);
        }
        int c = d;
        for (int i = 0; i < d; ++i)
          for (int j = i + 1; j < d; ++j) {
            const double2 p = make_double2(__ldg(ph + off + (size_t) c * cs), __ldg(ph + off + (size_t) (c + 1) * cs));
            c += 2;
            F[i].x = fma(p.x, uv[j].x, fma(-p.y, uv[j].y, F[i].x));
            F[i].y = fma(p.x, uv[j].y, fma(p.y, uv[j].x, F[i].y));
            F[j].x = fma(p.x, uv[i].x, fma(p.y, uv[i].y, F[j].x));
            F[j].y = fma(p.x, uv[i].y, fma(-p.y, uv[i].x, F[j].y));
          }
        for (int i = 0; i < d; ++i) stage_f[((size_t) i * g.kyb + kl) * nx + kx] = F[i];
      }
    } else if (DT > 0) {
      // M = Phi(q) + Cavg, full matrix from the packed planes
      double2 M[DT > 0 ? DT * DT : 1], b[DT > 0 ? DT : 1];
      int c = DT;
      for (int i = 0; i < DT; ++i) {
        M[i * DT + i] = make_double2(__ldg(ph + off + (size_t) i * cs) + __ldg(cavg + i * DT + i), 0.0);
        for (int j = i + 1; j < DT; ++j) {
          const double re = __ldg(ph + off + (size_t) c * cs), im = __ldg(ph + off + (size_t) (c + 1) * cs);
          c += 2;
          M[i * DT + j] = make_double2(re + __ldg(cavg + i * DT + j), im);
          M[j * DT + i] = make_double2(re + __ldg(cavg + j * DT + i), -im);
        }
      }
      for (int i = 0; i < DT; ++i) b[i] = smem[i * ld + kx];
      csolve_vec<DT>(M, b);
      for (int i = 0; i < DT; ++i)
        if (i < ncopy) smem[i * ld + kx] = b[i];
    }
  }

  __syncthreads();
  if (MODE == AUX_PREC) fft_batch<+1>(smem, ld, d, fd);

  for (int idx = threadIdx.x; idx < d * nx; idx += blockDim.x) {
    const int dof = idx / nx, ix = idx - dof * nx;
    stage[((size_t) dof * g.kyb + kl) * nx + ix] = smem[dof * ld + ix];
  }
}

// The per-q operation of k_cols_aux on a spectrum that lives in HBM (natural kx order, single rank):
// for column sets that do not fit one CTA, between the transform phases of the three-phase column stage
// (k_cols_split_fft, kernel_cols_split.cuh).  One thread per (kl, kx).
template <int DT, int MODE>
__global__ void __launch_bounds__(256)
k_aux_perq(double2 *__restrict__ stage, double2 *__restrict__ stage_f, GridDesc g, const double *__restrict__ phi,
           const double *__restrict__ cavg, int phi_mode, int top, int lognx, int ncopy)
{
  const int d = DT > 0 ? DT : g.d;
  const int nx = g.nx;
  const size_t dsq = (size_t) d * d;
  const long long total = (long long) g.nky_loc * nx;
  for (long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x) {
    const int kl = (int) (idx / nx), kx = (int) (idx - (long long) kl * nx);
    const double *ph = phi + (size_t) kl * dsq * nx;
    size_t off, cs;
    phi_slot(phi_mode, top, lognx, nx, dsq, kx, off, cs);
    auto at = [&](int i) -> size_t { return ((size_t) i * g.kyb + kl) * nx + kx; };
    if (MODE == AUX_SPECTRUM) {
      if (!stage_f) continue;
      if (DT > 0) {
        double2 uv[DT > 0 ? DT : 1], F[DT > 0 ? DT : 1];
#pragma unroll
        for (int i = 0; i < DT; ++i) uv[i] = stage[at(i)];
        phi_matvec<DT>(uv, F, [&](int c) { return __ldg(ph + off + (size_t) c * cs); });
#pragma unroll
        for (int i = 0; i < DT; ++i) stage_f[at(i)] = F[i];
      } else {
        double2 uv[24], F[24];
        for (int i = 0; i < d; ++i) uv[i] = stage[at(i)];
        for (int i = 0; i < d; ++i) {
          const double a = __ldg(ph + off + (size_t) i * cs);
          F[i] = make_double2(a * uv[i].x, a * uv[i].y);
        }
        int c = d;
        for (int i = 0; i < d; ++i)
          for (int j = i + 1; j < d; ++j) {
            const double2 p = make_double2(__ldg(ph + off + (size_t) c * cs), __ldg(ph + off + (size_t) (c + 1) * cs));
            c += 2;
            F[i].x = fma(p.x, uv[j].x, fma(-p.y, uv[j].y, F[i].x));
            F[i].y = fma(p.x, uv[j].y, fma(p.y, uv[j].x, F[i].y));
            F[j].x = fma(p.x, uv[i].x, fma(p.y, uv[i].y, F[j].x));
            F[j].y = fma(p.x, uv[i].y, fma(-p.y, uv[i].x, F[j].y));
          }
        for (int i = 0; i < d; ++i) stage_f[at(i)] = F[i];
      }
    } else if (DT > 0) {
      double2 M[DT > 0 ? DT * DT : 1], b[DT > 0 ? DT : 1];
      int c = DT;
      for (int i = 0; i < DT; ++i) {
        M[i * DT + i] = make_double2(__ldg(ph + off + (size_t) i * cs) + __ldg(cavg + i * DT + i), 0.0);
        for (int j = i + 1; j < DT; ++j) {
          const double re = __ldg(ph + off + (size_t) c * cs), im = __ldg(ph + off + (size_t) (c + 1) * cs);
          c += 2;
          M[i * DT + j] = make_double2(re + __ldg(cavg + i * DT + j), im);
          M[j * DT + i] = make_double2(re + __ldg(cavg + j * DT + i), -im);
        }
      }
      for (int i = 0; i < DT; ++i) b[i] = stage[at(i)];
      csolve_vec<DT>(M, b);
      for (int i = 0; i < DT; ++i)
        if (i < ncopy) stage[at(i)] = b[i];
    }
  }
}

}  // namespace gfmd
