/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library,
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call it.
 *
 * Plain-C restatement of the per-step elastic-force path of `fix gfmd`,
 * following the reference's order of operations line by line:
 *
 *   gfmd_oracle_gather      FixGFMD::pre_force        src/main/fix_gfmd.cpp:734-803
 *   gfmd_oracle_post_force  GFMDSolverStatic::post_force
 *                                                     src/solvers/gfmd_solver_static.cpp:145-249
 *       forward transform   GFMDSolverFFT::fft_forward src/solvers/gfmd_solver_fft.cpp:96-147
 *       per-q mat-vec       mat_mul_vec               src/mathutils/linearalgebra.h:197-206
 *       backward transform  GFMDSolverFFT::fft_reverse src/solvers/gfmd_solver_fft.cpp:150-195
 *   gfmd_oracle_scatter     FixGFMD::grid_to_list + post_force
 *                                                     src/main/fix_gfmd.cpp:952-1010, :896-902
 *
 * The FFT itself is LAMMPS FFT3d -> FFTW3 upstream (third-party, absent from
 * /root/reference, unpinned: "patch_10Feb2021 or later", README.md:36-38); it is
 * restated as the unnormalised complex 2-D DFT it is documented to be
 * (oracle/fft_plain.c).  PINNING: this file is checked against the reference's
 * own solver sources compiled unchanged (oracle/_ref/libgfmd_ref.so, see
 * oracle/ref_driver.cpp) in tests/test_oracle.py, and against the golden
 * vectors those sources produced (tests/golden/, made by
 * tests/golden/make_golden.py).  The reference's own tests hold no golden
 * vectors for this path.
 */
#include <stdlib.h>
#include <string.h>
#include "gfmd_oracle.h"
#include "fft_plain.h"

/* fix_gfmd.cpp:734-803 (orthogonal box branch).  x, xeq: [nall][3]; gid:
 * [nall][3] (ix, iy, iu), updated in place when a shift is applied exactly as
 * the reference does; u_xy: [ndof][nx_loc*ny_loc], y fastest.  Returns
 * natoms_cur (atoms stored on the local brick).  The velocity copy into f_xy
 * (:783-793) is dead for the static solver and not restated. */
int gfmd_oracle_gather(int nall, int nlocal, const double *x, const double *xeq,
                       int *gid, const int *mask, int groupbit,
                       int nx, int ny, int xlo_loc, int ylo_loc, int nx_loc, int ny_loc,
                       double xprd, double yprd, int dxshift, int dyshift,
                       double *u_xy)
{
  double xprd_half = 0.5 * xprd, yprd_half = 0.5 * yprd;
  size_t nxy_loc = (size_t) nx_loc * ny_loc;
  int natoms_cur = 0;
  (void) nlocal;
  for (int i = 0; i < nall; i++) {
    if (mask[i] & groupbit) {
      int ix = gid[3 * i] - dxshift;
      int iy = gid[3 * i + 1] - dyshift;
      int iu = gid[3 * i + 2];
      if (dxshift != 0 || dyshift != 0) {
        while (ix >= nx) ix -= nx;
        while (ix < 0) ix += nx;
        while (iy >= ny) iy -= ny;
        while (iy < 0) iy += ny;
        gid[3 * i] = ix;
        gid[3 * i + 1] = iy;
      }
      ix -= xlo_loc;
      iy -= ylo_loc;
      double uz = x[3 * i + 2] - xeq[3 * i + 2];
      if (ix >= 0 && ix < nx_loc && iy >= 0 && iy < ny_loc) {
        size_t iloc = (size_t) ix * ny_loc + iy;
        int idof = 3 * iu;
        double ux = x[3 * i] - xeq[3 * i];
        double uy = x[3 * i + 1] - xeq[3 * i + 1];
        while (ux > xprd_half) ux -= xprd;
        while (ux < -xprd_half) ux += xprd;
        while (uy > yprd_half) uy -= yprd;
        while (uy < -yprd_half) uy += yprd;
        u_xy[(size_t) idof * nxy_loc + iloc] = ux;
        u_xy[(size_t) (idof + 1) * nxy_loc + iloc] = uy;
        u_xy[(size_t) (idof + 2) * nxy_loc + iloc] = uz;
        natoms_cur++;
      }
    }
  }
  return natoms_cur;
}

/* gfmd_solver_static.cpp:145-249, single rank (xlo_loc = ylo_loc = 0, so
 * gammai_ = 0, gfmd_solver.cpp:132-135).  phi: [nx*ny][ndof*ndof] complex128
 * interleaved, row-major, already divided by nx*ny (gfmd_misc.cpp:88-99);
 * u, f: [ndof][nx*ny]; u0: [ndof].  fft_backend 0 = long-double DFT, 1 = fast. */
double gfmd_oracle_post_force(int nx, int ny, int ndof, const double *phi,
                              const double *linf, const double *u, double *f,
                              double *u0, int fft_backend)
{
  size_t nxy = (size_t) nx * ny;
  int nu = ndof / 3;
  double *fft_data = (double *) malloc(sizeof(double) * 2 * nxy);
  double *q = (double *) malloc(sizeof(double) * 2 * nxy * ndof); /* q_buffer_[idq][idim] */
  double *F_q = (double *) malloc(sizeof(double) * 2 * ndof);
  fftp_plan *plan = fft_backend ? fftp_plan_2d(nx, ny) : NULL;

  /* fft_forward, gfmd_solver_fft.cpp:96-147 */
  for (int idim = 0; idim < ndof; idim++) {
    for (size_t idx = 0; idx < nxy; idx++) {
      fft_data[2 * idx] = u[(size_t) idim * nxy + idx];
      fft_data[2 * idx + 1] = 0.0;
    }
    if (plan) fftp_exec_2d(plan, fft_data, -1);
    else fftp_dft2d_ld(nx, ny, fft_data, -1);
    for (size_t idq = 0; idq < nxy; idq++) {
      q[2 * (idq * ndof + idim)] = fft_data[2 * idq];
      q[2 * (idq * ndof + idim) + 1] = fft_data[2 * idq + 1];
    }
  }

  /* :160-176 q=0 displacement */
  for (int idim = 0; idim < ndof; idim++) u0[idim] = q[2 * idim];

  /* :187-196 gamma point energy */
  double epot = 0.0;
  for (int i = 0; i < nu; i++) epot -= 2 * linf[i] * q[2 * (3 * i + 2)];

  /* :202-208 F(q) = -Phi(q) U(q) and energy */
  for (size_t idq = 0; idq < nxy; idq++) {
    const double *M = phi + 2 * idq * ndof * ndof;
    double *v = q + 2 * idq * ndof;
    for (int i = 0; i < ndof; i++) {          /* mat_mul_vec, linearalgebra.h:197-206 */
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < ndof; j++) {
        double mr = M[2 * (i * ndof + j)], mi = M[2 * (i * ndof + j) + 1];
        /* std::complex operator*: (mr + i mi)(vr + i vi) */
        sr += mr * v[2 * j] - mi * v[2 * j + 1];
        si += mr * v[2 * j + 1] + mi * v[2 * j];
      }
      F_q[2 * i] = sr; F_q[2 * i + 1] = si;
    }
    for (int idim = 0; idim < ndof; idim++) {
      /* creal(conj(F) * u) */
      epot += F_q[2 * idim] * v[2 * idim] + F_q[2 * idim + 1] * v[2 * idim + 1];
      v[2 * idim] = -F_q[2 * idim];
      v[2 * idim + 1] = -F_q[2 * idim + 1];
    }
  }

  /* :228-230 gamma point force */
  for (int i = 0; i < nu; i++) q[2 * (3 * i + 2)] += linf[i];

  /* :236 */
  epot *= 0.5;

  /* fft_reverse, gfmd_solver_fft.cpp:150-195 */
  for (int idim = 0; idim < ndof; idim++) {
    for (size_t idq = 0; idq < nxy; idq++) {
      fft_data[2 * idq] = q[2 * (idq * ndof + idim)];
      fft_data[2 * idq + 1] = q[2 * (idq * ndof + idim) + 1];
    }
    if (plan) fftp_exec_2d(plan, fft_data, +1);
    else fftp_dft2d_ld(nx, ny, fft_data, +1);
    for (size_t idx = 0; idx < nxy; idx++)
      f[(size_t) idim * nxy + idx] = fft_data[2 * idx];
  }

  if (plan) fftp_destroy(plan);
  free(fft_data); free(q); free(F_q);
  return epot;
}

/* fix_gfmd.cpp:952-1010 (grid_to_list) followed by :896-902 (f += f_i).
 * f: [nall][3] accumulated in place; fsum_loc: sum over LOCAL atoms only.
 * Returns the number of atoms that received a force, or -1 on a duplicate. */
int gfmd_oracle_scatter(int nall, int nlocal, const int *gid, const int *mask,
                        int groupbit, int xlo_loc, int xhi_loc, int ylo_loc,
                        int yhi_loc, const double *f_xy, double *f,
                        double *fsum_loc)
{
  int nx_loc = xhi_loc - xlo_loc + 1, ny_loc = yhi_loc - ylo_loc + 1;
  size_t nxy_loc = (size_t) nx_loc * ny_loc;
  int nknown = 0;
  fsum_loc[0] = fsum_loc[1] = fsum_loc[2] = 0.0;
  for (int i = 0; i < nall; i++) {
    if (mask[i] & groupbit) {
      int ix = gid[3 * i], iy = gid[3 * i + 1];
      if (ix >= xlo_loc && ix <= xhi_loc && iy >= ylo_loc && iy <= yhi_loc) {
        ix -= xlo_loc; iy -= ylo_loc;
        int iu = gid[3 * i + 2];
        size_t iloc = (size_t) ix * ny_loc + iy;
        int idof = 3 * iu;
        double fx = f_xy[(size_t) idof * nxy_loc + iloc];
        double fy = f_xy[(size_t) (idof + 1) * nxy_loc + iloc];
        double fz = f_xy[(size_t) (idof + 2) * nxy_loc + iloc];
        if (i < nlocal) {
          fsum_loc[0] += fx; fsum_loc[1] += fy; fsum_loc[2] += fz;
        }
        f[3 * i] += fx; f[3 * i + 1] += fy; f[3 * i + 2] += fz;
        nknown++;
      }
    }
  }
  return nknown;
}
