/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for LAMMPS' KSPACE FFT3d wrapper (fft3d_wrap.h, NOT part of
 * /root/reference; upstream it dispatches to FFTW3).  Call sites it serves:
 * src/solvers/gfmd_solver_fft.cpp:72-80 (ctor, nfast=1, nmid=ny, nslow=nx),
 * :116 (compute(..., 1) = forward, exp(-i q r) since LAMMPS patch_10Feb2021,
 * reference README.md:36-38) and :181 (compute(..., -1) = backward).  Both
 * directions unnormalised (scaled = 0).  Single rank only: the brick must be
 * the whole grid.
 *
 * Backend is chosen at run time through oracle_fft_backend:
 *   0 = direct long-double DFT (arbiter, small grids), 1 = oracle/fft_plain.c
 */
#ifndef ORACLE_SHIM_FFT3D_WRAP_H
#define ORACLE_SHIM_FFT3D_WRAP_H

#include "pointers.h"
#include "../fft_plain.h"

extern "C" int oracle_fft_backend;

namespace LAMMPS_NS {

class FFT3d : protected Pointers {
 public:
  FFT3d(LAMMPS *lmp, MPI_Comm, int nfast, int nmid, int nslow,
        int in_ilo, int in_ihi, int in_jlo, int in_jhi, int in_klo, int in_khi,
        int, int, int, int, int, int,
        int scaled, int permute, int *nbuf, int /*usecollective*/)
      : Pointers(lmp), nx_(nslow), ny_(nmid), plan_(NULL) {
    if (nfast != 1 || in_ilo != 0 || in_ihi != 0 || in_jlo != 0 ||
        in_jhi != nmid - 1 || in_klo != 0 || in_khi != nslow - 1 ||
        scaled != 0 || permute != 0)
      error->all(FLERR, "FFT3d shim: only the single-rank full-grid "
                        "unscaled 1 x ny x nx case is supported.");
    *nbuf = nslow * nmid;
  }
  ~FFT3d() { if (plan_) fftp_destroy(plan_); }

  void compute(double *in, double *out, int flag) {
    if (in != out)
      error->all(FLERR, "FFT3d shim: in-place transforms only.");
    int sign = (flag == 1) ? -1 : +1;
    if (oracle_fft_backend == 0) {
      fftp_dft2d_ld(nx_, ny_, in, sign);
    } else {
      if (!plan_) plan_ = fftp_plan_2d(nx_, ny_);
      fftp_exec_2d(plan_, in, sign);
    }
  }

 private:
  int nx_, ny_;
  fftp_plan *plan_;
};

}

#endif
