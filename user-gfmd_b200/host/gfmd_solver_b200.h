/* GFMDSolverB200 -- `solver static/b200` for the LAMMPS `fix gfmd` (user-gfmd).
 *
 * Drop-in replacement of GFMDSolverStatic (reference src/solvers/gfmd_solver_static.{h,cpp})
 * behind the unchanged plugin interface GFMDSolver (reference src/main/gfmd_solver.h:34-123).
 * All arithmetic of the per-step path runs on a B200 through the C ABI of libgfmd_b200.so
 * (include/gfmd_b200.h); this file is host glue only: it owns no numerics.
 *
 * Installation into a LAMMPS + USER-GFMD tree: see INTEGRATION.md (copy this pair of files
 * next to gfmd_solver_static.*, add the two marked lines to gfmd_solver_factory, link
 * -lgfmd_b200).
 */
#ifndef GFMD_SOLVER_B200_H
#define GFMD_SOLVER_B200_H

#include "gfmd_solver.h"

struct gfmd_b200;

namespace LAMMPS_NS {

class GFMDSolverB200 : public GFMDSolver {
 public:
  GFMDSolverB200(LAMMPS *, int, int *, char **);
  virtual ~GFMDSolverB200();

  virtual void init();
  virtual void set_grid_size(int, int, int);
  virtual void set_kernel(StiffnessKernel *, bool normalize = true);

  /* asynchronous start: upload u and launch the step while LAMMPS computes pair forces
     (hook at src/main/fix_gfmd.cpp:853) */
  virtual void pre_force(void *, void *);
  virtual double post_force(void *, void *, char *);

  /* gP = IDFT[(Phi(q) + Cavg)^-1 DFT[g]] (GFMDSolverStatic::prec_gradient) */
  virtual void prec_gradient(double *, double **, double **);

  virtual double memory_usage();

  /* init-time diagnostics of `fix gfmd ... dump_stiffness / dump_greens_function`
     (GFMDSolverFFT::dump_stiffness, ::dump_greens_function) */
  virtual void dump_stiffness();
  virtual void dump_greens_function();

 protected:
  struct gfmd_b200 *handle_;
  int device_;
  bool async_;
  bool pin_;       /* `pin`: page-lock the fix's u_xy / f_xy (see the constructor) */
  StiffnessKernel *kernel_;     /* of the last set_kernel; owned by the fix */
  bool normalize_;
  void check(int rc, const char *what);
  void dump(char *dump_prefix, double *u);
  void dump_table(const char *stem, bool invert);
};

}

#endif
