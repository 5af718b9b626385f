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

  virtual double memory_usage();

 protected:
  struct gfmd_b200 *handle_;
  int device_;
  bool async_;
  void check(int rc, const char *what);
};

}

#endif
