/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Runs the reference's GFMDSolverStatic and this repository's GFMDSolverB200
 * (user-gfmd_b200/host/gfmd_solver_b200.cpp, the C++ glue a LAMMPS build would
 * compile) through the SAME plugin interface, GFMDSolver (src/main/gfmd_solver.h),
 * with the SAME StiffnessKernel object and the same u_xy, and reports the
 * differences.  Built by `make -C oracle hostshim` against oracle/shim.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pointers.h"
#include "surface_stiffness.h"
#include "gfmd_solver.h"
#include "gfmd_solver_static.h"
#include "gfmd_solver_b200.h"

using namespace LAMMPS_NS;

extern "C" int oracle_fft_backend;

static std::vector<char *> split(const char *s)
{
  std::vector<char *> out;
  std::string cur;
  for (const char *p = s;; p++) {
    if (*p == ' ' || *p == 0) {
      if (!cur.empty()) { out.push_back(strdup(cur.c_str())); cur.clear(); }
      if (*p == 0) break;
    } else cur.push_back(*p);
  }
  return out;
}

/* out[0] = max|f_b200 - f_ref| / max|f_ref|, out[1] = |e_b200 - e_ref| / |e_ref|,
 * out[2] = max|u0 diff| / max(1, max|u0|), out[3] = e_ref.  Returns 0 on success. */
extern "C" int hostshim_compare(const char *kernel_string, int nx, int ny, unsigned seed, int use_pre_force,
                                double *out)
{
  LAMMPS *lmp = new LAMMPS();
  lmp->domain->set_cell(nx, ny, 1);
  std::vector<char *> argv = split(kernel_string);
  int carg = 1;
  StiffnessKernel *kernel = stiffness_kernel_factory(argv[0], (int) argv.size(), &carg, argv.data(),
                                                     lmp->domain, lmp->force, lmp->memory, lmp->error);
  if (!kernel) return 1;
  const int ndof = kernel->get_dimension();

  oracle_fft_backend = (nx * ny <= 4096) ? 0 : 1;
  char kw[] = "static";
  int c0 = 0;
  GFMDSolver *ref = gfmd_solver_factory(kw, lmp, 0, &c0, NULL);
  GFMDSolver *b200 = new GFMDSolverB200(lmp, 0, &c0, NULL);
  if (strcmp(b200->get_name(), "static/b200")) return 2;

  GFMDSolver *solvers[2] = {ref, b200};
  double **u, **f[2];
  lmp->memory->create(u, ndof, nx * ny, "u");
  lmp->memory->create(f[0], ndof, nx * ny, "f0");
  lmp->memory->create(f[1], ndof, nx * ny, "f1");
  srand(seed);
  for (int i = 0; i < ndof * nx * ny; i++) u[0][i] = 0.2 * (rand() / (double) RAND_MAX) - 0.1;

  double e[2];
  std::vector<double> u0[2];
  for (int s = 0; s < 2; s++) {
    solvers[s]->set_grid_size(nx, ny, ndof);
    if (solvers[s]->get_nxy_loc() != nx * ny || solvers[s]->get_xlo_loc() != 0) return 3;
    solvers[s]->init();
    solvers[s]->set_kernel(kernel);
    if (use_pre_force) solvers[s]->pre_force(u, f[s]);
    e[s] = solvers[s]->post_force(u, f[s], NULL);
    u0[s].assign(solvers[s]->get_u0(), solvers[s]->get_u0() + ndof);
  }

  double fmax = 0, dmax = 0, u0max = 1, du0 = 0;
  for (int i = 0; i < ndof * nx * ny; i++) {
    fmax = std::fmax(fmax, std::fabs(f[0][0][i]));
    dmax = std::fmax(dmax, std::fabs(f[0][0][i] - f[1][0][i]));
  }
  for (int i = 0; i < ndof; i++) {
    u0max = std::fmax(u0max, std::fabs(u0[0][i]));
    du0 = std::fmax(du0, std::fabs(u0[0][i] - u0[1][i]));
  }
  out[0] = dmax / fmax;
  out[1] = std::fabs(e[1] - e[0]) / std::fabs(e[0]);
  out[2] = du0 / u0max;
  out[3] = e[0];

  delete ref;
  delete b200;
  lmp->memory->destroy(u);
  lmp->memory->destroy(f[0]);
  lmp->memory->destroy(f[1]);
  delete kernel;
  for (char *a : argv) free(a);
  delete lmp;
  return 0;
}


/* The off-path services through the same plugin interface: dump_stiffness,
 * dump_greens_function, post_force on a `dumpq_every` step (dump_prefix != NULL) and
 * prec_gradient, run by the reference's GFMDSolverStatic in directory dir_ref and by
 * GFMDSolverB200 in directory dir_b200 (the dumps go to the current directory); the test
 * compares the two sets of files.  out[0] = max|gP_b200 - gP_ref| / max|gP_ref|,
 * out[1] = max|f_b200 - f_ref| / max|f_ref| of the dump step.  Returns 0 on success. */
#include <unistd.h>

extern "C" int hostshim_compare_aux(const char *kernel_string, int nx, int ny, unsigned seed,
                                    const char *dir_ref, const char *dir_b200, double cdiag, double *out)
{
  LAMMPS *lmp = new LAMMPS();
  lmp->domain->set_cell(nx, ny, 1);
  std::vector<char *> argv = split(kernel_string);
  int carg = 1;
  StiffnessKernel *kernel = stiffness_kernel_factory(argv[0], (int) argv.size(), &carg, argv.data(),
                                                     lmp->domain, lmp->force, lmp->memory, lmp->error);
  if (!kernel) return 1;
  const int ndof = kernel->get_dimension();

  oracle_fft_backend = (nx * ny <= 4096) ? 0 : 1;
  char kw[] = "static";
  int c0 = 0;
  GFMDSolver *solvers[2] = {gfmd_solver_factory(kw, lmp, 0, &c0, NULL), new GFMDSolverB200(lmp, 0, &c0, NULL)};
  const char *dirs[2] = {dir_ref, dir_b200};

  double **u, **f[2], **gP[2];
  lmp->memory->create(u, ndof, nx * ny, "u");
  for (int s = 0; s < 2; s++) {
    lmp->memory->create(f[s], ndof, nx * ny, "f");
    lmp->memory->create(gP[s], ndof, nx * ny, "gP");
  }
  srand(seed);
  for (int i = 0; i < ndof * nx * ny; i++) u[0][i] = 0.2 * (rand() / (double) RAND_MAX) - 0.1;
  std::vector<double> cavg((size_t) ndof * ndof);
  for (int i = 0; i < ndof; i++)
    for (int j = 0; j < ndof; j++)
      cavg[i * ndof + j] = ((i == j ? cdiag : 0.0) + 0.05 * cdiag * (rand() / (double) RAND_MAX - 0.5)) / (nx * ny);

  char cwd[4096];
  if (!getcwd(cwd, sizeof(cwd))) return 4;
  for (int s = 0; s < 2; s++) {
    if (chdir(dirs[s])) return 5;
    solvers[s]->set_grid_size(nx, ny, ndof);
    solvers[s]->init();
    solvers[s]->set_kernel(kernel);
    solvers[s]->dump_stiffness();
    solvers[s]->dump_greens_function();
    char prefix[] = "dump";
    solvers[s]->post_force(u, f[s], prefix);
    solvers[s]->prec_gradient(cavg.data(), u, gP[s]);
    if (chdir(cwd)) return 5;
  }

  double gmax = 0, gd = 0, fmax = 0, fd = 0;
  for (int i = 0; i < ndof * nx * ny; i++) {
    gmax = std::fmax(gmax, std::fabs(gP[0][0][i]));
    gd = std::fmax(gd, std::fabs(gP[0][0][i] - gP[1][0][i]));
    fmax = std::fmax(fmax, std::fabs(f[0][0][i]));
    fd = std::fmax(fd, std::fabs(f[0][0][i] - f[1][0][i]));
  }
  out[0] = gd / gmax;
  out[1] = fd / fmax;

  delete solvers[0];
  delete solvers[1];
  lmp->memory->destroy(u);
  for (int s = 0; s < 2; s++) {
    lmp->memory->destroy(f[s]);
    lmp->memory->destroy(gP[s]);
  }
  delete kernel;
  for (char *a : argv) free(a);
  delete lmp;
  return 0;
}
