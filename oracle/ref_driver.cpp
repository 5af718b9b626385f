/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * C entry points around the UNMODIFIED reference sources, compiled in place
 * from /root/reference by oracle/Makefile into oracle/_ref/libgfmd_ref.so:
 *
 *   - stiffness kernels + fill_phi_buffer  (src/main/surface_stiffness.cpp,
 *     src/main/gfmd_misc.cpp:32-101, src/stiffness_kernels/...)
 *   - the solver itself: GFMDSolverStatic::post_force
 *     (src/solvers/gfmd_solver_static.cpp:145-249) on top of
 *     GFMDSolverFFT::fft_forward/fft_reverse (src/solvers/gfmd_solver_fft.cpp:96-195)
 *     created through gfmd_solver_factory (src/main/gfmd_solver.cpp:209-257),
 *     with LAMMPS' FFT3d replaced by oracle/shim/fft3d_wrap.h.
 *
 * The only code of ours on that path is the shim (LAMMPS object, FFT backend)
 * and the pair-potential-only force_constants_factory below (the reference's
 * own factory needs LAMMPS pair-style headers, src/main/force_constants.cpp:26-31).
 */
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "pointers.h"
#include "surface_stiffness.h"
#include "force_constants.h"
#include "fc_pair_potential.h"
#include "gfmd_misc.h"
#include "gfmd_solver.h"
#include "gfmd_solver_static.h"

using namespace LAMMPS_NS;

extern "C" { int oracle_fft_backend = 0; }

/* Replacement for src/main/force_constants.cpp:113-164, "pair-potential" only. */
ForceConstants *force_constants_factory(char *keyword, int narg, int *carg,
                                        char **arg, CrystalSurface *surface,
                                        Force *force, Memory *memory,
                                        Error *error)
{
  if (!strcmp(keyword, "pair-potential"))
    return new FCPairPotential(narg, carg, arg, surface, force, error);
  return NULL;
}

namespace {

struct RefKernel {
  LAMMPS *lmp;
  StiffnessKernel *kernel;
  std::vector<char *> argv;
};

struct RefSolver {
  LAMMPS *lmp;
  GFMDSolver *solver;
  int nx, ny, ndof;
  double **u, **f;
};

std::vector<char *> split(const char *s)
{
  std::vector<char *> out;
  std::string cur;
  for (const char *p = s;; p++) {
    if (*p == ' ' || *p == '\t' || *p == 0) {
      if (!cur.empty()) { out.push_back(strdup(cur.c_str())); cur.clear(); }
      if (*p == 0) break;
    } else cur.push_back(*p);
  }
  return out;
}

}

extern "C" {

void ref_set_fft_backend(int b) { oracle_fft_backend = b; }

/* kernel_string = what follows "fix ID group gfmd <prefix>" up to the fix
 * keywords, e.g. "ft sc100 1 1.0 pair-potential 2 1.0 1.0 height 128"
 * (parsed like src/main/fix_gfmd.cpp:190-200). */
void *ref_kernel_create(const char *kernel_string, int invariant)
{
  RefKernel *k = new RefKernel;
  k->lmp = new LAMMPS();
  k->argv = split(kernel_string);
  int narg = (int) k->argv.size();
  if (narg < 1) return NULL;
  int carg = 1;
  k->kernel = stiffness_kernel_factory(k->argv[0], narg, &carg, k->argv.data(),
                                       k->lmp->domain, k->lmp->force,
                                       k->lmp->memory, k->lmp->error);
  if (!k->kernel) return NULL;
  if (invariant) k->kernel->set_invariant(true);
  return k;
}

int ref_kernel_ndof(void *h)
{
  RefKernel *k = (RefKernel *) h;
  return k->kernel->get_dimension();
}

int ref_kernel_nu(void *h)
{
  return ((RefKernel *) h)->kernel->get_number_of_atoms();
}

/* phi_out: [nx*ny][ndof*ndof] complex128 interleaved, row-major, idq=ix*ny+iy */
void ref_fill_phi(void *h, int nx, int ny, double *phi_out, int normalize)
{
  RefKernel *k = (RefKernel *) h;
  int ndof = k->kernel->get_dimension();
  size_t nq = (size_t) nx * ny, dsq = (size_t) ndof * ndof;
  double_complex *data = reinterpret_cast<double_complex *>(phi_out);
  std::vector<double_complex *> rows(nq);
  for (size_t i = 0; i < nq; i++) rows[i] = data + i * dsq;
  fill_phi_buffer(ndof, nx, 0, nx - 1, ny, 0, ny - 1, k->kernel, rows.data(),
                  normalize != 0, k->lmp->error);
}

/* Phi at an arbitrary wavevector (unnormalised), for sampled checks. */
void ref_phi_at(void *h, int nx, int ny, double qx, double qy, double *phi_out)
{
  RefKernel *k = (RefKernel *) h;
  k->kernel->pre_compute();
  k->kernel->get_stiffness_matrix(nx, ny, qx, qy,
                                  reinterpret_cast<double_complex *>(phi_out));
  k->kernel->post_compute();
}

/* U0, U, V of StiffnessKernel::get_dynamical_matrices for kx in [0,nx), ky in
 * [ky_first, ky_first+nky), q as in fill_phi_buffer (gfmd_misc.cpp:43-48);
 * out: [nx][nky][3][ndof*ndof] complex128. */
void ref_dynamical_matrices_columns(void *h, int nx, int ny, int ky_first, int nky, double *out)
{
  RefKernel *k = (RefKernel *) h;
  int ndof = k->kernel->get_dimension();
  size_t dsq = (size_t) ndof * ndof;
  double_complex *o = reinterpret_cast<double_complex *>(out);
  k->kernel->pre_compute();
  for (int i = 0; i < nx; i++) {
    double qx = (i <= int((nx)/2)) ? (2.0*M_PI*(i)/nx) : (2.0*M_PI*(i-nx)/nx);
    for (int jj = 0; jj < nky; jj++) {
      int j = ky_first + jj;
      double qy = (j <= int((ny)/2)) ? (2.0*M_PI*(j)/ny) : (2.0*M_PI*(j-ny)/ny);
      double_complex *base = o + ((size_t) i * nky + jj) * 3 * dsq;
      k->kernel->get_dynamical_matrices(qx, qy, base, base + dsq, base + 2 * dsq);
    }
  }
  k->kernel->post_compute();
}

int ref_kernel_height(void *h) { return ((RefKernel *) h)->kernel->get_height(); }

void ref_get_linf(void *h, double *linf)
{
  RefKernel *k = (RefKernel *) h;
  k->kernel->get_force_at_gamma_point(linf);
}

void ref_kernel_destroy(void *h)
{
  RefKernel *k = (RefKernel *) h;
  delete k->kernel;
  for (char *a : k->argv) free(a);
  delete k->lmp;
  delete k;
}

/* The reference solver ("static") on the full nx x ny grid, single rank. */
void *ref_solver_create(int nx, int ny, int ndof)
{
  RefSolver *s = new RefSolver;
  s->lmp = new LAMMPS();
  s->lmp->domain->set_cell(nx, ny, 1);
  char kw[] = "static";
  int carg = 0;
  s->solver = gfmd_solver_factory(kw, s->lmp, 0, &carg, NULL);
  s->solver->set_grid_size(nx, ny, ndof);
  s->solver->init();
  s->nx = nx; s->ny = ny; s->ndof = ndof;
  s->lmp->memory->create(s->u, ndof, nx * ny, "ref:u");
  s->lmp->memory->create(s->f, ndof, nx * ny, "ref:f");
  return s;
}

void ref_solver_set_kernel(void *hs, void *hk, int normalize)
{
  RefSolver *s = (RefSolver *) hs;
  RefKernel *k = (RefKernel *) hk;
  s->solver->set_kernel(k->kernel, normalize != 0);
}

/* Overwrite the solver's Phi table / linf with caller data (used to feed
 * synthetic or golden tables through the reference arithmetic). */
void ref_solver_set_phi(void *hs, const double *phi, const double *linf);

/* u, f: [ndof][nx*ny] contiguous.  Returns epot; u0: [ndof]. */
double ref_solver_post_force(void *hs, const double *u, double *f, double *u0)
{
  RefSolver *s = (RefSolver *) hs;
  size_t n = (size_t) s->ndof * s->nx * s->ny;
  memcpy(s->u[0], u, n * sizeof(double));
  double epot = s->solver->post_force(s->u, s->f, NULL);
  memcpy(f, s->f[0], n * sizeof(double));
  memcpy(u0, s->solver->get_u0(), s->ndof * sizeof(double));
  return epot;
}

/* GFMDSolverStatic::prec_gradient (src/solvers/gfmd_solver_static.cpp:253-271).
 * cavg: [ndof*ndof]; g, gP: [ndof][nx*ny] contiguous. */
void ref_solver_prec_gradient(void *hs, const double *cavg, const double *g, double *gP)
{
  RefSolver *s = (RefSolver *) hs;
  size_t n = (size_t) s->ndof * s->nx * s->ny;
  std::vector<double> c(cavg, cavg + (size_t) s->ndof * s->ndof);
  memcpy(s->u[0], g, n * sizeof(double));
  s->solver->prec_gradient(c.data(), s->u, s->f);
  memcpy(gP, s->f[0], n * sizeof(double));
}

/* post_force with a dump prefix: GFMDSolverFFT::dump writes <prefix>.q.*.out
 * (src/solvers/gfmd_solver_fft.cpp:209-287). */
double ref_solver_post_force_dump(void *hs, const double *u, double *f, const char *prefix)
{
  RefSolver *s = (RefSolver *) hs;
  size_t n = (size_t) s->ndof * s->nx * s->ny;
  memcpy(s->u[0], u, n * sizeof(double));
  std::string p(prefix);
  double epot = s->solver->post_force(s->u, s->f, &p[0]);
  memcpy(f, s->f[0], n * sizeof(double));
  return epot;
}

/* dump_stiffness / dump_greens_function write phi??.*.out / g??.*.out into the
 * current directory (src/solvers/gfmd_solver_fft.cpp:294-355, :362-430). */
void ref_solver_dump_tables(void *hs)
{
  RefSolver *s = (RefSolver *) hs;
  s->solver->dump_stiffness();
  s->solver->dump_greens_function();
}

void ref_solver_destroy(void *hs)
{
  RefSolver *s = (RefSolver *) hs;
  s->lmp->memory->destroy(s->u);
  s->lmp->memory->destroy(s->f);
  delete s->solver;
  delete s->lmp;
  delete s;
}

}

/* Access to the protected phi / linf_ members without touching the reference
 * sources: a derived class in this translation unit. */
namespace {
class StaticPeek : public GFMDSolverStatic {
 public:
  static void overwrite(GFMDSolverStatic *s, int nxy, int ndof,
                        const double *phi_in, const double *linf) {
    StaticPeek *p = static_cast<StaticPeek *>(s);
    if (p->phi) p->destroy_complex_operator_buffer(p->phi);
    p->phi = p->create_complex_operator_buffer("ref:phi");
    memcpy(p->phi[0], phi_in,
           sizeof(double_complex) * (size_t) nxy * ndof * ndof);
    for (int i = 0; i < ndof / 3; i++) p->linf_[i] = linf ? linf[i] : 0.0;
  }
};
}

extern "C" void ref_solver_set_phi(void *hs, const double *phi,
                                   const double *linf)
{
  RefSolver *s = (RefSolver *) hs;
  StaticPeek::overwrite(static_cast<GFMDSolverStatic *>(s->solver),
                        s->nx * s->ny, s->ndof, phi, linf);
}
