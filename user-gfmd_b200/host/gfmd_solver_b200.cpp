/* Host glue of `solver static/b200`: see gfmd_solver_b200.h.  Mirrors, call by call,
 * what GFMDSolverStatic does (reference src/solvers/gfmd_solver_static.cpp):
 *
 *   ctor            :47-62    name, optional solver arguments
 *   set_grid_size   :79-87    -> gfmd_b200_create[_slab]; the brick comes from the library
 *   set_kernel      :90-136   fill_phi_buffer for this rank's q columns -> gfmd_b200_set_phi_columns,
 *                             kernel->get_force_at_gamma_point -> gfmd_b200_set_linf, same warning
 *   post_force      :145-249  -> gfmd_b200_post_force_host (u_xy/f_xy are the fix's host arrays);
 *                             on `dumpq_every` steps the q-space fields come from
 *                             gfmd_b200_spectrum_host and are written in the format of
 *                             GFMDSolverFFT::dump (src/solvers/gfmd_solver_fft.cpp:209-287)
 *   prec_gradient   :253-271  -> gfmd_b200_prec_gradient_host
 *   dump_stiffness / dump_greens_function (gfmd_solver_fft.cpp:294-355, :362-430): init-time,
 *                             host only, from the stiffness kernel of the last set_kernel
 */
#include <numeric>
#include <vector>
#include <string.h>
#include <stdlib.h>

#include "gfmd_solver_b200.h"

#include "pointers.h"
#include "comm.h"
#include "domain.h"
#include "memory.h"
#include "mpi.h"

#include "gfmd_misc.h"
#include "linearalgebra.h"
#include "gfmd_b200.h"

using namespace LAMMPS_NS;

GFMDSolverB200::GFMDSolverB200(LAMMPS *lmp, int narg, int *iarg, char **arg)
  : GFMDSolver(lmp), handle_(NULL), device_(-1), async_(true), pin_(false), kernel_(NULL), normalize_(true)
{
  strcpy(name, "static/b200");

  /* optional: `solver static/b200 device <id>`, `sync` (no launch from pre_force) and `pin`
     (page-lock the fix's u_xy / f_xy for full-rate PCIe copies).  `pin` is opt-in because the
     library would then hold a registration on memory the FIX owns: FixGFMD::~FixGFMD frees
     u_xy / f_xy (fix_gfmd.cpp:479-480) BEFORE it deletes the solver (:486), i.e. while they are
     still registered.  Use it together with the two-line reordering of that destructor shown in
     INTEGRATION.md (delete the solver first). */
  while (narg > 0 && *iarg < narg) {
    if (!strcmp(arg[*iarg], "device") && *iarg + 1 < narg) {
      device_ = atoi(arg[*iarg + 1]);
      (*iarg) += 2;
    }
    else if (!strcmp(arg[*iarg], "sync")) {
      async_ = false;
      (*iarg)++;
    }
    else if (!strcmp(arg[*iarg], "pin")) {
      pin_ = true;
      (*iarg)++;
    }
    else break;
  }
}


GFMDSolverB200::~GFMDSolverB200()
{
  if (handle_) gfmd_b200_destroy(handle_);
}


void GFMDSolverB200::check(int rc, const char *what)
{
  if (rc) {
    char errstr[1280];
    snprintf(errstr, sizeof(errstr), "fix gfmd solver static/b200: %s failed: %s", what,
             gfmd_b200_last_error(handle_));
    error->one(FLERR, errstr);
  }
}


void GFMDSolverB200::set_grid_size(int in_nx, int in_ny, int in_ndof)
{
  nx = in_nx;
  ny = in_ny;
  nu_ = in_ndof/3;
  ndof = in_ndof;
  ndof_sq = ndof*ndof;

  if (u0) delete [] u0;
  u0 = new double[ndof];

  /* The B200 path decomposes the grid into x-slabs, one per rank (procgrid = P x 1 x 1,
     which is what fix gfmd requires anyway for z: src/main/fix_gfmd.cpp:167-170). */
  if (comm->procgrid[1] != 1 || comm->procgrid[2] != 1)
    error->all(FLERR,"fix gfmd solver static/b200 needs a processor grid of P x 1 x 1 "
               "(use 'processors * 1 1').");

  int dev = device_;
  if (dev < 0) {
    const char *lr = getenv("OMPI_COMM_WORLD_LOCAL_RANK");
    if (!lr) lr = getenv("MV2_COMM_WORLD_LOCAL_RANK");
    if (!lr) lr = getenv("SLURM_LOCALID");
    dev = lr ? atoi(lr) : 0;
  }

  if (handle_) gfmd_b200_destroy(handle_);
  handle_ = NULL;
  if (nprocs == 1) {
    check(gfmd_b200_create(&handle_, nx, ny, ndof, dev), "gfmd_b200_create");
  }
  else {
    check(gfmd_b200_create_slab(&handle_, nx, ny, ndof, dev, me, nprocs), "gfmd_b200_create_slab");
    char id[GFMD_B200_UNIQUE_ID_BYTES];
    if (me == 0) check(gfmd_b200_get_unique_id(id), "gfmd_b200_get_unique_id");
    MPI_Bcast(id, GFMD_B200_UNIQUE_ID_BYTES, MPI_CHAR, 0, world);
    check(gfmd_b200_comm_init(handle_, id), "gfmd_b200_comm_init");

    /* NVLink peer mappings (CUDA IPC): every rank exposes its two receive buffers and the buffer
       its row kernels write, everyone maps everyone's -- the transposes then run as peer pushes
       (or, with GFMD_B200_PEER_DIRECT=1, inside the column kernels) instead of ncclSend/ncclRecv.
       Optional: if any rank cannot export or import (GPUs of other ranks not visible to this
       process, no peer access), all ranks stay on the NCCL exchange.  GFMD_B200_NO_PEER=1 skips it. */
    if (!getenv("GFMD_B200_NO_PEER")) {
      const int nb = 2*GFMD_B200_IPC_HANDLE_BYTES;
      char mine[3*GFMD_B200_IPC_HANDLE_BYTES];
      int ok = gfmd_b200_ipc_export(handle_, mine) == 0 &&
               gfmd_b200_ipc_export_stage(handle_, mine+nb) == 0;
      int allok = 0;
      MPI_Allreduce(&ok, &allok, 1, MPI_INT, MPI_MIN, world);
      if (allok) {
        char *recv = new char[(size_t) 2*nb*nprocs];
        char *stage = recv + (size_t) nb*nprocs;
        MPI_Allgather(mine, nb, MPI_CHAR, recv, nb, MPI_CHAR, world);
        MPI_Allgather(mine+nb, GFMD_B200_IPC_HANDLE_BYTES, MPI_CHAR, stage, GFMD_B200_IPC_HANDLE_BYTES,
                      MPI_CHAR, world);
        ok = gfmd_b200_ipc_import(handle_, recv) == 0;
        /* a rank that failed to import would wait at barriers the others never reach: agree first */
        MPI_Allreduce(&ok, &allok, 1, MPI_INT, MPI_MIN, world);
        if (!allok)
          error->all(FLERR,"fix gfmd solver static/b200: mapping the peers' GPU buffers failed on some rank; "
                     "make all GPUs visible to every rank or set GFMD_B200_NO_PEER=1.");
        ok = gfmd_b200_ipc_import_stage(handle_, stage) == 0;
        MPI_Allreduce(&ok, &allok, 1, MPI_INT, MPI_MIN, world);
        if (!allok)
          error->all(FLERR,"fix gfmd solver static/b200: mapping the peers' row buffers failed on some rank; "
                     "set GFMD_B200_NO_PEER=1.");
        delete [] recv;
      }
    }
  }

  check(gfmd_b200_get_brick(handle_, &xlo_loc, &xhi_loc, &ylo_loc, &yhi_loc, &nxy_loc, &gammai_),
        "gfmd_b200_get_brick");
  nx_loc = xhi_loc-xlo_loc+1;
  ny_loc = yhi_loc-ylo_loc+1;

  /* the brick of the library must be the one LAMMPS' decomposition implies
     (src/main/gfmd_solver.cpp:95-99) or atoms and grid would not line up */
  int xlo_ref = (int) round(nx*(domain->sublo[0]-domain->boxlo[0])/domain->xprd);
  int xhi_ref = (int) round(nx*(domain->subhi[0]-domain->boxlo[0])/domain->xprd)-1;
  if (xlo_ref != xlo_loc || xhi_ref != xhi_loc) {
    char errstr[256];
    snprintf(errstr, sizeof(errstr), "fix gfmd solver static/b200: LAMMPS sub-domain maps to grid rows "
             "%i..%i but the slab decomposition assigns %i..%i; use a uniform decomposition with nx "
             "divisible by the number of ranks.", xlo_ref, xhi_ref, xlo_loc, xhi_loc);
    error->one(FLERR, errstr);
  }

  /* never register caller-owned memory unless asked to (see the constructor) */
  gfmd_b200_pin_host_buffers(handle_, pin_ ? 1 : 0);
}


void GFMDSolverB200::init()
{
}


void GFMDSolverB200::set_kernel(StiffnessKernel *kernel, bool normalize)
{
  kernel_ = kernel;
  normalize_ = normalize;

  if (screen && comm->me == 0)
    fprintf(screen, "USER-GFMD: Computing stiffness matrices...\n");

  /*
   * Stiffness matrices for the q columns this rank owns after the transpose:
   * all kx, kylo <= ky < kylo+nky of the half spectrum.  Streamed in chunks.
   */
  int kylo, nky;
  check(gfmd_b200_get_q_columns(handle_, &kylo, &nky), "gfmd_b200_get_q_columns");
  const int chunk = 32;
  double_complex **phi = NULL;
  for (int k0 = 0; k0 < nky; k0 += chunk) {
    int nk = MIN(chunk, nky-k0);
    memory->create(phi, nx*nk, ndof_sq, "GFMDSolverB200::phi");
    fill_phi_buffer(ndof, nx, 0, nx-1, ny, kylo+k0, kylo+k0+nk-1, kernel, phi, normalize, error);
    /* fill_phi_buffer has applied exactly the normalisation the caller asked for; the library
       must not add one (with normalize = false the reference uses the raw table too) */
    check(gfmd_b200_set_phi_columns(handle_, reinterpret_cast<double*>(phi[0]), kylo+k0, nk, 1),
          "gfmd_b200_set_phi_columns");
    memory->destroy(phi);
  }

  /*
   * Linear force contributions, with the reference's sanity check
   * (src/solvers/gfmd_solver_static.cpp:107-122)
   */
  double *linf = new double[nu_];
  kernel->get_force_at_gamma_point(linf);
  if (std::abs(std::accumulate(linf, linf+nu_, 0.0)) > 1e-6 && screen) {
    fprintf(screen, "USER-GFMD: Warning: Forces do not sum to zero at the "
            "surface. Sum is = %e.\n", std::accumulate(linf, linf+nu_, 0.0));
  }
  check(gfmd_b200_set_linf(handle_, linf), "gfmd_b200_set_linf");
  delete [] linf;

  if (screen && comm->me == 0)
    fprintf(screen, "USER-GFMD: ...done\n");
}


void GFMDSolverB200::pre_force(void *input_buffer_ptr, void *)
{
  if (!async_) return;
  double **input_buffer = static_cast<double**>(input_buffer_ptr);
  check(gfmd_b200_pre_force_async_host(handle_, input_buffer[0]), "gfmd_b200_pre_force_async_host");
}


double GFMDSolverB200::post_force(void *input_buffer_ptr, void *output_buffer_ptr, char *dump_prefix)
{
  double **input_buffer = static_cast<double**>(input_buffer_ptr);
  double **output_buffer = static_cast<double**>(output_buffer_ptr);

  double epot;
  check(gfmd_b200_post_force_host(handle_, input_buffer[0], output_buffer[0], &epot, u0),
        "gfmd_b200_post_force_host");

  /* the reference dumps between the forward transform and the contraction
     (src/solvers/gfmd_solver_static.cpp:181-182); the fields are the same afterwards */
  if (dump_prefix)
    dump(dump_prefix, input_buffer[0]);

  return epot;
}


/* ----------------------------------------------------------------------
 * q-space dump of a `dumpq_every` step: same files, same layout (rows = iy,
 * columns = ix) and same number format as GFMDSolverFFT::dump
 * (src/solvers/gfmd_solver_fft.cpp:209-287)
 * --------------------------------------------------------------------*/

void GFMDSolverB200::dump(char *dump_prefix, double *u)
{
  if (nprocs > 1)
    error->all(FLERR,"Can only dump from single processor run.");

  /* u~(q) and Phi(q).u~(q) in the q_buffer layout [ix*ny + iy][idof] */
  const size_t nq = (size_t) nx*ny;
  std::vector<double_complex> uq(nq*ndof), fq(nq*ndof);
  check(gfmd_b200_spectrum_host(handle_, u, reinterpret_cast<double*>(&uq[0]), reinterpret_cast<double*>(&fq[0])),
        "gfmd_b200_spectrum_host");

  /* one file per field; columns of a text row run over ix, rows over iy */
  enum { UR, UI, FR, FI, NPER };
  const char *per_dof[NPER] = { "u%i.real", "u%i.imag", "f%i.real", "f%i.imag" };
  const char *sums[3] = { "uP", "fP", "e" };
  std::vector<FILE*> out(NPER*ndof + 3);
  char stem[64], fn[1024];
  for (int idof = 0; idof < ndof; idof++) {
    for (int k = 0; k < NPER; k++) {
      snprintf(stem, sizeof(stem), per_dof[k], idof);
      snprintf(fn, sizeof(fn), "%s.q.%s.out", dump_prefix, stem);
      out[NPER*idof + k] = fopen(fn, "w");
    }
  }
  for (int k = 0; k < 3; k++) {
    snprintf(fn, sizeof(fn), "%s.q.%s.out", dump_prefix, sums[k]);
    out[NPER*ndof + k] = fopen(fn, "w");
  }
  for (size_t k = 0; k < out.size(); k++) {
    if (!out[k]) error->one(FLERR,"fix gfmd solver static/b200: cannot open q-space dump file.");
  }

  const char *fmt = " %20.10e ";
  for (int iy = 0; iy < ny; iy++) {
    for (int ix = 0; ix < nx; ix++) {
      const double_complex *uv = &uq[((size_t) ix*ny + iy)*ndof];
      const double_complex *fv = &fq[((size_t) ix*ny + iy)*ndof];
      double u2 = 0.0, f2 = 0.0, uf = 0.0;
      for (int idof = 0; idof < ndof; idof++) {
        const double ur = creal(uv[idof]), ui = cimag(uv[idof]);
        const double fr = creal(fv[idof]), fi = cimag(fv[idof]);
        fprintf(out[NPER*idof + UR], fmt, ur);
        fprintf(out[NPER*idof + UI], fmt, ui);
        fprintf(out[NPER*idof + FR], fmt, fr);
        fprintf(out[NPER*idof + FI], fmt, fi);
        u2 += ur*ur + ui*ui;
        f2 += fr*fr + fi*fi;
        uf += ur*fr + ui*fi;          /* Re(u conj(F)) */
      }
      fprintf(out[NPER*ndof + 0], fmt, u2);
      fprintf(out[NPER*ndof + 1], fmt, f2);
      fprintf(out[NPER*ndof + 2], fmt, uf);
    }
    for (size_t k = 0; k < out.size(); k++) fputc('\n', out[k]);
  }
  for (size_t k = 0; k < out.size(); k++) fclose(out[k]);
}


void GFMDSolverB200::prec_gradient(double *cavg, double **g, double **gP)
{
  /* first3_only = 1: for ndof > 3 the reference replaces only the first three components
     (src/main/gfmd_misc.h:113-115); a drop-in does the same */
  check(gfmd_b200_prec_gradient_host(handle_, cavg, g[0], gP[0], 1), "gfmd_b200_prec_gradient_host");
}


/* ----------------------------------------------------------------------
 * dump stiffness coefficients / Green's function for plotting in gnuplot.
 * The reference prints its stored table entry by entry in storage order
 * (n = ix*ny + iy) with a line break after every nx entries, and opens but
 * never fills the two trace files (src/solvers/gfmd_solver_fft.cpp:294-355,
 * :362-430); the table is recomputed here row block by row block because
 * this solver keeps no host copy of it.
 * --------------------------------------------------------------------*/

void GFMDSolverB200::dump_table(const char *stem, bool invert)
{
  if (nprocs > 1) {
    char errstr[256];
    snprintf(errstr, sizeof(errstr), "fix gfmd/static: Dump %s only works from a single processor run.",
             invert ? "Green's function" : "stiffness");
    error->all(FLERR, errstr);
  }
  if (!kernel_)
    error->all(FLERR,"fix gfmd solver static/b200: dump requested before set_kernel.");
  if (me != 0) return;

  /* <stem><i><j>.real.out / .imag.out for every matrix element, plus the two trace files the
     reference opens and only ever writes line breaks to */
  char fn[1024];
  std::vector<FILE*> re(ndof_sq), im(ndof_sq);
  for (int k = 0; k < ndof_sq; k++) {
    snprintf(fn, sizeof(fn), "%s%i%i.real.out", stem, k/ndof, k%ndof);
    re[k] = fopen(fn, "w");
    snprintf(fn, sizeof(fn), "%s%i%i.imag.out", stem, k/ndof, k%ndof);
    im[k] = fopen(fn, "w");
  }
  snprintf(fn, sizeof(fn), "%str.real.out", stem);
  FILE *trre = fopen(fn, "w");
  snprintf(fn, sizeof(fn), "%str.imag.out", stem);
  FILE *trim = fopen(fn, "w");

  /* storage order n = ix*ny + iy, a line break after every nx entries */
  double_complex **phi = NULL;
  std::vector<double_complex> G(ndof_sq);
  size_t n = 0;
  for (int ix = 0; ix < nx; ix++) {
    memory->create(phi, ny, ndof_sq, "GFMDSolverB200::phi");
    fill_phi_buffer(ndof, nx, ix, ix, ny, 0, ny-1, kernel_, phi, normalize_, error);
    for (int iy = 0; iy < ny; iy++) {
      memcpy(&G[0], phi[iy], ndof_sq*sizeof(double_complex));
      if (invert) GaussJordan(ndof, &G[0], error);
      for (int k = 0; k < ndof_sq; k++) {
        fprintf(re[k], " %e ", creal(G[k]));
        fprintf(im[k], " %e ", cimag(G[k]));
      }
      if (++n % nx == 0) {
        for (int k = 0; k < ndof_sq; k++) {
          fputc('\n', re[k]);
          fputc('\n', im[k]);
        }
        fputc('\n', trre);
        fputc('\n', trim);
      }
    }
    memory->destroy(phi);
  }

  for (int k = 0; k < ndof_sq; k++) {
    fclose(re[k]);
    fclose(im[k]);
  }
  fclose(trre);
  fclose(trim);
}


void GFMDSolverB200::dump_stiffness()
{
  dump_table("phi", false);
}


void GFMDSolverB200::dump_greens_function()
{
  dump_table("g", true);
}


double GFMDSolverB200::memory_usage()
{
  return handle_ ? gfmd_b200_memory_usage(handle_) : 0.0;
}
