/* gfmd_b200.h -- C ABI of libgfmd_b200.so: the per-step elastic-force path of the
 * LAMMPS `fix gfmd` (user-gfmd) on NVIDIA B200 (sm_100a), FP64.
 *
 * Boundary.  The entry points are what a `GFMDSolver` subclass
 * (reference: src/main/gfmd_solver.h:34-123) needs in order to replace
 * `GFMDSolverStatic` (src/solvers/gfmd_solver_static.{h,cpp}) -- see
 * INTEGRATION.md for the ~150-line `GFMDSolverB200` host shim and the one-line
 * registration in gfmd_solver_factory (src/main/gfmd_solver.cpp:232-242).
 * Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions.  Every function returns 0 on success and a non-zero GFMD_B200_E*
 * code otherwise; gfmd_b200_last_error() then holds a message (the host shim
 * forwards it to LAMMPS error->one(FLERR, ...), the reference's only error path,
 * src/main/gfmd_solver.h).  The caller owns all host and atom arrays; the
 * library owns its device buffers, plans, stream and NCCL communicator.  A
 * handle is not re-entrant: one caller thread per handle (LAMMPS: one per MPI
 * rank).  There is NO CPU fallback: without a usable GPU gfmd_b200_create fails.
 *
 * Layouts (all the reference's own):
 *   u, f   [ndof][nx_loc*ny] double, index [idof][ix*ny + iy]   (u_xy / f_xy,
 *          src/main/fix_gfmd.cpp:613-616, :763-780)
 *   phi    [..][ndof*ndof] complex128 row-major (MEL, src/mathutils/mat.h:37),
 *          q ordering of fill_phi_buffer (src/main/gfmd_misc.cpp:32-101)
 *   x, xeq [nall][3] double;  gid [nall][3] int (ix, iy, iu);  mask [nall] int
 *          (atom_style gfmd, src/main/atom_vec_gfmd.cpp:55-66)
 *
 * Environment variables, read when a handle is created (tuning and debugging only; none of
 * them changes results beyond rounding):
 *   GFMD_B200_NO_FAST=1          generic (any-size) kernels even where specialised ones exist
 *   GFMD_B200_ROWS_VARIANT=<id>  a specific row-kernel variant (csrc/kernels_fast.cuh)
 *   GFMD_B200_COLS_SPLIT=<n>     three-phase column stage with at most n dofs per CTA even where a
 *                                column set fits one CTA (csrc/kernel_cols_split.cuh; for tests)
 *   GFMD_B200_CHUNKS=<n>         column chunks of the multi-GPU pipelines (default 4; 8 without transposes)
 *   GFMD_B200_PEER_STORE=1       read by gfmd_b200_ipc_import: the column stage stores its results
 *                                straight into the peers' return buffers (opt-in)
 *   GFMD_B200_PEER_DIRECT=0/1    with gfmd_b200_ipc_import_stage: no transposes at all, the column stage loads
 *                                its input from the peers and stores its results into them, overlapped chunk by
 *                                chunk (default for >= 4 ranks with nx >= 8192; 0 = copy-engine pushes)
 *   GFMD_B200_XCHG_SMS=<a>[,<b>] SMs of the pulling / pushing top-digit passes of that mode (default 40,32)
 *   GFMD_B200_SYNC=nccl          one-element NCCL all-reduces as cross-rank barriers instead of flag words
 *   GFMD_B200_TIMELINE=1         with profiling on: one per-chunk timeline of that mode on stderr
 *   GFMD_B200_COLS_PIPE=0        the plain instead of the software-pipelined fused column kernel
 *   GFMD_B200_ROWS_PREFETCH=<n>  L2 prefetch distance (CTAs) of the radix-16 row kernels (0 = off)
 *   GFMD_B200_HOST_PIPE=0        no per-dof upload / download pipeline on the host path
 *   GFMD_B200_NCCL_LIB=<path>    the NCCL build to dlopen before libnccl.so.2
 */
#ifndef GFMD_B200_H
#define GFMD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfmd_b200 gfmd_b200_t;

enum {
  GFMD_B200_OK = 0,
  GFMD_B200_EINVAL = 1,      /* bad argument                                   */
  GFMD_B200_ECUDA = 2,       /* CUDA runtime error                             */
  GFMD_B200_ENOGPU = 3,      /* no usable CUDA device                          */
  GFMD_B200_EUNSUPPORTED = 4,/* grid size / ndof outside what the kernels do   */
  GFMD_B200_ENCCL = 5,       /* NCCL error or libnccl not loadable             */
  GFMD_B200_ESTATE = 6,      /* call order (e.g. step before set_phi)          */
  GFMD_B200_EPHI = 7         /* Phi table not Hermitian / not conj-symmetric   */
};

#define GFMD_B200_MAX_NDOF 24            /* MAX_NDOF, src/main/gfmd_solver.h:30 */
#define GFMD_B200_UNIQUE_ID_BYTES 128

/* ---- life cycle --------------------------------------------------------- */

/* Single GPU, whole nx x ny grid.  Replaces GFMDSolverStatic ctor +
 * set_grid_size (gfmd_solver_static.cpp:47-87, gfmd_solver_fft.cpp:66-93). */
int gfmd_b200_create(gfmd_b200_t **h, int nx, int ny, int ndof, int device);

/* One handle per rank/GPU of a slab decomposition along x (the reference's
 * brick decomposition with procgrid = nranks x 1 x 1, gfmd_solver.cpp:95-101).
 * nx must be divisible by nranks.  Follow with gfmd_b200_comm_init. */
int gfmd_b200_create_slab(gfmd_b200_t **h, int nx, int ny, int ndof, int device,
                          int rank, int nranks);

/* NCCL bootstrap: rank 0 obtains an id, the HOST distributes the 128 bytes
 * (MPI_Bcast in LAMMPS, torch.distributed in bench.py), every rank calls
 * comm_init.  libnccl.so.2 is dlopen'ed on first use. */
int gfmd_b200_get_unique_id(char id[GFMD_B200_UNIQUE_ID_BYTES]);
int gfmd_b200_comm_init(gfmd_b200_t *h, const char id[GFMD_B200_UNIQUE_ID_BYTES]);

/* Optional, after comm_init: replace the NCCL send/recv transposes by direct pushes into
 * the peers' receive buffers over NVLink (CUDA IPC + copy engines; NCCL then only provides
 * the barrier).  Every rank exports two 64-byte memory handles, the HOST gathers the
 * 2*64*nranks bytes in rank order (MPI_Allgather / torch.distributed.all_gather) and every
 * rank imports them.  All GPUs must be visible to every process (peer access). */
#define GFMD_B200_IPC_HANDLE_BYTES 64
int gfmd_b200_ipc_export(gfmd_b200_t *h, char *handles /* [2][64] */);
int gfmd_b200_ipc_import(gfmd_b200_t *h, const char *all_handles /* [nranks][2][64] */);

/* Optional third mapping, after gfmd_b200_ipc_import: every rank also exposes the buffer its row
 * kernels write (one handle), gathered and imported the same way.  With it and
 * GFMD_B200_PEER_DIRECT=1 the transposes disappear altogether: after a barrier the column stage
 * loads its pieces straight from the ranks that produced them and stores its results straight
 * into their return buffers, NVLink traffic issued by the kernels themselves (
 * replaces, like the pushes, what MPI does inside the reference's FFT3d remap,
 * src/solvers/gfmd_solver_fft.cpp:72-80). */
int gfmd_b200_ipc_export_stage(gfmd_b200_t *h, char *handle /* [64] */);
int gfmd_b200_ipc_import_stage(gfmd_b200_t *h, const char *all_handles /* [nranks][64] */);

void gfmd_b200_destroy(gfmd_b200_t *h);

/* Last error message of this handle (or of the failed create when h == NULL). */
const char *gfmd_b200_last_error(const gfmd_b200_t *h);

/* ---- grid partition (GFMDSolver::get_xlo_loc ... get_nxy_loc, gfmd_solver.h:72-86)
 * real-space brick [xlo,xhi] x [ylo,yhi] owned by this handle, gammai as in
 * gfmd_solver.cpp:132-135 (-1 if q=0 is elsewhere). */
int gfmd_b200_get_brick(const gfmd_b200_t *h, int *xlo, int *xhi, int *ylo, int *yhi,
                        int *nxy_loc, int *gammai);

/* q-space ownership after the transpose: all kx in [0,nx), ky in
 * [*kylo, *kylo + *nky) of the half spectrum 0 <= ky <= ny/2.  This is the
 * range set_kernel must pass to fill_phi_buffer for this rank. */
int gfmd_b200_get_q_columns(const gfmd_b200_t *h, int *kylo, int *nky);

/* ---- stiffness table (GFMDSolverStatic::set_kernel, gfmd_solver_static.cpp:90-136) */

/* Full table in the reference layout: phi[(ix*ny+iy)*ndof*ndof + i*ndof+j]
 * (complex128 as re,im pairs) for the WHOLE grid, as fill_phi_buffer(ndof, nx,
 * 0,nx-1, ny, 0,ny-1, ...) produces it.  already_normalised != 0 if the 1/(nx*ny)
 * factor is in (normalize = true, gfmd_misc.cpp:88-99).  linf: [ndof/3] from
 * kernel->get_force_at_gamma_point (may be NULL = zeros).  The half spectrum is
 * formed as (Phi(q) + conj Phi(-q))/2 and Hermitian-packed; deviations above
 * 1e-6 * max|Phi| (the reference's own Hermiticity tolerance, gfmd_misc.cpp:56)
 * are an error. */
int gfmd_b200_set_phi(gfmd_b200_t *h, const double *phi_ri, int already_normalised,
                      const double *linf);

/* Streaming variant for large grids / slab ranks: phi for kx in [0,nx) and ky in
 * [ky_first, ky_first+nky) within this handle's q columns, laid out as
 * fill_phi_buffer(ndof, nx, 0,nx-1, ny, ky_first,ky_first+nky-1, ...) does:
 * index ((ix*nky + j)*ndof*ndof + i*ndof + jdof).  Relies on
 * Phi(-q) = conj Phi(q).  May be called repeatedly to cover the range. */
int gfmd_b200_set_phi_columns(gfmd_b200_t *h, const double *phi_ri, int ky_first, int nky,
                              int already_normalised);
int gfmd_b200_set_linf(gfmd_b200_t *h, const double *linf);

/* Device-side table builder: runs the transfer-matrix recursion of
 * greens_function_transfer_matrix_stiffness (src/main/surface_stiffness.cpp:811-873, iterate_Gnn
 * :493-548) on the GPU.  uuv holds, for kx in [0,nx) and ky in [ky_first, ky_first+nky), the three
 * matrices StiffnessKernel::get_dynamical_matrices returns (U0, U, V; each ndof*ndof complex128,
 * row-major): index (((ix*nky + j)*3 + m)*ndof*ndof + i*ndof + jdof).  height as in the kernel
 * arguments (`height N`): N-1 iterations, 0 means Phi = U0; negative: iterate to the reference's
 * 1e-8 convergence, at most 100000 times -- a wavevector that does not converge fails the call with
 * GFMD_B200_EPHI, where the reference aborts ("Out of iterations ...").  normalise != 0 applies the 1/(nx*ny) of fill_phi_buffer.  ndof 3, 6, 9, 12. */
int gfmd_b200_build_phi_columns(gfmd_b200_t *h, const double *uuv_ri, int ky_first, int nky, int height,
                                int normalise);
/* Same, with the (U0, U, V) blocks already in DEVICE memory (large grids: the host never holds them). */
int gfmd_b200_build_phi_columns_device(gfmd_b200_t *h, const double *d_uuv_ri, int ky_first, int nky, int height,
                                       int normalise);

/* Deviations found by the last gfmd_b200_set_phi: max |Phi - Phi^H| and
 * max |Phi(q) - conj Phi(-q)|, both relative to max |Phi|. */
int gfmd_b200_phi_deviation(const gfmd_b200_t *h, double *herm_dev, double *conj_dev);

/* ---- the solver boundary: GFMDSolver::post_force (gfmd_solver.h:51;
 *      gfmd_solver_static.cpp:145-249) ------------------------------------ */

/* CPU contract: u and f are HOST arrays [ndof][nxy_loc] (the fix's u_xy/f_xy
 * contiguous blocks, i.e. u_xy[0]).  Copies u to the GPU, runs the step, copies
 * f back, returns this rank's epot and the q=0 displacement u0[ndof] (summed
 * over ranks like the reference's MPI_Allreduce, gfmd_solver_static.cpp:176).
 * See gfmd_b200_pin_host_buffers for faster copies from long-lived buffers. */
int gfmd_b200_post_force_host(gfmd_b200_t *h, const double *u, double *f, double *epot,
                              double *u0);

/* GFMDSolver::pre_force hook (gfmd_solver.h:44-50, called at fix_gfmd.cpp:853):
 * start the upload and the whole step asynchronously so that it overlaps
 * LAMMPS' pair computation.  The next post_force_host with the same u then only
 * waits and downloads. */
int gfmd_b200_pre_force_async_host(gfmd_b200_t *h, const double *u);

/* GPU contract (the reference's legacy `static/cuda` contract,
 * src/cuda/gfmd_solver_cuda.cpp:394-416): raw DEVICE pointers, plane per dof,
 * pitch nxy_loc, d_u != d_f.  Asynchronous on the handle's stream; fetch
 * epot/u0 with gfmd_b200_get_results.  NULL selects the library-owned grids. */
int gfmd_b200_post_force_device(gfmd_b200_t *h, const double *d_u, double *d_f);

/* ---- off-path services of the reference solver (single-rank handles) ----- */

/* What GFMDSolverFFT::dump writes on `dumpq_every` steps (gfmd_solver_fft.cpp:209-287,
 * called from gfmd_solver_static.cpp:181-182): uq = u~(q), the unnormalised forward DFT of
 * the HOST field u [ndof][nx*ny], and fq = Phi(q).u~(q) with the (normalised) table of this
 * handle, no sign flip.  Both come back in the reference's q_buffer layout,
 * [nx*ny][ndof] complex128 interleaved, idq = ix*ny + iy, full spectrum.  fq may be NULL.
 * Column sets that fit one CTA (nx * ndof * 16 B <= 227 KB) take one fused kernel, larger ones the
 * three-phase column stage (transform, per-q kernel, transform); one column must fit: nx <= 8192. */
int gfmd_b200_spectrum_host(gfmd_b200_t *h, const double *u, double *uq_ri, double *fq_ri);

/* GFMDSolverStatic::prec_gradient (gfmd_solver_static.cpp:253-271) with
 * precondition_gradient<DEF_G> (gfmd_misc.h:39-133):
 *   gP = IDFT[ (Phi(q) + Cavg)^-1 DFT[g] ],  both transforms unnormalised, Phi the
 * normalised table.  cavg: real [ndof][ndof] row-major; g, gP: HOST [ndof][nx*ny].
 * ndof 3, 6, 9, 12.  first3_only != 0 reproduces the reference for ndof > 3, whose general
 * branch copies only components 0..2 of the result back (`idim < 3`, gfmd_misc.h:113-115)
 * and leaves the others un-preconditioned -- what the drop-in host glue passes;
 * first3_only == 0 replaces all components. */
int gfmd_b200_prec_gradient_host(gfmd_b200_t *h, const double *cavg, const double *g, double *gP,
                                 int first3_only);

/* ---- device-resident fix-side stages ------------------------------------ */

/* FixGFMD::pre_force list->grid (fix_gfmd.cpp:734-803): u = x - xeq with x/y
 * minimum-image wrap, optional lattice shift of gid (written back, :748-759).
 * All pointers are DEVICE pointers.  d_u NULL = library-owned grid. */
int gfmd_b200_gather(gfmd_b200_t *h, const double *d_x, const double *d_xeq, int *d_gid,
                     const int *d_mask, int groupbit, int nall, double xprd, double yprd,
                     int dxshift, int dyshift, double *d_u);

/* FixGFMD::grid_to_list + f += f_i (fix_gfmd.cpp:952-1010, :896-902).  fsum is
 * accumulated over the first nlocal atoms only (:997-1001). */
int gfmd_b200_scatter(gfmd_b200_t *h, const double *d_fgrid, const int *d_gid,
                      const int *d_mask, int groupbit, int nall, int nlocal, double *d_f);

/* Optional: cell -> atom map for the FUSED form of gfmd_b200_full_step (radix-16 row kernels, ny =
 * 4096 / 8192 / 16384): the forward row transform reads x and xeq itself and the inverse one adds the
 * forces to the atoms itself, so the grids u_xy / f_xy never exist in memory (one HBM pass less on
 * either side).  Same index arithmetic as gfmd_b200_gather incl. the lattice shift (fix_gfmd.cpp:
 * 734-760).  *usable = 1 if every cell of this handle's brick holds exactly one atom of the group --
 * only then are gather / scatter through the map the reference's loops exactly (a local atom AND its
 * ghost image in one cell, or an empty cell, keep the separate kernels) -- and the handle's row
 * kernels have the fused form.  Call again whenever gid / mask / the atom order change (LAMMPS:
 * reneighbouring steps); full_step uses the map only when called with the same arrays and counts. */
int gfmd_b200_build_cell_map(gfmd_b200_t *h, int *d_gid, const int *d_mask, int groupbit, int nall,
                             int nlocal, int dxshift, int dyshift, int *usable);
int gfmd_b200_drop_cell_map(gfmd_b200_t *h);

/* gather + post_force + scatter on the library-owned grids, one call. */
int gfmd_b200_full_step(gfmd_b200_t *h, const double *d_x, const double *d_xeq, int *d_gid,
                        const int *d_mask, int groupbit, int nall, int nlocal, double xprd,
                        double yprd, double *d_f);

/* Waits for the stream and returns the last step's scalars.  Any pointer may be
 * NULL.  counters = {natoms_gathered, natoms_scattered, n_out_of_range}. */
int gfmd_b200_get_results(gfmd_b200_t *h, double *epot, double *u0, double fsum[3],
                          int counters[3]);

/* ---- plumbing ----------------------------------------------------------- */

double *gfmd_b200_device_u(gfmd_b200_t *h);      /* library-owned [ndof][nxy_loc] */
double *gfmd_b200_device_f(gfmd_b200_t *h);
void *gfmd_b200_stream(gfmd_b200_t *h);          /* cudaStream_t */
int gfmd_b200_set_stream(gfmd_b200_t *h, void *cuda_stream);
int gfmd_b200_synchronize(gfmd_b200_t *h);
/* Page-lock (cudaHostRegister) the caller's u/f host arrays the first time
 * post_force_host sees them; they are released by on=0 or destroy.  Only for
 * long-lived buffers such as the fix's u_xy/f_xy.  Default off. */
int gfmd_b200_pin_host_buffers(gfmd_b200_t *h, int on);
/* Host pipeline of gfmd_b200_pre_force_async_host / _post_force_host (default on; single rank
 * with the specialised row kernels, i.e. ny a power of two >= 2048; otherwise ignored): u is
 * uploaded and f downloaded dof by dof on two copy streams, so that the row transforms of one
 * dof overlap the PCIe transfer of the next.  Same kernels, bit-identical results.
 * on = 1 / 0 sets it, on < 0 only queries; returns whether it takes effect for this handle
 * (0 / 1) -- NOT an error code -- or GFMD_B200_EINVAL (>1) for a null handle.  Environment
 * GFMD_B200_HOST_PIPE=0 disables it at creation. */
int gfmd_b200_host_pipeline(gfmd_b200_t *h, int on);

/* replay the solver step through a captured CUDA graph (single GPU) */
int gfmd_b200_use_graph(gfmd_b200_t *h, int on);

/* kernels launched by this handle since creation */
long long gfmd_b200_launch_count(const gfmd_b200_t *h);

/* Per-stage device times (CUDA events on the handle's stream), accumulated while
 * profiling is on.  stage ids: 0 gather, 1 rows_fwd, 2 exchange_fwd, 3 cols_fused,
 * 4 exchange_inv, 5 rows_inv, 6 scatter, 7 cols_top_fwd, 8 cols_top_inv (the top-digit passes
 * of long columns, nx > 4096, when the column stage runs unchunked; otherwise inside stage 3).
 * ms[9], counts[9]. */
#define GFMD_B200_NSTAGES 9
int gfmd_b200_profile(gfmd_b200_t *h, int on);
int gfmd_b200_get_stage_times(gfmd_b200_t *h, double ms[GFMD_B200_NSTAGES],
                              long long counts[GFMD_B200_NSTAGES]);
/* one line naming the kernel variants this handle selected */
const char *gfmd_b200_describe(gfmd_b200_t *h);

/* bytes of device memory held by the handle (GFMDSolver::memory_usage) */
double gfmd_b200_memory_usage(const gfmd_b200_t *h);

const char *gfmd_b200_version(void);

#ifdef __cplusplus
}
#endif

#endif
