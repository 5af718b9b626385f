// Generic (any grid size) kernels of the fix-gfmd elastic-force path.
//
//   k_gather      FixGFMD::pre_force list->grid     src/main/fix_gfmd.cpp:734-803
//   k_rows_fwd    y-direction real->half-spectrum   src/solvers/gfmd_solver_fft.cpp:96-147 (y part)
//   k_cols_fused  x-direction FFT, Phi(q).u(q), energy, gamma point, x-direction
//                 inverse FFT                       gfmd_solver_fft.cpp (x part) +
//                                                   src/solvers/gfmd_solver_static.cpp:160-236
//   k_finalize    epot = 0.5 (sum_q ... - 2 linf u0) gfmd_solver_static.cpp:192-196,236
//   k_rows_inv    half-spectrum->real along y       gfmd_solver_fft.cpp:150-195 (y part)
//   k_scatter     grid_to_list + f += f_i           fix_gfmd.cpp:952-1010, :896-902
//
// Data layout in HBM (P = number of slab ranks, 1 on a single GPU):
//   u, f    : [d][nx_loc][ny] real, y fastest -- the reference's u_xy / f_xy
//   stage   : [P][d][kyb][nx_loc] complex -- half spectrum (ky <= ny/2) TRANSPOSED
//             so that x is the fastest index; block r holds the ky-range owned
//             by rank r after the exchange (kyb = ceil(nyh / P)).  With P = 1
//             this is simply u~[d][ky][kx] and the column kernel works in place.
//   phi     : [nky_loc][d*d][nx] real -- Hermitian-packed Phi(kx, ky) planes:
//             c < d: Re Phi_cc; then for i < j: Re Phi_ij, Im Phi_ij.
#pragma once

#include "fft_engine.cuh"

namespace gfmd {

struct GridDesc {
  int nx, ny, nyh, d;
  int nx_loc, x0;        // this rank's x-slab [x0, x0 + nx_loc)
  int P, rank;
  int kyb;               // ky block length per rank
  int ky0, nky_loc;      // this rank's ky range [ky0, ky0 + nky_loc)
};

struct StepResults {     // device-resident, copied to the host on request
  double epot;           // this rank's potential energy
  double esum;           // sum_q w(q) Re(u~^H Phi u~) over this rank's columns
  double egamma;         // -2 sum_i linf_i Re u~_{3i+2}(0) (gamma rank only)
  double u0[24];         // Re u~(q=0) (gamma rank only, zero elsewhere)
  double fsum[3];        // sum of scattered forces over local atoms
  int natoms_gathered;   // fix_gfmd.cpp:796 natoms_cur
  int natoms_scattered;  // atoms that received a force
  int n_out_of_range;    // atoms whose grid index was invalid (skipped)
  int pad;
};

__device__ __forceinline__ size_t stage_index(const GridDesc &g, int ky, int dof, int ix)
{
  if (g.P == 1) return ((size_t) dof * g.kyb + ky) * g.nx_loc + ix;
  const int r = ky / g.kyb;
  const int kl = ky - r * g.kyb;
  return ((size_t) (r * g.d + dof) * g.kyb + kl) * g.nx_loc + ix;
}

// ------------------------------------------------------------------ gather ---

constexpr int kAtomTile = 256;   // atoms per tile = threads per block of gather / scatter

// Grid-stride over tiles of 256 atoms.  The AoS atom arrays ([nall][3]) are staged
// through shared memory with fully coalesced loads; thread t then owns atom t of the
// tile.  x, xeq: [nall][3]; gid: [nall][3] = (ix, iy, iu).
__global__ void __launch_bounds__(kAtomTile)
k_gather(const double *__restrict__ x, const double *__restrict__ xeq, int *__restrict__ gid,
         const int *__restrict__ mask, int groupbit, int nall, GridDesc g, double xprd, double yprd,
         int dxshift, int dyshift, double *__restrict__ u, StepResults *res)
{
  __shared__ double sx[3 * kAtomTile], sq[3 * kAtomTile];
  __shared__ int sg[3 * kAtomTile];
  __shared__ int scount[2];
  if (threadIdx.x < 2) scount[threadIdx.x] = 0;
  int stored = 0, bad = 0;
  const double xh = 0.5 * xprd, yh = 0.5 * yprd;
  const size_t nxy = (size_t) g.nx_loc * g.ny;
  for (long long t0 = (long long) blockIdx.x * kAtomTile; t0 < nall; t0 += (long long) gridDim.x * kAtomTile) {
    const int n = nall - t0 < kAtomTile ? (int) (nall - t0) : kAtomTile;
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * n; k += kAtomTile) {
      sx[k] = x[3 * t0 + k];
      sq[k] = xeq[3 * t0 + k];
      sg[k] = gid[3 * t0 + k];
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < n && (mask[t0 + t] & groupbit)) {
      int ix = sg[3 * t] - dxshift;
      int iy = sg[3 * t + 1] - dyshift;
      const int iu = sg[3 * t + 2];
      if (dxshift != 0 || dyshift != 0) {
        ix %= g.nx; if (ix < 0) ix += g.nx;
        iy %= g.ny; if (iy < 0) iy += g.ny;
        gid[3 * (t0 + t)] = ix;
        gid[3 * (t0 + t) + 1] = iy;
      }
      ix -= g.x0;
      if (ix >= 0 && ix < g.nx_loc && iy >= 0 && iy < g.ny) {
        if (iu < 0 || 3 * iu + 2 >= g.d) {
          bad++;
        } else {
          double ux = sx[3 * t] - sq[3 * t];
          double uy = sx[3 * t + 1] - sq[3 * t + 1];
          const double uz = sx[3 * t + 2] - sq[3 * t + 2];
          while (ux > xh) ux -= xprd;
          while (ux < -xh) ux += xprd;
          while (uy > yh) uy -= yprd;
          while (uy < -yh) uy += yprd;
          const size_t iloc = (size_t) ix * g.ny + iy;
          u[(size_t) (3 * iu) * nxy + iloc] = ux;
          u[(size_t) (3 * iu + 1) * nxy + iloc] = uy;
          u[(size_t) (3 * iu + 2) * nxy + iloc] = uz;
          stored++;
        }
      }
    }
  }
  // one atomic per block and counter
  __syncthreads();
  if (stored) atomicAdd(&scount[0], stored);
  if (bad) atomicAdd(&scount[1], bad);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (scount[0]) atomicAdd(&res->natoms_gathered, scount[0]);
    if (scount[1]) atomicAdd(&res->n_out_of_range, scount[1]);
  }
}

// ---------------------------------------------------------------- cell map ---

// cell -> atom map of the fused gather / scatter row kernels (AtomIO, kernels_rows_r16.cuh): the same
// index arithmetic as k_gather (lattice shift, wrap, brick test; fix_gfmd.cpp:734-760).  cmap must be
// preset to -1.  cnt[0] atoms stored, cnt[1] atoms that hit an occupied cell (e.g. a local atom and
// its own ghost image), cnt[2] atoms with an invalid sublattice index.
__global__ void __launch_bounds__(kAtomTile)
k_build_cellmap(int *__restrict__ gid, const int *__restrict__ mask, int groupbit, int nall, GridDesc g,
                int dxshift, int dyshift, int *__restrict__ cmap, int *cnt)
{
  __shared__ int sc[3];
  if (threadIdx.x < 3) sc[threadIdx.x] = 0;
  __syncthreads();
  int stored = 0, dup = 0, bad = 0;
  const size_t nxy = (size_t) g.nx_loc * g.ny;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += (long long) gridDim.x * blockDim.x) {
    if (!(mask[i] & groupbit)) continue;
    int ix = gid[3 * i] - dxshift;
    int iy = gid[3 * i + 1] - dyshift;
    const int iu = gid[3 * i + 2];
    if (dxshift != 0 || dyshift != 0) {
      ix %= g.nx; if (ix < 0) ix += g.nx;
      iy %= g.ny; if (iy < 0) iy += g.ny;
      gid[3 * i] = ix;
      gid[3 * i + 1] = iy;
    }
    ix -= g.x0;
    if (ix >= 0 && ix < g.nx_loc && iy >= 0 && iy < g.ny) {
      if (iu < 0 || 3 * iu + 2 >= g.d) {
        bad++;
      } else {
        const int old = atomicExch(&cmap[(size_t) iu * nxy + (size_t) ix * g.ny + iy], (int) i);
        stored++;
        if (old >= 0) dup++;
      }
    }
  }
  if (stored) atomicAdd(&sc[0], stored);
  if (dup) atomicAdd(&sc[1], dup);
  if (bad) atomicAdd(&sc[2], bad);
  __syncthreads();
  if (threadIdx.x < 3 && sc[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], sc[threadIdx.x]);
}

// force sum of the fused scatter: part[ntile][d] -> out[c] = sum over tiles and sublattices, fixed order
__global__ void k_sum_fsum_io(const double *__restrict__ part, int ntile, int d, double *out)
{
  __shared__ double sh[256];
  for (int c = 0; c < 3; ++c) {
    double a = 0.0;
    for (int k = threadIdx.x; k < ntile; k += blockDim.x)
      for (int iu = 0; iu < d / 3; ++iu) a += part[(size_t) k * d + 3 * iu + c];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = sh[0];
    __syncthreads();
  }
}

// ----------------------------------------------------------------- scatter ---

// f[nall][3] += grid force of the atom's cell; per-block partial sums of the forces on
// the first nlocal atoms go to fsum_part[3*blockIdx + c] (summed in fixed order later).
__global__ void __launch_bounds__(kAtomTile)
k_scatter(const double *__restrict__ fgrid, const int *__restrict__ gid, const int *__restrict__ mask,
          int groupbit, int nall, int nlocal, GridDesc g, double *__restrict__ f,
          double *__restrict__ fsum_part, StepResults *res)
{
  __shared__ double sf[3 * kAtomTile];
  __shared__ int sg[3 * kAtomTile];
  __shared__ double sh[3][kAtomTile / 32];
  __shared__ int scount;
  if (threadIdx.x == 0) scount = 0;
  double ax = 0.0, ay = 0.0, az = 0.0;
  int known = 0;
  const size_t nxy = (size_t) g.nx_loc * g.ny;
  for (long long t0 = (long long) blockIdx.x * kAtomTile; t0 < nall; t0 += (long long) gridDim.x * kAtomTile) {
    const int n = nall - t0 < kAtomTile ? (int) (nall - t0) : kAtomTile;
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * n; k += kAtomTile) {
      sf[k] = f[3 * t0 + k];
      sg[k] = gid[3 * t0 + k];
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < n && (mask[t0 + t] & groupbit)) {
      const int ix = sg[3 * t] - g.x0;
      const int iy = sg[3 * t + 1];
      const int iu = sg[3 * t + 2];
      if (ix >= 0 && ix < g.nx_loc && iy >= 0 && iy < g.ny && iu >= 0 && 3 * iu + 2 < g.d) {
        const size_t iloc = (size_t) ix * g.ny + iy;
        const double fx = fgrid[(size_t) (3 * iu) * nxy + iloc];
        const double fy = fgrid[(size_t) (3 * iu + 1) * nxy + iloc];
        const double fz = fgrid[(size_t) (3 * iu + 2) * nxy + iloc];
        sf[3 * t] += fx;
        sf[3 * t + 1] += fy;
        sf[3 * t + 2] += fz;
        known++;
        if (t0 + t < nlocal) { ax += fx; ay += fy; az += fz; }   // fsum_loc: local atoms only (:997-1001)
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * n; k += kAtomTile) f[3 * t0 + k] = sf[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ax += __shfl_down_sync(0xffffffffu, ax, o);
    ay += __shfl_down_sync(0xffffffffu, ay, o);
    az += __shfl_down_sync(0xffffffffu, az, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = ax; sh[1][w] = ay; sh[2][w] = az; }
  __syncthreads();
  if (known) atomicAdd(&scount, known);
  __syncthreads();
  if (threadIdx.x < 3) {
    double a = 0.0;
    for (int k = 0; k < kAtomTile / 32; ++k) a += sh[threadIdx.x][k];
    fsum_part[3 * blockIdx.x + threadIdx.x] = a;
  }
  if (threadIdx.x == 0 && scount) atomicAdd(&res->natoms_scattered, scount);
}

// fixed-order sum of per-block partials: part[nblk][ncomp] -> out[ncomp]
__global__ void k_sum_partials(const double *__restrict__ part, int nblk, int ncomp, double *out)
{
  __shared__ double sh[256];
  for (int c = 0; c < ncomp; ++c) {
    double a = 0.0;
    for (int k = threadIdx.x; k < nblk; k += blockDim.x) a += part[(size_t) k * ncomp + c];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = sh[0];
    __syncthreads();
  }
}

// ---------------------------------------------------------------- rows fwd ---

// One CTA transforms RB consecutive rows (same dof) along y and writes the half
// spectrum transposed into the staging layout.  EVEN: ny even, a real row is
// read as ny/2 complex numbers and un-mixed after a half-length transform.
// ODD: full-length complex transform of (row, 0).
template <bool EVEN>
__global__ void __launch_bounds__(512)
k_rows_fwd(const double *__restrict__ u, double2 *__restrict__ stage, GridDesc g, FftDesc fd,
           const double2 *__restrict__ tw_ny, int RB, int ld)
{
  extern __shared__ double2 smem[];
  const int nblk = (g.nx_loc + RB - 1) / RB;
  const int dof = blockIdx.x / nblk;
  const int ix0 = (blockIdx.x - dof * nblk) * RB;
  const int nr = g.nx_loc - ix0 < RB ? g.nx_loc - ix0 : RB;
  const int ny = g.ny;
  const double *src = u + ((size_t) dof * g.nx_loc + ix0) * ny;

  if (EVEN) {
    const int h = ny >> 1;
    for (int idx = threadIdx.x; idx < nr * h; idx += blockDim.x) {
      const int r = idx / h, j = idx - r * h;
      smem[r * ld + j] = reinterpret_cast<const double2 *>(src + (size_t) r * ny)[j];
    }
  } else {
    for (int idx = threadIdx.x; idx < nr * ny; idx += blockDim.x) {
      const int r = idx / ny, j = idx - r * ny;
      smem[r * ld + j] = make_double2(src[(size_t) r * ny + j], 0.0);
    }
  }
  __syncthreads();
  fft_batch<-1>(smem, ld, nr, fd);

  const int nyh = g.nyh;
  for (int idx = threadIdx.x; idx < nr * nyh; idx += blockDim.x) {
    const int ky = idx / nr, r = idx - ky * nr;
    double2 X;
    if (EVEN) {
      const int h = ny >> 1;
      const int ka = ky == h ? 0 : ky;
      const int kb = ky == 0 ? 0 : h - ky;
      const double2 zk = smem[r * ld + ka];
      const double2 zc = cconj(smem[r * ld + kb]);
      const double2 A = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
      const double2 B = make_double2(0.5 * (zk.x - zc.x), 0.5 * (zk.y - zc.y));
      const double2 t = cmul(__ldg(tw_ny + ky), B);
      X = make_double2(A.x + t.y, A.y - t.x);      // A - i w B
    } else {
      X = smem[r * ld + ky];
    }
    stage[stage_index(g, ky, dof, ix0 + r)] = X;
  }
}

// ---------------------------------------------------------------- rows inv ---

template <bool EVEN>
__global__ void __launch_bounds__(512)
k_rows_inv(const double2 *__restrict__ stage, double *__restrict__ f, GridDesc g, FftDesc fd,
           const double2 *__restrict__ tw_ny, int RB, int ld)
{
  extern __shared__ double2 smem[];
  const int nblk = (g.nx_loc + RB - 1) / RB;
  const int dof = blockIdx.x / nblk;
  const int ix0 = (blockIdx.x - dof * nblk) * RB;
  const int nr = g.nx_loc - ix0 < RB ? g.nx_loc - ix0 : RB;
  const int ny = g.ny, nyh = g.nyh;
  double *dst = f + ((size_t) dof * g.nx_loc + ix0) * ny;

  for (int idx = threadIdx.x; idx < nr * nyh; idx += blockDim.x) {
    const int ky = idx / nr, r = idx - ky * nr;
    smem[r * ld + ky] = stage[stage_index(g, ky, dof, ix0 + r)];
  }
  __syncthreads();

  if (EVEN) {
    const int h = ny >> 1;
    const int np = h / 2 + 1;
    // Z'[k] = (Y[k] + conj Y[h-k]) + i e^{+2 pi i k/ny} (Y[k] - conj Y[h-k])
    for (int idx = threadIdx.x; idx < nr * np; idx += blockDim.x) {
      const int r = idx / np, k = idx - r * np;
      const int k2 = h - k;
      const double2 yk = smem[r * ld + k];
      const double2 y2 = smem[r * ld + k2];
      {
        const double2 c2 = cconj(y2);
        const double2 S = cadd(yk, c2), D = csub(yk, c2);
        const double2 t = cmulc(D, __ldg(tw_ny + k));           // e_k D
        smem[r * ld + k] = make_double2(S.x - t.y, S.y + t.x);  // S + i t
      }
      if (k2 != k && k2 < h) {
        const double2 ck = cconj(yk);
        const double2 S = cadd(y2, ck), D = csub(y2, ck);
        const double2 t = cmulc(D, __ldg(tw_ny + k2));
        smem[r * ld + k2] = make_double2(S.x - t.y, S.y + t.x);
      }
    }
    __syncthreads();
    fft_batch<+1>(smem, ld, nr, fd);
    for (int idx = threadIdx.x; idx < nr * h; idx += blockDim.x) {
      const int r = idx / h, j = idx - r * h;
      reinterpret_cast<double2 *>(dst + (size_t) r * ny)[j] = smem[r * ld + j];
    }
  } else {
    const int nfill = ny - nyh;
    for (int idx = threadIdx.x; idx < nr * nfill; idx += blockDim.x) {
      const int r = idx / nfill, k = nyh + (idx - r * nfill);
      smem[r * ld + k] = cconj(smem[r * ld + ny - k]);
    }
    __syncthreads();
    fft_batch<+1>(smem, ld, nr, fd);
    for (int idx = threadIdx.x; idx < nr * ny; idx += blockDim.x) {
      const int r = idx / ny, j = idx - r * ny;
      dst[(size_t) r * ny + j] = smem[r * ld + j].x;
    }
  }
}

// ------------------------------------------------------------ cols (fused) ---

// Hermitian-packed Phi times vector for one q.  ph(c) returns plane c.
template <int D, typename PhiLoad>
__device__ __forceinline__ void phi_matvec(const double2 *uv, double2 *F, PhiLoad ph)
{
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const double a = ph(i);
    F[i] = make_double2(a * uv[i].x, a * uv[i].y);
  }
  int c = D;
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = i + 1; j < D; ++j) {
      const double2 p = make_double2(ph(c), ph(c + 1));
      c += 2;
      // F_i += Phi_ij u_j ; F_j += conj(Phi_ij) u_i
      F[i].x = fma(p.x, uv[j].x, fma(-p.y, uv[j].y, F[i].x));
      F[i].y = fma(p.x, uv[j].y, fma(p.y, uv[j].x, F[i].y));
      F[j].x = fma(p.x, uv[i].x, fma(p.y, uv[i].y, F[j].x));
      F[j].y = fma(p.x, uv[i].y, fma(-p.y, uv[i].x, F[j].y));
    }
  }
}

// One CTA per ky column of this rank: all d dofs of the column sit in shared
// memory (d * ld complex).  Forward x-transform, contraction with Phi(q),
// energy partial, gamma-point terms, backward x-transform, store.
// DT = compile-time ndof (3, 6, 9, 12) or 0 for the run-time generic version.
template <int DT>
__global__ void __launch_bounds__(512)
k_cols_fused(const double2 *__restrict__ stage_in, double2 *__restrict__ stage_out, GridDesc g,
             FftDesc fd, const double *__restrict__ phi, const double *__restrict__ linf,
             double *__restrict__ epart, StepResults *res, int ld)
{
  extern __shared__ double2 smem[];
  const int d = DT > 0 ? DT : g.d;
  const int kl = blockIdx.x;              // local ky index
  const int ky = g.ky0 + kl;
  const int nx = g.nx;

  // load: piece p of the column comes from source rank p's block
  for (int idx = threadIdx.x; idx < d * nx; idx += blockDim.x) {
    const int dof = idx / nx, ix = idx - dof * nx;
    const int p = ix / g.nx_loc, il = ix - p * g.nx_loc;
    smem[dof * ld + ix] = stage_in[((size_t) (p * d + dof) * g.kyb + kl) * g.nx_loc + il];
  }
  __syncthreads();
  fft_batch<-1>(smem, ld, d, fd);

  const double wgt = (ky == 0 || (2 * ky == g.ny)) ? 1.0 : 2.0;
  const double *ph = phi + (size_t) kl * d * d * nx;
  double e = 0.0;
  for (int kx = threadIdx.x; kx < nx; kx += blockDim.x) {
    if (DT > 0) {
      double2 uv[DT > 0 ? DT : 1], F[DT > 0 ? DT : 1];
#pragma unroll
      for (int i = 0; i < DT; ++i) uv[i] = smem[i * ld + kx];
      phi_matvec<DT>(uv, F, [&](int c) { return __ldg(ph + (size_t) c * nx + kx); });
      double eq = 0.0;
#pragma unroll
      for (int i = 0; i < DT; ++i) {
        eq = fma(F[i].x, uv[i].x, fma(F[i].y, uv[i].y, eq));
        F[i] = make_double2(-F[i].x, -F[i].y);
      }
      e = fma(wgt, eq, e);
      if (ky == 0 && kx == 0) {
        double eg = 0.0;
#pragma unroll
        for (int i = 0; i < DT; ++i) res->u0[i] = uv[i].x;
#pragma unroll
        for (int a = 0; a < DT / 3; ++a) {
          eg -= 2.0 * linf[a] * uv[3 * a + 2].x;
          F[3 * a + 2].x += linf[a];
        }
        res->egamma = eg;
      }
#pragma unroll
      for (int i = 0; i < DT; ++i) smem[i * ld + kx] = F[i];
    } else {
      double2 uv[24], F[24];
      for (int i = 0; i < d; ++i) uv[i] = smem[i * ld + kx];
      for (int i = 0; i < d; ++i) {
        const double a = __ldg(ph + (size_t) i * nx + kx);
        F[i] = make_double2(a * uv[i].x, a * uv[i].y);
      }
      int c = d;
      for (int i = 0; i < d; ++i)
        for (int j = i + 1; j < d; ++j) {
          const double2 p = make_double2(__ldg(ph + (size_t) c * nx + kx),
                                         __ldg(ph + (size_t) (c + 1) * nx + kx));
          c += 2;
          F[i].x = fma(p.x, uv[j].x, fma(-p.y, uv[j].y, F[i].x));
          F[i].y = fma(p.x, uv[j].y, fma(p.y, uv[j].x, F[i].y));
          F[j].x = fma(p.x, uv[i].x, fma(p.y, uv[i].y, F[j].x));
          F[j].y = fma(p.x, uv[i].y, fma(-p.y, uv[i].x, F[j].y));
        }
      double eq = 0.0;
      for (int i = 0; i < d; ++i) {
        eq = fma(F[i].x, uv[i].x, fma(F[i].y, uv[i].y, eq));
        F[i] = make_double2(-F[i].x, -F[i].y);
      }
      e = fma(wgt, eq, e);
      if (ky == 0 && kx == 0) {
        double eg = 0.0;
        for (int i = 0; i < d; ++i) res->u0[i] = uv[i].x;
        for (int a = 0; a < d / 3; ++a) {
          eg -= 2.0 * linf[a] * uv[3 * a + 2].x;
          F[3 * a + 2].x += linf[a];
        }
        res->egamma = eg;
      }
      for (int i = 0; i < d; ++i) smem[i * ld + kx] = F[i];
    }
  }

  // block energy partial (fixed order -> deterministic)
  __shared__ double she[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0) she[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int k = 0; k < nw; ++k) a += she[k];
    epart[kl] = a;
  }

  fft_batch<+1>(smem, ld, d, fd);

  for (int idx = threadIdx.x; idx < d * nx; idx += blockDim.x) {
    const int dof = idx / nx, ix = idx - dof * nx;
    const int p = ix / g.nx_loc, il = ix - p * g.nx_loc;
    stage_out[((size_t) (p * d + dof) * g.kyb + kl) * g.nx_loc + il] = smem[dof * ld + ix];
  }
}

// epot = 0.5 (esum + egamma), esum from the per-column partials in fixed order.
__global__ void k_finalize(const double *__restrict__ epart, int ncols, StepResults *res)
{
  __shared__ double sh[256];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;   // independent chains: loads overlap
  int k = threadIdx.x;
  for (; k + 3 * (int) blockDim.x < ncols; k += 4 * blockDim.x) {
    a0 += epart[k];
    a1 += epart[k + blockDim.x];
    a2 += epart[k + 2 * blockDim.x];
    a3 += epart[k + 3 * blockDim.x];
  }
  for (; k < ncols; k += blockDim.x) a0 += epart[k];
  const double a = (a0 + a1) + (a2 + a3);
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    res->esum = sh[0];
    res->epot = 0.5 * (sh[0] + res->egamma);
  }
}

}  // namespace gfmd
