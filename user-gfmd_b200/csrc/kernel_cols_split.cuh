// Column stage for column sets that do not fit one CTA's shared memory
// (ndof * nx * 16 B > 227 KB and no specialised kernel, e.g. 4096 x 4096 with two atoms
// per cell): the fused kernel k_cols_fused is split into its three phases, each streaming
// the staging buffer once:
//
//   k_cols_split_fft<-1>  x-direction forward transform of `db` dofs of one ky column per CTA
//                         (GFMDSolverFFT::fft_forward, x part; src/solvers/gfmd_solver_fft.cpp:96-147)
//   k_cols_contract<D>    per q: F = -Phi(q).u~(q), energy partial, u0, gamma-point terms
//                         (GFMDSolverStatic::post_force, src/solvers/gfmd_solver_static.cpp:160-236)
//   k_cols_split_fft<+1>  x-direction backward transform (gfmd_solver_fft.cpp:150-195, x part)
//
// Same arithmetic, same order of operations per q and per transform as k_cols_fused; only the
// energy partials are grouped differently (per tile of kContractTile q instead of per column).
// Algorithmic traffic per cell: 48 d + 4 d^2 B (the fused kernel: 16 d + 4 d^2).
// Layouts as in kernels_generic.cuh: stage [P][d][kyb][nx_loc], phi plane-major [kl][c][kx].
#pragma once

#include "fft_pow2.cuh"
#include "kernels_generic.cuh"

namespace gfmd {

// The transform phases on the specialised power-of-two passes (nx = N = 4096, single rank, ndof a multiple
// of 3): one CTA transforms the three dofs of one sublattice of one ky column, the passes of
// k_cols_fused_p2_lr (kernel_cols_lr.cuh) without the contraction in between.  Forward leaves the spectrum
// in HBM in POSITION order (digit-reversed, fft_pow2.cuh) -- nothing is un-permuted: the stiffness table
// is stored in the same order (phi_slot mode 2) and k_cols_contract reads both with one index.  Round 2:
// replaces the run-time mixed-radix engine for two atoms per cell on 4096-wide surfaces (4.68 ms per column
// stage at 4096 x 4096 on the generic engine, profiles/r2_stage_times_call1.txt).
// grid = min(nky_loc * ngrp, #SMs), persistent over (column, sublattice) pairs.
template <int N, int T, int DIR>
__global__ void __launch_bounds__(T, 1)
k_cols_fft_p2(const double2 *__restrict__ sin, double2 *__restrict__ sout, GridDesc g, const double2 *__restrict__ tw)
{
  constexpr int D = 3;
  constexpr int NW = T / 32;
  extern __shared__ double2 sm[];
  double2 *tws = sm + D * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ngrp = g.d / 3;
  p2_fill_tws<N>(tws, tw);
  __syncthreads();
  for (int vc = blockIdx.x; vc < g.nky_loc * ngrp; vc += gridDim.x) {
    const int kl = vc / ngrp, grp = vc - kl * ngrp;
    // single rank: dof a of column kl starts at ((3 grp + a) * kyb + kl) * nx
    const size_t col0 = ((size_t) (3 * grp) * g.kyb + kl) * N, dstride = (size_t) g.kyb * N;
    if (DIR < 0) {
      p2_pass0_fwd_blk<N, T, D, 0>(sm, tw, tws, [&](int a, int base, int off) { return sin[col0 + a * dstride + base + off]; });
      __syncthreads();
      p2_groupA_rest_seq<N, NW, -1, D, 0>(sm, tw, tws, lane, warp);
      __syncthreads();
#pragma unroll 1
      for (int idx = threadIdx.x; idx < N / 8; idx += T) p2_groupB_first_seq<N, -1, D, 0>(sm, tws, idx);
      __syncwarp();
#pragma unroll 1
      for (int idx = threadIdx.x; idx < N / 8; idx += T) {
        const int pos = p2_last_base(idx);
        const int key = swz_key(pos);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          double2 t[8];
          p2_last_fwd_load(sm + a * N, pos, key, t);
#pragma unroll
          for (int r = 0; r < 8; ++r) sout[col0 + a * dstride + pos + r] = t[r];      // 8 positions = one 128-byte line
        }
      }
      __syncthreads();            // the next column's pass 0 overwrites what other threads may still read
    } else {
#pragma unroll 1
      for (int idx = threadIdx.x; idx < N / 8; idx += T) {
        const int pos = p2_last_base(idx);
        const int key = swz_key(pos);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          double2 t[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) t[r] = sin[col0 + a * dstride + pos + r];
          p2_last_inv_store(sm + a * N, pos, key, t);
        }
      }
      __syncwarp();
#pragma unroll 1
      for (int idx = threadIdx.x; idx < N / 8; idx += T) p2_groupB_first_seq<N, +1, D, 0>(sm, tws, idx);
      __syncthreads();
      p2_groupA_rest_seq<N, NW, +1, D, 0>(sm, tw, tws, lane, warp);
      __syncthreads();
      p2_pass0_inv_blk<N, T, D, 0>(sm, tw, tws,
                                   [&](int a, int base, int off, double2 v) { sout[col0 + a * dstride + base + off] = v; });
      __syncthreads();
    }
  }
}

constexpr int kContractThreads = 256;
constexpr int kContractTile = 1024;   // q per CTA of k_cols_contract (4 per thread)

// element ix of dof `dof` of local column kl in the staging layout (piece p came from rank p)
__device__ __forceinline__ size_t col_index(const GridDesc &g, int d, int kl, int dof, int ix)
{
  const int p = ix / g.nx_loc, il = ix - p * g.nx_loc;
  return ((size_t) (p * d + dof) * g.kyb + kl) * g.nx_loc + il;
}

// grid = nky_loc * ngrp CTAs, ngrp = ceil(d / db); CTA (kl, grp) transforms dofs
// [grp*db, min(d, (grp+1)*db)) of column kl.  in == out is allowed (a CTA only touches its own data).
// __launch_bounds__(512, 1): with (512) alone ptxas 12.9 caps the kernel at 64 registers, which the
// 16 complex values a thread keeps across a pass barrier (fft_engine.cuh) already fill.
template <int DIR>
__global__ void __launch_bounds__(512, 1)
k_cols_split_fft(const double2 *stage_in, double2 *stage_out, GridDesc g, FftDesc fd, int ld, int db)
{
  extern __shared__ double2 smem[];
  const int d = g.d, nx = g.nx;
  const int ngrp = (d + db - 1) / db;
  const int kl = blockIdx.x / ngrp;
  const int dof0 = (blockIdx.x - kl * ngrp) * db;
  const int nd = d - dof0 < db ? d - dof0 : db;

  for (int idx = threadIdx.x; idx < nd * nx; idx += blockDim.x) {
    const int j = idx / nx, ix = idx - j * nx;
    smem[j * ld + ix] = stage_in[col_index(g, d, kl, dof0 + j, ix)];
  }
  __syncthreads();
  fft_batch<DIR>(smem, ld, nd, fd);
  for (int idx = threadIdx.x; idx < nd * nx; idx += blockDim.x) {
    const int j = idx / nx, ix = idx - j * nx;
    stage_out[col_index(g, d, kl, dof0 + j, ix)] = smem[j * ld + ix];
  }
}

// grid = nky_loc * ntile CTAs, ntile = ceil(nx / kContractTile); in place on `stage`.
// epart[kl * ntile + tile] receives the tile's energy partial (fixed order -> deterministic).
// DT = compile-time ndof (3, 6, 9, 12) or 0 for the run-time version (ndof <= 24).
template <int DT>
__global__ void __launch_bounds__(kContractThreads)
k_cols_contract(double2 *stage, GridDesc g, const double *__restrict__ phi, const double *__restrict__ linf,
                double *__restrict__ epart, StepResults *res)
{
  constexpr int DA = DT > 0 ? DT : 24;
  const int d = DT > 0 ? DT : g.d;
  const int nx = g.nx;
  const int ntile = (nx + kContractTile - 1) / kContractTile;
  const int kl = blockIdx.x / ntile;
  const int tile = blockIdx.x - kl * ntile;
  const int ky = g.ky0 + kl;
  const double wgt = (ky == 0 || (2 * ky == g.ny)) ? 1.0 : 2.0;
  const double *ph = phi + (size_t) kl * d * d * nx;
  const int kx_end = (tile + 1) * kContractTile < nx ? (tile + 1) * kContractTile : nx;

  double e = 0.0;
  for (int kx = tile * kContractTile + threadIdx.x; kx < kx_end; kx += kContractThreads) {
    double2 uv[DA], F[DA];
    if (DT > 0) {
#pragma unroll
      for (int i = 0; i < DA; ++i) uv[i] = stage[col_index(g, d, kl, i, kx)];
      phi_matvec<DA>(uv, F, [&](int c) { return __ldg(ph + (size_t) c * nx + kx); });
    } else {
      for (int i = 0; i < d; ++i) uv[i] = stage[col_index(g, d, kl, i, kx)];
      for (int i = 0; i < d; ++i) {
        const double a = __ldg(ph + (size_t) i * nx + kx);
        F[i] = make_double2(a * uv[i].x, a * uv[i].y);
      }
      int c = d;
      for (int i = 0; i < d; ++i)
        for (int j = i + 1; j < d; ++j) {
          const double2 p = make_double2(__ldg(ph + (size_t) c * nx + kx), __ldg(ph + (size_t) (c + 1) * nx + kx));
          c += 2;
          F[i].x = fma(p.x, uv[j].x, fma(-p.y, uv[j].y, F[i].x));
          F[i].y = fma(p.x, uv[j].y, fma(p.y, uv[j].x, F[i].y));
          F[j].x = fma(p.x, uv[i].x, fma(p.y, uv[i].y, F[j].x));
          F[j].y = fma(p.x, uv[i].y, fma(-p.y, uv[i].x, F[j].y));
        }
    }
    double eq = 0.0;
#pragma unroll
    for (int i = 0; i < DA; ++i)
      if (i < d) {
        eq = fma(F[i].x, uv[i].x, fma(F[i].y, uv[i].y, eq));
        F[i] = make_double2(-F[i].x, -F[i].y);
      }
    e = fma(wgt, eq, e);
    if (ky == 0 && kx == 0) {          // gamma point: u0, -2 linf u0z, +linf (gfmd_solver_static.cpp:168-196,213-225)
      double eg = 0.0;
      for (int i = 0; i < d; ++i) res->u0[i] = uv[i].x;
      for (int a = 0; a < d / 3; ++a) {
        eg -= 2.0 * linf[a] * uv[3 * a + 2].x;
        F[3 * a + 2].x += linf[a];
      }
      res->egamma = eg;
    }
#pragma unroll
    for (int i = 0; i < DA; ++i)
      if (i < d) stage[col_index(g, d, kl, i, kx)] = F[i];
  }

  __shared__ double she[kContractThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0) she[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int k = 0; k < kContractThreads / 32; ++k) a += she[k];
    epart[blockIdx.x] = a;
  }
}

}  // namespace gfmd
