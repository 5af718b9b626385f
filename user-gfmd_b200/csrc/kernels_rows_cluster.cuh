// Row kernels for ny = 16384 as thread-block CLUSTERS of two CTAs (variant id ny + 9).
//
// One row of 16384 reals is 8192 packed complex numbers = 128 KB of shared memory: a CTA holds ONE
// row, and a lone row can only touch the transposed staging buffer [ky][ix] with 16-byte accesses,
// half of a 32-byte sector each (k_rows_*_r16w: 0.32 of the HBM rate against 0.65 of the two-row
// ny = 4096 kernels, profiles/r2_rows_variants.txt).  Here the two CTAs of a cluster transform the
// neighbouring rows ix, ix + 1 with the very same passes as k_rows_*_r16w and split the staging
// accesses by FREQUENCY instead of by row:
//
//   forward   each thread has the 16 spectrum points of its unit in registers (8 "A" frequencies
//             klow + q S, 8 "B" frequencies klow2 + q S).  CTA 0 stores all A points of BOTH rows,
//             CTA 1 all B points: a thread writes the 8 points its partner stores straight into the
//             partner's shared memory (distributed shared memory), one cluster barrier, then
//             stores 8 x 32 bytes (two rows of one wavevector): whole sectors.
//   backward  the transposed flow: CTA 0 loads the A points of both rows with 32-byte loads, keeps
//             its own row's and hands the other row's to its partner through DSMEM; CTA 1 likewise
//             with the B points.
//
// Arithmetic per element is that of k_rows_*_r16w (same functions), so the results are bit-identical
// to variant ny + 8.  Not compiled for the CPU emulation build (no clusters there); verified on the
// GPU against the oracle and against variant ny + 8 (tests/test_gpu_parity.py).
//
// Reference: GFMDSolverFFT::fft_forward / fft_reverse, y part (src/solvers/gfmd_solver_fft.cpp:96-147, :150-195).
#pragma once

#ifndef GFMD_CUDA_EMU
#include <cooperative_groups.h>

namespace gfmd {

namespace cgx = cooperative_groups;

// exchange buffer behind the row: slot q of thread p at xb[q * T + p], the Nyquist point at xb[8 * T]
template <int NR, int T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T, 1)
k_rows_fwd_r16c(const double *__restrict__ u, double2 *__restrict__ stage, GridDesc g,
                const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0)
{
  static_assert(NR == 8192 && NR / 16 == T, "k_rows_fwd_r16c: NR = 16 * 8 * 8 * 8, one row per CTA, two CTAs per cluster");
  constexpr int S = NR / 8;
  extern __shared__ double2 sm[];
  double2 *xb = sm + NR;
  cgx::cluster_group cluster = cgx::this_cluster();
  const unsigned crank = cluster.block_rank();            // 0: stores the A frequencies, 1: the B frequencies
  double2 *xb_peer = cluster.map_shared_rank(xb, crank ^ 1);
  const int dof = dof0 + blockIdx.x / g.nx_loc;
  const int ix = blockIdx.x % g.nx_loc;                   // nx_loc even: the pair (ix & ~1, ix | 1) is one cluster
  const int t = threadIdx.x;
  cluster.barrier_arrive();                               // "I am running": awaited before the first remote store
  {
    const double2 *src = reinterpret_cast<const double2 *>(u + ((size_t) dof * g.nx_loc + ix) * (2 * NR));
    double2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = src[t + T * j];
    dft16<-1>(v);
    r16_twiddle<-1, true>(v, __ldg(tw + t));
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) sm[q0 * T + t] = v[r16_out(q0)];
  }
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < NR / 8; i += T) {
    const int t1 = i & 63;
    double2 *blk = sm + (i >> 6) * 512 + t1;
    double2 v[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = blk[64 * j];
    Butterfly<8, -1>::run(v);
    tw_powers<8>(__ldg(tw + 16 * t1), w);
    p2_apply_tw<8, -1>(v, w);
#pragma unroll
    for (int q = 0; q < 8; ++q) blk[64 * q] = v[q];
  }
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < NR / 8; i += T) {
    const int t2 = i & 7, q1 = (i >> 3) & 7;
    double2 *blk = sm + (i >> 3) * 64;
    double2 v[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = blk[t2 + 8 * j];
    Butterfly<8, -1>::run(v);
    tw_powers<8>(__ldg(tw + 128 * t2), w);
    p2_apply_tw<8, -1>(v, w);
    __syncwarp();
#pragma unroll
    for (int q2 = 0; q2 < 8; ++q2) blk[(q2 << 3) + (t2 ^ ((q2 & 3) | ((q1 & 1) << 2)))] = v[q2];
  }
  __syncthreads();
  {
    const int p = t;
    const int klow = r16w_klow(p);
    const int klow2 = p == 0 ? S / 2 : S - klow;
    double2 v1[8], v2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v1[e] = sm[r16w_slot(klow, e)];
      v2[e] = sm[r16w_slot(klow2, e)];
    }
    Butterfly<8, -1>::run(v1);
    Butterfly<8, -1>::run(v2);
    const RowsFuseTw<NR, 2> w(tw_ny, klow, klow2, p == 0);
    // A points: frequencies klow + q S (p = 0: 0, S, ..., 7 S and the Nyquist point NR); B points: klow2 + q S
    double2 keep[8], nyq = make_double2(0.0, 0.0);
    cluster.barrier_wait();                                 // the partner CTA has started: its shared memory is valid
    if (p == 0) {
      double2 a[8], b[8];
      a[0] = rows_unmix(v1[0], v1[0], w.get(0, 0));
#pragma unroll
      for (int q = 1; q < 8; ++q) a[q] = rows_unmix(v1[q], v1[8 - q], w.get(0, q));
#pragma unroll
      for (int q = 0; q < 8; ++q) b[q] = rows_unmix(v2[q], v2[7 - q], w.get(1, q));
      nyq = rows_unmix(v1[0], v1[0], w.nyquist());
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        keep[q] = crank == 0 ? a[q] : b[q];
        xb_peer[q * T + p] = crank == 0 ? b[q] : a[q];
      }
      if (crank == 1) xb_peer[8 * T] = nyq;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double2 a = rows_unmix(v1[q], v2[7 - q], w.get(0, q));
        const double2 b = rows_unmix(v2[q], v1[7 - q], w.get(1, q));
        keep[q] = crank == 0 ? a : b;
        xb_peer[q * T + p] = crank == 0 ? b : a;
      }
    }
    cluster.sync();                                         // the partner's points have landed in xb
    const int ixe = ix & ~1;
    const int kbase = crank == 0 ? klow : klow2;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double2 other = xb[q * T + p];                  // the same wavevector of the partner's row
      double2 *dst = stage + stage_index(g, kbase + q * S, dof, ixe);
      if (crank == 0) st_global_256(dst, keep[q], other);   // rows (ixe, ixe + 1) = (mine, partner's)
      else st_global_256(dst, other, keep[q]);
    }
    if (p == 0 && crank == 0) st_global_256(stage + stage_index(g, NR, dof, ixe), nyq, xb[8 * T]);
  }
}

template <int NR, int T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T, 1)
k_rows_inv_r16c(const double2 *__restrict__ stage, double *__restrict__ f, GridDesc g,
                const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0)
{
  static_assert(NR == 8192 && NR / 16 == T, "k_rows_inv_r16c: NR = 16 * 8 * 8 * 8, one row per CTA, two CTAs per cluster");
  constexpr int S = NR / 8;
  extern __shared__ double2 sm[];
  double2 *xb = sm + NR;
  cgx::cluster_group cluster = cgx::this_cluster();
  const unsigned crank = cluster.block_rank();
  double2 *xb_peer = cluster.map_shared_rank(xb, crank ^ 1);
  const int dof = dof0 + blockIdx.x / g.nx_loc;
  const int ix = blockIdx.x % g.nx_loc;
  const int t = threadIdx.x;
  cluster.sync();                                         // both CTAs run: remote shared memory is valid
  {
    const int p = t;
    const int klow = r16w_klow(p);
    const int klow2 = p == 0 ? S / 2 : S - klow;
    const int ixe = ix & ~1;
    const int kbase = crank == 0 ? klow : klow2;
    // this CTA's frequencies of BOTH rows, 32 bytes per load; the partner row's points go to the partner
    double2 mine[8], yh0 = make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double2 r0, r1;
      ld_global_256(stage + stage_index(g, kbase + q * S, dof, ixe), r0, r1);
      mine[q] = crank == 0 ? r0 : r1;
      xb_peer[q * T + p] = crank == 0 ? r1 : r0;
    }
    if (p == 0 && crank == 0) {
      double2 r0, r1;
      ld_global_256(stage + stage_index(g, NR, dof, ixe), r0, r1);
      yh0 = r0;
      xb_peer[8 * T] = r1;
    }
    cluster.sync();
    double2 y1[8], y2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double2 other = xb[q * T + p];
      y1[q] = crank == 0 ? mine[q] : other;                 // A points of my row
      y2[q] = crank == 0 ? other : mine[q];                 // B points of my row
    }
    if (p == 0 && crank == 1) yh0 = xb[8 * T];
    const RowsFuseTw<NR, 2> w(tw_ny, klow, klow2, p == 0);
    if (p == 0) {
      double2 z[8];
      z[0] = rows_premix(y1[0], yh0, w.get(0, 0));
#pragma unroll
      for (int q = 1; q < 8; ++q) z[q] = rows_premix(y1[q], y1[8 - q], w.get(0, q));
#pragma unroll
      for (int q = 0; q < 8; ++q) y1[q] = z[q];
#pragma unroll
      for (int q = 0; q < 8; ++q) z[q] = rows_premix(y2[q], y2[7 - q], w.get(1, q));
#pragma unroll
      for (int q = 0; q < 8; ++q) y2[q] = z[q];
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double2 a = y1[q], b = y2[7 - q];
        y1[q] = rows_premix(a, b, w.get(0, q));
        y2[7 - q] = rows_premix(b, a, w.get(1, 7 - q));
      }
    }
    Butterfly<8, +1>::run(y1);
    Butterfly<8, +1>::run(y2);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sm[r16w_slot(klow, e)] = y1[e];
      sm[r16w_slot(klow2, e)] = y2[e];
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < NR / 8; i += T) {
    const int t2 = i & 7, q1 = (i >> 3) & 7;
    double2 *blk = sm + (i >> 3) * 64;
    double2 v[8], w[8];
#pragma unroll
    for (int q2 = 0; q2 < 8; ++q2) v[q2] = blk[(q2 << 3) + (t2 ^ ((q2 & 3) | ((q1 & 1) << 2)))];
    tw_powers<8>(__ldg(tw + 128 * t2), w);
    p2_apply_tw<8, +1>(v, w);
    Butterfly<8, +1>::run(v);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) blk[t2 + 8 * j] = v[j];
  }
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < NR / 8; i += T) {
    const int t1 = i & 63;
    double2 *blk = sm + (i >> 6) * 512 + t1;
    double2 v[8], w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = blk[64 * q];
    tw_powers<8>(__ldg(tw + 16 * t1), w);
    p2_apply_tw<8, +1>(v, w);
    Butterfly<8, +1>::run(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) blk[64 * j] = v[j];
  }
  __syncthreads();
  {
    double2 v[16];
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) v[q0] = sm[q0 * T + t];
    r16_twiddle<+1, false>(v, __ldg(tw + t));
    dft16<+1>(v);
    double2 *dst = reinterpret_cast<double2 *>(f + ((size_t) dof * g.nx_loc + ix) * (2 * NR));
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[t + T * j] = v[r16_out(j)];
  }
  // a CTA may not exit while its partner can still write into its exchange buffer: the only remote
  // writes precede the cluster barrier above, so nothing is pending here
}

}  // namespace gfmd
#endif   // !GFMD_CUDA_EMU
