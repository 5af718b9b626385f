// Fused column kernel, 16-warp / low-register variant (ndof 3, nx = 4096):
// x-FFT + Phi(q).u(q) + energy + gamma point + x-IFFT with one column set
// (3 x 4096 complex = 192 KB) resident in shared memory.
//
// Differences to k_cols_fused_p2 (kernels_fast.cuh, 8 warps x 255 registers):
// 512 threads x <= 128 registers, i.e. twice the warps to hide shared-memory and
// L2/DRAM latency.  The three dofs go through a two-deep register pipeline in
// every pass, and in the contraction dofs 0 and 1 stay in registers while dof 2
// is kept in place in shared memory (same thread, no barrier).
#pragma once

#include "fft_pow2.cuh"
#include "kernels_generic.cuh"

namespace gfmd {

#ifdef GFMD_PHASE_TIMING
__device__ long long g_phase_cycles[16];
#define PHASE_MARK(i)                                                  \
  do {                                                                 \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                         \
      const long long t__ = clock64();                                 \
      g_phase_cycles[i] += t__ - tprev__;                              \
      tprev__ = t__;                                                   \
    }                                                                  \
  } while (0)
#else
#define PHASE_MARK(i)
#endif

// LP = log2(number of slab ranks): the column of dof a is made of 2^LP pieces of
// nx_loc = N >> LP elements, piece p at sin + ((p*D + a)*kyb + kl) * nx_loc.  The piece of
// a butterfly element is known at compile time (offset >> LNXL), so addresses are one
// base pointer per dof plus constants.
template <int N, int T, int LP>
__global__ void __launch_bounds__(T, 1)
k_cols_fused_p2_lr(const double2 *__restrict__ sin, double2 *__restrict__ sout, GridDesc g,
                   const double2 *__restrict__ tw, const double *__restrict__ phi,
                   const double *__restrict__ linf, double *__restrict__ epart, StepResults *res)
{
  constexpr int D = 3;
  constexpr int NW = T / 32;
  constexpr int LNXL = P2<N>::LOG - LP;
  constexpr int XMASK = (1 << LNXL) - 1;
  extern __shared__ double2 sm[];
  double2 *tws = sm + D * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t pstride = ((size_t) D * g.kyb) << LNXL;      // elements between pieces of one dof
  p2_fill_tws<N>(tws, tw);
  __syncthreads();

#ifdef GFMD_PHASE_TIMING
  long long tprev__ = clock64();
#endif
  for (int kl = blockIdx.x; kl < g.nky_loc; kl += gridDim.x) {
    const int ky = g.ky0 + kl;
    const size_t col0 = ((size_t) kl) << LNXL;
    const size_t dstride = ((size_t) g.kyb) << LNXL;          // elements between dofs of one piece
    auto addr = [&](int a, int base, int off) -> size_t {
      return col0 + a * dstride + (size_t) (off >> LNXL) * pstride + (size_t) ((off & XMASK) + base);
    };
    const double *ph = phi + (size_t) kl * D * D * N;

    // ---- group A forward: pass 0 straight from global memory
    p2_pass0_fwd_blk<N, T, D, 0>(sm, tw, tws, [&](int a, int base, int off) { return sin[addr(a, base, off)]; });
    __syncthreads();
    PHASE_MARK(0);
    p2_groupA_rest_seq<N, NW, -1, D, 0>(sm, tw, tws, lane, warp);
    PHASE_MARK(1);
    __syncthreads();
    PHASE_MARK(2);
    // just-in-time L2 prefetch of this column's Phi planes (contiguous D*D*N doubles): the DRAM
    // fetch runs while the in-shared-memory passes compute; a much longer distance would be
    // evicted by the streaming traffic of the other SMs before use
    for (int i = threadIdx.x; i < D * D * N / 16; i += T) prefetch_l2(ph + (size_t) i * 16);

    // ---- group B forward, contraction, group B backward
#pragma unroll 1
    for (int idx = threadIdx.x; idx < N / 8; idx += T) p2_groupB_first_seq<N, -1, D, 0>(sm, tws, idx);
    __syncwarp();
    PHASE_MARK(3);

    const double wgt = (ky == 0 || (2 * ky == g.ny)) ? 1.0 : 2.0;
    double e = 0.0;
#pragma unroll 1
    for (int idx = threadIdx.x; idx < N / 8; idx += T) {
      const int pos = p2_last_base(idx);
      const int key = swz_key(pos);
      // interleaved Phi planes of this item (see phi_slot in gfmd_b200.cu)
      const double *phi_item = ph + (size_t) (idx >> 3) * (64 * D * D) + (idx & 7) * 2;
      double2 *s2 = sm + 2 * N + pos;               // dof 2 of this item, element r at s2[r ^ key]
      double2 u0[8], u1[8];
      {
        double2 t[8];
        p2_last_fwd_load(sm + 2 * N, pos, key, t);
#pragma unroll
        for (int r = 0; r < 8; ++r) s2[r ^ key] = t[r];
      }
      p2_last_fwd_load(sm, pos, key, u0);
      p2_last_fwd_load(sm + N, pos, key, u1);
#pragma unroll
      for (int rp = 0; rp < 4; ++rp) {
        double2 pl[D * D];
#pragma unroll
        for (int c = 0; c < D * D; ++c)
          pl[c] = __ldg(reinterpret_cast<const double2 *>(phi_item + rp * (16 * D * D) + c * 16));
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int r = 2 * rp + s;
          double2 uv[D], F[D];
          uv[0] = u0[r];
          uv[1] = u1[r];
          uv[2] = s2[r ^ key];
          phi_matvec<D>(uv, F, [&](int c) { return s == 0 ? pl[c].x : pl[c].y; });
          double eq = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            eq = fma(F[i].x, uv[i].x, fma(F[i].y, uv[i].y, eq));
            F[i] = make_double2(-F[i].x, -F[i].y);
          }
          e = fma(wgt, eq, e);
          if (ky == 0 && pos + r == 0) {            // gamma point: kx = 0 sits at position 0
#pragma unroll
            for (int i = 0; i < D; ++i) res->u0[i] = uv[i].x;
            res->egamma = -2.0 * linf[0] * uv[2].x;
            F[2].x += linf[0];
          }
          u0[r] = F[0];
          u1[r] = F[1];
          s2[r ^ key] = F[2];
        }
      }
      p2_last_inv_store(sm, pos, key, u0);
      p2_last_inv_store(sm + N, pos, key, u1);
      {
        double2 t[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) t[r] = s2[r ^ key];
        p2_last_inv_store(sm + 2 * N, pos, key, t);
      }
    }
    __syncwarp();
    PHASE_MARK(4);
#pragma unroll 1
    for (int idx = threadIdx.x; idx < N / 8; idx += T) p2_groupB_first_seq<N, +1, D, 0>(sm, tws, idx);
    PHASE_MARK(5);

    // energy partial of this warp (fixed order -> deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if (lane == 0) epart[(size_t) kl * NW + warp] = e;
    __syncthreads();
    PHASE_MARK(6);
    // next column's data -> L2 while the backward passes run
    if (kl + (int) gridDim.x < g.nky_loc) {
      const size_t ncol0 = ((size_t) (kl + gridDim.x)) << LNXL;
      for (int i = threadIdx.x; i < D * N / 8; i += T) {
        const int a = i / (N / 8), x = (i - a * (N / 8)) * 8;
        prefetch_l2(sin + ncol0 + a * dstride + (size_t) (x >> LNXL) * pstride + (size_t) (x & XMASK));
      }
    }
    // ---- group A backward, last pass straight to global memory
    p2_groupA_rest_seq<N, NW, +1, D, 0>(sm, tw, tws, lane, warp);
    __syncthreads();
    PHASE_MARK(7);
    p2_pass0_inv_blk<N, T, D, 0>(sm, tw, tws,
                                 [&](int a, int base, int off, double2 v) { sout[addr(a, base, off)] = v; });
    PHASE_MARK(8);
    // no barrier: the next column's pass 0 writes exactly what this thread just read
  }
}

}  // namespace gfmd
