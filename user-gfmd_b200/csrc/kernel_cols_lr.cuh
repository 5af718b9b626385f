// Fused column kernel, 16-warp / low-register variant (ndof 3, sub-column length 4096):
// x-FFT + Phi(q).u(q) + energy + gamma point + x-IFFT with one column set
// (3 x 4096 complex = 192 KB) resident in shared memory.
//
// Differences to k_cols_fused_p2 (kernels_fast.cuh, 8 warps x 255 registers):
// 512 threads x <= 128 registers, i.e. twice the warps to hide shared-memory and
// L2/DRAM latency.  The three dofs go through a two-deep register pipeline in
// every pass, and in the contraction dofs 0 and 1 stay in registers while dof 2
// is kept in place in shared memory (same thread, no barrier).
//
// Long columns (nx = 8192, 16384): the column transform is split once more,
// decimation in frequency over the top digit (radix R = nx / 4096):
//   k_cols_top_pass<-1>  y_q[n] = (sum_r x[r*4096 + n] w_R^{rq}) w_nx^{qn}, in place in HBM
//   k_cols_fused_p2_lr   on the R * nky "virtual columns" (ky, q) of length 4096, which
//                        hold the frequencies kx = R k' + q
//   k_cols_top_pass<+1>  the transposed step on the way back.
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

// Where the pieces of this rank's columns live in the PEERS' memory (PEER variants, NVLink loads and
// stores issued by the kernel itself): p[r] = a buffer of rank r + this rank's block offset (own
// rank: the local buffer), so that piece r of local column kl, dof a, is
// p[r][(a * kyb + kl) * nx_loc + x % nx_loc] -- the layout the copy-engine pushes produce,
// without the staging write, the copy and its wait.  As destination (`outp`): the ranks' return
// buffers, read by their backward row kernels.  As source (`inp`): the ranks' row-kernel outputs.
struct PeerOut {
  double2 *p[16];
};

// LP = log2(pieces per 4096-element sub-column) = max(0, 12 - log2(nx_loc)): element
// x = x0 + xi of dof a lives in source-rank piece p = x / nx_loc at
// ((p*D + a)*kyb + kl) * nx_loc + x % nx_loc.  The piece of a butterfly element relative to
// the first one is known at compile time (offset >> LNXLC).
// PIPE (the default; GFMD_B200_COLS_PIPE=0 selects the plain form): the last
// backward pass of column c is fused with
// pass 0 of column c + gridDim.x (p2_pass0_inv_fwd_blk): the loads of the next column are in
// flight while the finished column is transformed and stored, instead of after it.
// PEER (slab mode with peer mappings, ltop == 0; opt-in).  1 (GFMD_B200_PEER_STORE=1): the
// last backward pass stores each piece straight into its owner's return buffer (`outp`) -- full
// 128-byte lines over NVLink -- instead of the local staging buffer.  2 (GFMD_B200_PEER_DIRECT=1):
// pass 0 also LOADS each piece straight from the row-kernel output of the rank that produced it
// (`inp`), so the column stage has no transfer before or after it at all.
template <int N, int T, int LP, bool PIPE = false, int PEER = 0>
__global__ void __launch_bounds__(T, 1)
k_cols_fused_p2_lr(const double2 *__restrict__ sin, double2 *__restrict__ sout, GridDesc g, int lnxl, int ltop,
                   int kl0, int kl1,   // local ky range of this launch (chunked multi-GPU pipeline)
                   const double2 *__restrict__ tw, const double *__restrict__ phi,
                   const double *__restrict__ linf, double *__restrict__ epart, StepResults *res,
                   PeerOut outp = PeerOut(), PeerOut inp = PeerOut())
{
  static_assert(!(PIPE && PEER != 0), "the pipelined variant has no peer form");
  constexpr int D = 3;
  constexpr int NW = T / 32;
  constexpr int LNXLC = P2<N>::LOG - LP;                      // min(log2 N, lnxl)
  constexpr int XMASK = (1 << LNXLC) - 1;
  extern __shared__ double2 sm[];
  __shared__ double warp_e[NW];
  double2 *tws = sm + D * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t pstride = ((size_t) D * g.kyb) << lnxl;        // elements between pieces of one dof
  const size_t dstride = ((size_t) g.kyb) << lnxl;            // elements between dofs of one piece
  const int nvc = kl1 << ltop;                                // virtual columns (ky, q) end
  p2_fill_tws<N>(tws, tw);
  __syncthreads();

  auto column_base = [&](int vc) -> size_t {
    const int kl = vc >> ltop;
    const int x0 = (vc & ((1 << ltop) - 1)) << P2<N>::LOG;
    const int p0 = x0 >> lnxl;
    return ((((size_t) p0 * D) * g.kyb + kl) << lnxl) + (size_t) (x0 & ((1 << lnxl) - 1));
  };

#ifdef GFMD_PHASE_TIMING
  long long tprev__ = clock64();
#endif
  for (int vc = (kl0 << ltop) + blockIdx.x; vc < nvc; vc += gridDim.x) {
    const int ky = g.ky0 + (vc >> ltop);
    const size_t col0 = column_base(vc);
    auto addr = [&](int a, int base, int off) -> size_t {
      return col0 + a * dstride + (size_t) (off >> LNXLC) * pstride + (size_t) ((off & XMASK) + base);
    };
    const double *ph = phi + (size_t) vc * D * D * N;

    // ---- pass 0 straight from global memory (block-wide mapping: full 128-byte lines)
    if (!PIPE || vc == (kl0 << ltop) + (int) blockIdx.x) {
    if constexpr (PEER == 2) {
      const size_t rel0 = ((size_t) (vc >> ltop)) << lnxl;
      p2_pass0_fwd_blk<N, T, D, 0>(sm, tw, tws, [&](int a, int base, int off) {
        return inp.p[off >> LNXLC][rel0 + a * dstride + (size_t) ((off & XMASK) + base)];
      });
    } else
    p2_pass0_fwd_blk<N, T, D, 0>(sm, tw, tws, [&](int a, int base, int off) { return sin[addr(a, base, off)]; });
    }
    __syncthreads();
    PHASE_MARK(0);
    p2_groupA_rest_seq<N, NW, -1, D, 0>(sm, tw, tws, lane, warp);
    PHASE_MARK(1);
    __syncthreads();
    PHASE_MARK(2);
    // just-in-time L2 prefetch of this column's Phi planes (contiguous D*D*N doubles): the DRAM
    // fetch runs while the stride-8 pass computes; a much longer distance would be evicted by
    // the streaming traffic of the other SMs before use
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
          if (vc == 0 && g.ky0 == 0 && pos + r == 0) {   // gamma point: kx = 0 is position 0 of (ky 0, q 0)
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
    if (lane == 0) warp_e[warp] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) a += warp_e[k];
      epart[vc] = a;
    }
    PHASE_MARK(6);
    // next column's data -> L2 while the backward passes run (not for data in a peer's memory)
    if (PEER != 2 && vc + (int) gridDim.x < nvc) {
      const size_t ncol0 = column_base(vc + gridDim.x);
      for (int i = threadIdx.x; i < D * N / 8; i += T) {
        const int a = i / (N / 8), x = (i - a * (N / 8)) * 8;
        prefetch_l2(sin + ncol0 + a * dstride + (size_t) (x >> LNXLC) * pstride + (size_t) (x & XMASK));
      }
    }

    // ---- group A backward, last pass straight to global memory
    p2_groupA_rest_seq<N, NW, +1, D, 0>(sm, tw, tws, lane, warp);
    __syncthreads();
    PHASE_MARK(7);
    if (PIPE && vc + (int) gridDim.x < nvc) {
      const size_t ncol = column_base(vc + gridDim.x);
      p2_pass0_inv_fwd_blk<N, T, D, 0>(
          sm, tw, tws, [&](int a, int base, int off, double2 v) { sout[addr(a, base, off)] = v; },
          [&](int a, int base, int off) {
            return sin[ncol + a * dstride + (size_t) (off >> LNXLC) * pstride + (size_t) ((off & XMASK) + base)];
          });
    } else
    if constexpr (PEER != 0) {
      // ltop == 0: the sub-column is the column, piece = off >> LNXLC is a compile-time constant
      const size_t rel0 = ((size_t) (vc >> ltop)) << lnxl;
      p2_pass0_inv_blk<N, T, D, 0>(sm, tw, tws, [&](int a, int base, int off, double2 v) {
        outp.p[off >> LNXLC][rel0 + a * dstride + (size_t) ((off & XMASK) + base)] = v;
      });
    } else
    p2_pass0_inv_blk<N, T, D, 0>(sm, tw, tws,
                                 [&](int a, int base, int off, double2 v) { sout[addr(a, base, off)] = v; });
    PHASE_MARK(8);
    // no barrier: the next column's pass 0 writes exactly what this thread just read
  }
}

// Top-digit pass of a long column transform, in place in the staging buffer.
// One thread per (dof, kl, n), n < S = nx >> LR; elements x = n + r*S, r < R = 2^LR.
// PEER, DIR = +1: the results go to the pieces' owners (`peer`, see PeerOut) instead of back into
// `stage`.  PEER, DIR = -1: the inputs come from the ranks that produced them (`peer`) and the
// results go to `stage`.
// U = items per thread and iteration: the PEER forms run on a FEW SMs next to the fused column kernel
// (direct_pipelined_step) and need U * R independent 16-byte NVLink loads in flight per thread.
template <int LR, int DIR, bool PEER = false, int T = 256, int U = 1>
__global__ void __launch_bounds__(T)
k_cols_top_pass(double2 *__restrict__ stage, GridDesc g, int lnxl, const double2 *__restrict__ tw_nx, int kl0,
                int kl1, PeerOut peer = PeerOut(), int dof0 = 0, int nd = -1)
{
  // dofs [dof0, dof0 + nd) of the columns kl0 .. kl1 - 1 (nd < 0: all)
  constexpr int R = 1 << LR;
  const int S = g.nx >> LR;
  const int xmask = (1 << lnxl) - 1;
  if (nd < 0) nd = g.d - dof0;
  const long long total = (long long) nd * (kl1 - kl0) * S;
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long idx0 = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx0 < total; idx0 += stride * U) {
    double2 v[U][R];
    size_t a[U][R];
    int nn[U];
    size_t rel[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long idx = idx0 + k * stride;
      if (idx < total) {
        const int n = (int) (idx % S);
        const int col = (int) (idx / S);
        const int dof = dof0 + col % nd, kl = kl0 + col / nd;
        nn[k] = n;
        rel[k] = ((((size_t) dof) * g.kyb + kl) << lnxl);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int x = n + r * S;
          a[k][r] = ((((size_t) (x >> lnxl) * g.d + dof) * g.kyb + kl) << lnxl) + (size_t) (x & xmask);
          if (PEER && DIR < 0) v[k][r] = peer.p[x >> lnxl][rel[k] + (size_t) (x & xmask)];
          else v[k][r] = stage[a[k][r]];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long idx = idx0 + k * stride;
      if (idx < total) {
        const int n = nn[k];
        if (DIR > 0) {
#pragma unroll
          for (int q = 1; q < R; ++q) v[k][q] = cmulc(v[k][q], __ldg(tw_nx + (size_t) q * n));
        }
        Butterfly<R, DIR>::run(v[k]);
        if (DIR < 0) {
#pragma unroll
          for (int q = 1; q < R; ++q) v[k][q] = cmul(v[k][q], __ldg(tw_nx + (size_t) q * n));
        }
        if constexpr (PEER && DIR > 0) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int x = n + r * S;
            peer.p[x >> lnxl][rel[k] + (size_t) (x & xmask)] = v[k][r];
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) stage[a[k][r]] = v[k][r];
        }
      }
    }
  }
}

}  // namespace gfmd
