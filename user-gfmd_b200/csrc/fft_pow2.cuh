// Power-of-two FP64 complex FFT passes in shared memory, in place, compile-time
// length.  Forward = decimation in frequency (natural order in, digit-reversed
// order out), backward = the transposed flow, decimation in time (digit-reversed
// in, natural out).  Because every pass reads and writes the SAME 8 (or 2/4)
// locations, a pass needs no read/write barrier, and because
//   * the passes with stride >= 64 ("group A") exchange data only between items
//     that share the low 6 position bits, and
//   * the last two passes, strides 8 and 1 ("group B"), only inside one aligned
//     block of 64 elements,
// the work is mapped so that each of those groups stays inside one warp: a whole
// transform costs ONE __syncthreads (between group A and group B); everything
// else is __syncwarp.  The spectrum is left digit-reversed -- the Phi table is
// stored in the same permuted order (host: p2_freq_to_pos), so it never needs to
// be un-permuted.
//
// All pass functions work on A arrays at once (the ndof columns of one q-column,
// or the rows of one row tile): the twiddles are fetched once per item and the
// loads of all arrays are issued before the first butterfly.
//
// Shared-memory addressing is XOR-swizzled: the 16-byte column of an element is
// XORed with a rotation of the XOR of all its higher 3-bit digits, which makes
// every access pattern used here (8 lanes varying one digit, or 4 consecutive
// elements x 2 values of a digit bit) bank-conflict free.  For an item with base
// position `base` the key is computed once (swz_key); element r of the butterfly
// then sits at (base ^ key ^ rot2(r)) + (r << stride).
#pragma once

#include "fft_engine.cuh"

namespace gfmd {

constexpr int ilog2_c(int n) { return n <= 1 ? 0 : 1 + ilog2_c(n >> 1); }

template <int N> struct P2 {
  static constexpr int LOG = ilog2_c(N);
  static constexpr int NP = (LOG + 2) / 3;            // passes
  static constexpr int LR0 = LOG - 3 * (NP - 1);      // log2 radix of pass 0 (1, 2 or 3)
  static constexpr int NA = NP - 2;                   // group-A passes (stride >= 64)
  __host__ __device__ static constexpr int lr(int p) { return p == 0 ? LR0 : 3; }
  __host__ __device__ static constexpr int ls(int p) { return 3 * (NP - 1 - p); }
  // shared-memory twiddle tables for the passes with stride 64 and 8: [q-1][n_lo]
  static constexpr int TW64 = 0;                      // 7 * 64 entries
  static constexpr int TW8 = 7 * 64;                  // 7 * 8 entries
  static constexpr int TWS = 7 * 64 + 7 * 8;          // 504 double2 = 8064 B
  static_assert(NP >= 4 && NP <= 5, "fast path covers 1024 <= N <= 8192");
};

__host__ __device__ constexpr int rot2c(int x) { return ((x << 2) | (x >> 1)) & 7; }

// 3-bit swizzle key of a position whose low 3 bits do not matter
__device__ __forceinline__ int swz_key(int pos)
{
  const int x = ((pos >> 3) ^ (pos >> 6) ^ (pos >> 9) ^ (pos >> 12)) & 7;
  return ((x << 2) | (x >> 1)) & 7;
}
__host__ __device__ __forceinline__ int swz(int pos)
{
  int x = ((pos >> 3) ^ (pos >> 6) ^ (pos >> 9) ^ (pos >> 12)) & 7;
  x = ((x << 2) | (x >> 1)) & 7;
  return pos ^ x;
}

// frequency index k -> position in the digit-reversed spectrum (host + device)
__host__ __device__ inline int p2_freq_to_pos(int log, int k)
{
  const int np = (log + 2) / 3;
  const int lr0 = log - 3 * (np - 1);
  int pos = (k & ((1 << lr0) - 1)) << (3 * (np - 1));
  k >>= lr0;
  for (int p = 1; p < np; ++p) {
    pos |= (k & 7) << (3 * (np - 1 - p));
    k >>= 3;
  }
  return pos;
}

// inverse of p2_freq_to_pos
__host__ __device__ inline int p2_pos_to_freq(int log, int pos)
{
  const int np = (log + 2) / 3;
  const int lr0 = log - 3 * (np - 1);
  int k = pos >> (3 * (np - 1));
  int sh = lr0;
  for (int p = 1; p < np; ++p) {
    k |= ((pos >> (3 * (np - 1 - p))) & 7) << sh;
    sh += 3;
  }
  return k;
}

// Where plane c of wavevector kx lives inside one column's block of d*d*nx doubles:
// off + c * cstride.  Generic kernels: plane-major [c][kx].  Specialised column kernels:
// the spectrum is digit-reversed (pos) and the planes are interleaved item by item,
// [pos / 64][(pos & 7) / 2][c][(pos >> 3) & 7][pos & 1], so that the 9 x 16-byte loads of
// one contraction round of 8 neighbouring threads form one contiguous 1152-byte chunk.
// mode 0: generic kernels, plane-major [c][kx].  mode 1: specialised fused column kernels (below).
// mode 2: plane-major in POSITION order [c][pos(kx)] -- the three-phase column stage on the
// power-of-two passes (k_cols_fft_p2, kernel_cols_split.cuh) leaves the spectrum digit-reversed in HBM,
// and the contraction kernel then reads table and spectrum with the same (coalesced) index.
__host__ __device__ inline void phi_slot(int mode, int top, int lognx, int nx, size_t dsq, int kx, size_t &off, size_t &cstride)
{
  if (mode == 0) {
    off = (size_t) kx;
    cstride = (size_t) nx;
    return;
  }
  if (mode == 2) {
    off = (size_t) p2_freq_to_pos(lognx, kx);
    cstride = (size_t) nx;
    return;
  }
  // long columns: kx = R k' + q lives in sub-column q (a block of dsq * nx / R doubles)
  const int q = kx & ((1 << top) - 1), ksub = kx >> top, logsub = lognx - top;
  const int pos = p2_freq_to_pos(logsub, ksub);
  const int blk = pos >> 6, a = (pos >> 3) & 7, r = pos & 7;
  off = (size_t) q * dsq * ((size_t) nx >> top) + (size_t) blk * 64 * dsq + (size_t) (r >> 1) * 16 * dsq +
        (size_t) a * 2 + (r & 1);
  cstride = 16;
}

__device__ __forceinline__ double2 csqr(double2 a)
{
  return make_double2(fma(a.x, a.x, -(a.y * a.y)), 2.0 * a.x * a.y);
}

// w[q] = w1^q, q = 1..R-1  (w[0] unused)
template <int R> __device__ __forceinline__ void tw_powers(double2 w1, double2 *w)
{
  w[1] = w1;
  if (R > 2) {
    w[2] = csqr(w1);
    w[3] = cmul(w[2], w1);
  }
  if (R > 4) {
    w[4] = csqr(w[2]);
    w[5] = cmul(w[4], w1);
    w[6] = csqr(w[3]);
    w[7] = cmul(w[4], w[3]);
  }
}

// fill the shared twiddle tables from the global table tw[k] = exp(-2 pi i k/N)
template <int N> __device__ __forceinline__ void p2_fill_tws(double2 *tws, const double2 *__restrict__ tw)
{
  // stride-64 pass: sub-transform length 512, exponent q*n_lo*(N/512)
  for (int i = threadIdx.x; i < 7 * 64; i += blockDim.x) {
    const int q = i / 64 + 1, n = i % 64;
    tws[P2<N>::TW64 + i] = tw[q * n * (N / 512)];
  }
  for (int i = threadIdx.x; i < 7 * 8; i += blockDim.x) {
    const int q = i / 8 + 1, n = i % 8;
    tws[P2<N>::TW8 + i] = tw[q * n * (N / 64)];
  }
}

// ------------------------------------------------------------ group A items ---
// Item m of this warp for pass P: returns the base position (digit P = 0).
template <int N, int P, int NW> struct GA {
  static constexpr int LOG = P2<N>::LOG;
  static constexpr int lr = P2<N>::lr(P);
  static constexpr int ls = P2<N>::ls(P);
  static constexpr int FH = LOG - 6 - lr;             // free high bits
  static constexpr int CW = 16 / NW;                  // low-6 combos per warp
  static constexpr int M = CW << FH;                  // items per warp per (i0 & 3)
  static_assert(16 % NW == 0 && M % 8 == 0, "warp count does not fit the group-A mapping");
  __device__ static __forceinline__ int base(int m, int lane, int warp)
  {
    const int e = m & ((1 << FH) - 1);
    const int c = warp * CW + (m >> FH);
    const int low6 = (c << 2) | (lane & 3);
    const int lowbits = e & ((1 << (ls - 6)) - 1);
    const int high = e >> (ls - 6);
    return (high << (ls + lr)) | (lowbits << 6) | low6;
  }
};

// twiddles w[q] = w_L^(q n_lo), q = 1..R-1, of pass P (forward sign)
template <int N, int P>
__device__ __forceinline__ void p2_get_tw(double2 *w, int n_lo, const double2 *__restrict__ tw, const double2 *tws)
{
  constexpr int lr = P2<N>::lr(P), ls = P2<N>::ls(P), R = 1 << lr;
  if (ls == 6) {
#pragma unroll
    for (int q = 1; q < R; ++q) w[q] = tws[P2<N>::TW64 + (q - 1) * 64 + n_lo];
  } else if (ls == 3) {
#pragma unroll
    for (int q = 1; q < R; ++q) w[q] = tws[P2<N>::TW8 + (q - 1) * 8 + n_lo];
  } else {
    // long strides: one table load, powers in registers
    tw_powers<R>(__ldg(tw + (n_lo << (P2<N>::LOG - ls - lr))), w);
  }
}

template <int R, int DIR> __device__ __forceinline__ void p2_apply_tw(double2 *v, const double2 *w)
{
#pragma unroll
  for (int q = 1; q < R; ++q) v[q] = DIR < 0 ? cmul(v[q], w[q]) : cmulc(v[q], w[q]);
}

// element r of an item: physical index inside one array
template <int LS> __device__ __forceinline__ int p2_elem(int sb, int r) { return (sb ^ rot2c(r)) + (r << LS); }

// One in-place group-A pass over A arrays (array a at sm + a*N).  Forward: butterfly
// then twiddle; backward: conj twiddle then inverse butterfly.
template <int N, int P, int NW, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupA_pass(double2 *sm, const double2 *__restrict__ tw, const double2 *tws,
                                               int lane, int warp)
{
  using G = GA<N, P, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    const int sb = base ^ swz_key(base);
    double2 w[8];
    p2_get_tw<N, P>(w, base & ((1 << G::ls) - 1), tw, tws);
    double2 v[A][R];
#pragma unroll
    for (int a = 0; a < A; ++a)
#pragma unroll
      for (int r = 0; r < R; ++r) v[a][r] = sm[a * N + (p2_elem<G::ls>(sb, r) ^ ((a * AX) & 7))];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if (DIR > 0) p2_apply_tw<R, DIR>(v[a], w);
      Butterfly<R, DIR>::run(v[a]);
      if (DIR < 0) p2_apply_tw<R, DIR>(v[a], w);
#pragma unroll
      for (int r = 0; r < R; ++r) sm[a * N + (p2_elem<G::ls>(sb, r) ^ ((a * AX) & 7))] = v[a][r];
    }
  }
}

// Pass 0 forward, inputs from global memory through load(a, base, offset).
template <int N, int NW, int A, int AX, typename Load>
__device__ __forceinline__ void p2_pass0_fwd(double2 *sm, const double2 *__restrict__ tw, const double2 *tws,
                                             int lane, int warp, Load load)
{
  using G = GA<N, 0, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    const int sb = base ^ swz_key(base);
    double2 v[A][R];
#pragma unroll
    for (int a = 0; a < A; ++a)
#pragma unroll
      for (int r = 0; r < R; ++r) v[a][r] = load(a, base, r << G::ls);
    double2 w[8];
    p2_get_tw<N, 0>(w, base, tw, tws);                // n_lo = base (digit 0 is the top digit)
#pragma unroll
    for (int a = 0; a < A; ++a) {
      Butterfly<R, -1>::run(v[a]);
      p2_apply_tw<R, -1>(v[a], w);
#pragma unroll
      for (int r = 0; r < R; ++r) sm[a * N + (p2_elem<G::ls>(sb, r) ^ ((a * AX) & 7))] = v[a][r];
    }
  }
}

// Pass 0 backward, outputs to global memory through store(a, base, offset, value).
template <int N, int NW, int A, int AX, typename Store>
__device__ __forceinline__ void p2_pass0_inv(const double2 *sm, const double2 *__restrict__ tw, const double2 *tws,
                                             int lane, int warp, Store store)
{
  using G = GA<N, 0, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    const int sb = base ^ swz_key(base);
    double2 w[8];
    p2_get_tw<N, 0>(w, base, tw, tws);
    double2 v[A][R];
#pragma unroll
    for (int a = 0; a < A; ++a)
#pragma unroll
      for (int r = 0; r < R; ++r) v[a][r] = sm[a * N + (p2_elem<G::ls>(sb, r) ^ ((a * AX) & 7))];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      p2_apply_tw<R, +1>(v[a], w);
      Butterfly<R, +1>::run(v[a]);
#pragma unroll
      for (int r = 0; r < R; ++r) store(a, base, r << G::ls, v[a][r]);
    }
  }
}

// remaining group-A passes 1 .. NA-1 (forward) / NA-1 .. 1 (backward), warp-synchronous
template <int N, int NW, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupA_rest(double2 *sm, const double2 *__restrict__ tw, const double2 *tws,
                                               int lane, int warp)
{
  constexpr int P2nd = P2<N>::NA > 2 ? 2 : 1;
  if (DIR < 0) {
    __syncwarp();
    p2_groupA_pass<N, 1, NW, -1, A, AX>(sm, tw, tws, lane, warp);
    if (P2<N>::NA > 2) {
      __syncwarp();
      p2_groupA_pass<N, P2nd, NW, -1, A, AX>(sm, tw, tws, lane, warp);
    }
  } else {
    if (P2<N>::NA > 2) {
      p2_groupA_pass<N, P2nd, NW, +1, A, AX>(sm, tw, tws, lane, warp);
      __syncwarp();
    }
    p2_groupA_pass<N, 1, NW, +1, A, AX>(sm, tw, tws, lane, warp);
    __syncwarp();
  }
}

// ------------------------------------------------------------ group B items ---
// item idx in [0, N/8): block of 64 = idx >> 3, al = idx & 7 (the same 8 lanes own a
// block in both passes).

// pass NP-2 (stride 8) in place over A arrays
template <int N, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupB_first(double2 *sm, const double2 *tws, int idx)
{
  const int al = idx & 7, base = (idx >> 3) * 64 + al;
  const int sb = base ^ swz_key(base);
  double2 w[8];
#pragma unroll
  for (int q = 1; q < 8; ++q) w[q] = tws[P2<N>::TW8 + (q - 1) * 8 + al];
  double2 v[A][8];
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int r = 0; r < 8; ++r) v[a][r] = sm[a * N + (p2_elem<3>(sb, r) ^ ((a * AX) & 7))];
#pragma unroll
  for (int a = 0; a < A; ++a) {
    if (DIR > 0) p2_apply_tw<8, DIR>(v[a], w);
    Butterfly<8, DIR>::run(v[a]);
    if (DIR < 0) p2_apply_tw<8, DIR>(v[a], w);
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[a * N + (p2_elem<3>(sb, r) ^ ((a * AX) & 7))] = v[a][r];
  }
}

// last pass (stride 1): element r of item idx sits at lastbase + (r ^ lastkey)
__device__ __forceinline__ int p2_last_base(int idx) { return (idx >> 3) * 64 + 8 * (idx & 7); }

__device__ __forceinline__ void p2_last_fwd_load(const double2 *a, int base, int key, double2 *v)
{
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = a[base + (r ^ key)];
  Butterfly<8, -1>::run(v);
}

__device__ __forceinline__ void p2_last_inv_store(double2 *a, int base, int key, double2 *v)
{
  Butterfly<8, +1>::run(v);
#pragma unroll
  for (int r = 0; r < 8; ++r) a[base + (r ^ key)] = v[r];
}

// ---------------------------------------------- low-register pass variants ---
// Same passes, but the A arrays go through a two-deep register pipeline (array
// a+1 is being loaded while array a is transformed) instead of being all live at
// once: ~64 instead of 32*A data registers, for kernels that run 16 warps per SM.

template <int N, int P, int NW, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupA_pass_seq(double2 *sm, const double2 *__restrict__ tw,
                                                   const double2 *tws, int lane, int warp)
{
  using G = GA<N, P, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    const int sb = base ^ swz_key(base);
    double2 w[8];
    p2_get_tw<N, P>(w, base & ((1 << G::ls) - 1), tw, tws);
    double2 v[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[0][r] = sm[p2_elem<G::ls>(sb, r)];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if (a + 1 < A) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          v[(a + 1) & 1][r] = sm[(a + 1) * N + (p2_elem<G::ls>(sb, r) ^ (((a + 1) * AX) & 7))];
      }
      if (DIR > 0) p2_apply_tw<R, DIR>(v[a & 1], w);
      Butterfly<R, DIR>::run(v[a & 1]);
      if (DIR < 0) p2_apply_tw<R, DIR>(v[a & 1], w);
#pragma unroll
      for (int r = 0; r < R; ++r) sm[a * N + (p2_elem<G::ls>(sb, r) ^ ((a * AX) & 7))] = v[a & 1][r];
    }
  }
}

template <int N, int NW, int A, int AX, typename Load>
__device__ __forceinline__ void p2_pass0_fwd_seq(double2 *sm, const double2 *__restrict__ tw, const double2 *tws,
                                                 int lane, int warp, Load load)
{
  using G = GA<N, 0, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    const int sb = base ^ swz_key(base);
    double2 v[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[0][r] = load(0, base, r << G::ls);
    if (A > 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[1][r] = load(1, base, r << G::ls);
    }
    double2 w[8];
    p2_get_tw<N, 0>(w, base, tw, tws);
#pragma unroll
    for (int a = 0; a < A; ++a) {
      Butterfly<R, -1>::run(v[a & 1]);
      p2_apply_tw<R, -1>(v[a & 1], w);
#pragma unroll
      for (int r = 0; r < R; ++r) sm[a * N + (p2_elem<G::ls>(sb, r) ^ ((a * AX) & 7))] = v[a & 1][r];
      if (a + 2 < A) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[a & 1][r] = load(a + 2, base, r << G::ls);
      }
    }
  }
}

template <int N, int NW, int A, int AX, typename Store>
__device__ __forceinline__ void p2_pass0_inv_seq(const double2 *sm, const double2 *__restrict__ tw,
                                                 const double2 *tws, int lane, int warp, Store store)
{
  using G = GA<N, 0, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    const int sb = base ^ swz_key(base);
    double2 w[8];
    p2_get_tw<N, 0>(w, base, tw, tws);
    double2 v[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[0][r] = sm[p2_elem<G::ls>(sb, r)];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if (a + 1 < A) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          v[(a + 1) & 1][r] = sm[(a + 1) * N + (p2_elem<G::ls>(sb, r) ^ (((a + 1) * AX) & 7))];
      }
      p2_apply_tw<R, +1>(v[a & 1], w);
      Butterfly<R, +1>::run(v[a & 1]);
#pragma unroll
      for (int r = 0; r < R; ++r) store(a, base, r << G::ls, v[a & 1][r]);
    }
  }
}

// Pass 0 with a block-wide item mapping: item = thread (base position = item index), so
// that 8 neighbouring lanes touch 8 consecutive elements = one full 128-byte line in
// global memory.  NOT warp-local: needs a __syncthreads towards the other group-A passes.
template <int N, int T, int A, int AX, typename Load>
__device__ __forceinline__ void p2_pass0_fwd_blk(double2 *sm, const double2 *__restrict__ tw, const double2 *tws,
                                                 Load load)
{
  constexpr int lr = P2<N>::lr(0), ls = P2<N>::ls(0), R = 1 << lr;
#pragma unroll 1
  for (int base = threadIdx.x; base < (N >> lr); base += T) {
    const int sb = base ^ swz_key(base);
    double2 v[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[0][r] = load(0, base, r << ls);
    if (A > 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[1][r] = load(1, base, r << ls);
    }
    double2 w[8];
    p2_get_tw<N, 0>(w, base, tw, tws);
#pragma unroll
    for (int a = 0; a < A; ++a) {
      Butterfly<R, -1>::run(v[a & 1]);
      p2_apply_tw<R, -1>(v[a & 1], w);
#pragma unroll
      for (int r = 0; r < R; ++r) sm[a * N + (p2_elem<ls>(sb, r) ^ ((a * AX) & 7))] = v[a & 1][r];
      if (a + 2 < A) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[a & 1][r] = load(a + 2, base, r << ls);
      }
    }
  }
}

template <int N, int T, int A, int AX, typename Store>
__device__ __forceinline__ void p2_pass0_inv_blk(const double2 *sm, const double2 *__restrict__ tw,
                                                 const double2 *tws, Store store)
{
  constexpr int lr = P2<N>::lr(0), ls = P2<N>::ls(0), R = 1 << lr;
#pragma unroll 1
  for (int base = threadIdx.x; base < (N >> lr); base += T) {
    const int sb = base ^ swz_key(base);
    double2 w[8];
    p2_get_tw<N, 0>(w, base, tw, tws);
    double2 v[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[0][r] = sm[p2_elem<ls>(sb, r)];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if (a + 1 < A) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[(a + 1) & 1][r] = sm[(a + 1) * N + (p2_elem<ls>(sb, r) ^ (((a + 1) * AX) & 7))];
      }
      p2_apply_tw<R, +1>(v[a & 1], w);
      Butterfly<R, +1>::run(v[a & 1]);
#pragma unroll
      for (int r = 0; r < R; ++r) store(a, base, r << ls, v[a & 1][r]);
    }
  }
}

// Last backward pass of one data set fused with pass 0 (forward) of the NEXT one, block-wide item
// mapping: per array the thread reads its R elements of the finished set from shared memory,
// issues the R global loads of the next set, finishes and stores the old elements while those
// loads are in flight, then transforms the new elements into the very same shared-memory slots.
// No barrier between the two uses (same thread, same slots).  The caller needs a __syncthreads
// before (towards the previous backward pass) and after (towards the next forward pass).
template <int N, int T, int A, int AX, typename Store, typename Load>
__device__ __forceinline__ void p2_pass0_inv_fwd_blk(double2 *sm, const double2 *__restrict__ tw,
                                                     const double2 *tws, Store store, Load load)
{
  constexpr int lr = P2<N>::lr(0), ls = P2<N>::ls(0), R = 1 << lr;
#pragma unroll 1
  for (int base = threadIdx.x; base < (N >> lr); base += T) {
    const int sb = base ^ swz_key(base);
    double2 w[8];
    p2_get_tw<N, 0>(w, base, tw, tws);
#pragma unroll
    for (int a = 0; a < A; ++a) {
      double2 v[R], nv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = sm[a * N + (p2_elem<ls>(sb, r) ^ ((a * AX) & 7))];
#pragma unroll
      for (int r = 0; r < R; ++r) nv[r] = load(a, base, r << ls);
      p2_apply_tw<R, +1>(v, w);
      Butterfly<R, +1>::run(v);
#pragma unroll
      for (int r = 0; r < R; ++r) store(a, base, r << ls, v[r]);
      Butterfly<R, -1>::run(nv);
      p2_apply_tw<R, -1>(nv, w);
#pragma unroll
      for (int r = 0; r < R; ++r) sm[a * N + (p2_elem<ls>(sb, r) ^ ((a * AX) & 7))] = nv[r];
    }
  }
}

// In-place group-A pass P with the block-wide item mapping (item j of N/R: the digit of
// pass P is spliced out of j).  Needs __syncthreads before and after.
template <int N, int P, int T, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupA_pass_blk(double2 *sm, const double2 *__restrict__ tw,
                                                   const double2 *tws)
{
  constexpr int lr = P2<N>::lr(P), ls = P2<N>::ls(P), R = 1 << lr;
#pragma unroll 1
  for (int j = threadIdx.x; j < (N >> lr); j += T) {
    const int base = ((j >> ls) << (ls + lr)) | (j & ((1 << ls) - 1));
    const int sb = base ^ swz_key(base);
    double2 w[8];
    p2_get_tw<N, P>(w, base & ((1 << ls) - 1), tw, tws);
    double2 v[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[0][r] = sm[p2_elem<ls>(sb, r)];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if (a + 1 < A) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[(a + 1) & 1][r] = sm[(a + 1) * N + (p2_elem<ls>(sb, r) ^ (((a + 1) * AX) & 7))];
      }
      if (DIR > 0) p2_apply_tw<R, DIR>(v[a & 1], w);
      Butterfly<R, DIR>::run(v[a & 1]);
      if (DIR < 0) p2_apply_tw<R, DIR>(v[a & 1], w);
#pragma unroll
      for (int r = 0; r < R; ++r) sm[a * N + (p2_elem<ls>(sb, r) ^ ((a * AX) & 7))] = v[a & 1][r];
    }
  }
}

// group-A passes 1 .. NA-1 (forward) / NA-1 .. 1 (backward) with block-wide mapping;
// starts and ends with a __syncthreads
template <int N, int T, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupA_rest_blk(double2 *sm, const double2 *__restrict__ tw,
                                                   const double2 *tws)
{
  constexpr int P2nd = P2<N>::NA > 2 ? 2 : 1;
  __syncthreads();
  if (DIR < 0) {
    p2_groupA_pass_blk<N, 1, T, -1, A, AX>(sm, tw, tws);
    __syncthreads();
    if (P2<N>::NA > 2) {
      p2_groupA_pass_blk<N, P2nd, T, -1, A, AX>(sm, tw, tws);
      __syncthreads();
    }
  } else {
    if (P2<N>::NA > 2) {
      p2_groupA_pass_blk<N, P2nd, T, +1, A, AX>(sm, tw, tws);
      __syncthreads();
    }
    p2_groupA_pass_blk<N, 1, T, +1, A, AX>(sm, tw, tws);
    __syncthreads();
  }
}

template <int N, int NW, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupA_rest_seq(double2 *sm, const double2 *__restrict__ tw,
                                                   const double2 *tws, int lane, int warp)
{
  constexpr int P2nd = P2<N>::NA > 2 ? 2 : 1;
  if (DIR < 0) {
    __syncwarp();
    p2_groupA_pass_seq<N, 1, NW, -1, A, AX>(sm, tw, tws, lane, warp);
    if (P2<N>::NA > 2) {
      __syncwarp();
      p2_groupA_pass_seq<N, P2nd, NW, -1, A, AX>(sm, tw, tws, lane, warp);
    }
  } else {
    if (P2<N>::NA > 2) {
      p2_groupA_pass_seq<N, P2nd, NW, +1, A, AX>(sm, tw, tws, lane, warp);
      __syncwarp();
    }
    p2_groupA_pass_seq<N, 1, NW, +1, A, AX>(sm, tw, tws, lane, warp);
    __syncwarp();
  }
}

template <int N, int DIR, int A, int AX>
__device__ __forceinline__ void p2_groupB_first_seq(double2 *sm, const double2 *tws, int idx)
{
  const int al = idx & 7, base = (idx >> 3) * 64 + al;
  const int sb = base ^ swz_key(base);
  double2 w[8];
#pragma unroll
  for (int q = 1; q < 8; ++q) w[q] = tws[P2<N>::TW8 + (q - 1) * 8 + al];
  double2 v[2][8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[0][r] = sm[p2_elem<3>(sb, r)];
#pragma unroll
  for (int a = 0; a < A; ++a) {
    if (a + 1 < A) {
#pragma unroll
      for (int r = 0; r < 8; ++r) v[(a + 1) & 1][r] = sm[(a + 1) * N + (p2_elem<3>(sb, r) ^ (((a + 1) * AX) & 7))];
    }
    if (DIR > 0) p2_apply_tw<8, DIR>(v[a & 1], w);
    Butterfly<8, DIR>::run(v[a & 1]);
    if (DIR < 0) p2_apply_tw<8, DIR>(v[a & 1], w);
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[a * N + (p2_elem<3>(sb, r) ^ ((a * AX) & 7))] = v[a & 1][r];
  }
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): two adjacent complex numbers, i.e. the
// same wavevector of two neighbouring rows in the transposed staging layout.  p must be
// 32-byte aligned.
__device__ __forceinline__ void st_global_256(double2 *p, double2 a, double2 b)
{
#ifdef GFMD_CUDA_EMU
  p[0] = a;
  p[1] = b;
#else
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
#endif
}

__device__ __forceinline__ void ld_global_256(const double2 *p, double2 &a, double2 &b)
{
#ifdef GFMD_CUDA_EMU
  a = p[0];
  b = p[1];
#else
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
#endif
}

}  // namespace gfmd
