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
// Shared-memory addressing is XOR-swizzled (swz): the 16-byte column of an
// element is XORed with a rotation of the XOR of all its higher 3-bit digits,
// which makes every access pattern used here (8 lanes varying one digit, or
// 4 consecutive elements x 2 values of a digit bit) bank-conflict free.
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

// twiddles of a forward pass: v[q] *= w^(q * n_lo); backward: conj, applied to inputs
template <int N, int P, int DIR>
__device__ __forceinline__ void p2_twiddle(double2 *v, int n_lo, const double2 *__restrict__ tw, const double2 *tws)
{
  constexpr int lr = P2<N>::lr(P), ls = P2<N>::ls(P), R = 1 << lr;
  if (ls == 0) return;
  if (ls == 6) {
#pragma unroll
    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], twid<DIR>(tws[P2<N>::TW64 + (q - 1) * 64 + n_lo]));
  } else if (ls == 3) {
#pragma unroll
    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], twid<DIR>(tws[P2<N>::TW8 + (q - 1) * 8 + n_lo]));
  } else {
    // long strides: one table load, powers in registers
    double2 w[8];
    tw_powers<R>(__ldg(tw + (n_lo << (P2<N>::LOG - ls - lr))), w);
#pragma unroll
    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], twid<DIR>(w[q]));
  }
}

// One in-place group-A pass over one array in shared memory (forward: butterfly
// then twiddle; backward: conj twiddle then inverse butterfly).
template <int N, int P, int NW, int DIR>
__device__ __forceinline__ void p2_groupA_pass(double2 *a, const double2 *__restrict__ tw, const double2 *tws,
                                               int lane, int warp)
{
  using G = GA<N, P, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    const int n_lo = base & ((1 << G::ls) - 1);
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = a[swz(base + (r << G::ls))];
    if (DIR > 0) p2_twiddle<N, P, DIR>(v, n_lo, tw, tws);
    Butterfly<R, DIR>::run(v);
    if (DIR < 0) p2_twiddle<N, P, DIR>(v, n_lo, tw, tws);
#pragma unroll
    for (int r = 0; r < R; ++r) a[swz(base + (r << G::ls))] = v[r];
  }
}

// Pass 0 forward with the inputs coming from global memory through `load(pos)`.
template <int N, int NW, typename Load>
__device__ __forceinline__ void p2_pass0_fwd(double2 *a, const double2 *__restrict__ tw, const double2 *tws,
                                             int lane, int warp, Load load)
{
  using G = GA<N, 0, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = load(base + (r << G::ls));
    Butterfly<R, -1>::run(v);
    p2_twiddle<N, 0, -1>(v, base, tw, tws);          // n_lo = base (digit 0 is the top digit)
#pragma unroll
    for (int r = 0; r < R; ++r) a[swz(base + (r << G::ls))] = v[r];
  }
}

// Pass 0 backward with the outputs going to global memory through `store(pos, value)`.
template <int N, int NW, typename Store>
__device__ __forceinline__ void p2_pass0_inv(const double2 *a, const double2 *__restrict__ tw, const double2 *tws,
                                             int lane, int warp, Store store)
{
  using G = GA<N, 0, NW>;
  constexpr int R = 1 << G::lr;
#pragma unroll 1
  for (int m = lane >> 2; m < G::M; m += 8) {
    const int base = G::base(m, lane, warp);
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = a[swz(base + (r << G::ls))];
    p2_twiddle<N, 0, +1>(v, base, tw, tws);
    Butterfly<R, +1>::run(v);
#pragma unroll
    for (int r = 0; r < R; ++r) store(base + (r << G::ls), v[r]);
  }
}

// remaining group-A passes 1 .. NA-1 (forward) / NA-1 .. 1 (backward), warp-synchronous
template <int N, int NW, int DIR>
__device__ __forceinline__ void p2_groupA_rest(double2 *a, const double2 *__restrict__ tw, const double2 *tws,
                                               int lane, int warp)
{
  if (DIR < 0) {
    __syncwarp();
    p2_groupA_pass<N, 1, NW, -1>(a, tw, tws, lane, warp);
    if (P2<N>::NA > 2) {
      __syncwarp();
      p2_groupA_pass<N, (P2<N>::NA > 2 ? 2 : 1), NW, -1>(a, tw, tws, lane, warp);
    }
  } else {
    if (P2<N>::NA > 2) {
      p2_groupA_pass<N, (P2<N>::NA > 2 ? 2 : 1), NW, +1>(a, tw, tws, lane, warp);
      __syncwarp();
    }
    p2_groupA_pass<N, 1, NW, +1>(a, tw, tws, lane, warp);
    __syncwarp();
  }
}

// ------------------------------------------------------------ group B items ---
// item idx in [0, N/8): block of 64 = idx >> 3, a = idx & 7 (the same 8 lanes own a
// block in both passes).

// forward pass NP-2 (stride 8) in place
template <int N> __device__ __forceinline__ void p2_groupB_first_fwd(double2 *a, const double2 *tws, int idx)
{
  const int al = idx & 7, base = (idx >> 3) * 64 + al;
  double2 v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = a[swz(base + 8 * r)];
  Butterfly<8, -1>::run(v);
#pragma unroll
  for (int q = 1; q < 8; ++q) v[q] = cmul(v[q], tws[P2<N>::TW8 + (q - 1) * 8 + al]);
#pragma unroll
  for (int r = 0; r < 8; ++r) a[swz(base + 8 * r)] = v[r];
}

// backward pass NP-2 in place
template <int N> __device__ __forceinline__ void p2_groupB_first_inv(double2 *a, const double2 *tws, int idx)
{
  const int al = idx & 7, base = (idx >> 3) * 64 + al;
  double2 v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = a[swz(base + 8 * r)];
#pragma unroll
  for (int q = 1; q < 8; ++q) v[q] = cmulc(v[q], tws[P2<N>::TW8 + (q - 1) * 8 + al]);
  Butterfly<8, +1>::run(v);
#pragma unroll
  for (int r = 0; r < 8; ++r) a[swz(base + 8 * r)] = v[r];
}

// forward last pass (stride 1): result in registers, positions (idx>>3)*64 + 8*(idx&7) + r
__device__ __forceinline__ void p2_last_fwd_load(const double2 *a, int idx, double2 *v)
{
  const int base = (idx >> 3) * 64 + 8 * (idx & 7);
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = a[swz(base + r)];
  Butterfly<8, -1>::run(v);
}

__device__ __forceinline__ void p2_last_inv_store(double2 *a, int idx, double2 *v)
{
  const int base = (idx >> 3) * 64 + 8 * (idx & 7);
  Butterfly<8, +1>::run(v);
#pragma unroll
  for (int r = 0; r < 8; ++r) a[swz(base + r)] = v[r];
}

}  // namespace gfmd
