// Row kernels with radix-16 passes (variant ids ny + 8 of GFMD_B200_ROWS_VARIANT; the default since round 2):
// the same transforms as k_rows_fwd_p2 / k_rows_inv_p2 (kernels_fast.cuh) -- a real row of
// ny = 2 NR values as NR packed complex numbers, half-length FFT, real/complex (un)mixing,
// transposed access to the staging buffer -- with FOUR shared-memory sweeps over the tile
// instead of about ten, because that is what bounds the radix-8 kernels (profiles/README.md:
// l1tex 82 %, DRAM 34 %).
//
//   NR = 16 * 16 * 8 (ny = 4096).  Forward, decimation in frequency, frequency
//   k = q0 + 16 q1 + 256 q2:
//     pass 1  thread t < NR/16:  z[t + (NR/16) j], j < 16, straight from global memory ->
//             16-point DFT -> times w_NR^(q0 t) -> shared Y[q0][t]                       (1 write)
//     pass 2  thread (q0, t1 < 8): Y[q0][t1 + 8 j] -> 16-point DFT -> times w_128^(q1 t1)
//             -> shared Y[q0][q1][t1]                                                    (1 read, 1 write)
//     pass 3  one thread runs the 8-point DFTs of BOTH groups klow = q0 + 16 q1 and 256 - klow:
//             output q2 of one pairs with output 7 - q2 of the other (k <-> NR - k), so the
//             real/complex un-mixing happens in registers and X goes straight to the staging
//             buffer (1 read); groups 0 and 128 pair with themselves and form one unit.
//   Backward: the transposed flow (decimation in time, conjugate twiddles).
//
// Shared memory is [row][q0][q1][t1 ^ s], s = (q1 & 3) | (row & 1) << 2: passes 1 and 2 touch 8
// consecutive 16-byte words per quarter warp; in pass 3 the quarter warp is 4 units x 2 rows whose
// groups differ in q1 & 3 (unit -> group map r16_klow), so every access is conflict-free.
// Twiddles: one table load per thread and pass, powers by squaring (depth 4); the un-mixing
// twiddles in closed form (RowsFuseTw<NR, 2>).  Results agree with the default kernels to
// rounding (not bit for bit: other radix, other twiddle roundings).
//
// Reference: GFMDSolverFFT::fft_forward / fft_reverse, y part
// (src/solvers/gfmd_solver_fft.cpp:96-147, :150-195).
//
// Emulator-verified (tests/test_emulated_kernels.py) and measured on B200s (profiles/r2_rows_variants.txt):
// 4096 x 4096 rows_fwd / rows_inv 0.192 / 0.218 ms against 0.272 / 0.263 ms of the radix-8 kernels; with the
// L2 prefetch of the next tile (pf) 0.181 / 0.200 ms = 0.69 / 0.62 of the measured HBM rate.
//
// Included by kernels_fast.cuh after the helpers it shares with the FUSE variants
// (rows_unmix, rows_premix, RowsFuseTw).
#pragma once

namespace gfmd {

// 16-point DFT, in place.  Output X[q] is left at v[4 * (q & 3) + (q >> 2)] (see r16_out).
template <int DIR> __device__ __forceinline__ void dft16(double2 *v)
{
  constexpr double c1 = 0.92387953251128673848, s1 = 0.38268343236508978178, h = 0.70710678118654752440;
  // columns: j = j0 + 4 j1 -> y[j0][q1] at v[j0 + 4 q1]
#pragma unroll
  for (int j0 = 0; j0 < 4; ++j0) bf4<DIR>(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12]);
  // twiddles w16^(j0 q1), w16 = exp(DIR * 2 pi i / 16)
  auto rot = [&](double2 a, double c, double s) {      // a * (c + i DIR s)
    return DIR < 0 ? make_double2(fma(a.x, c, a.y * s), fma(a.y, c, -(a.x * s)))
                   : make_double2(fma(a.x, c, -(a.y * s)), fma(a.y, c, a.x * s));
  };
  v[1 + 4 * 1] = rot(v[1 + 4 * 1], c1, s1);            // w^1
  v[2 + 4 * 1] = rot(v[2 + 4 * 1], h, h);              // w^2
  v[3 + 4 * 1] = rot(v[3 + 4 * 1], s1, c1);            // w^3
  v[1 + 4 * 2] = rot(v[1 + 4 * 2], h, h);              // w^2
  v[2 + 4 * 2] = mul_mi<DIR>(v[2 + 4 * 2]);            // w^4
  v[3 + 4 * 2] = rot(v[3 + 4 * 2], -h, h);             // w^6
  v[1 + 4 * 3] = rot(v[1 + 4 * 3], s1, c1);            // w^3
  v[2 + 4 * 3] = rot(v[2 + 4 * 3], -h, h);             // w^6
  v[3 + 4 * 3] = rot(v[3 + 4 * 3], -c1, -s1);          // w^9
  // rows: X[q1 + 4 q0] = sum_j0 y[j0][q1] w4^(j0 q0), left at v[4 q1 + q0]
#pragma unroll
  for (int q1 = 0; q1 < 4; ++q1) bf4<DIR>(v[4 * q1], v[4 * q1 + 1], v[4 * q1 + 2], v[4 * q1 + 3]);
}
__host__ __device__ constexpr int r16_out(int q) { return 4 * (q & 3) + (q >> 2); }

// v[r16_out(q)] *= w1^q (DIR < 0) or conj(w1)^q, q = 1..15; powers by squaring, depth 4
template <int DIR, bool OUT_ORDER> __device__ __forceinline__ void r16_twiddle(double2 *v, double2 w1)
{
  double2 w[16];
  w[1] = w1;
  w[2] = csqr(w[1]);
  w[3] = cmul(w[2], w[1]);
  w[4] = csqr(w[2]);
  w[5] = cmul(w[4], w[1]);
  w[6] = csqr(w[3]);
  w[7] = cmul(w[4], w[3]);
  w[8] = csqr(w[4]);
#pragma unroll
  for (int q = 9; q < 16; ++q) w[q] = cmul(w[8], w[q - 8]);
#pragma unroll
  for (int q = 1; q < 16; ++q) {
    const int i = OUT_ORDER ? r16_out(q) : q;
    v[i] = DIR < 0 ? cmul(v[i], w[q]) : cmulc(v[i], w[q]);
  }
}


// ---- fused atom I/O of the row kernels (MAPPED variants) -----------------------------------------
// gather o forward rows and inverse rows o scatter in ONE kernel each, through a cell -> atom map
// (gfmd_b200_build_cell_map; valid only if every cell of the brick holds exactly one atom of the
// group, otherwise the library keeps the separate k_gather / k_scatter kernels): the displacement
// grid u_xy and the force grid f_xy never exist in memory.  Same arithmetic per element as
// k_gather / k_scatter (kernels_generic.cuh), i.e. FixGFMD::pre_force list -> grid
// (src/main/fix_gfmd.cpp:734-803: u = x - xeq, x / y minimum image) and grid_to_list + f += f_i
// (:952-1010, :896-902; the force sum runs over atoms i < nlocal only, :997-1001).
struct AtomIO {
  const double *x, *xeq;      // [nall][3]
  double *fat;                // [nall][3], forces accumulate (+=)
  const int *cmap;            // [nu][nx_loc * ny] atom index of every cell
  double *fsum_part;          // [ntiles][ndof] per-CTA partial sums of the scattered forces
  double xprd, yprd;
  int nlocal;
};

// packed complex element e of row ixl of dof: (u[2e], u[2e + 1])
__device__ __forceinline__ double2 atomio_load(const AtomIO &io, const GridDesc &g, int dof, int ixl, int e)
{
  const int iu = dof / 3, c = dof - 3 * iu;
  const int2 a = __ldg(reinterpret_cast<const int2 *>(io.cmap + ((size_t) iu * g.nx_loc + ixl) * g.ny) + e);
  double u0 = __ldg(io.x + 3 * (size_t) a.x + c) - __ldg(io.xeq + 3 * (size_t) a.x + c);
  double u1 = __ldg(io.x + 3 * (size_t) a.y + c) - __ldg(io.xeq + 3 * (size_t) a.y + c);
  if (c < 2) {                                   // x, y: minimum image (z is not wrapped, fix_gfmd.cpp:758)
    const double prd = c == 0 ? io.xprd : io.yprd, half = 0.5 * prd;
    while (u0 > half) u0 -= prd;
    while (u0 < -half) u0 += prd;
    while (u1 > half) u1 -= prd;
    while (u1 < -half) u1 += prd;
  }
  return make_double2(u0, u1);
}

// f[atom] += grid force; returns this pair's contribution to the force sum over local atoms
__device__ __forceinline__ double atomio_store(const AtomIO &io, const GridDesc &g, int dof, int ixl, int e, double2 v)
{
  const int iu = dof / 3, c = dof - 3 * iu;
  const int2 a = __ldg(reinterpret_cast<const int2 *>(io.cmap + ((size_t) iu * g.nx_loc + ixl) * g.ny) + e);
  double *f0 = io.fat + 3 * (size_t) a.x + c, *f1 = io.fat + 3 * (size_t) a.y + c;
  *f0 += v.x;
  *f1 += v.y;
  return (a.x < io.nlocal ? v.x : 0.0) + (a.y < io.nlocal ? v.y : 0.0);
}

// fixed-order block sum of `acc`, written to *dst by thread 0 (T threads, all must call)
template <int T> __device__ __forceinline__ void atomio_block_sum(double acc, double *dst)
{
  __shared__ double red[T / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < T / 32; ++k) a += red[k];
    *dst = a;
  }
}

// block -> (dof, first row): MAPPED kernels put the dofs of one tile in neighbouring CTAs, which run
// at the same time and share the atoms' cache lines in L2 (x, xeq and f are [atom][3])
template <bool MAPPED> __device__ __forceinline__ void rows_block_map(int ntiles, int dof0, int &dof, int &tile)
{
  if (MAPPED) {
    const int nd = gridDim.x / ntiles;
    dof = dof0 + blockIdx.x % nd;
    tile = blockIdx.x / nd;
  } else {
    dof = dof0 + blockIdx.x / ntiles;
    tile = blockIdx.x % ntiles;
  }
}

// unit p in [0, 128) -> klow in [0, 128), 0 -> 0: the four units of a quarter warp take four
// values of q1 & 3 (q1 = klow >> 4); their partner groups 256 - klow then do too
__device__ __forceinline__ int r16_klow(int p) { return ((p >> 2) & 15) | ((p & 3) << 4) | ((p >> 6) << 6); }

// shared-memory index of element t1 of group (q0, q1) = klow of row r (rows are NR apart)
__device__ __forceinline__ int r16_slot(int klow, int t1, int r)
{
  const int q1 = klow >> 4;
  return ((klow & 15) << 7) + (q1 << 3) + (t1 ^ ((q1 & 3) | ((r & 1) << 2)));
}

template <int NR, int RB, int T, bool MAPPED = false>
__global__ void __launch_bounds__(T, 2)
k_rows_fwd_r16(const double *__restrict__ u, double2 *__restrict__ stage, GridDesc g,
               const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0, AtomIO io = AtomIO(), int pf = 0)
{
  static_assert(NR == 2048 && RB * (NR / 16) == T, "k_rows_fwd_r16: NR = 16 * 16 * 8, one pass-1 item per thread");
  constexpr int M1 = NR / 16;                 // 128: pass-1 threads per row = length of a pass-2 block
  constexpr int S = NR / 8;                   // 256: frequency stride of the last pass
  extern __shared__ double2 sm[];
  int dof, tile;
  rows_block_map<MAPPED>(g.nx_loc / RB, dof0, dof, tile);
  const int ix0 = tile * RB;
  const int r = threadIdx.x / M1, m = threadIdx.x % M1;
  double2 *row = sm + r * NR;

  // ---- pass 1: 16 independent 16-byte loads per thread, coalesced over the threads of a row
  {
    double2 v[16];
    if constexpr (MAPPED) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = atomio_load(io, g, dof, ix0 + r, m + M1 * j);
    } else {
      const double2 *src = reinterpret_cast<const double2 *>(u + ((size_t) dof * g.nx_loc + ix0 + r) * (2 * NR));
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = src[m + M1 * j];
      // L2 prefetch of the tile of the CTA that runs `pf` blocks later (behind this tile's own loads)
      if (pf > 0 && blockIdx.x + pf < gridDim.x) {
        const int b = blockIdx.x + pf, nt = g.nx_loc / RB;
        const char *pb = reinterpret_cast<const char *>(u + ((size_t) (dof0 + b / nt) * g.nx_loc + (b % nt) * RB) * (2 * NR));
        for (int i = threadIdx.x; i < RB * 2 * NR * 8 / 128; i += T) prefetch_l2(pb + (size_t) i * 128);
      }
    }
    dft16<-1>(v);
    r16_twiddle<-1, true>(v, __ldg(tw + m));
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) row[q0 * M1 + m] = v[r16_out(q0)];
  }
  __syncthreads();

  // ---- pass 2: thread (q0, t1); the 8 threads of a q0 block sit in one quarter warp
  {
    const int q0 = m >> 3, t1 = m & 7;
    double2 *blk = row + q0 * M1;
    double2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = blk[t1 + 8 * j];
    dft16<-1>(v);
    r16_twiddle<-1, true>(v, __ldg(tw + 16 * t1));
    __syncwarp();                             // the swizzled slots below belong to the other threads of the block
#pragma unroll
    for (int q1 = 0; q1 < 16; ++q1) blk[(q1 << 3) + (t1 ^ ((q1 & 3) | ((r & 1) << 2)))] = v[r16_out(q1)];
  }
  __syncthreads();

  // ---- pass 3 + un-mixing + transposed store: adjacent lanes = adjacent rows (whole 32-byte sectors)
  {
    const int rr = threadIdx.x % RB, p = threadIdx.x / RB;
    const double2 *rw = sm + rr * NR;
    const int klow = r16_klow(p);
    const int klow2 = p == 0 ? S / 2 : S - klow;
    double2 v1[8], v2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v1[e] = rw[r16_slot(klow, e, rr)];
      v2[e] = rw[r16_slot(klow2, e, rr)];
    }
    Butterfly<8, -1>::run(v1);
    Butterfly<8, -1>::run(v2);
    const RowsFuseTw<NR, 2> w(tw_ny, klow, klow2, p == 0);
    if (p == 0) {
      // groups 0 and S/2 pair with themselves: k = q S <-> (8 - q) S, and S/2 + q S <-> S/2 + (7 - q) S
      stage[stage_index(g, 0, dof, ix0 + rr)] = rows_unmix(v1[0], v1[0], w.get(0, 0));
      stage[stage_index(g, NR, dof, ix0 + rr)] = rows_unmix(v1[0], v1[0], w.nyquist());
#pragma unroll
      for (int q = 1; q < 8; ++q)
        stage[stage_index(g, q * S, dof, ix0 + rr)] = rows_unmix(v1[q], v1[8 - q], w.get(0, q));
#pragma unroll
      for (int q = 0; q < 8; ++q)
        stage[stage_index(g, S / 2 + q * S, dof, ix0 + rr)] = rows_unmix(v2[q], v2[7 - q], w.get(1, q));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        stage[stage_index(g, klow + q * S, dof, ix0 + rr)] = rows_unmix(v1[q], v2[7 - q], w.get(0, q));
        stage[stage_index(g, klow2 + q * S, dof, ix0 + rr)] = rows_unmix(v2[q], v1[7 - q], w.get(1, q));
      }
    }
  }
}

template <int NR, int RB, int T, bool MAPPED = false>
__global__ void __launch_bounds__(T, 2)
k_rows_inv_r16(const double2 *__restrict__ stage, double *__restrict__ f, GridDesc g,
               const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0, AtomIO io = AtomIO(), int pf = 0)
{
  static_assert(NR == 2048 && RB * (NR / 16) == T, "k_rows_inv_r16: NR = 16 * 16 * 8, one pass-1 item per thread");
  constexpr int M1 = NR / 16;
  constexpr int S = NR / 8;
  extern __shared__ double2 sm[];
  int dof, tile;
  rows_block_map<MAPPED>(g.nx_loc / RB, dof0, dof, tile);
  const int ix0 = tile * RB;

  // ---- transposed load of both groups of a unit (16 independent loads), pre-mix, inverse 8-point DFTs
  {
    const int rr = threadIdx.x % RB, p = threadIdx.x / RB;
    double2 *rw = sm + rr * NR;
    const int klow = r16_klow(p);
    const int klow2 = p == 0 ? S / 2 : S - klow;
    double2 y1[8], y2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      y1[q] = stage[stage_index(g, klow + q * S, dof, ix0 + rr)];
      y2[q] = stage[stage_index(g, klow2 + q * S, dof, ix0 + rr)];
    }
    if (!MAPPED && pf > 0 && blockIdx.x + pf < gridDim.x && rr == 0) {      // one lane per 32-byte sector
      const int b = blockIdx.x + pf, nt = g.nx_loc / RB;
      const int pdof = dof0 + b / nt, pix = (b % nt) * RB;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        prefetch_l2(stage + stage_index(g, klow + q * S, pdof, pix));
        prefetch_l2(stage + stage_index(g, klow2 + q * S, pdof, pix));
      }
    }
    const RowsFuseTw<NR, 2> w(tw_ny, klow, klow2, p == 0);
    if (p == 0) {
      const double2 yh0 = stage[stage_index(g, NR, dof, ix0 + rr)];
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
      // the pair (y1[q], y2[7-q]) yields (z1[q], z2[7-q]): in place
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
      rw[r16_slot(klow, e, rr)] = y1[e];
      rw[r16_slot(klow2, e, rr)] = y2[e];
    }
  }
  __syncthreads();

  const int r = threadIdx.x / M1, m = threadIdx.x % M1;
  double2 *row = sm + r * NR;
  // ---- pass 2 backward: conjugate twiddles, inverse 16-point DFT over q1
  {
    const int q0 = m >> 3, t1 = m & 7;
    double2 *blk = row + q0 * M1;
    double2 v[16];
#pragma unroll
    for (int q1 = 0; q1 < 16; ++q1) v[q1] = blk[(q1 << 3) + (t1 ^ ((q1 & 3) | ((r & 1) << 2)))];
    r16_twiddle<+1, false>(v, __ldg(tw + 16 * t1));
    dft16<+1>(v);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) blk[t1 + 8 * j] = v[r16_out(j)];
  }
  __syncthreads();

  // ---- pass 1 backward, straight to global memory: packed complex = pairs of reals
  {
    double2 v[16];
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) v[q0] = row[q0 * M1 + m];
    r16_twiddle<+1, false>(v, __ldg(tw + m));
    dft16<+1>(v);
    if constexpr (MAPPED) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += atomio_store(io, g, dof, ix0 + r, m + M1 * j, v[r16_out(j)]);
      atomio_block_sum<T>(acc, io.fsum_part + (size_t) tile * g.d + dof);
    } else {
      double2 *dst = reinterpret_cast<double2 *>(f + ((size_t) dof * g.nx_loc + ix0 + r) * (2 * NR));
#pragma unroll
      for (int j = 0; j < 16; ++j) dst[m + M1 * j] = v[r16_out(j)];
    }
  }
}

// ------------------------------------------------------------------ NR = 4096 ---
// NR = 16 * 16 * 16 (ny = 8192).  Passes 1 and 2 as above (256 threads per row, blocks of 256,
// groups of 16).  A last radix-16 pass with both groups of a pair in one thread would need 32
// complex registers, so the pairs are cut by output parity instead ("half units"): of group
// g = q0 + 16 q1 (frequencies g + 256 m, m < 16) the outputs m = e + 2 m' are the 8-point DFT of
//   c[t] = (y[t] + (-1)^e y[t + 8]) w16^(t e),  t < 8,
// i.e. frequencies (g + 256 e) + 512 m'.  A thread computes parity e of group klow and parity
// 1 - e of group 256 - klow: exactly the frequencies kA + 512 m' and (512 - kA) + 512 m'' of ONE
// fused unit of the radix-8 kernels (S = 512), so un-mixing and stores are the same code.  Groups
// 0 and 128 pair with themselves; two threads take both parities of one of them each.  Cost: the
// last pass reads the tile twice -- five sweeps instead of four (the radix-8 kernels: ~twelve).
// Backward, each thread leaves d_e[t] = c'_e[t] conj(w16^(t e)) in slot t + 8 e of its groups and
// pass 2 forms y[t1] = d_0[t1 & 7] +- d_1[t1 & 7] while loading.

__device__ __forceinline__ int r16h_slot(int grp, int t, int r)
{
  const int q1 = grp >> 4;
  return ((grp & 15) << 8) + (q1 << 4) + (t ^ ((q1 & 3) | ((r & 1) << 2)));
}

// thread p of a row -> (group, parity) of its two halves and the unit's base frequencies
struct R16HalfUnit {
  int gA, eA, gB, eB, kA, kB;
  bool special;
  __device__ __forceinline__ R16HalfUnit(int p)
  {
    const int e = p >> 7, pp = p & 127;
    const int klow = r16_klow(pp);
    special = p == 0;
    if (pp == 0) {              // groups 0 (p = 0) and 128 (p = 128): both parities of one group
      gA = gB = e ? 128 : 0;
      eA = 0;
    } else {
      gA = klow;
      gB = 256 - klow;
      eA = e;
    }
    eB = 1 - eA;
    kA = gA + 256 * eA;
    kB = gB + 256 * eB;         // = 512 - kA, or 256 for p = 0
  }
};

// 8 outputs of parity e of one group: loads 16 values, combines, 8-point DFT
__device__ __forceinline__ void r16h_half_fwd(const double2 *rw, int grp, int e, int rr, double2 *c)
{
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const double2 lo = rw[r16h_slot(grp, t, rr)], hi = rw[r16h_slot(grp, t + 8, rr)];
    c[t] = e ? cmul(csub(lo, hi), rot16(t)) : cadd(lo, hi);
  }
  Butterfly<8, -1>::run(c);
}

// transposed step: inverse 8-point DFT, conjugate twiddle, slot t + 8 e of the group
__device__ __forceinline__ void r16h_half_inv(double2 *rw, int grp, int e, int rr, double2 *z)
{
  Butterfly<8, +1>::run(z);
#pragma unroll
  for (int t = 0; t < 8; ++t) rw[r16h_slot(grp, t + 8 * e, rr)] = e ? cmulc(z[t], rot16(t)) : z[t];
}

template <int NR, int RB, int T, bool MAPPED = false>
__global__ void __launch_bounds__(T, 1)
k_rows_fwd_r16h(const double *__restrict__ u, double2 *__restrict__ stage, GridDesc g,
                const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0, AtomIO io = AtomIO())
{
  static_assert(NR == 4096 && RB * (NR / 16) == T, "k_rows_fwd_r16h: NR = 16 * 16 * 16, one pass-1 item per thread");
  constexpr int M1 = NR / 16;                 // 256
  constexpr int S = NR / 8;                   // 512: frequency stride of a unit's outputs
  extern __shared__ double2 sm[];
  int dof, tile;
  rows_block_map<MAPPED>(g.nx_loc / RB, dof0, dof, tile);
  const int ix0 = tile * RB;
  const int r = threadIdx.x / M1, m = threadIdx.x % M1;
  double2 *row = sm + r * NR;

  {
    double2 v[16];
    if constexpr (MAPPED) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = atomio_load(io, g, dof, ix0 + r, m + M1 * j);
    } else {
      const double2 *src = reinterpret_cast<const double2 *>(u + ((size_t) dof * g.nx_loc + ix0 + r) * (2 * NR));
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = src[m + M1 * j];
    }
    dft16<-1>(v);
    r16_twiddle<-1, true>(v, __ldg(tw + m));
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) row[q0 * M1 + m] = v[r16_out(q0)];
  }
  __syncthreads();

  {
    const int q0 = m >> 4, t1 = m & 15;      // the 16 threads of a q0 block sit in one warp
    double2 *blk = row + q0 * M1;
    double2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = blk[t1 + 16 * j];
    dft16<-1>(v);
    r16_twiddle<-1, true>(v, __ldg(tw + 16 * t1));
    __syncwarp();
#pragma unroll
    for (int q1 = 0; q1 < 16; ++q1) blk[(q1 << 4) + (t1 ^ ((q1 & 3) | ((r & 1) << 2)))] = v[r16_out(q1)];
  }
  __syncthreads();

  {
    const int rr = threadIdx.x % RB;
    const R16HalfUnit hu(threadIdx.x / RB);
    const double2 *rw = sm + rr * NR;
    double2 v1[8], v2[8];
    r16h_half_fwd(rw, hu.gA, hu.eA, rr, v1);
    r16h_half_fwd(rw, hu.gB, hu.eB, rr, v2);
    const RowsFuseTw<NR, 2> w(tw_ny, hu.kA, hu.kB, hu.special);
    if (hu.special) {
      stage[stage_index(g, 0, dof, ix0 + rr)] = rows_unmix(v1[0], v1[0], w.get(0, 0));
      stage[stage_index(g, NR, dof, ix0 + rr)] = rows_unmix(v1[0], v1[0], w.nyquist());
#pragma unroll
      for (int q = 1; q < 8; ++q)
        stage[stage_index(g, q * S, dof, ix0 + rr)] = rows_unmix(v1[q], v1[8 - q], w.get(0, q));
#pragma unroll
      for (int q = 0; q < 8; ++q)
        stage[stage_index(g, S / 2 + q * S, dof, ix0 + rr)] = rows_unmix(v2[q], v2[7 - q], w.get(1, q));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        stage[stage_index(g, hu.kA + q * S, dof, ix0 + rr)] = rows_unmix(v1[q], v2[7 - q], w.get(0, q));
        stage[stage_index(g, hu.kB + q * S, dof, ix0 + rr)] = rows_unmix(v2[q], v1[7 - q], w.get(1, q));
      }
    }
  }
}

template <int NR, int RB, int T, bool MAPPED = false>
__global__ void __launch_bounds__(T, 1)
k_rows_inv_r16h(const double2 *__restrict__ stage, double *__restrict__ f, GridDesc g,
                const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0, AtomIO io = AtomIO())
{
  static_assert(NR == 4096 && RB * (NR / 16) == T, "k_rows_inv_r16h: NR = 16 * 16 * 16, one pass-1 item per thread");
  constexpr int M1 = NR / 16;
  constexpr int S = NR / 8;
  extern __shared__ double2 sm[];
  int dof, tile;
  rows_block_map<MAPPED>(g.nx_loc / RB, dof0, dof, tile);
  const int ix0 = tile * RB;

  {
    const int rr = threadIdx.x % RB;
    const R16HalfUnit hu(threadIdx.x / RB);
    double2 *rw = sm + rr * NR;
    double2 y1[8], y2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      y1[q] = stage[stage_index(g, hu.kA + q * S, dof, ix0 + rr)];
      y2[q] = stage[stage_index(g, hu.kB + q * S, dof, ix0 + rr)];
    }
    const RowsFuseTw<NR, 2> w(tw_ny, hu.kA, hu.kB, hu.special);
    if (hu.special) {
      const double2 yh0 = stage[stage_index(g, NR, dof, ix0 + rr)];
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
    r16h_half_inv(rw, hu.gA, hu.eA, rr, y1);
    r16h_half_inv(rw, hu.gB, hu.eB, rr, y2);
  }
  __syncthreads();

  const int r = threadIdx.x / M1, m = threadIdx.x % M1;
  double2 *row = sm + r * NR;
  {
    const int q0 = m >> 4, t1 = m & 15;
    double2 *blk = row + q0 * M1;
    double2 v[16];
#pragma unroll
    for (int q1 = 0; q1 < 16; ++q1) {
      const int sl = (q1 << 4) + ((t1 & 7) ^ ((q1 & 3) | ((r & 1) << 2)));
      const double2 d0 = blk[sl], d1 = blk[sl + 8];
      v[q1] = (t1 & 8) ? csub(d0, d1) : cadd(d0, d1);
    }
    r16_twiddle<+1, false>(v, __ldg(tw + 16 * t1));
    dft16<+1>(v);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) blk[t1 + 16 * j] = v[r16_out(j)];
  }
  __syncthreads();

  {
    double2 v[16];
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) v[q0] = row[q0 * M1 + m];
    r16_twiddle<+1, false>(v, __ldg(tw + m));
    dft16<+1>(v);
    if constexpr (MAPPED) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += atomio_store(io, g, dof, ix0 + r, m + M1 * j, v[r16_out(j)]);
      atomio_block_sum<T>(acc, io.fsum_part + (size_t) tile * g.d + dof);
    } else {
      double2 *dst = reinterpret_cast<double2 *>(f + ((size_t) dof * g.nx_loc + ix0 + r) * (2 * NR));
#pragma unroll
      for (int j = 0; j < 16; ++j) dst[m + M1 * j] = v[r16_out(j)];
    }
  }
}

// ------------------------------------------------------------------ NR = 8192 ---
// NR = 16 * 8 * 8 * 8 (ny = 16384, one row per CTA: 128 KB).  Pass 1 as above (512 threads); two
// radix-8 passes in shared memory, two items per thread (blocks of 512: in place; blocks of 64: into
// the swizzled group layout); the last radix-8 pass of both partner groups klow and 1024 - klow in
// one thread, fused with the (un)mixing.  W, RW, RW, R = six sweeps (the radix-8 kernel: ~twelve).
// Shared memory is [q0][q1][q2][t ^ s], s = (q2 & 3) | (q1 & 1) << 2; the unit -> group map r16w_klow
// gives the eight units of a quarter warp eight different s, and their partner groups too.

__device__ __forceinline__ int r16w_slot(int grp, int t)
{
  const int q1 = (grp >> 4) & 7, q2 = grp >> 7;
  return ((grp & 15) << 9) + (q1 << 6) + (q2 << 3) + (t ^ ((q2 & 3) | ((q1 & 1) << 2)));
}

// unit p in [0, 512) -> klow in [0, 512), 0 -> 0: q2 = p & 3, q1 = (p >> 7) << 1 | (p >> 2) & 1, q0 = (p >> 3) & 15
__device__ __forceinline__ int r16w_klow(int p)
{
  const int q1 = ((p >> 7) << 1) | ((p >> 2) & 1);
  return ((p >> 3) & 15) | (q1 << 4) | ((p & 3) << 7);
}

template <int NR, int T, bool MAPPED = false>
__global__ void __launch_bounds__(T, 1)
k_rows_fwd_r16w(const double *__restrict__ u, double2 *__restrict__ stage, GridDesc g,
                const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0, AtomIO io = AtomIO(), int pf = 0)
{
  static_assert(NR == 8192 && NR / 16 == T, "k_rows_fwd_r16w: NR = 16 * 8 * 8 * 8, one row per CTA");
  constexpr int S = NR / 8;                   // 1024: frequency stride of the last pass
  extern __shared__ double2 sm[];
  int dof, ix;
  rows_block_map<MAPPED>(g.nx_loc, dof0, dof, ix);
  const int t = threadIdx.x;
  {
    double2 v[16];
    if constexpr (MAPPED) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = atomio_load(io, g, dof, ix, t + T * j);
    } else {
      const double2 *src = reinterpret_cast<const double2 *>(u + ((size_t) dof * g.nx_loc + ix) * (2 * NR));
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = src[t + T * j];
      if (pf > 0 && blockIdx.x + pf < gridDim.x) {
        const int b = blockIdx.x + pf;
        const char *pb = reinterpret_cast<const char *>(u + ((size_t) (dof0 + b / g.nx_loc) * g.nx_loc + b % g.nx_loc) * (2 * NR));
        for (int i = t; i < 2 * NR * 8 / 128; i += T) prefetch_l2(pb + (size_t) i * 128);
      }
    }
    dft16<-1>(v);
    r16_twiddle<-1, true>(v, __ldg(tw + t));
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) sm[q0 * T + t] = v[r16_out(q0)];
  }
  __syncthreads();
  // blocks of 512 = 8 x 64, in place
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
  // blocks of 64 = 8 x 8, into the swizzled group layout (the 8 threads of a block share a quarter warp)
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
    if (p == 0) {
      stage[stage_index(g, 0, dof, ix)] = rows_unmix(v1[0], v1[0], w.get(0, 0));
      stage[stage_index(g, NR, dof, ix)] = rows_unmix(v1[0], v1[0], w.nyquist());
#pragma unroll
      for (int q = 1; q < 8; ++q) stage[stage_index(g, q * S, dof, ix)] = rows_unmix(v1[q], v1[8 - q], w.get(0, q));
#pragma unroll
      for (int q = 0; q < 8; ++q)
        stage[stage_index(g, S / 2 + q * S, dof, ix)] = rows_unmix(v2[q], v2[7 - q], w.get(1, q));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        stage[stage_index(g, klow + q * S, dof, ix)] = rows_unmix(v1[q], v2[7 - q], w.get(0, q));
        stage[stage_index(g, klow2 + q * S, dof, ix)] = rows_unmix(v2[q], v1[7 - q], w.get(1, q));
      }
    }
  }
}

template <int NR, int T, bool MAPPED = false>
__global__ void __launch_bounds__(T, 1)
k_rows_inv_r16w(const double2 *__restrict__ stage, double *__restrict__ f, GridDesc g,
                const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0, AtomIO io = AtomIO(), int pf = 0)
{
  static_assert(NR == 8192 && NR / 16 == T, "k_rows_inv_r16w: NR = 16 * 8 * 8 * 8, one row per CTA");
  constexpr int S = NR / 8;
  extern __shared__ double2 sm[];
  int dof, ix;
  rows_block_map<MAPPED>(g.nx_loc, dof0, dof, ix);
  const int t = threadIdx.x;
  {
    const int p = t;
    const int klow = r16w_klow(p);
    const int klow2 = p == 0 ? S / 2 : S - klow;
    double2 y1[8], y2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      y1[q] = stage[stage_index(g, klow + q * S, dof, ix)];
      y2[q] = stage[stage_index(g, klow2 + q * S, dof, ix)];
    }
    if (!MAPPED && pf > 0 && blockIdx.x + pf < gridDim.x) {
      const int b = blockIdx.x + pf;
      const int pdof = dof0 + b / g.nx_loc, pix = b % g.nx_loc;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        prefetch_l2(stage + stage_index(g, klow + q * S, pdof, pix));
        prefetch_l2(stage + stage_index(g, klow2 + q * S, pdof, pix));
      }
    }
    const RowsFuseTw<NR, 2> w(tw_ny, klow, klow2, p == 0);
    if (p == 0) {
      const double2 yh0 = stage[stage_index(g, NR, dof, ix)];
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
    if constexpr (MAPPED) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += atomio_store(io, g, dof, ix, t + T * j, v[r16_out(j)]);
      atomio_block_sum<T>(acc, io.fsum_part + (size_t) ix * g.d + dof);
    } else {
      double2 *dst = reinterpret_cast<double2 *>(f + ((size_t) dof * g.nx_loc + ix) * (2 * NR));
#pragma unroll
      for (int j = 0; j < 16; ++j) dst[t + T * j] = v[r16_out(j)];
    }
  }
}

}  // namespace gfmd
