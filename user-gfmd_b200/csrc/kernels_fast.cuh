// Specialised kernels for large power-of-two grids, selected by fast_plan():
//
//   k_rows_fwd_p2   ny in {2048 .. 16384}: real rows -> half spectrum, transposed store
//   k_cols_fused_p2 nx in {2048, 4096}, ndof 3: x-FFT + Phi(q).u(q) + energy +
//                   gamma point + x-IFFT, one column set resident in shared memory,
//                   the contraction done in registers between the last forward and the
//                   first backward pass
//   k_rows_inv_p2   half spectrum -> real rows
//
// Same data layout and same role as the generic kernels (kernels_generic.cuh); the
// column spectrum is kept digit-reversed (see fft_pow2.cuh) and the Phi planes are
// stored in that order by set_phi when this path is active.
#pragma once

#include "fft_engine.cuh"
#include "fft_pow2.cuh"
#include "kernels_generic.cuh"
#include "kernel_cols_lr.cuh"

namespace gfmd {

constexpr int kColsNW = 16;     // epart slots allocated per column (max warps of a column kernel)
inline int fast_cols_nw(int) { return 1; }        // energy partials per (virtual) column
constexpr int kColsNW8 = 8;      // warps of the fused column kernel (epart has kColsNW slots per column)

// ------------------------------------------------------------------ columns ---

template <int D, int N, int T>
__global__ void __launch_bounds__(T, 1)
k_cols_fused_p2(const double2 *__restrict__ sin, double2 *__restrict__ sout, GridDesc g, int lnxl,
                const double2 *__restrict__ tw, const double *__restrict__ phi,
                const double *__restrict__ linf, double *__restrict__ epart, StepResults *res)
{
  constexpr int NW = T / 32;
  extern __shared__ double2 sm[];
  __shared__ double warp_e[NW];
  double2 *tws = sm + D * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int xmask = (1 << lnxl) - 1;
  p2_fill_tws<N>(tws, tw);
  __syncthreads();

  for (int kl = blockIdx.x; kl < g.nky_loc; kl += gridDim.x) {
    const int ky = g.ky0 + kl;
    auto addr = [&](int kcol, int dof, int x) -> size_t {
      const int p = x >> lnxl;
      return ((((size_t) (p * D + dof)) * g.kyb + kcol) << lnxl) + (x & xmask);
    };
    const double *ph = phi + (size_t) kl * D * D * N;

    // ---- group A forward: pass 0 straight from global memory, all dofs in flight
    p2_pass0_fwd<N, NW, D, 0>(sm, tw, tws, lane, warp, [&](int a, int base, int off) { return sin[addr(kl, a, base + off)]; });

    p2_groupA_rest<N, NW, -1, D, 0>(sm, tw, tws, lane, warp);
    __syncthreads();

    // ---- group B forward, contraction in registers, group B backward
#pragma unroll 1
    for (int idx = threadIdx.x; idx < N / 8; idx += T) p2_groupB_first<N, -1, D, 0>(sm, tws, idx);
    __syncwarp();

    const double wgt = (ky == 0 || (2 * ky == g.ny)) ? 1.0 : 2.0;
    double e = 0.0;
#pragma unroll 1
    for (int idx = threadIdx.x; idx < N / 8; idx += T) {
      const int pos = p2_last_base(idx);
      const int key = swz_key(pos);
      // interleaved Phi planes of this item (see phi_slot in gfmd_b200.cu)
      const double *phi_item = ph + (size_t) (idx >> 3) * (64 * D * D) + (idx & 7) * 2;
      double2 pl[2][D * D];
#pragma unroll
      for (int c = 0; c < D * D; ++c) pl[0][c] = __ldg(reinterpret_cast<const double2 *>(phi_item + c * 16));
      double2 u[D][8];
#pragma unroll
      for (int dof = 0; dof < D; ++dof) p2_last_fwd_load(sm + dof * N, pos, key, u[dof]);
#pragma unroll
      for (int rp = 0; rp < 4; ++rp) {
        if (rp < 3) {
#pragma unroll
          for (int c = 0; c < D * D; ++c)
            pl[(rp + 1) & 1][c] = __ldg(reinterpret_cast<const double2 *>(phi_item + (rp + 1) * (16 * D * D) + c * 16));
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int r = 2 * rp + s;
          double2 uv[D], F[D];
#pragma unroll
          for (int i = 0; i < D; ++i) uv[i] = u[i][r];
          phi_matvec<D>(uv, F, [&](int c) { return s == 0 ? pl[rp & 1][c].x : pl[rp & 1][c].y; });
          double eq = 0.0;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            eq = fma(F[i].x, uv[i].x, fma(F[i].y, uv[i].y, eq));
            F[i] = make_double2(-F[i].x, -F[i].y);
          }
          e = fma(wgt, eq, e);
          if (ky == 0 && pos + r == 0) {            // gamma point: kx = 0 sits at position 0
            double eg = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) res->u0[i] = uv[i].x;
#pragma unroll
            for (int a = 0; a < D / 3; ++a) {
              eg -= 2.0 * linf[a] * uv[3 * a + 2].x;
              F[3 * a + 2].x += linf[a];
            }
            res->egamma = eg;
          }
#pragma unroll
          for (int i = 0; i < D; ++i) u[i][r] = F[i];
        }
      }
#pragma unroll
      for (int dof = 0; dof < D; ++dof) p2_last_inv_store(sm + dof * N, pos, key, u[dof]);
    }
    __syncwarp();
#pragma unroll 1
    for (int idx = threadIdx.x; idx < N / 8; idx += T) p2_groupB_first<N, +1, D, 0>(sm, tws, idx);

    // energy partial of this warp (fixed order -> deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if (lane == 0) warp_e[warp] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) a += warp_e[k];
      epart[kl] = a;
    }

    // ---- group A backward, last pass straight to global memory
    p2_groupA_rest<N, NW, +1, D, 0>(sm, tw, tws, lane, warp);
    p2_pass0_inv<N, NW, D, 0>(sm, tw, tws, lane, warp,
                              [&](int a, int base, int off, double2 v) { sout[addr(kl, a, base + off)] = v; });
    // no barrier: the next column's group A touches only what this warp owns
  }
}

// --------------------------------------------------------------------- rows ---

// Lane -> wavevector map of the 256-bit transposed accesses: the eight lanes of a quarter
// warp take eight values of the SECOND spectrum digit (the first one has only LR0 bits), so
// that their shared-memory columns differ (swz_key) -- in global memory every lane touches
// its own line anyway.  Bijection on [0, NR).
template <int NR> __device__ __forceinline__ int rows256_ky(int i)
{
  constexpr int L0 = P2<NR>::LR0;
  return (i & ~((8 << L0) - 1)) | ((i & 7) << L0) | ((i >> 3) & ((1 << L0) - 1));
}

// ---- fused last pass + real/complex (un)mixing of the row kernels (FUSE variants) ----
// The packed half-length transform Z (length h = NR) and the spectrum X of the real row are
// related pairwise, k <-> h - k.  The last radix-8 pass (stride 1) of butterfly `klow`
// (frequencies klow + q S, S = NR / 8) therefore pairs with the one of butterfly S - klow, output
// q with output 7 - q: a thread that runs BOTH butterflies has every pair in registers and can
// un-mix and store straight to global memory (forward), or load, pre-mix and run the first
// inverse butterflies (backward) -- the spectrum never goes through shared memory a second
// time: 3 (forward) and 4 (backward) of the ~10 sweeps over the tile disappear, with two block
// barriers.  Butterflies 0 and S / 2 pair with themselves and form one unit together, so a row
// has exactly NR / 16 units.  Arithmetic per element is that of the unfused code (bit-identical).
// Measured at 4096 x 4096 (profiles/r1_rows_variants.txt): with one row per warp (16-byte stores
// to 32 different lines per instruction) the forward kernel LOST 32 % (0.361 vs 0.273 ms) and the
// backward one 4 %: half-written 32-byte sectors cost more than three shared-memory sweeps save.
// The unit -> (row, pair) map now keeps the RB rows of a pair in adjacent lanes, i.e. the same
// sector pattern as the unfused kernels.

// unit p in [1, NR/16) -> klow in [1, NR/16): the eight lanes of a quarter warp take eight values
// of the SECOND spectrum digit, so that both butterflies they load sit in different
// shared-memory columns (cf. rows256_ky)
template <int NR> __device__ __forceinline__ int rowsfuse_klow(int p) { return rows256_ky<NR>(p); }

// X[k] = A - i w B,  A = (Z[k] + conj Z[h-k]) / 2,  B = (Z[k] - conj Z[h-k]) / 2
__device__ __forceinline__ double2 rows_unmix(double2 zk, double2 zpartner, double2 w)
{
  const double2 zc = cconj(zpartner);
  const double2 A = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
  const double2 B = make_double2(0.5 * (zk.x - zc.x), 0.5 * (zk.y - zc.y));
  const double2 t = cmul(w, B);
  return make_double2(A.x + t.y, A.y - t.x);
}

// Z'[k] = (Y[k] + conj Y[h-k]) + i conj(w) (Y[k] - conj Y[h-k])
__device__ __forceinline__ double2 rows_premix(double2 yk, double2 ypartner, double2 w)
{
  const double2 c2 = cconj(ypartner);
  const double2 S = cadd(yk, c2), Dv = csub(yk, c2);
  const double2 t = cmulc(Dv, w);
  return make_double2(S.x - t.y, S.y + t.x);
}

// exp(-2 pi i q / 16): the frequencies of one last-pass butterfly are klow + q S with S = ny / 16,
// so their un-mixing twiddles are exp(-2 pi i klow / ny) times these constants (FUSE == 2: one
// table load per unit instead of sixteen scattered ones; not bit-identical to the table values,
// the products carry one more rounding)
__device__ __forceinline__ double2 rot16(int q)
{
  constexpr double c1 = 0.92387953251128673848, s1 = 0.38268343236508978178, h = 0.70710678118654752440;
  switch (q & 7) {
    case 0: return make_double2(1.0, 0.0);
    case 1: return make_double2(c1, -s1);
    case 2: return make_double2(h, -h);
    case 3: return make_double2(s1, -c1);
    case 4: return make_double2(0.0, -1.0);
    case 5: return make_double2(-s1, -c1);
    case 6: return make_double2(-h, -h);
    default: return make_double2(-c1, -s1);
  }
}

// twiddles of a fused unit: base values of its two butterflies (klow, klow2)
template <int NR, int FUSE>
struct RowsFuseTw {
  double2 wa, wb;
  const double2 *tw_ny;
  int klow, klow2;
  __device__ __forceinline__ RowsFuseTw(const double2 *__restrict__ t, int k1, int k2, bool special)
      : tw_ny(t), klow(k1), klow2(k2)
  {
    if (FUSE == 2) {
      if (special) {                                   // klow = 0, klow2 = S / 2: exp(-2 pi i / 32)
        wa = make_double2(1.0, 0.0);
        wb = make_double2(0.98078528040323044913, -0.19509032201612826785);
      } else {
        wa = __ldg(t + k1);
        wb = cmul(rot16(1), cconj(wa));                // exp(-2 pi i (S - klow) / ny)
      }
    }
  }
  // twiddle of frequency klow + q S (which = 0) or klow2 + q S (which = 1)
  __device__ __forceinline__ double2 get(int which, int q) const
  {
    constexpr int S = NR / 8;
    if (FUSE == 2) return cmul(which ? wb : wa, rot16(q));
    return __ldg(tw_ny + (which ? klow2 : klow) + q * S);
  }
  // twiddle of frequency ny / 2
  __device__ __forceinline__ double2 nyquist() const
  {
    if (FUSE == 2) return make_double2(-1.0, 0.0);
    return __ldg(tw_ny + NR);
  }
};

// RB rows (same dof) of ny = 2 NR reals per CTA; array a = row a of the tile.
// MB = CTAs per SM the register allocation aims at; W256 = the transposed stores move the
// two rows of a pair with one 256-bit instruction per wavevector (RB even).
template <int NR, int RB, int T, int MB = 0, bool W256 = false, int FUSE = 0>
__global__ void __launch_bounds__(T, MB)
k_rows_fwd_p2(const double *__restrict__ u, double2 *__restrict__ stage, GridDesc g,
              const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0)
{
  constexpr int LOG = P2<NR>::LOG;
  constexpr int AX = RB == 2 ? 2 : 1;       // per-row column XOR: rows of a tile never collide
  extern __shared__ double2 sm[];
  double2 *tws = sm + RB * NR;
  const int nblk = g.nx_loc / RB;
  const int dof = dof0 + blockIdx.x / nblk;
  const int ix0 = (blockIdx.x % nblk) * RB;
  const int ny = 2 * NR;
  p2_fill_tws<NR>(tws, tw);
  __syncthreads();

  const double2 *src = reinterpret_cast<const double2 *>(u + ((size_t) dof * g.nx_loc + ix0) * ny);
  p2_pass0_fwd_blk<NR, T, RB, AX>(sm, tw, tws,
                                  [&](int a, int base, int off) { return src[(size_t) a * NR + base + off]; });
  p2_groupA_rest_blk<NR, T, -1, RB, AX>(sm, tw, tws);
#pragma unroll 1
  for (int idx = threadIdx.x; idx < NR / 8; idx += T) p2_groupB_first_seq<NR, -1, RB, AX>(sm, tws, idx);
  if constexpr (FUSE == 2) {
    constexpr int S = NR / 8, HU = NR / 16;
    __syncthreads();
#pragma unroll 1
    for (int ui = threadIdx.x; ui < RB * HU; ui += T) {
      const int r = ui % RB, p = ui / RB;      // adjacent lanes = adjacent rows: 32-byte sectors stay whole
      const int rx = (r * AX) & 7;
      const double2 *row = sm + r * NR;
      const int klow = p == 0 ? 0 : rowsfuse_klow<NR>(p);
      const int klow2 = p == 0 ? S / 2 : S - klow;
      const int b1 = p2_freq_to_pos(LOG, klow), b2 = p2_freq_to_pos(LOG, klow2);
      double2 v1[8], v2[8];
      p2_last_fwd_load(row, b1, swz_key(b1) ^ rx, v1);
      p2_last_fwd_load(row, b2, swz_key(b2) ^ rx, v2);
      const RowsFuseTw<NR, FUSE> w(tw_ny, klow, klow2, p == 0);
      if (p == 0) {
        // butterflies 0 and S/2 pair with themselves: k = q S <-> (8 - q) S, and S/2 + q S <-> S/2 + (7 - q) S
        stage[stage_index(g, 0, dof, ix0 + r)] = rows_unmix(v1[0], v1[0], w.get(0, 0));
        stage[stage_index(g, NR, dof, ix0 + r)] = rows_unmix(v1[0], v1[0], w.nyquist());
#pragma unroll
        for (int q = 1; q < 8; ++q)
          stage[stage_index(g, q * S, dof, ix0 + r)] = rows_unmix(v1[q], v1[8 - q], w.get(0, q));
#pragma unroll
        for (int q = 0; q < 8; ++q)
          stage[stage_index(g, S / 2 + q * S, dof, ix0 + r)] = rows_unmix(v2[q], v2[7 - q], w.get(1, q));
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = klow + q * S, k2 = klow2 + q * S;
          stage[stage_index(g, k, dof, ix0 + r)] = rows_unmix(v1[q], v2[7 - q], w.get(0, q));
          stage[stage_index(g, k2, dof, ix0 + r)] = rows_unmix(v2[q], v1[7 - q], w.get(1, q));
        }
      }
    }
    return;
  }
  if constexpr (FUSE == 1) {
    constexpr int S = NR / 8, HU = NR / 16;
    __syncthreads();
#pragma unroll 1
    for (int ui = threadIdx.x; ui < RB * HU; ui += T) {
      const int r = ui % RB, p = ui / RB;      // adjacent lanes = adjacent rows: 32-byte sectors stay whole
      const int rx = (r * AX) & 7;
      const double2 *row = sm + r * NR;
      const int klow = p == 0 ? 0 : rowsfuse_klow<NR>(p);
      const int klow2 = p == 0 ? S / 2 : S - klow;
      const int b1 = p2_freq_to_pos(LOG, klow), b2 = p2_freq_to_pos(LOG, klow2);
      double2 v1[8], v2[8];
      p2_last_fwd_load(row, b1, swz_key(b1) ^ rx, v1);
      p2_last_fwd_load(row, b2, swz_key(b2) ^ rx, v2);
      if (p == 0) {
        // butterflies 0 and S/2 pair with themselves: k = q S <-> (8 - q) S, and S/2 + q S <-> S/2 + (7 - q) S
        stage[stage_index(g, 0, dof, ix0 + r)] = rows_unmix(v1[0], v1[0], __ldg(tw_ny));
        stage[stage_index(g, NR, dof, ix0 + r)] = rows_unmix(v1[0], v1[0], __ldg(tw_ny + NR));
#pragma unroll
        for (int q = 1; q < 8; ++q)
          stage[stage_index(g, q * S, dof, ix0 + r)] = rows_unmix(v1[q], v1[8 - q], __ldg(tw_ny + q * S));
#pragma unroll
        for (int q = 0; q < 8; ++q)
          stage[stage_index(g, S / 2 + q * S, dof, ix0 + r)] =
              rows_unmix(v2[q], v2[7 - q], __ldg(tw_ny + S / 2 + q * S));
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = klow + q * S, k2 = klow2 + q * S;
          stage[stage_index(g, k, dof, ix0 + r)] = rows_unmix(v1[q], v2[7 - q], __ldg(tw_ny + k));
          stage[stage_index(g, k2, dof, ix0 + r)] = rows_unmix(v2[q], v1[7 - q], __ldg(tw_ny + k2));
        }
      }
    }
    return;
  }
  __syncwarp();
#pragma unroll 1
  for (int idx = threadIdx.x; idx < NR / 8; idx += T) {
    const int base = p2_last_base(idx);
    const int key = swz_key(base);
#pragma unroll
    for (int a = 0; a < RB; ++a) {
      double2 v[8];
      p2_last_fwd_load(sm + a * NR, base, key ^ ((a * AX) & 7), v);
#pragma unroll
      for (int q = 0; q < 8; ++q) sm[a * NR + base + (q ^ key ^ ((a * AX) & 7))] = v[q];
    }
  }
  __syncthreads();

  // un-mix the packed transform and store transposed: X[ky] = A - i w B
  constexpr int h = NR;
  if constexpr (W256) {
    static_assert(RB % 2 == 0, "256-bit stores pair two rows");
    constexpr int RP = RB / 2;
#pragma unroll 2
    for (int i = threadIdx.x; i < RP * (h + 1); i += T) {
      const int rp = i % RP, kk = i / RP;
      const int ky = kk == h ? h : rows256_ky<NR>(kk);
      const int ka = ky == h ? 0 : ky;
      const int kb = ky == 0 ? 0 : h - ky;
      const int pa = swz(p2_freq_to_pos(LOG, ka)), pb = swz(p2_freq_to_pos(LOG, kb));
      const double2 w = __ldg(tw_ny + ky);
      double2 X[2];
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int r = 2 * rp + s2;
        const int rx = (r * AX) & 7;
        const double2 zk = sm[r * NR + (pa ^ rx)];
        const double2 zc = cconj(sm[r * NR + (pb ^ rx)]);
        const double2 A = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
        const double2 B = make_double2(0.5 * (zk.x - zc.x), 0.5 * (zk.y - zc.y));
        const double2 t = cmul(w, B);
        X[s2] = make_double2(A.x + t.y, A.y - t.x);
      }
      st_global_256(stage + stage_index(g, ky, dof, ix0 + 2 * rp), X[0], X[1]);
    }
    return;
  }
#pragma unroll 4
  for (int i = threadIdx.x; i < RB * (h + 1); i += T) {
    const int r = i % RB, ky = i / RB;
    const int ka = ky == h ? 0 : ky;
    const int kb = ky == 0 ? 0 : h - ky;
    const int rx = (r * AX) & 7;
    const double2 zk = sm[r * NR + (swz(p2_freq_to_pos(LOG, ka)) ^ rx)];
    const double2 zc = cconj(sm[r * NR + (swz(p2_freq_to_pos(LOG, kb)) ^ rx)]);
    const double2 A = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
    const double2 B = make_double2(0.5 * (zk.x - zc.x), 0.5 * (zk.y - zc.y));
    const double2 t = cmul(__ldg(tw_ny + ky), B);
    stage[stage_index(g, ky, dof, ix0 + r)] = make_double2(A.x + t.y, A.y - t.x);
  }
}

template <int NR, int RB, int T, int MB = 0, bool W256 = false, int FUSE = 0>
__global__ void __launch_bounds__(T, MB)
k_rows_inv_p2(const double2 *__restrict__ stage, double *__restrict__ f, GridDesc g,
              const double2 *__restrict__ tw, const double2 *__restrict__ tw_ny, int dof0)
{
  constexpr int LOG = P2<NR>::LOG;
  constexpr int AX = RB == 2 ? 2 : 1;
  extern __shared__ double2 sm[];
  double2 *tws = sm + RB * NR;
  double2 *yh = tws + P2<NR>::TWS;          // Y[h] of each row
  const int nblk = g.nx_loc / RB;
  const int dof = dof0 + blockIdx.x / nblk;
  const int ix0 = (blockIdx.x % nblk) * RB;
  const int ny = 2 * NR;
  constexpr int h = NR;
  p2_fill_tws<NR>(tws, tw);

  if constexpr (FUSE == 2) {
    // transposed load of both butterflies of a unit (16 independent 16-byte loads), pre-mix in
    // registers, first inverse butterflies, one store sweep to shared memory
    constexpr int S = NR / 8, HU = NR / 16;
#pragma unroll 1
    for (int ui = threadIdx.x; ui < RB * HU; ui += T) {
      const int r = ui % RB, p = ui / RB;      // adjacent lanes = adjacent rows: 32-byte sectors stay whole
      const int rx = (r * AX) & 7;
      double2 *row = sm + r * NR;
      const int klow = p == 0 ? 0 : rowsfuse_klow<NR>(p);
      const int klow2 = p == 0 ? S / 2 : S - klow;
      const int b1 = p2_freq_to_pos(LOG, klow), b2 = p2_freq_to_pos(LOG, klow2);
      double2 y1[8], y2[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        y1[q] = stage[stage_index(g, klow + q * S, dof, ix0 + r)];
        y2[q] = stage[stage_index(g, klow2 + q * S, dof, ix0 + r)];
      }
      const RowsFuseTw<NR, FUSE> w(tw_ny, klow, klow2, p == 0);
      if (p == 0) {
        const double2 yh0 = stage[stage_index(g, NR, dof, ix0 + r)];
        double2 z[8];
        z[0] = rows_premix(y1[0], yh0, w.get(0, 0));
#pragma unroll
        for (int q = 1; q < 8; ++q) z[q] = rows_premix(y1[q], y1[8 - q], w.get(0, q));
        p2_last_inv_store(row, b1, swz_key(b1) ^ rx, z);
#pragma unroll
        for (int q = 0; q < 8; ++q) z[q] = rows_premix(y2[q], y2[7 - q], w.get(1, q));
        p2_last_inv_store(row, b2, swz_key(b2) ^ rx, z);
      } else {
        // the pair (y1[q], y2[7-q]) yields (z1[q], z2[7-q]): in place
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double2 a = y1[q], b = y2[7 - q];
          y1[q] = rows_premix(a, b, w.get(0, q));
          y2[7 - q] = rows_premix(b, a, w.get(1, 7 - q));
        }
        p2_last_inv_store(row, b1, swz_key(b1) ^ rx, y1);
        p2_last_inv_store(row, b2, swz_key(b2) ^ rx, y2);
      }
    }
    __syncthreads();
  } else if constexpr (FUSE == 1) {
    // transposed load of both butterflies of a unit (16 independent 16-byte loads), pre-mix in
    // registers, first inverse butterflies, one store sweep to shared memory
    constexpr int S = NR / 8, HU = NR / 16;
#pragma unroll 1
    for (int ui = threadIdx.x; ui < RB * HU; ui += T) {
      const int r = ui % RB, p = ui / RB;      // adjacent lanes = adjacent rows: 32-byte sectors stay whole
      const int rx = (r * AX) & 7;
      double2 *row = sm + r * NR;
      const int klow = p == 0 ? 0 : rowsfuse_klow<NR>(p);
      const int klow2 = p == 0 ? S / 2 : S - klow;
      const int b1 = p2_freq_to_pos(LOG, klow), b2 = p2_freq_to_pos(LOG, klow2);
      double2 y1[8], y2[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        y1[q] = stage[stage_index(g, klow + q * S, dof, ix0 + r)];
        y2[q] = stage[stage_index(g, klow2 + q * S, dof, ix0 + r)];
      }
      if (p == 0) {
        const double2 yh0 = stage[stage_index(g, NR, dof, ix0 + r)];
        double2 z[8];
        z[0] = rows_premix(y1[0], yh0, __ldg(tw_ny));
#pragma unroll
        for (int q = 1; q < 8; ++q) z[q] = rows_premix(y1[q], y1[8 - q], __ldg(tw_ny + q * S));
        p2_last_inv_store(row, b1, swz_key(b1) ^ rx, z);
#pragma unroll
        for (int q = 0; q < 8; ++q) z[q] = rows_premix(y2[q], y2[7 - q], __ldg(tw_ny + S / 2 + q * S));
        p2_last_inv_store(row, b2, swz_key(b2) ^ rx, z);
      } else {
        // the pair (y1[q], y2[7-q]) yields (z1[q], z2[7-q]): in place
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double2 a = y1[q], b = y2[7 - q];
          y1[q] = rows_premix(a, b, __ldg(tw_ny + klow + q * S));
          y2[7 - q] = rows_premix(b, a, __ldg(tw_ny + klow2 + (7 - q) * S));
        }
        p2_last_inv_store(row, b1, swz_key(b1) ^ rx, y1);
        p2_last_inv_store(row, b2, swz_key(b2) ^ rx, y2);
      }
    }
    __syncthreads();
  } else {
  if constexpr (W256) {
    // transposed load, 4 independent 32-byte loads (both rows of a pair) in flight per thread
    static_assert(RB % 2 == 0, "256-bit loads pair two rows");
    constexpr int RP = RB / 2;
    constexpr int UL2 = 4;
#pragma unroll 1
    for (int i0 = threadIdx.x; i0 < RP * (h + 1); i0 += T * UL2) {
      double2 y[UL2][2];
#pragma unroll
      for (int j = 0; j < UL2; ++j) {
        const int i = i0 + j * T;
        if (i < RP * (h + 1)) {
          const int kk = i / RP;
          const int ky = kk == h ? h : rows256_ky<NR>(kk);
          ld_global_256(stage + stage_index(g, ky, dof, ix0 + 2 * (i % RP)), y[j][0], y[j][1]);
        }
      }
#pragma unroll
      for (int j = 0; j < UL2; ++j) {
        const int i = i0 + j * T;
        if (i < RP * (h + 1)) {
          const int rp = i % RP, kk = i / RP;
          const int ky = kk == h ? h : rows256_ky<NR>(kk);
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int r = 2 * rp + s2;
            if (ky < h) sm[r * NR + (swz(p2_freq_to_pos(LOG, ky)) ^ ((r * AX) & 7))] = y[j][s2];
            else yh[r] = y[j][s2];
          }
        }
      }
    }
  } else {
  // transposed load, 8 independent 16-byte loads in flight per thread
  constexpr int UL = 8;
#pragma unroll 1
  for (int i0 = threadIdx.x; i0 < RB * (h + 1); i0 += T * UL) {
    double2 y[UL];
#pragma unroll
    for (int j = 0; j < UL; ++j) {
      const int i = i0 + j * T;
      if (i < RB * (h + 1)) y[j] = stage[stage_index(g, i / RB, dof, ix0 + i % RB)];
    }
#pragma unroll
    for (int j = 0; j < UL; ++j) {
      const int i = i0 + j * T;
      if (i < RB * (h + 1)) {
        const int r = i % RB, ky = i / RB;
        if (ky < h) sm[r * NR + (swz(p2_freq_to_pos(LOG, ky)) ^ ((r * AX) & 7))] = y[j];
        else yh[r] = y[j];
      }
    }
  }
  }
  __syncthreads();

  // Z'[k] = (Y[k] + conj Y[h-k]) + i e^{+2 pi i k/ny} (Y[k] - conj Y[h-k]), in place pairwise
  constexpr int np = h / 2 + 1;
#pragma unroll 2
  for (int i = threadIdx.x; i < RB * np; i += T) {
    const int r = i % RB, k = i / RB;
    const int k2 = h - k;
    const int rx = (r * AX) & 7;
    const int pk = r * NR + (swz(p2_freq_to_pos(LOG, k)) ^ rx);
    const int p2i = r * NR + (swz(p2_freq_to_pos(LOG, k2 & (h - 1))) ^ rx);
    const double2 yk = sm[pk];
    const double2 y2 = k == 0 ? yh[r] : sm[p2i];
    {
      const double2 c2 = cconj(y2);
      const double2 S = cadd(yk, c2), Dv = csub(yk, c2);
      const double2 t = cmulc(Dv, __ldg(tw_ny + k));
      sm[pk] = make_double2(S.x - t.y, S.y + t.x);
    }
    if (k2 != k && k2 < h) {
      const double2 ck = cconj(yk);
      const double2 S = cadd(y2, ck), Dv = csub(y2, ck);
      const double2 t = cmulc(Dv, __ldg(tw_ny + k2));
      sm[p2i] = make_double2(S.x - t.y, S.y + t.x);
    }
  }
  __syncthreads();

#pragma unroll 1
  for (int idx = threadIdx.x; idx < NR / 8; idx += T) {
    const int base = p2_last_base(idx);
    const int key = swz_key(base);
#pragma unroll
    for (int a = 0; a < RB; ++a) {
      const int ka = key ^ ((a * AX) & 7);
      double2 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = sm[a * NR + base + (q ^ ka)];
      p2_last_inv_store(sm + a * NR, base, ka, v);
    }
  }
  __syncwarp();
  }   // !FUSE
#pragma unroll 1
  for (int idx = threadIdx.x; idx < NR / 8; idx += T) p2_groupB_first_seq<NR, +1, RB, AX>(sm, tws, idx);
  p2_groupA_rest_blk<NR, T, +1, RB, AX>(sm, tw, tws);
  double2 *dst = reinterpret_cast<double2 *>(f + ((size_t) dof * g.nx_loc + ix0) * ny);
  p2_pass0_inv_blk<NR, T, RB, AX>(sm, tw, tws,
                                  [&](int a, int base, int off, double2 v) { dst[(size_t) a * NR + base + off] = v; });
}

}  // namespace gfmd

#include "kernels_rows_r16.cuh"
#include "kernels_rows_cluster.cuh"

namespace gfmd {

// ---------------------------------------------------------------- selection ---

struct FastRowsCfg { int nr, rb, t; };

// variant id = ny + k (k = 8, the radix-16 kernels, is the default where it exists: fast_rows_default);
// the others are selected with GFMD_B200_ROWS_VARIANT=<id> at handle creation (see ROWS_VARIANTS below):
//   ny = 4096: +1 four rows per CTA, +3 256-bit transposed accesses;  ny = 8192: +3 likewise;
//   every ny: +5 last pass fused with the real/complex (un)mixing (FUSE, see above) in both
//   directions;  ny = 4096, 8192: +6 fused backward, unfused forward (the default at 4096),
//   +7 fused both ways with closed-form twiddles (one table load per unit);
//   ny = 4096, 8192, 16384: +8 radix-16 passes, four / five / six shared-memory sweeps
//   (kernels_rows_r16.cuh);  ny = 16384: +9 the same as two-CTA clusters (kernels_rows_cluster.cuh).
// Measured on a B200 at 4096 x 4096 (tools/rows_variants_ab.py, profiles/r1_rows_variants.txt):
// rows_fwd / rows_inv 0.273 / 0.308 ms default, 0.266 / 0.339 (+1), 0.268 / 0.304 (+3); forcing
// three CTAs per SM with __launch_bounds__(256, 3) (85 registers, spills) was 25 % slower and
// is not kept.  The template parameter MB stays for such experiments.
inline bool fast_rows_cfg(int variant, FastRowsCfg &c)
{
  switch (variant) {
    case 2048: case 2053: c = {1024, 4, 128}; return true;
    case 4096: case 4099: case 4101: case 4102: case 4103: case 4104: c = {2048, 2, 256}; return true;
    case 4097: c = {2048, 4, 512}; return true;
    case 8192: case 8195: case 8197: case 8198: case 8199: case 8200: c = {4096, 2, 512}; return true;
    case 16384: case 16389: case 16392: c = {8192, 1, 512}; return true;
#ifndef GFMD_CUDA_EMU
    case 16393: c = {8192, 1, 512}; return true;      // two-CTA clusters (kernels_rows_cluster.cuh)
#endif
    default: return false;
  }
}

// X(id, NR, RB, T, MB, W256, FUSE_FWD, FUSE_INV) for every instantiated row-kernel variant;
// FUSE: 0 unfused, 1 fused with table twiddles (bit-identical to 0), 2 fused with closed-form twiddles
#define ROWS_VARIANTS(X)                  \
  X(2048, 1024, 4, 128, 0, false, 0, 0)   \
  X(2053, 1024, 4, 128, 0, false, 1, 1)   \
  X(4096, 2048, 2, 256, 0, false, 0, 0)   \
  X(4097, 2048, 4, 512, 0, false, 0, 0)   \
  X(4099, 2048, 2, 256, 0, true, 0, 0)    \
  X(4101, 2048, 2, 256, 2, false, 1, 1)   \
  X(4102, 2048, 2, 256, 2, false, 0, 1)   \
  X(4103, 2048, 2, 256, 2, false, 2, 2)   \
  X(8192, 4096, 2, 512, 0, false, 0, 0)   \
  X(8195, 4096, 2, 512, 0, true, 0, 0)    \
  X(8197, 4096, 2, 512, 0, false, 1, 1)   \
  X(8198, 4096, 2, 512, 0, false, 0, 1)   \
  X(8199, 4096, 2, 512, 0, false, 2, 2)   \
  X(16384, 8192, 1, 512, 0, false, 0, 0)  \
  X(16389, 8192, 1, 512, 0, false, 1, 1)

// the variant a grid gets when GFMD_B200_ROWS_VARIANT does not say otherwise: the radix-16 kernels
// (kernels_rows_r16.cuh) wherever they exist.  Measured on a B200 (profiles/r2_rows_variants.txt),
// rows_fwd / rows_inv in ms: 4096 x 4096: 0.192 / 0.218 against 0.272 / 0.263 (ny + 6, the round-1
// default) and 0.244 / 0.243 (ny + 7); 4096 x 8192: 0.514 / 0.612 against 0.564 / 0.746 (ny + 0)
// and 0.539 / 0.596 (ny + 7); 2048 x 16384: 0.769 / 0.711 against 0.962 / 0.890
inline int fast_rows_default(int ny) { return (ny == 4096 || ny == 8192 || ny == 16384) ? ny + 8 : ny; }

inline size_t fast_rows_smem(const FastRowsCfg &c)
{
  return sizeof(double2) * ((size_t) c.rb * c.nr + 504 + c.rb);
}
// cluster variants: the row(s) plus the exchange buffer (8 points per thread and the Nyquist point)
inline size_t fast_rows_smem_cluster(const FastRowsCfg &c)
{
  return sizeof(double2) * ((size_t) c.rb * c.nr + 8 * (size_t) c.t + 8);
}

inline size_t fast_cols_smem(int d, int nx) { return sizeof(double2) * ((size_t) d * nx + 504); }

inline int ilog2_rt(int n)
{
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}

// returns 0 on success; fast_rows / fast_cols = 0 (generic kernels) or the variant id
inline int fast_plan(const GridDesc &g, int &fast_rows, int &fast_cols, int &cols_top)
{
  fast_rows = 0;
  fast_cols = 0;
  if (getenv("GFMD_B200_NO_FAST")) return 0;
  FastRowsCfg rc;
  int rv = fast_rows_default(g.ny);
  if (g.ny == 4096 && getenv("GFMD_B200_ROWS_RB4") && g.nx_loc % 4 == 0) rv = 4097;
  if (const char *e = getenv("GFMD_B200_ROWS_VARIANT")) {
    const int want = atoi(e);
    if (want >= g.ny && want < g.ny + 16 && fast_rows_cfg(want, rc) && 2 * rc.nr == g.ny) rv = want;
  }
  if (fast_rows_cfg(rv, rc) && g.nx_loc % rc.rb == 0) {
    fast_rows = rv;
    cudaError_t e = cudaSuccess;
    switch (rv) {
#define ROWS_ATTR(ID, NR, RB, T, MB, W, FF, FI)                                                                      \
  case ID:                                                                                                    \
    e = cudaFuncSetAttribute(k_rows_fwd_p2<NR, RB, T, FF ? MB : 0, W, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int) fast_rows_smem(rc));                                                       \
    if (e == cudaSuccess)                                                                                     \
      e = cudaFuncSetAttribute(k_rows_inv_p2<NR, RB, T, FI ? MB : 0, W, FI>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                               (int) fast_rows_smem(rc));                                                     \
    break;
      ROWS_VARIANTS(ROWS_ATTR)
#undef ROWS_ATTR
#define R16_ATTR(K)                                                                                             \
  if (e == cudaSuccess)                                                                                         \
    e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fast_rows_smem(rc));
      case 4104:
        R16_ATTR((k_rows_fwd_r16<2048, 2, 256>)) R16_ATTR((k_rows_inv_r16<2048, 2, 256>))
        R16_ATTR((k_rows_fwd_r16<2048, 2, 256, true>)) R16_ATTR((k_rows_inv_r16<2048, 2, 256, true>))
        break;
      case 16392:
        R16_ATTR((k_rows_fwd_r16w<8192, 512>)) R16_ATTR((k_rows_inv_r16w<8192, 512>))
        R16_ATTR((k_rows_fwd_r16w<8192, 512, true>)) R16_ATTR((k_rows_inv_r16w<8192, 512, true>))
        break;
#ifndef GFMD_CUDA_EMU
      case 16393:
        if (g.nx_loc % 2) { fast_rows = 0; break; }
        e = cudaFuncSetAttribute(k_rows_fwd_r16c<8192, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) fast_rows_smem_cluster(rc));
        if (e == cudaSuccess)
          e = cudaFuncSetAttribute(k_rows_inv_r16c<8192, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int) fast_rows_smem_cluster(rc));
        break;
#endif
      case 8200:
        R16_ATTR((k_rows_fwd_r16h<4096, 2, 512>)) R16_ATTR((k_rows_inv_r16h<4096, 2, 512>))
        R16_ATTR((k_rows_fwd_r16h<4096, 2, 512, true>)) R16_ATTR((k_rows_inv_r16h<4096, 2, 512, true>))
        break;
#undef R16_ATTR
    }
    if (e != cudaSuccess) return 1;
  }
  cols_top = 0;
  const bool p2ranks = (g.P & (g.P - 1)) == 0;
  if (g.d == 3 && p2ranks && g.nx == 2048 && g.nx_loc >= 64) {
    fast_cols = 2048;
    if (cudaFuncSetAttribute(k_cols_fused_p2<3, 2048, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int) fast_cols_smem(3, 2048)) != cudaSuccess)
      return 1;
  } else if (g.d == 3 && p2ranks && (g.nx == 4096 || g.nx == 8192 || g.nx == 16384) && g.nx_loc >= 512) {
    // sub-columns of 4096; nx = 8192 / 16384 add one top-digit pass in HBM (kernel_cols_lr.cuh)
    fast_cols = 4096;
    cols_top = ilog2_rt(g.nx) - 12;
    const int lnxl = ilog2_rt(g.nx_loc);
    const int lp = lnxl >= 12 ? 0 : 12 - lnxl;
    cudaError_t e = cudaSuccess;
    switch (lp) {
#define LR_ATTR_PIPE(LP)                                                                                     \
  if (e == cudaSuccess)                                                                                      \
    e = cudaFuncSetAttribute(k_cols_fused_p2_lr<4096, 512, LP, true>,                                        \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fast_cols_smem(3, 4096));
#define LR_ATTR(LP)                                                                                          \
  case LP:                                                                                                   \
    e = cudaFuncSetAttribute(k_cols_fused_p2_lr<4096, 512, LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int) fast_cols_smem(3, 4096));                                                 \
    if (e == cudaSuccess && g.P > 1)                                                                         \
      e = cudaFuncSetAttribute(k_cols_fused_p2_lr<4096, 512, LP, false, 1>,                                  \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fast_cols_smem(3, 4096));  \
    if (e == cudaSuccess && g.P > 1)                                                                         \
      e = cudaFuncSetAttribute(k_cols_fused_p2_lr<4096, 512, LP, false, 2>,                                  \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fast_cols_smem(3, 4096));  \
    LR_ATTR_PIPE(LP)                                                                                         \
    break;
      LR_ATTR(0) LR_ATTR(1) LR_ATTR(2) LR_ATTR(3)
#undef LR_ATTR
      default: fast_cols = 0; cols_top = 0; break;
    }
    if (e != cudaSuccess) return 1;
  }
  return 0;
}

// variants whose kernels have a fused atom I/O form (AtomIO, kernels_rows_r16.cuh)
// L2 prefetch distance (in CTAs) of the radix-16 row kernels: a CTA prefetches the input tile of the
// CTA `pf` blocks later, behind its own loads, so that tile's loads hit L2.  Measured (one B200,
// profiles/r2_rows_prefetch_ab.txt; distance 148 = one wave of SMs): 4096 x 4096 rows_fwd 0.193 ->
// 0.182 ms, rows_inv 0.219 -> 0.201 ms (296: the same, 592: worse); 2048 x 16384 rows_fwd 0.721 ->
// 0.692, rows_inv 0.716 -> 0.736 (worse: off).  GFMD_B200_ROWS_PREFETCH=<n> overrides (0 = off).
inline int fast_rows_prefetch(int variant, int dir)
{
  static const int env = getenv("GFMD_B200_ROWS_PREFETCH") ? atoi(getenv("GFMD_B200_ROWS_PREFETCH")) : -1;
  if (env >= 0) return env;
  if (variant == 4104) return 148;
  if (variant == 16392) return dir < 0 ? 148 : 0;
  return 0;
}

inline bool fast_rows_has_atomio(int variant) { return variant == 4104 || variant == 8200 || variant == 16392; }

// io != nullptr: gather fused into the transform (u is not read)
inline int fast_rows_fwd(int variant, const double *u, double2 *stage, const GridDesc &g, const double2 *tw_ny,
                         const FftDesc &fd, cudaStream_t s, long long *launches, int dof0 = 0, int ndofs = -1,
                         const AtomIO *io = nullptr)
{
  FastRowsCfg rc;
  if (!fast_rows_cfg(variant, rc)) return 1;
  if (ndofs < 0) ndofs = g.d - dof0;
  const int grid = ndofs * (g.nx_loc / rc.rb);
  const size_t smem = fast_rows_smem(rc);
  if (io) {
    switch (variant) {
      case 4104: k_rows_fwd_r16<2048, 2, 256, true><<<grid, 256, smem, s>>>(u, stage, g, fd.core.tw, tw_ny, dof0, *io); break;
      case 8200: k_rows_fwd_r16h<4096, 2, 512, true><<<grid, 512, smem, s>>>(u, stage, g, fd.core.tw, tw_ny, dof0, *io); break;
      case 16392: k_rows_fwd_r16w<8192, 512, true><<<grid, 512, smem, s>>>(u, stage, g, fd.core.tw, tw_ny, dof0, *io); break;
      default: return 1;
    }
    ++*launches;
    return 0;
  }
  switch (variant) {
#define ROWS_LAUNCH(ID, NR, RB, T, MB, W, FF, FI) \
  case ID: k_rows_fwd_p2<NR, RB, T, FF ? MB : 0, W, FF><<<grid, T, smem, s>>>(u, stage, g, fd.core.tw, tw_ny, dof0); break;
    ROWS_VARIANTS(ROWS_LAUNCH)
#undef ROWS_LAUNCH
    case 4104: k_rows_fwd_r16<2048, 2, 256><<<grid, 256, smem, s>>>(u, stage, g, fd.core.tw, tw_ny, dof0, AtomIO(), fast_rows_prefetch(4104, -1)); break;
    case 8200: k_rows_fwd_r16h<4096, 2, 512><<<grid, 512, smem, s>>>(u, stage, g, fd.core.tw, tw_ny, dof0); break;
    case 16392: k_rows_fwd_r16w<8192, 512><<<grid, 512, smem, s>>>(u, stage, g, fd.core.tw, tw_ny, dof0, AtomIO(), fast_rows_prefetch(16392, -1)); break;
#ifndef GFMD_CUDA_EMU
    case 16393: k_rows_fwd_r16c<8192, 512><<<grid, 512, fast_rows_smem_cluster(rc), s>>>(u, stage, g, fd.core.tw, tw_ny, dof0); break;
#endif
    default: return 1;
  }
  ++*launches;
  return 0;
}

// io != nullptr: scatter fused into the transform (f is not written; io->fsum_part receives
// [nx_loc / rb][ndof] partial force sums)
inline int fast_rows_inv(int variant, const double2 *stage, double *f, const GridDesc &g, const double2 *tw_ny,
                         const FftDesc &fd, cudaStream_t s, long long *launches, int dof0 = 0, int ndofs = -1,
                         const AtomIO *io = nullptr)
{
  FastRowsCfg rc;
  if (!fast_rows_cfg(variant, rc)) return 1;
  if (ndofs < 0) ndofs = g.d - dof0;
  const int grid = ndofs * (g.nx_loc / rc.rb);
  const size_t smem = fast_rows_smem(rc);
  if (io) {
    switch (variant) {
      case 4104: k_rows_inv_r16<2048, 2, 256, true><<<grid, 256, smem, s>>>(stage, f, g, fd.core.tw, tw_ny, dof0, *io); break;
      case 8200: k_rows_inv_r16h<4096, 2, 512, true><<<grid, 512, smem, s>>>(stage, f, g, fd.core.tw, tw_ny, dof0, *io); break;
      case 16392: k_rows_inv_r16w<8192, 512, true><<<grid, 512, smem, s>>>(stage, f, g, fd.core.tw, tw_ny, dof0, *io); break;
      default: return 1;
    }
    ++*launches;
    return 0;
  }
  switch (variant) {
#define ROWS_LAUNCH(ID, NR, RB, T, MB, W, FF, FI) \
  case ID: k_rows_inv_p2<NR, RB, T, FI ? MB : 0, W, FI><<<grid, T, smem, s>>>(stage, f, g, fd.core.tw, tw_ny, dof0); break;
    ROWS_VARIANTS(ROWS_LAUNCH)
#undef ROWS_LAUNCH
    case 4104: k_rows_inv_r16<2048, 2, 256><<<grid, 256, smem, s>>>(stage, f, g, fd.core.tw, tw_ny, dof0, AtomIO(), fast_rows_prefetch(4104, +1)); break;
    case 8200: k_rows_inv_r16h<4096, 2, 512><<<grid, 512, smem, s>>>(stage, f, g, fd.core.tw, tw_ny, dof0); break;
    case 16392: k_rows_inv_r16w<8192, 512><<<grid, 512, smem, s>>>(stage, f, g, fd.core.tw, tw_ny, dof0, AtomIO(), fast_rows_prefetch(16392, +1)); break;
#ifndef GFMD_CUDA_EMU
    case 16393: k_rows_inv_r16c<8192, 512><<<grid, 512, fast_rows_smem_cluster(rc), s>>>(stage, f, g, fd.core.tw, tw_ny, dof0); break;
#endif
    default: return 1;
  }
  ++*launches;
  return 0;
}

// sin: staging buffer holding the received columns, sout: where the result columns go
// (the same buffer on a single GPU).  tw_sub: twiddles of the in-shared-memory length
// (2048 or 4096); tw_nx: twiddles of the full column length (top pass only).
inline int fast_cols_fused(int variant, int top, double2 *sin, double2 *sout, const GridDesc &g,
                           const double2 *tw_sub, const double2 *tw_nx, const double *phi, const double *linf,
                           double *epart, StepResults *res, int num_sms, cudaStream_t s, long long *launches,
                           int kl0 = 0, int kl1 = -1, const PeerOut *peer_out = nullptr,
                           const PeerOut *peer_in = nullptr, cudaEvent_t *ev_split = nullptr, int phases = 7,
                           int top_sms = 0, int pull_dof0 = 0, int pull_nd = -1)
{
  // pull_dof0 / pull_nd: dof range of the confined PULLING top-digit pass (phases == 1, top_sms > 0)
  // phases: bit 0 the forward top-digit pass, bit 1 the fused kernel, bit 2 the backward top-digit pass
  // (the overlapped multi-GPU step launches them on different streams); top_sms > 0 caps the grid of
  // the top-digit passes to that many SMs' worth of CTAs
  // ev_split (profiling): [0] recorded behind the forward top-digit pass, [1] before the backward one
  // peer_out (variant 4096 in slab mode): the last kernel of the stage stores the result pieces
  // straight into their owners' return buffers instead of sout.  peer_in (only together with
  // peer_out): the first kernel of the stage loads the pieces straight from the ranks that produced
  // them instead of sin; with a top-digit pass, sin then receives the assembled columns.
  if (peer_in && !peer_out) return 1;
  if (kl1 < 0) kl1 = g.nky_loc;
  if (kl1 > g.nky_loc) kl1 = g.nky_loc;
  if (kl0 >= kl1) return 0;
  if (variant != 4096 && (kl0 != 0 || kl1 != g.nky_loc || peer_out)) return 1;
  const int nvc = (kl1 - kl0) << top;
  const int grid = nvc < num_sms ? nvc : num_sms;
  // software-pipelined column kernel (kernel_cols_lr.cuh): the default since it was measured
  // (0.376 -> 0.345 ms at 4096 x 4096, profiles/r2_cols_pipe.txt); GFMD_B200_COLS_PIPE=0 selects the plain one
  static const bool cols_pipe = !(getenv("GFMD_B200_COLS_PIPE") && atoi(getenv("GFMD_B200_COLS_PIPE")) == 0);
  const int lnxl = ilog2_rt(g.nx_loc);
  const size_t smem = fast_cols_smem(3, variant);
  const long long top_items = (long long) g.d * (kl1 - kl0) * (g.nx >> top);
  const long long top_cap = (long long) (top_sms > 0 ? top_sms : num_sms) * 8;
  const int top_grid = (int) ((top_items + 255) / 256 < top_cap ? (top_items + 255) / 256 : top_cap);
  // overlapped multi-GPU step: the exchanging passes are confined to top_sms SMs, ONE fat CTA each (CTAs
  // are placed breadth-first: many small CTAs would spread over all SMs and keep the fused kernel, which
  // needs a whole SM per CTA, from starting), with two items = 2 R NVLink loads in flight per thread
  if (!(phases & 1)) {
  } else if (top_sms > 0 && peer_in && top == 1) {
    k_cols_top_pass<1, -1, true, 1024, 2><<<top_sms, 1024, 0, s>>>(sin, g, lnxl, tw_nx, kl0, kl1, *peer_in, pull_dof0, pull_nd);
    ++*launches;
  } else if (top_sms > 0 && peer_in && top == 2) {
    k_cols_top_pass<2, -1, true, 1024, 2><<<top_sms, 1024, 0, s>>>(sin, g, lnxl, tw_nx, kl0, kl1, *peer_in, pull_dof0, pull_nd);
    ++*launches;
  } else if (top == 1) {
    if (peer_in) k_cols_top_pass<1, -1, true><<<top_grid, 256, 0, s>>>(sin, g, lnxl, tw_nx, kl0, kl1, *peer_in);
    else k_cols_top_pass<1, -1><<<top_grid, 256, 0, s>>>(sin, g, lnxl, tw_nx, kl0, kl1);
    ++*launches;
  } else if (top == 2) {
    if (peer_in) k_cols_top_pass<2, -1, true><<<top_grid, 256, 0, s>>>(sin, g, lnxl, tw_nx, kl0, kl1, *peer_in);
    else k_cols_top_pass<2, -1><<<top_grid, 256, 0, s>>>(sin, g, lnxl, tw_nx, kl0, kl1);
    ++*launches;
  }
  if (ev_split && top > 0) cudaEventRecord(ev_split[0], s);
  if (phases & 2) {
  switch (variant) {
    case 2048:
      k_cols_fused_p2<3, 2048, 256><<<grid, 256, smem, s>>>(sin, sout, g, lnxl, tw_sub, phi, linf, epart, res);
      break;
    case 4096:
      switch (lnxl >= 12 ? 0 : 12 - lnxl) {
#define LR_LAUNCH_PIPE(LP)                                                                                  \
  if (cols_pipe) {                                                                                          \
    k_cols_fused_p2_lr<4096, 512, LP, true><<<grid, 512, smem, s>>>(sin, sout, g, lnxl, top, kl0, kl1,      \
                                                                    tw_sub, phi, linf, epart, res);         \
    break;                                                                                                  \
  }
#define LR_LAUNCH(LP)                                                                                       \
  case LP:                                                                                                  \
    if (peer_out && top == 0) {                                                                             \
      if (peer_in)                                                                                          \
        k_cols_fused_p2_lr<4096, 512, LP, false, 2><<<grid, 512, smem, s>>>(                                \
            sin, sout, g, lnxl, top, kl0, kl1, tw_sub, phi, linf, epart, res, *peer_out, *peer_in);         \
      else                                                                                                  \
        k_cols_fused_p2_lr<4096, 512, LP, false, 1><<<grid, 512, smem, s>>>(                                \
            sin, sout, g, lnxl, top, kl0, kl1, tw_sub, phi, linf, epart, res, *peer_out);                   \
      break;                                                                                                \
    }                                                                                                       \
    LR_LAUNCH_PIPE(LP)                                                                                      \
    k_cols_fused_p2_lr<4096, 512, LP><<<grid, 512, smem, s>>>(sin, sout, g, lnxl, top, kl0, kl1, tw_sub,    \
                                                              phi, linf, epart, res);                       \
    break;
        LR_LAUNCH(0) LR_LAUNCH(1) LR_LAUNCH(2) LR_LAUNCH(3)
#undef LR_LAUNCH
        default: return 1;
      }
      break;
    default: return 1;
  }
  ++*launches;
  }
  if (ev_split && top > 0) cudaEventRecord(ev_split[1], s);
  if (!(phases & 4)) {
  } else if (top_sms > 0 && peer_out && top == 1) {
    k_cols_top_pass<1, +1, true, 1024, 2><<<top_sms, 1024, 0, s>>>(sout, g, lnxl, tw_nx, kl0, kl1, *peer_out);
    ++*launches;
  } else if (top_sms > 0 && peer_out && top == 2) {
    k_cols_top_pass<2, +1, true, 1024, 2><<<top_sms, 1024, 0, s>>>(sout, g, lnxl, tw_nx, kl0, kl1, *peer_out);
    ++*launches;
  } else if (top == 1) {
    if (peer_out) k_cols_top_pass<1, +1, true><<<top_grid, 256, 0, s>>>(sout, g, lnxl, tw_nx, kl0, kl1, *peer_out);
    else k_cols_top_pass<1, +1><<<top_grid, 256, 0, s>>>(sout, g, lnxl, tw_nx, kl0, kl1);
    ++*launches;
  } else if (top == 2) {
    if (peer_out) k_cols_top_pass<2, +1, true><<<top_grid, 256, 0, s>>>(sout, g, lnxl, tw_nx, kl0, kl1, *peer_out);
    else k_cols_top_pass<2, +1><<<top_grid, 256, 0, s>>>(sout, g, lnxl, tw_nx, kl0, kl1);
    ++*launches;
  }
  return 0;
}

}  // namespace gfmd
