// Shared-memory FP64 complex FFT engine (device side), generic in the length.
//
// One CTA transforms a small batch of sequences that live in shared memory:
// in-place Stockham autosort passes of radix 2/3/4/5/7/8, every thread
// keeping up to EPT = 16 complex values in registers across the barrier that
// separates "all reads" from "all writes" of a pass, so no second buffer is
// needed.  Lengths with a prime factor > 7 (e.g. the reference's 64x37 test
// grid) go through Bluestein's chirp-z on a power-of-two core length.
//
// This replaces what LAMMPS' FFT3d/FFTW3 does for the reference
// (src/solvers/gfmd_solver_fft.cpp:72-80,116,181): unnormalised transforms,
// DIR = -1 is exp(-i q r) ("forward"), DIR = +1 is exp(+i q r).
#pragma once

#include <cuda_runtime.h>

namespace gfmd {

constexpr int kMaxPass = 16;
constexpr int kEPT = 16;          // complex elements a thread holds per pass

struct FftCore {                  // smooth-length transform run in shared memory
  int len;
  int npass;
  int radix[kMaxPass];
  const double2 *tw;              // tw[k] = exp(-2 pi i k / len), k < len
};

struct FftDesc {
  int n;                          // logical transform length
  int ld_min;                     // shared-memory elements one transform needs
  int bluestein;                  // 0: core.len == n;  1: chirp-z, core.len == m
  FftCore core;
  const double2 *chirp;           // [n]  exp(-i pi k^2 / n)
  const double2 *bhat;            // [m]  FFT_m(wrapped conj chirp) / m
};

// ---------------------------------------------------------------- complex ---

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
  return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ double2 cmulc(double2 a, double2 b)
{
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }

template <int DIR> __device__ __forceinline__ double2 twid(double2 w)
{
  return DIR < 0 ? w : make_double2(w.x, -w.y);
}
// multiply by -i (forward) / +i (backward)
template <int DIR> __device__ __forceinline__ double2 mul_mi(double2 a)
{
  return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

// ------------------------------------------------------------ butterflies ---

template <int DIR> __device__ __forceinline__ void bf2(double2 &a, double2 &b)
{
  double2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <int DIR> __device__ __forceinline__ void bf4(double2 &v0, double2 &v1, double2 &v2, double2 &v3)
{
  double2 t0 = cadd(v0, v2), t1 = csub(v0, v2);
  double2 t2 = cadd(v1, v3), t3 = mul_mi<DIR>(csub(v1, v3));
  v0 = cadd(t0, t2);
  v1 = cadd(t1, t3);
  v2 = csub(t0, t2);
  v3 = csub(t1, t3);
}

template <int R, int DIR> struct Butterfly;

template <int DIR> struct Butterfly<2, DIR> {
  static __device__ __forceinline__ void run(double2 *v) { bf2<DIR>(v[0], v[1]); }
};

template <int DIR> struct Butterfly<4, DIR> {
  static __device__ __forceinline__ void run(double2 *v) { bf4<DIR>(v[0], v[1], v[2], v[3]); }
};

template <int DIR> struct Butterfly<3, DIR> {
  static __device__ __forceinline__ void run(double2 *v)
  {
    const double s3 = 0.86602540378443864676;
    double2 t = cadd(v[1], v[2]);
    double2 d = cscale(mul_mi<DIR>(csub(v[1], v[2])), s3);
    double2 m = make_double2(fma(-0.5, t.x, v[0].x), fma(-0.5, t.y, v[0].y));
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, d);
    v[2] = csub(m, d);
  }
};

template <int DIR> struct Butterfly<8, DIR> {
  static __device__ __forceinline__ void run(double2 *v)
  {
    const double h = 0.70710678118654752440;
    // even / odd radix-4
    bf4<DIR>(v[0], v[2], v[4], v[6]);   // E0..E3 in v0,v2,v4,v6
    bf4<DIR>(v[1], v[3], v[5], v[7]);   // O0..O3 in v1,v3,v5,v7
    double2 o1, o2, o3;
    if (DIR < 0) {
      o1 = make_double2((v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h);
      o3 = make_double2((v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h);
    } else {
      o1 = make_double2((v[3].x - v[3].y) * h, (v[3].x + v[3].y) * h);
      o3 = make_double2(-(v[7].x + v[7].y) * h, (v[7].x - v[7].y) * h);
    }
    o2 = mul_mi<DIR>(v[5]);
    double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0);
    v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1);
    v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2);
    v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3);
    v[7] = csub(e3, o3);
  }
};

// small-prime DFT by definition (rarely used lengths; correctness first)
template <int R> struct Roots;
template <> struct Roots<5> {
  static __device__ __forceinline__ double c(int k)
  {
    const double t[5] = {1.0, 0.30901699437494742410, -0.80901699437494742410, -0.80901699437494742410,
                         0.30901699437494742410};
    return t[k];
  }
  static __device__ __forceinline__ double s(int k)
  {
    const double t[5] = {0.0, 0.95105651629515357212, 0.58778525229247312917, -0.58778525229247312917,
                         -0.95105651629515357212};
    return t[k];
  }
};
template <> struct Roots<7> {
  static __device__ __forceinline__ double c(int k)
  {
    const double t[7] = {1.0,
                         0.62348980185873353053,
                         -0.22252093395631440429,
                         -0.90096886790241912624,
                         -0.90096886790241912624,
                         -0.22252093395631440429,
                         0.62348980185873353053};
    return t[k];
  }
  static __device__ __forceinline__ double s(int k)
  {
    const double t[7] = {0.0,
                         0.78183148246802980871,
                         0.97492791218182360702,
                         0.43388373911755812048,
                         -0.43388373911755812048,
                         -0.97492791218182360702,
                         -0.78183148246802980871};
    return t[k];
  }
};

template <int R, int DIR> struct ButterflyPrime {
  static __device__ __forceinline__ void run(double2 *v)
  {
    double2 o[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      double2 acc = v[0];
#pragma unroll
      for (int r = 1; r < R; ++r) {
        const int k = (q * r) % R;
        // root = cos + i*DIR*sin  (DIR=-1: exp(-2 pi i k/R))
        double2 w = make_double2(Roots<R>::c(k), DIR < 0 ? -Roots<R>::s(k) : Roots<R>::s(k));
        double2 t = cmul(v[r], w);
        acc = cadd(acc, t);
      }
      o[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = o[q];
  }
};
template <int DIR> struct Butterfly<5, DIR> : ButterflyPrime<5, DIR> {};
template <int DIR> struct Butterfly<7, DIR> : ButterflyPrime<7, DIR> {};

// ------------------------------------------------------------------ passes ---

// One Stockham pass of radix R over `nb` whole transforms (nb*len/R <= U*T).
// Transform b lives at s + b*ld.  Ns = product of the radices already done.
template <int R, int DIR>
__device__ __forceinline__ void fft_pass(double2 *s, int ld, int nb, const FftCore &c, int Ns)
{
  constexpr int U = kEPT / R;
  const int nbf = c.len / R;
  const int total = nb * nbf;
  const int tstride = c.len / (Ns * R);
  double2 v[U][R];
  int obase[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int g = threadIdx.x + u * blockDim.x;
    obase[u] = -1;
    if (g < total) {
      const int b = g / nbf;
      const int j = g - b * nbf;
      const int k = j % Ns;
      const double2 *p = s + b * ld + j;
#pragma unroll
      for (int r = 0; r < R; ++r) v[u][r] = p[r * nbf];
      if (Ns > 1) {
        const int t = k * tstride;
#pragma unroll
        for (int r = 1; r < R; ++r) v[u][r] = cmul(v[u][r], twid<DIR>(__ldg(c.tw + t * r)));
      }
      Butterfly<R, DIR>::run(v[u]);
      obase[u] = b * ld + (j - k) * R + k;
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (obase[u] >= 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) s[obase[u] + r * Ns] = v[u][r];
    }
  }
  __syncthreads();
}

template <int R, int DIR>
__device__ __forceinline__ void fft_run_pass(double2 *s, int ld, int nb, const FftCore &c, int Ns)
{
  constexpr int U = kEPT / R;
  const int nbf = c.len / R;
  int gb = (U * (int) blockDim.x) / nbf;   // whole transforms per group (host guarantees >= 1)
  if (gb < 1) gb = 1;
  for (int b0 = 0; b0 < nb; b0 += gb) {
    const int cnt = nb - b0 < gb ? nb - b0 : gb;
    fft_pass<R, DIR>(s + b0 * ld, ld, cnt, c, Ns);
  }
}

// nb smooth-length transforms in shared memory, in place.  Callers must have
// synchronised after filling s; returns synchronised.
template <int DIR>
__device__ void fft_core_batch(double2 *s, int ld, int nb, const FftCore &c)
{
  int Ns = 1;
  for (int p = 0; p < c.npass; ++p) {
    const int R = c.radix[p];
    switch (R) {
      case 8: fft_run_pass<8, DIR>(s, ld, nb, c, Ns); break;
      case 4: fft_run_pass<4, DIR>(s, ld, nb, c, Ns); break;
      case 2: fft_run_pass<2, DIR>(s, ld, nb, c, Ns); break;
      case 3: fft_run_pass<3, DIR>(s, ld, nb, c, Ns); break;
      case 5: fft_run_pass<5, DIR>(s, ld, nb, c, Ns); break;
      case 7: fft_run_pass<7, DIR>(s, ld, nb, c, Ns); break;
      default: break;
    }
    Ns *= R;
  }
}

// nb transforms of logical length d.n (any length), in place, ld >= d.ld_min.
template <int DIR>
__device__ void fft_batch(double2 *s, int ld, int nb, const FftDesc &d)
{
  if (!d.bluestein) {
    fft_core_batch<DIR>(s, ld, nb, d.core);
    return;
  }
  const int n = d.n, m = d.core.len;
  for (int idx = threadIdx.x; idx < nb * m; idx += blockDim.x) {
    const int b = idx / m, k = idx - b * m;
    double2 x = make_double2(0.0, 0.0);
    if (k < n) {
      x = s[b * ld + k];
      if (DIR > 0) x.y = -x.y;
      x = cmul(x, __ldg(d.chirp + k));
    }
    s[b * ld + k] = x;
  }
  __syncthreads();
  fft_core_batch<-1>(s, ld, nb, d.core);
  for (int idx = threadIdx.x; idx < nb * m; idx += blockDim.x) {
    const int b = idx / m, k = idx - b * m;
    s[b * ld + k] = cmul(s[b * ld + k], __ldg(d.bhat + k));
  }
  __syncthreads();
  fft_core_batch<+1>(s, ld, nb, d.core);
  for (int idx = threadIdx.x; idx < nb * n; idx += blockDim.x) {
    const int b = idx / n, k = idx - b * n;
    double2 x = cmul(s[b * ld + k], __ldg(d.chirp + k));
    if (DIR > 0) x.y = -x.y;
    s[b * ld + k] = x;
  }
  __syncthreads();
}

}  // namespace gfmd
