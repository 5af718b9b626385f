/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Plain-C complex FFT used (a) as the FFT backend of the FFT3d shim
 * (oracle/shim/fft3d_wrap.h) that lets the reference's solver sources run
 * without LAMMPS/FFTW, and (b) by the C restatement of the hot path
 * (oracle/gfmd_oracle.c).  LAMMPS' FFT3d (KSPACE package, not part of
 * /root/reference) wraps FFTW3; what the reference relies on
 * (src/solvers/gfmd_solver_fft.cpp:72-80,116,181) is an UNNORMALISED complex
 * 2-D DFT over [nslow=nx][nmid=ny], forward = exp(-i q r), backward =
 * exp(+i q r).  This file restates exactly that, two ways:
 *
 *   fftp_dft2d_ld   : direct O(N (nx+ny)) separable DFT in long double.
 *                     Slow, algorithm-free -- the arbiter for small grids.
 *   fftp_exec_2d    : mixed-radix (4,2,3,5,7) Stockham + Bluestein for other
 *                     prime factors, OpenMP over rows / column blocks -- the
 *                     substitute FFT for the CPU baseline at large grids.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "fft_plain.h"

typedef struct { double re, im; } cplx;

typedef struct fftp_plan1d {
  int n;
  int npass;
  int radix[64];
  cplx *tw;              /* tw[k] = exp(-2 pi i k / n), k < n */
  /* Bluestein */
  int bluestein, m;
  struct fftp_plan1d *sub;  /* plan of length m */
  cplx *chirp;           /* chirp[k] = exp(-i pi k^2 / n), k < n */
  cplx *bhat;            /* FFT_m of the chirp filter, pre-divided by m */
} plan1d;

struct fftp_plan {
  int nx, ny;
  plan1d *px, *py;
};

static const long double PI_L = 3.141592653589793238462643383279502884L;

static void fill_twiddles(cplx *tw, int n)
{
  for (int k = 0; k < n; k++) {
    long double a = -2.0L * PI_L * (long double) k / (long double) n;
    tw[k].re = (double) cosl(a);
    tw[k].im = (double) sinl(a);
  }
}

static int factor(int n, int *radix)
{
  /* returns number of passes, or -1 if a prime factor > 7 is left */
  int np = 0;
  while (n % 4 == 0) { radix[np++] = 4; n /= 4; }
  while (n % 2 == 0) { radix[np++] = 2; n /= 2; }
  while (n % 3 == 0) { radix[np++] = 3; n /= 3; }
  while (n % 5 == 0) { radix[np++] = 5; n /= 5; }
  while (n % 7 == 0) { radix[np++] = 7; n /= 7; }
  return n == 1 ? np : -1;
}

static void plan1d_destroy(plan1d *p)
{
  if (!p) return;
  free(p->tw); free(p->chirp); free(p->bhat);
  plan1d_destroy(p->sub);
  free(p);
}

static void exec1d(const plan1d *p, cplx *x, cplx *scratch, int sign);

static plan1d *plan1d_create(int n)
{
  plan1d *p = (plan1d *) calloc(1, sizeof(plan1d));
  p->n = n;
  p->npass = factor(n, p->radix);
  if (p->npass >= 0) {
    p->tw = (cplx *) malloc(sizeof(cplx) * (size_t) n);
    fill_twiddles(p->tw, n);
    return p;
  }
  /* Bluestein: X_k = c_k sum_j (x_j c_j) conj(c)_{k-j},  c_k = e^{-i pi k^2/n} */
  p->bluestein = 1;
  int m = 1;
  while (m < 2 * n - 1) m *= 2;
  p->m = m;
  p->sub = plan1d_create(m);
  p->chirp = (cplx *) malloc(sizeof(cplx) * (size_t) n);
  p->bhat = (cplx *) calloc((size_t) m, sizeof(cplx));
  for (int k = 0; k < n; k++) {
    long long k2 = ((long long) k * k) % (2LL * n);
    long double a = -PI_L * (long double) k2 / (long double) n;
    p->chirp[k].re = (double) cosl(a);
    p->chirp[k].im = (double) sinl(a);
  }
  for (int k = 0; k < n; k++) {
    cplx b; b.re = p->chirp[k].re; b.im = -p->chirp[k].im;   /* conj(c_k) */
    p->bhat[k] = b;
    if (k > 0) p->bhat[m - k] = b;
  }
  cplx *scr = (cplx *) malloc(sizeof(cplx) * (size_t) m);
  exec1d(p->sub, p->bhat, scr, -1);
  free(scr);
  for (int k = 0; k < m; k++) { p->bhat[k].re /= m; p->bhat[k].im /= m; }
  return p;
}

static inline cplx cmul(cplx a, cplx b)
{
  cplx c; c.re = a.re * b.re - a.im * b.im; c.im = a.re * b.im + a.im * b.re;
  return c;
}

/* One Stockham pass, radix R generic (small DFT by definition). */
static void pass_generic(int n, int R, int Ns, const cplx *tw, const cplx *in,
                         cplx *out, int sign)
{
  int m = n / R;
  int tstride = n / (Ns * R);
  int rstride = n / R;           /* root of unity of order R: tw[r*rstride] */
  for (int jb = 0; jb < m; jb += Ns) {
    for (int k = 0; k < Ns; k++) {
      int j = jb + k;
      cplx v[8];
      for (int r = 0; r < R; r++) {
        cplx w = tw[(size_t) k * tstride * r];
        if (sign > 0) w.im = -w.im;
        v[r] = cmul(in[j + r * m], w);
      }
      int j0 = (j / Ns) * Ns * R + k;
      for (int q = 0; q < R; q++) {
        cplx s; s.re = 0; s.im = 0;
        for (int r = 0; r < R; r++) {
          cplx w = tw[(size_t) ((q * r) % R) * rstride];
          if (sign > 0) w.im = -w.im;
          cplx t = cmul(v[r], w);
          s.re += t.re; s.im += t.im;
        }
        out[j0 + q * Ns] = s;
      }
    }
  }
}

static void pass_radix2(int n, int Ns, const cplx *tw, const cplx *in,
                        cplx *out, int sign)
{
  int m = n / 2;
  int tstride = n / (Ns * 2);
  for (int jb = 0; jb < m; jb += Ns) {
    int j0b = 2 * jb;
    for (int k = 0; k < Ns; k++) {
      cplx w = tw[(size_t) k * tstride];
      if (sign > 0) w.im = -w.im;
      cplx a = in[jb + k];
      cplx b = cmul(in[jb + k + m], w);
      cplx s, d;
      s.re = a.re + b.re; s.im = a.im + b.im;
      d.re = a.re - b.re; d.im = a.im - b.im;
      out[j0b + k] = s;
      out[j0b + k + Ns] = d;
    }
  }
}

static void pass_radix4(int n, int Ns, const cplx *tw, const cplx *in,
                        cplx *out, int sign)
{
  int m = n / 4;
  int tstride = n / (Ns * 4);
  for (int jb = 0; jb < m; jb += Ns) {
    int j0b = 4 * jb;
    for (int k = 0; k < Ns; k++) {
      cplx w1 = tw[(size_t) k * tstride];
      cplx w2 = tw[(size_t) k * tstride * 2];
      cplx w3 = tw[(size_t) k * tstride * 3];
      if (sign > 0) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
      cplx a = in[jb + k];
      cplx b = cmul(in[jb + k + m], w1);
      cplx c = cmul(in[jb + k + 2 * m], w2);
      cplx d = cmul(in[jb + k + 3 * m], w3);
      cplx t0, t1, t2, t3;
      t0.re = a.re + c.re; t0.im = a.im + c.im;
      t1.re = a.re - c.re; t1.im = a.im - c.im;
      t2.re = b.re + d.re; t2.im = b.im + d.im;
      t3.re = b.re - d.re; t3.im = b.im - d.im;
      /* multiply t3 by -i (forward) or +i (backward) */
      cplx t3r;
      if (sign < 0) { t3r.re = t3.im; t3r.im = -t3.re; }
      else          { t3r.re = -t3.im; t3r.im = t3.re; }
      cplx o;
      o.re = t0.re + t2.re; o.im = t0.im + t2.im; out[j0b + k] = o;
      o.re = t1.re + t3r.re; o.im = t1.im + t3r.im; out[j0b + k + Ns] = o;
      o.re = t0.re - t2.re; o.im = t0.im - t2.im; out[j0b + k + 2 * Ns] = o;
      o.re = t1.re - t3r.re; o.im = t1.im - t3r.im; out[j0b + k + 3 * Ns] = o;
    }
  }
}

/* In-place on x (length p->n); scratch must hold max(n, m) elements (for
 * Bluestein 2*m). */
static void exec1d(const plan1d *p, cplx *x, cplx *scratch, int sign)
{
  int n = p->n;
  if (n == 1) return;
  if (!p->bluestein) {
    cplx *in = x, *out = scratch;
    int Ns = 1;
    for (int ip = 0; ip < p->npass; ip++) {
      int R = p->radix[ip];
      if (R == 4) pass_radix4(n, Ns, p->tw, in, out, sign);
      else if (R == 2) pass_radix2(n, Ns, p->tw, in, out, sign);
      else pass_generic(n, R, Ns, p->tw, in, out, sign);
      cplx *t = in; in = out; out = t;
      Ns *= R;
    }
    if (in != x) memcpy(x, in, sizeof(cplx) * (size_t) n);
    return;
  }
  /* Bluestein; backward transform = conj(forward(conj(x))) */
  int m = p->m;
  cplx *a = scratch, *scr2 = scratch + m;
  for (int k = 0; k < n; k++) {
    cplx xi = x[k];
    if (sign > 0) xi.im = -xi.im;
    a[k] = cmul(xi, p->chirp[k]);
  }
  for (int k = n; k < m; k++) { a[k].re = 0; a[k].im = 0; }
  exec1d(p->sub, a, scr2, -1);
  for (int k = 0; k < m; k++) a[k] = cmul(a[k], p->bhat[k]);
  exec1d(p->sub, a, scr2, +1);
  for (int k = 0; k < n; k++) {
    cplx r = cmul(a[k], p->chirp[k]);
    if (sign > 0) r.im = -r.im;
    x[k] = r;
  }
}

static size_t scratch_len(const plan1d *p)
{
  return p->bluestein ? (size_t) 2 * p->m + 16 : (size_t) p->n + 16;
}

fftp_plan *fftp_plan_2d(int nx, int ny)
{
  fftp_plan *p = (fftp_plan *) calloc(1, sizeof(fftp_plan));
  p->nx = nx; p->ny = ny;
  p->px = plan1d_create(nx);
  p->py = plan1d_create(ny);
  return p;
}

void fftp_destroy(fftp_plan *p)
{
  if (!p) return;
  plan1d_destroy(p->px); plan1d_destroy(p->py);
  free(p);
}

/* Threads of the 2-D transform: 0 = the OpenMP default.  Set explicitly by callers that must not
 * depend on the environment (torchrun exports OMP_NUM_THREADS=1 to its children). */
static int g_fftp_threads = 0;
void fftp_set_threads(int n) { g_fftp_threads = n > 0 ? n : 0; }
int fftp_get_threads(void)
{
#ifdef _OPENMP
  return g_fftp_threads > 0 ? g_fftp_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* data: interleaved complex, row-major [nx][ny] (index ix*ny+iy); in place;
 * sign -1: forward exp(-i q r); +1: backward exp(+i q r); unnormalised. */
void fftp_exec_2d(const fftp_plan *p, double *data, int sign)
{
  int nx = p->nx, ny = p->ny;
  cplx *d = (cplx *) data;
  enum { CB = 8 };
#ifdef _OPENMP
#pragma omp parallel num_threads(fftp_get_threads())
#endif
  {
    size_t sl = scratch_len(p->py);
    size_t sx = scratch_len(p->px);
    cplx *scr = (cplx *) malloc(sizeof(cplx) * (sl > sx ? sl : sx));
    cplx *col = (cplx *) malloc(sizeof(cplx) * (size_t) nx * CB);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int ix = 0; ix < nx; ix++)
      exec1d(p->py, d + (size_t) ix * ny, scr, sign);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int iy0 = 0; iy0 < ny; iy0 += CB) {
      int nb = ny - iy0 < CB ? ny - iy0 : CB;
      for (int ix = 0; ix < nx; ix++)
        for (int b = 0; b < nb; b++)
          col[(size_t) b * nx + ix] = d[(size_t) ix * ny + iy0 + b];
      for (int b = 0; b < nb; b++)
        exec1d(p->px, col + (size_t) b * nx, scr, sign);
      for (int ix = 0; ix < nx; ix++)
        for (int b = 0; b < nb; b++)
          d[(size_t) ix * ny + iy0 + b] = col[(size_t) b * nx + ix];
    }
    free(scr); free(col);
  }
}

/* Direct separable DFT in long double: the algorithm-free arbiter. */
void fftp_dft2d_ld(int nx, int ny, double *data, int sign)
{
  size_t N = (size_t) nx * ny;
  long double *tr = (long double *) malloc(sizeof(long double) * 2 * N);
  long double *wyr = (long double *) malloc(sizeof(long double) * 2 * (size_t) ny);
  long double *wxr = (long double *) malloc(sizeof(long double) * 2 * (size_t) nx);
  for (int k = 0; k < ny; k++) {
    long double a = sign * 2.0L * PI_L * (long double) k / (long double) ny;
    wyr[2 * k] = cosl(a); wyr[2 * k + 1] = sinl(a);
  }
  for (int k = 0; k < nx; k++) {
    long double a = sign * 2.0L * PI_L * (long double) k / (long double) nx;
    wxr[2 * k] = cosl(a); wxr[2 * k + 1] = sinl(a);
  }
  /* along y */
  for (int ix = 0; ix < nx; ix++)
    for (int q = 0; q < ny; q++) {
      long double sr = 0, si = 0;
      for (int j = 0; j < ny; j++) {
        int t = (int) (((long long) q * j) % ny);
        long double c = wyr[2 * t], s = wyr[2 * t + 1];
        long double xr = data[2 * ((size_t) ix * ny + j)];
        long double xi = data[2 * ((size_t) ix * ny + j) + 1];
        sr += xr * c - xi * s; si += xr * s + xi * c;
      }
      tr[2 * ((size_t) ix * ny + q)] = sr; tr[2 * ((size_t) ix * ny + q) + 1] = si;
    }
  /* along x */
  for (int q = 0; q < nx; q++)
    for (int iy = 0; iy < ny; iy++) {
      long double sr = 0, si = 0;
      for (int j = 0; j < nx; j++) {
        int t = (int) (((long long) q * j) % nx);
        long double c = wxr[2 * t], s = wxr[2 * t + 1];
        long double xr = tr[2 * ((size_t) j * ny + iy)];
        long double xi = tr[2 * ((size_t) j * ny + iy) + 1];
        sr += xr * c - xi * s; si += xr * s + xi * c;
      }
      data[2 * ((size_t) q * ny + iy)] = (double) sr;
      data[2 * ((size_t) q * ny + iy) + 1] = (double) si;
    }
  free(tr); free(wyr); free(wxr);
}
