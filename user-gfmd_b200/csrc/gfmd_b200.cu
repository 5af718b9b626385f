// libgfmd_b200: C ABI (include/gfmd_b200.h) + orchestration of the CUDA kernels.
// No CPU fallback anywhere in this file: every compute entry point launches
// kernels on the handle's device or fails with an error code.
#include <cuda_runtime.h>
#ifndef GFMD_CUDA_EMU
#include <cuda.h>       // types of the stream memory operations; entry points are fetched at run time
#endif
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gfmd_b200.h"
#include "fft_engine.cuh"
#include "kernels_generic.cuh"
#include "kernels_fast.cuh"
#include "kernel_phi_build.cuh"
#include "kernel_cols_aux.cuh"
#include "kernel_cols_split.cuh"

using namespace gfmd;

namespace {

thread_local std::string g_create_error;

constexpr size_t kMaxSmem = 232448;   // 227 KB opt-in dynamic shared memory per CTA
constexpr size_t kFlagTail = 256;     // complex elements (4096 B) behind d_stage2: flag words, then the u0 zone
constexpr size_t kFlagU0Bytes = 3072; // byte offset of the u0 landing zone inside that tail

// ------------------------------------------------------------------- NCCL ---

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi &nccl()
{
  static NcclApi api;
  return api;
}

bool nccl_load()
{
  NcclApi &a = nccl();
  if (a.lib) return true;
  // GFMD_B200_NCCL_LIB names a specific build (path or soname) to try first
  const char *names[] = {getenv("GFMD_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n || !*n) continue;
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) {
    a.err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    return false;
  }
#define LOAD(sym)                                                       \
  a.sym = reinterpret_cast<decltype(a.sym)>(dlsym(a.lib, "nccl" #sym)); \
  if (!a.sym) {                                                         \
    a.err = "libnccl lacks nccl" #sym;                                  \
    a.lib = nullptr;                                                    \
    return false;                                                       \
  }
  LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Send) LOAD(Recv) LOAD(AllReduce)
  LOAD(GroupStart) LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
  return true;
}

// ------------------------------------------------- stream memory operations ---
// Cross-rank ordering without a collective: a rank WRITES a sequence number into a flag word in a
// peer's memory from the stream that carried the data (cuStreamWriteValue32 -- ordered after the
// copies / kernels queued before it), the peer's stream WAITS for it (cuStreamWaitValue32, >=).
// Both are executed by the GPU's front end: no SM is occupied by a spinning kernel and no host
// thread takes part.  The driver entry points come from cudaGetDriverEntryPoint, so the library
// has no link-time dependency on libcuda.

struct MemOps {
  bool tried = false, ok = false;
#ifndef GFMD_CUDA_EMU
  CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
#endif
  std::string err;
};

MemOps &memops()
{
  static MemOps m;
  return m;
}

bool memops_load()
{
  MemOps &m = memops();
  if (m.tried) return m.ok;
  m.tried = true;
#ifdef GFMD_CUDA_EMU
  m.ok = true;
#else
  void *fw = nullptr, *fq = nullptr;
  cudaDriverEntryPointQueryResult q1, q2;
  cudaError_t e1 = cudaGetDriverEntryPoint("cuStreamWriteValue32", &fw, cudaEnableDefault, &q1);
  cudaError_t e2 = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fq, cudaEnableDefault, &q2);
  if (e1 != cudaSuccess || e2 != cudaSuccess || !fw || !fq || q1 != cudaDriverEntryPointSuccess ||
      q2 != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    m.err = "the driver does not provide cuStreamWriteValue32 / cuStreamWaitValue32";
    return false;
  }
  m.write32 = reinterpret_cast<decltype(m.write32)>(fw);
  m.wait32 = reinterpret_cast<decltype(m.wait32)>(fq);
  m.ok = true;
#endif
  return m.ok;
}

// 0 on success
int stream_write32(cudaStream_t s, unsigned *addr, unsigned value)
{
#ifdef GFMD_CUDA_EMU
  return emuStreamWriteValue32(s, addr, value) == cudaSuccess ? 0 : 1;
#else
  return memops().write32((CUstream) s, (CUdeviceptr) addr, value, CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS ? 0 : 1;
#endif
}

int stream_wait32_geq(cudaStream_t s, unsigned *addr, unsigned value)
{
#ifdef GFMD_CUDA_EMU
  return emuStreamWaitValue32Geq(s, addr, value) == cudaSuccess ? 0 : 1;
#else
  return memops().wait32((CUstream) s, (CUdeviceptr) addr, value, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS ? 0 : 1;
#endif
}

// --------------------------------------------------------------- FFT plans ---

struct DevFft {
  FftDesc desc;
  std::vector<void *> allocs;
  size_t bytes = 0;
};

void fill_tw(std::vector<double2> &tw, int n)
{
  const long double pi = 3.141592653589793238462643383279502884L;
  tw.resize(n);
  for (int k = 0; k < n; ++k) {
    // exact octant symmetry is not needed: long double gives < 1 ulp double error
    long double a = -2.0L * pi * (long double) k / (long double) n;
    tw[k] = make_double2((double) cosl(a), (double) sinl(a));
  }
}

bool factor_smooth(int n, FftCore &c)
{
  c.len = n;
  c.npass = 0;
  int rem = n;
  const int rad[] = {8, 4, 2, 3, 5, 7};
  for (int r : rad)
    while (rem % r == 0 && c.npass < kMaxPass) {
      c.radix[c.npass++] = r;
      rem /= r;
    }
  return rem == 1;
}

// Blocking host-to-device copy for set-up paths.  cudaMemcpy from pageable memory may return
// once the data sits in the driver's staging buffer, before the DMA has landed (CUDA API
// synchronisation notes), and the legacy stream it runs on is not ordered against this
// library's non-blocking streams -- so wait for the device before any kernel may read it.
cudaError_t h2d_blocking(void *dst, const void *src, size_t bytes)
{
  cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  return cudaDeviceSynchronize();
}

cudaError_t upload(const void *src, size_t bytes, void **dst, DevFft &f)
{
  cudaError_t e = cudaMalloc(dst, bytes);
  if (e != cudaSuccess) return e;
  f.allocs.push_back(*dst);
  f.bytes += bytes;
  return h2d_blocking(*dst, src, bytes);
}

// host-side complex FFT by definition-free recursion is not needed: the chirp
// filter spectrum is computed with a simple O(m log m) radix-2 in long double.
void host_fft_pow2(std::vector<long double> &re, std::vector<long double> &im, int sign)
{
  const int n = (int) re.size();
  const long double pi = 3.141592653589793238462643383279502884L;
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (int len = 2; len <= n; len <<= 1) {
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        long double a = sign * 2.0L * pi * k / len;
        long double wr = cosl(a), wi = sinl(a);
        long double xr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
        long double xi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
        re[i + k + len / 2] = re[i + k] - xr;
        im[i + k + len / 2] = im[i + k] - xi;
        re[i + k] += xr;
        im[i + k] += xi;
      }
  }
}

cudaError_t build_fft(int n, DevFft &out)
{
  memset(&out.desc, 0, sizeof(out.desc));
  FftDesc &d = out.desc;
  d.n = n;
  std::vector<double2> tw;
  if (factor_smooth(n, d.core)) {
    d.bluestein = 0;
    d.ld_min = n;
    fill_tw(tw, n);
    return upload(tw.data(), sizeof(double2) * n, (void **) &d.core.tw, out);
  }
  d.bluestein = 1;
  int m = 1;
  while (m < 2 * n - 1) m *= 2;
  factor_smooth(m, d.core);
  d.ld_min = m;
  fill_tw(tw, m);
  cudaError_t e = upload(tw.data(), sizeof(double2) * m, (void **) &d.core.tw, out);
  if (e != cudaSuccess) return e;
  const long double pi = 3.141592653589793238462643383279502884L;
  std::vector<double2> chirp(n);
  std::vector<long double> br(m, 0.0L), bi(m, 0.0L);
  for (int k = 0; k < n; ++k) {
    long long k2 = ((long long) k * k) % (2LL * n);
    long double a = -pi * (long double) k2 / (long double) n;
    long double c = cosl(a), s = sinl(a);
    chirp[k] = make_double2((double) c, (double) s);
    br[k] = c; bi[k] = -s;
    if (k > 0) { br[m - k] = c; bi[m - k] = -s; }
  }
  host_fft_pow2(br, bi, -1);
  std::vector<double2> bhat(m);
  for (int k = 0; k < m; ++k) bhat[k] = make_double2((double) (br[k] / m), (double) (bi[k] / m));
  e = upload(chirp.data(), sizeof(double2) * n, (void **) &d.chirp, out);
  if (e != cudaSuccess) return e;
  return upload(bhat.data(), sizeof(double2) * m, (void **) &d.bhat, out);
}

void free_fft(DevFft &f)
{
  for (void *p : f.allocs) cudaFree(p);
  f.allocs.clear();
}

// threads needed so that every pass fits kEPT elements per thread
int min_threads_for(const FftCore &c)
{
  int t = 32;
  for (int p = 0; p < c.npass; ++p) {
    const int R = c.radix[p];
    const int U = kEPT / R;
    const int need = (c.len / R + U - 1) / U;
    if (need > t) t = need;
  }
  return t;
}

int round_up_pow2(int v)
{
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

// ----------------------------------------------------------------- handle ---

struct gfmd_b200 {
  int nx = 0, ny = 0, d = 0, device = 0;
  GridDesc g{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;

  double *d_u = nullptr, *d_f = nullptr;
  double2 *d_stage = nullptr, *d_stage2 = nullptr, *d_stage3 = nullptr;
  // peer-copy exchange over CUDA IPC (gfmd_b200_ipc_export / _import)
  static constexpr int kMaxRanks = 16;
  bool ipc_on = false;
  bool peer_store = false;                    // GFMD_B200_PEER_STORE=1: the column stage stores its result pieces
                                              // straight into the peers' return buffers (no return pushes)
  bool peer_direct = false;                   // GFMD_B200_PEER_DIRECT=1: it also loads its input pieces straight
                                              // from the peers' row-kernel outputs (no transfers at all)
  double2 *peer_stage[kMaxRanks] = {};        // peers' row-output buffers (d_stage), gfmd_b200_ipc_import_stage
  double2 *peer_recv[2][kMaxRanks] = {};      // peers' receive buffers (forward, return), mapped here
  cudaStream_t copy_stream[kMaxRanks] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxRanks] = {};
  double *d_barrier = nullptr;
  // flag words + u0 landing zone at the end of d_stage2 (mapped by every peer together with it):
  // peers write sequence numbers there with stream memory operations, this rank's stream waits
  bool sync_flags = false;                    // cross-rank ordering by flags instead of NCCL all-reduces
  unsigned seq = 0;                           // step counter, the value the flags carry
  unsigned *flags = nullptr, *peer_flags[kMaxRanks] = {};
  double *u0_in = nullptr, *peer_u0_in[kMaxRanks] = {};
  // chunked pipeline: column chunks overlap their own transposes
  static constexpr int kMaxChunks = 8;
  int nchunks = 1, chunk_kl = 0;
  cudaEvent_t ev_chunk[kMaxChunks][kMaxRanks] = {}, ev_k2[kMaxChunks] = {}, ev_row[GFMD_B200_MAX_NDOF] = {};
  // overlapped step without transposes (long columns): the top-digit passes pull / push the pieces over
  // NVLink on two side streams, chunk by chunk, next to the fused column kernel on a reduced grid
  cudaStream_t pull_stream = nullptr, push_stream = nullptr;
  cudaEvent_t ev_pull[kMaxChunks] = {}, ev_rows_done = nullptr, ev_push_done = nullptr;
  bool timeline = false;                      // GFMD_B200_TIMELINE=1: per-chunk event times of the overlapped step on stderr
  cudaEvent_t ev_tl[kMaxChunks][6] = {};      // pull start / end, fused start / end, push start / end
  long long tl_steps = 0;
  int push_sms = 32;                          // SMs of the pushing pass (second number of GFMD_B200_XCHG_SMS=pull,push)
  int xchg_sms = 40;                          // SMs of the pulling pass (GFMD_B200_XCHG_SMS=pull,push).  Measured at 8 GPUs,
                                              // 16384^2, solver step (profiles/r2_stage_times_8gpu_sm_split.txt): 24,24 -> 4.04 ms,
                                              // 40,16 -> 3.88, 48,24 -> 3.69, 40,32 -> 3.69, 56,16 -> 3.74, 64,16 -> 3.78, 24,40 -> 3.91
  double *d_phi = nullptr, *d_linf = nullptr, *d_epart = nullptr, *d_fsum_part = nullptr;
  int fsum_part_cap = 0;
  StepResults *d_res = nullptr, *h_res = nullptr;
  double2 *d_tw_ny = nullptr;
  DevFft fft_rows, fft_cols;

  // launch parameters
  bool even = true;
  int rows_RB = 1, rows_ld = 0, rows_T = 64;
  size_t rows_smem = 0;
  int cols_ld = 0, cols_T = 64;
  size_t cols_smem = 0;
  int fast_rows = 0, fast_cols = 0;   // specialised kernels selected
  bool cols_split_fast = false;       // three-phase column stage on the power-of-two passes (k_cols_fft_p2; nx = 4096,
                                      // single rank): the spectrum stays in position order, Phi is stored likewise
  int phi_mode() const { return fast_cols ? 1 : (cols_split_fast ? 2 : 0); }      // layout of d_phi, see phi_slot
  int cols_split_db = 0;              // > 0: column set too large for one CTA, three-phase column stage
                                      // with this many dofs per CTA (kernel_cols_split.cuh)
  int cols_top = 0;                   // log2(nx / 4096): top-digit pass of long columns
  DevFft fft_sub;                     // twiddles of the 4096 sub-columns (long columns only)
  int num_sms = 148;
  // auxiliary (off-path) column kernel: q-space dumps and the preconditioner
  size_t aux_cols_smem = 0;           // 0: a column set does not fit one CTA
  int aux_split_db = 0, aux_split_T = 0;   // then: the services run through the three-phase column stage (nx <= 8192)
  size_t aux_split_smem = 0;
  bool aux_attr_set = false;
  double2 *d_spec = nullptr;          // Phi.u~ of the last spectrum request
  double *d_cavg = nullptr;

  // fused atom I/O (gfmd_b200_build_cell_map): cell -> atom map and what it was built from
  int *d_cmap = nullptr, *d_cmap_cnt = nullptr;
  bool cmap_valid = false;
  const int *cmap_gid = nullptr, *cmap_mask = nullptr;
  int cmap_groupbit = 0, cmap_nall = 0, cmap_nlocal = 0;
  int cmap_cnt[3] = {0, 0, 0};
  double *d_fsum_io = nullptr;          // [nx_loc / rb][ndof] partial force sums of the fused scatter
  const AtomIO *io_fwd = nullptr, *io_inv = nullptr;   // set around one enqueue_solver by the fused full step

  bool phi_set = false;
  std::vector<char> phi_cols_set;
  double herm_dev = 0.0, conj_dev = 0.0;

  ncclComm_t comm = nullptr;

  // CUDA graph of the solver step
  bool want_graph = false;
  cudaGraphExec_t graph_exec = nullptr;
  const double *graph_u = nullptr;
  double *graph_f = nullptr;
  long long graph_launches = 0;

  // async host path
  const double *pending_u = nullptr;
  // host pipeline: u arrives and f leaves dof by dof on two copy streams, so that the row
  // transforms of a dof overlap the PCIe transfer of the next one (single rank, specialised rows)
  bool hp_enabled = true, hp_pending = false;
  cudaStream_t hp_stream[2] = {};
  cudaEvent_t hp_in[GFMD_B200_MAX_NDOF] = {}, hp_out[GFMD_B200_MAX_NDOF] = {}, hp_fork = nullptr, hp_done = nullptr;
  std::map<const void *, size_t> pinned;   // host ranges this handle page-locked
  bool pin_host = false;

  bool profiling = false;
  cudaEvent_t ev[8] = {};                  // marks 0..7 around stages 0..6
  cudaEvent_t ev_top[2] = {};              // long columns: behind the forward / before the backward top-digit pass
  bool top_split = false;                  // ev_top were recorded in the last profiled step
  double stage_ms[GFMD_B200_NSTAGES] = {};
  long long stage_cnt[GFMD_B200_NSTAGES] = {};

  long long launches = 0;
  double bytes = 0.0;
  std::string err;
  std::string desc;
};

namespace {

int fail(gfmd_b200 *h, int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU(h, call)                                                                     \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return fail(h, GFMD_B200_ECUDA, "%s failed: %s (%s:%d)", #call,                   \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                         \
  } while (0)

#define NC(h, call)                                                                     \
  do {                                                                                  \
    ncclResult_t r__ = (call);                                                          \
    if (r__ != ncclSuccess)                                                             \
      return fail(h, GFMD_B200_ENCCL, "%s failed: %s", #call, nccl().GetErrorString(r__)); \
  } while (0)

// The dynamic shared-memory limit is an attribute of the kernel (per device), not of a handle:
// only ever raise it, or a second handle with a smaller grid would lower it under the first.
cudaError_t grow_dyn_smem(const void *func, size_t bytes)
{
  static std::mutex m;
  static std::map<std::pair<int, const void *>, size_t> granted;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(m);
  size_t &have = granted[std::make_pair(dev, func)];
  if (bytes <= have) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

template <typename T> cudaError_t dmalloc(gfmd_b200 *h, T **p, size_t count)
{
  size_t bytes = sizeof(T) * (count ? count : 1);
  cudaError_t e = cudaMalloc((void **) p, bytes);
  if (e == cudaSuccess) h->bytes += (double) bytes;
  return e;
}

int set_device(gfmd_b200 *h)
{
  CU(h, cudaSetDevice(h->device));
  return 0;
}

int plan(gfmd_b200 *h)
{
  const GridDesc &g = h->g;
  // rows
  h->even = (g.ny % 2 == 0);
  const int nrow_len = h->even ? g.ny / 2 : g.ny;
  CU(h, build_fft(nrow_len, h->fft_rows));
  CU(h, build_fft(g.nx, h->fft_cols));
  h->bytes += (double) (h->fft_rows.bytes + h->fft_cols.bytes);
  {
    std::vector<double2> tw;
    fill_tw(tw, g.ny);
    CU(h, dmalloc(h, &h->d_tw_ny, (size_t) g.ny));
    CU(h, h2d_blocking(h->d_tw_ny, tw.data(), sizeof(double2) * g.ny));
  }
  int ld = h->fft_rows.desc.ld_min;
  const int need = h->even ? g.ny / 2 + 1 : g.ny;
  if (ld < need) ld = need;
  if ((ld & 1) == 0) ld += 1;            // odd stride: transposed accesses spread over banks
  h->rows_ld = ld;
  int RB = 8192 / ld;
  if (RB < 1) RB = 1;
  if (RB > 32) RB = 32;
  if (RB > g.nx_loc) RB = g.nx_loc;
  // small grids (BASELINE configs C1-C3): a step is latency-bound, and 32 rows per CTA leave a 128 x 128 or
  // 64 x 37 surface on a dozen SMs.  Fewer rows per CTA until the row kernels fill the GPU (round 2:
  // config.latency of the bench line).
  {
    const long long total_rows = (long long) g.d * g.nx_loc;
    long long fill = (total_rows + h->num_sms - 1) / h->num_sms;
    if (fill < 1) fill = 1;
    if ((long long) RB > fill) RB = (int) fill;
  }
  while (RB > 1 && (size_t) RB * ld * sizeof(double2) > 200 * 1024) --RB;
  h->rows_RB = RB;
  h->rows_smem = (size_t) RB * ld * sizeof(double2);
  if (h->rows_smem > kMaxSmem)
    return fail(h, GFMD_B200_EUNSUPPORTED, "ny = %d: one row (%zu B) exceeds shared memory", g.ny,
                h->rows_smem);
  int tmin = min_threads_for(h->fft_rows.desc.core);
  if (tmin > 512)
    return fail(h, GFMD_B200_EUNSUPPORTED, "ny = %d: row transform of length %d too long for one CTA",
                g.ny, h->fft_rows.desc.core.len);
  int t = round_up_pow2(RB * h->fft_rows.desc.core.len / 8);
  if (t < tmin) t = round_up_pow2(tmin);
  if (t < 64) t = 64;
  if (t > 512) t = 512;
  h->rows_T = t;

  // columns
  {
    int rc = fast_plan(h->g, h->fast_rows, h->fast_cols, h->cols_top);
    if (rc) return fail(h, GFMD_B200_ECUDA, "fast kernel setup failed");
    if (h->cols_top > 0) {
      CU(h, build_fft(4096, h->fft_sub));
      h->bytes += (double) h->fft_sub.bytes;
    }
  }
  int cld = h->fft_cols.desc.ld_min;
  if (cld < g.nx) cld = g.nx;
  h->cols_ld = cld;
  h->cols_smem = (size_t) g.d * cld * sizeof(double2);
  tmin = min_threads_for(h->fft_cols.desc.core);
  if (tmin > 512 && !h->fast_cols)
    return fail(h, GFMD_B200_EUNSUPPORTED, "nx = %d: column transform too long for one CTA", g.nx);
  h->aux_cols_smem = (h->cols_smem <= kMaxSmem && tmin <= 512 && g.P == 1) ? h->cols_smem : 0;
  if (!h->aux_cols_smem && g.P == 1 && tmin <= 512 && (size_t) cld * sizeof(double2) <= kMaxSmem) {
    const size_t one = (size_t) cld * sizeof(double2);
    int db = (int) (kMaxSmem / one);
    if (db > g.d) db = g.d;
    h->aux_split_db = db;
    h->aux_split_smem = (size_t) db * one;
    int ta = round_up_pow2(db * h->fft_cols.desc.core.len / 8);
    if (ta < tmin) ta = round_up_pow2(tmin);
    if (ta < 64) ta = 64;
    if (ta > 512) ta = 512;
    h->aux_split_T = ta;
  }
  int cols_nb = g.d;                     // transforms a column CTA holds
  if (!h->fast_cols) {
    // GFMD_B200_COLS_SPLIT=<n> forces the three-phase column stage with at most n dofs per CTA
    int force_db = 0;
    if (const char *e = getenv("GFMD_B200_COLS_SPLIT")) force_db = atoi(e);
    if (h->cols_smem > kMaxSmem || force_db > 0) {
      const size_t one = (size_t) cld * sizeof(double2);
      if (one > kMaxSmem)
        return fail(h, GFMD_B200_EUNSUPPORTED,
                    "nx = %d: one column (%zu B) exceeds the %zu B of shared memory per CTA", g.nx, one, kMaxSmem);
      int db = (int) (kMaxSmem / one);
      if (db > g.d) db = g.d;
      if (force_db > 0 && force_db < db) db = force_db;
      h->cols_split_db = db;
      h->cols_smem = (size_t) db * one;
      cols_nb = db;
      // the transform phases on the specialised power-of-two passes where they exist
      if (g.nx == 4096 && g.P == 1 && force_db == 0 && !getenv("GFMD_B200_NO_FAST")) {
        h->cols_split_fast = true;
        h->cols_split_db = 3;
        h->cols_smem = fast_cols_smem(3, 4096);
        cols_nb = 3;
      }
    }
  }
  if (h->fast_cols) h->cols_smem = fast_cols_smem(3, h->fast_cols);
  t = round_up_pow2(cols_nb * h->fft_cols.desc.core.len / 8);
  if (t < tmin) t = round_up_pow2(tmin);
  if (t < 64) t = 64;
  if (t > 512) t = 512;
  h->cols_T = t;

#define SET_SMEM(k, bytes) CU(h, grow_dyn_smem((const void *) k, bytes))
  if (h->even) {
    SET_SMEM(k_rows_fwd<true>, h->rows_smem);
    SET_SMEM(k_rows_inv<true>, h->rows_smem);
  } else {
    SET_SMEM(k_rows_fwd<false>, h->rows_smem);
    SET_SMEM(k_rows_inv<false>, h->rows_smem);
  }
  if (h->cols_split_fast) {
    SET_SMEM((k_cols_fft_p2<4096, 512, -1>), h->cols_smem);
    SET_SMEM((k_cols_fft_p2<4096, 512, +1>), h->cols_smem);
  } else if (h->cols_split_db) {
    SET_SMEM(k_cols_split_fft<-1>, h->cols_smem);
    SET_SMEM(k_cols_split_fft<+1>, h->cols_smem);
  } else if (!h->fast_cols) switch (g.d) {
    case 3: SET_SMEM(k_cols_fused<3>, h->cols_smem); break;
    case 6: SET_SMEM(k_cols_fused<6>, h->cols_smem); break;
    case 9: SET_SMEM(k_cols_fused<9>, h->cols_smem); break;
    case 12: SET_SMEM(k_cols_fused<12>, h->cols_smem); break;
    default: SET_SMEM(k_cols_fused<0>, h->cols_smem); break;
  }
#undef SET_SMEM
  return 0;
}

int create_common(gfmd_b200_t **out, int nx, int ny, int ndof, int device, int rank, int nranks)
{
  if (!out) return fail(nullptr, GFMD_B200_EINVAL, "null handle pointer");
  *out = nullptr;
  if (nx < 1 || ny < 1) return fail(nullptr, GFMD_B200_EINVAL, "grid %d x %d invalid", nx, ny);
  if (ndof < 3 || ndof % 3 != 0 || ndof > GFMD_B200_MAX_NDOF)
    return fail(nullptr, GFMD_B200_EINVAL, "ndof = %d must be a multiple of 3 in [3, %d]", ndof,
                GFMD_B200_MAX_NDOF);
  if (nranks < 1 || rank < 0 || rank >= nranks)
    return fail(nullptr, GFMD_B200_EINVAL, "rank %d of %d invalid", rank, nranks);
  if (nranks > gfmd_b200::kMaxRanks)
    return fail(nullptr, GFMD_B200_EUNSUPPORTED, "at most %d slab ranks", gfmd_b200::kMaxRanks);
  if (nx % nranks != 0)
    return fail(nullptr, GFMD_B200_EUNSUPPORTED, "nx = %d not divisible by %d slab ranks", nx, nranks);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, GFMD_B200_ENOGPU, "no CUDA device: %s (this library has no CPU fallback)",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev)
    return fail(nullptr, GFMD_B200_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);

  gfmd_b200 *h = new gfmd_b200;
  h->nx = nx; h->ny = ny; h->d = ndof; h->device = device;
  GridDesc &g = h->g;
  g.nx = nx; g.ny = ny; g.nyh = ny / 2 + 1; g.d = ndof;
  g.P = nranks; g.rank = rank;
  g.nx_loc = nx / nranks; g.x0 = rank * g.nx_loc;
  g.kyb = (g.nyh + nranks - 1) / nranks;
  g.ky0 = rank * g.kyb;
  g.nky_loc = g.nyh - g.ky0;
  if (g.nky_loc > g.kyb) g.nky_loc = g.kyb;
  if (g.nky_loc < 0) g.nky_loc = 0;

  int rc = 0;
  auto bail = [&](int code) {
    g_create_error = h->err;
    gfmd_b200_destroy(h);
    return code;
  };
  if ((rc = set_device(h))) return bail(rc);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    h->err = "cudaStreamCreate failed";
    return bail(GFMD_B200_ECUDA);
  }
  h->own_stream = true;
  if (const char *e = getenv("GFMD_B200_HOST_PIPE")) h->hp_enabled = atoi(e) != 0;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  if ((rc = plan(h))) return bail(rc);

  const size_t nxy = (size_t) g.nx_loc * g.ny;
  const size_t nstage = (size_t) g.P * g.d * g.kyb * g.nx_loc;
  const size_t nphi = (size_t) g.nky_loc * g.d * g.d * g.nx;
  cudaError_t ce = cudaSuccess;
  if (ce == cudaSuccess) ce = dmalloc(h, &h->d_u, nxy * g.d);
  if (ce == cudaSuccess) ce = dmalloc(h, &h->d_f, nxy * g.d);
  if (ce == cudaSuccess) ce = dmalloc(h, &h->d_stage, nstage);
  if (ce == cudaSuccess && g.P > 1) ce = dmalloc(h, &h->d_stage2, nstage + kFlagTail);
  if (ce == cudaSuccess) ce = dmalloc(h, &h->d_phi, nphi);
  if (ce == cudaSuccess) ce = dmalloc(h, &h->d_linf, (size_t) GFMD_B200_MAX_NDOF);
  if (ce == cudaSuccess) ce = dmalloc(h, &h->d_epart, ((size_t) g.kyb + 1) * kColsNW * 4);
  if (ce == cudaSuccess) ce = dmalloc(h, &h->d_res, (size_t) 1);
  if (ce == cudaSuccess) ce = cudaMallocHost((void **) &h->h_res, sizeof(StepResults));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_u, 0, sizeof(double) * nxy * g.d);
  if (ce == cudaSuccess) ce = cudaMemset(h->d_f, 0, sizeof(double) * nxy * g.d);
  if (ce == cudaSuccess) ce = cudaMemset(h->d_stage, 0, sizeof(double2) * nstage);
  if (ce == cudaSuccess && g.P > 1) ce = cudaMemset(h->d_stage2, 0, sizeof(double2) * (nstage + kFlagTail));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_linf, 0, sizeof(double) * GFMD_B200_MAX_NDOF);
  if (ce == cudaSuccess) ce = cudaMemset(h->d_epart, 0, sizeof(double) * (g.kyb + 1) * kColsNW * 4);
  if (ce == cudaSuccess) ce = cudaMemset(h->d_res, 0, sizeof(StepResults));
  if (ce == cudaSuccess) ce = cudaDeviceSynchronize();     // memsets ran on the legacy stream
  if (ce != cudaSuccess) {
    h->err = std::string("device allocation failed: ") + cudaGetErrorString(ce);
    return bail(GFMD_B200_ECUDA);
  }
  memset(h->h_res, 0, sizeof(StepResults));
  if (g.P > 1) {
    h->flags = reinterpret_cast<unsigned *>(h->d_stage2 + nstage);
    h->u0_in = reinterpret_cast<double *>(reinterpret_cast<char *>(h->flags) + kFlagU0Bytes);
  }
  h->phi_cols_set.assign(g.nky_loc > 0 ? g.nky_loc : 0, 0);
  h->phi_set = g.nky_loc == 0;      // a rank without q columns has no table to wait for
  for (int i = 0; i < 8; ++i) cudaEventCreate(&h->ev[i]);
  for (int i = 0; i < 2; ++i) cudaEventCreate(&h->ev_top[i]);

  char rows[160], cols[200], buf[512];
  FastRowsCfg frc;
  if (h->fast_rows && fast_rows_cfg(h->fast_rows, frc))
    snprintf(rows, sizeof(rows), "k_rows_*_%s half-length len %d, %d rows/CTA, %d threads, smem %zu [fast, variant %d%s]",
             h->fast_rows == h->g.ny + 9 ? "r16c (2-CTA clusters)" : h->fast_rows == h->g.ny + 8 ? "r16" : "p2", frc.nr, frc.rb,
             frc.t, h->fast_rows == h->g.ny + 9 ? fast_rows_smem_cluster(frc) : fast_rows_smem(frc), h->fast_rows,
             h->fast_rows == fast_rows_default(h->g.ny) ? " = default" : "");
  else
    snprintf(rows, sizeof(rows), "k_rows_* %s len %d%s, %d rows/CTA, %d threads, smem %zu",
             h->even ? "half-length" : "full-length", h->fft_rows.desc.n,
             h->fft_rows.desc.bluestein ? " (bluestein)" : "", h->rows_RB, h->rows_T, h->rows_smem);
  if (h->fast_cols == 4096)
    snprintf(cols, sizeof(cols), "k_cols_fused_p2_lr sub-column 4096 x %d (top radix %d in HBM), 512 threads, "
             "smem %zu [fast]", 1 << h->cols_top, 1 << h->cols_top, h->cols_smem);
  else if (h->fast_cols)
    snprintf(cols, sizeof(cols), "k_cols_fused_p2 len %d, 256 threads, smem %zu [fast]", h->fast_cols,
             h->cols_smem);
  else if (h->cols_split_fast)
    snprintf(cols, sizeof(cols), "k_cols_split_fft -> k_cols_fft_p2 len %d (3 dofs/CTA, 512 threads, smem %zu, spectrum in "
             "position order) + k_cols_contract [fast]", h->g.nx, h->cols_smem);
  else if (h->cols_split_db)
    snprintf(cols, sizeof(cols), "k_cols_split_fft len %d%s (%d dofs/CTA, %d threads, smem %zu) + k_cols_contract",
             h->fft_cols.desc.n, h->fft_cols.desc.bluestein ? " (bluestein)" : "", h->cols_split_db, h->cols_T,
             h->cols_smem);
  else
    snprintf(cols, sizeof(cols), "k_cols_fused len %d%s, %d threads, smem %zu", h->fft_cols.desc.n,
             h->fft_cols.desc.bluestein ? " (bluestein)" : "", h->cols_T, h->cols_smem);
  snprintf(buf, sizeof(buf), "grid %dx%d ndof %d rank %d/%d | rows: %s | cols: %s", nx, ny, ndof, rank, nranks,
           rows, cols);
  h->desc = buf;
  *out = h;
  return 0;
}

void drop_graph(gfmd_b200 *h)
{
  if (h->graph_exec) {
    cudaGraphExecDestroy(h->graph_exec);
    h->graph_exec = nullptr;
  }
}

inline void stage_mark(gfmd_b200 *h, int i)
{
  if (h->profiling) cudaEventRecord(h->ev[i], h->stream);
}

int atom_blocks(const gfmd_b200 *h, int nall)
{
  const int tiles = (nall + kAtomTile - 1) / kAtomTile;
  const int cap = h->num_sms * 8;           // 8 resident blocks of 256 threads per SM
  return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

// ---- cross-rank ordering by flags (see MemOps above) ----
// Flag words of a rank live behind its forward receive buffer (d_stage2), which every peer maps:
//   kFlagFwd + src * kMaxChunks + c   chunk c of src's forward blocks has landed here
//   kFlagRet + src                    src's return blocks (and, from rank 0, u0) have landed here
//   kFlagRows + src + 16 * dof        src's row transforms of that dof are complete (mode without transposes;
//                                     where the rows are not launched dof by dof: dof 0 stands for all)
// Values are the step counter h->seq; a rank can never be more than one step ahead of a peer it
// exchanges with (it needs that peer's return blocks to finish its own step), so ">= seq" is exact.
constexpr int kFlagFwd = 0;
constexpr int kFlagRet = gfmd_b200::kMaxRanks * gfmd_b200::kMaxChunks;
constexpr int kFlagRows = kFlagRet + gfmd_b200::kMaxRanks;      // + src + kMaxRanks * dof (dof 0 alone where rows are not split)
static_assert((kFlagRows + gfmd_b200::kMaxRanks * GFMD_B200_MAX_NDOF) * sizeof(unsigned) <= kFlagU0Bytes,
              "flag words overlap the u0 zone");

int signal_peer(gfmd_b200 *h, cudaStream_t s, int r, int slot)
{
  if (stream_write32(s, h->peer_flags[r] + slot, h->seq))
    return fail(h, GFMD_B200_ECUDA, "cuStreamWriteValue32 into rank %d's flags failed", r);
  return 0;
}

// this rank's stream waits until every peer has signalled slot_base + src * stride for this step
int wait_peers(gfmd_b200 *h, int slot_base, int stride)
{
  for (int k = 1; k < h->g.P; ++k) {
    const int p = (h->g.rank + k) % h->g.P;
    if (stream_wait32_geq(h->stream, h->flags + slot_base + p * stride, h->seq))
      return fail(h, GFMD_B200_ECUDA, "cuStreamWaitValue32 on the flag of rank %d failed", p);
  }
  return 0;
}

// u0 = Re u~(q = 0) lives on rank 0 only (the ky = 0 column): the reference's MPI_Allreduce
// (gfmd_solver_static.cpp:176) is a broadcast here.  Rank 0 pushes it ahead of its return flag.
int push_u0(gfmd_b200 *h, cudaStream_t s, int r)
{
  if (h->g.rank != 0) return 0;
  CU(h, cudaMemcpyAsync(h->peer_u0_in[r], h->d_res->u0, sizeof(double) * h->g.d, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int take_u0(gfmd_b200 *h)
{
  if (h->g.rank == 0) return 0;
  CU(h, cudaMemcpyAsync(h->d_res->u0, h->u0_in, sizeof(double) * h->g.d, cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

// All-to-all of the staging blocks: block r of src goes to rank r, block p of dst comes
// from rank p.  which = 0 forward transpose, 1 return transpose.
//
// Peer-copy path (after gfmd_b200_ipc_import): every rank PUSHES its blocks straight into
// the peers' receive buffers with copy-engine transfers over NVLink (one stream per peer),
// then a tiny NCCL all-reduce on the compute stream serves as the cross-rank barrier:
// it completes only when every rank's pushes, which precede its contribution in stream
// order, have landed.  Forward and return use different receive buffers, so a buffer is
// never overwritten before its reader of the previous step has passed a later barrier.
// Without IPC handles the same exchange is a grouped ncclSend/ncclRecv.
int exchange(gfmd_b200 *h, const double2 *src, double2 *dst, int which)
{
  const GridDesc &g = h->g;
  if (!h->comm && !h->sync_flags)
    return fail(h, GFMD_B200_ESTATE, "slab handle used before gfmd_b200_comm_init / gfmd_b200_ipc_import");
  const size_t blk = (size_t) g.d * g.kyb * g.nx_loc;   // complex elements per peer block
  NcclApi &a = nccl();
  CU(h, cudaMemcpyAsync(dst + g.rank * blk, src + g.rank * blk, blk * sizeof(double2), cudaMemcpyDeviceToDevice,
                        h->stream));
  if (h->ipc_on) {
    CU(h, cudaEventRecord(h->ev_fork, h->stream));
    for (int k = 1; k < g.P; ++k) {
      const int r = (g.rank + k) % g.P;                 // staggered: no two ranks hit one peer first
      CU(h, cudaStreamWaitEvent(h->copy_stream[r], h->ev_fork, 0));
      CU(h, cudaMemcpyAsync(h->peer_recv[which][r] + g.rank * blk, src + r * blk, blk * sizeof(double2),
                            cudaMemcpyDeviceToDevice, h->copy_stream[r]));
      if (h->sync_flags) {
        int rc = which == 1 ? push_u0(h, h->copy_stream[r], r) : 0;
        if (!rc) rc = signal_peer(h, h->copy_stream[r], r, which == 0 ? kFlagFwd + g.rank * gfmd_b200::kMaxChunks
                                                                       : kFlagRet + g.rank);
        if (rc) return rc;
      }
      CU(h, cudaEventRecord(h->ev_join[r], h->copy_stream[r]));
      CU(h, cudaStreamWaitEvent(h->stream, h->ev_join[r], 0));
    }
    if (h->sync_flags) {
      int rc = which == 0 ? wait_peers(h, kFlagFwd, gfmd_b200::kMaxChunks) : wait_peers(h, kFlagRet, 1);
      if (!rc && which == 1) rc = take_u0(h);
      return rc;
    }
    if (which == 0)   // the return transpose is followed by the u0 all-reduce, which is its barrier
      NC(h, a.AllReduce(h->d_barrier, h->d_barrier, 1, ncclDouble, ncclSum, h->comm, h->stream));
    return 0;
  }
  NC(h, a.GroupStart());
  for (int r = 0; r < g.P; ++r) {
    if (r == g.rank) continue;
    NC(h, a.Send(src + r * blk, blk * 2, ncclDouble, r, h->comm, h->stream));
    NC(h, a.Recv(dst + r * blk, blk * 2, ncclDouble, r, h->comm, h->stream));
  }
  NC(h, a.GroupEnd());
  return 0;
}

// Column stage of the generic kernels: `in` holds the received columns, the result goes to `out`
// (in == out on a single rank).  Either the fused kernel (a column set fits one CTA) or the
// three-phase form of kernel_cols_split.cuh.  *nepart receives the number of energy partials written.
int launch_generic_cols(gfmd_b200 *h, double2 *in, double2 *out, int *nepart)
{
  const GridDesc &g = h->g;
  *nepart = g.nky_loc;
  if (g.nky_loc <= 0) return 0;
  if (h->cols_split_db) {
    const int db = h->cols_split_db;
    const int ngrp = (g.d + db - 1) / db;
    const int ntile = (g.nx + kContractTile - 1) / kContractTile;
    const int fgrid = g.nky_loc * ngrp < h->num_sms ? g.nky_loc * ngrp : h->num_sms;
    if (h->cols_split_fast)
      k_cols_fft_p2<4096, 512, -1><<<fgrid, 512, h->cols_smem, h->stream>>>(in, in, g, h->fft_cols.desc.core.tw);
    else
    k_cols_split_fft<-1><<<g.nky_loc * ngrp, h->cols_T, h->cols_smem, h->stream>>>(in, in, g, h->fft_cols.desc,
                                                                                  h->cols_ld, db);
#define LAUNCH_CONTRACT(DT)                                                                   \
  k_cols_contract<DT><<<g.nky_loc * ntile, kContractThreads, 0, h->stream>>>(in, g, h->d_phi, \
                                                                             h->d_linf, h->d_epart, h->d_res)
    switch (g.d) {
      case 3: LAUNCH_CONTRACT(3); break;
      case 6: LAUNCH_CONTRACT(6); break;
      case 9: LAUNCH_CONTRACT(9); break;
      case 12: LAUNCH_CONTRACT(12); break;
      default: LAUNCH_CONTRACT(0); break;
    }
#undef LAUNCH_CONTRACT
    if (h->cols_split_fast)
      k_cols_fft_p2<4096, 512, +1><<<fgrid, 512, h->cols_smem, h->stream>>>(in, out, g, h->fft_cols.desc.core.tw);
    else
    k_cols_split_fft<+1><<<g.nky_loc * ngrp, h->cols_T, h->cols_smem, h->stream>>>(in, out, g, h->fft_cols.desc,
                                                                                  h->cols_ld, db);
    h->launches += 3;
    *nepart = g.nky_loc * ntile;
    return 0;
  }
#define LAUNCH_COLS(DT)                                                                          \
  k_cols_fused<DT><<<g.nky_loc, h->cols_T, h->cols_smem, h->stream>>>(                           \
      in, out, g, h->fft_cols.desc, h->d_phi, h->d_linf, h->d_epart, h->d_res, h->cols_ld)
  switch (g.d) {
    case 3: LAUNCH_COLS(3); break;
    case 6: LAUNCH_COLS(6); break;
    case 9: LAUNCH_COLS(9); break;
    case 12: LAUNCH_COLS(12); break;
    default: LAUNCH_COLS(0); break;
  }
#undef LAUNCH_COLS
  h->launches++;
  return 0;
}

// Multi-GPU step with the transposes overlapped (peer pushes on copy-engine streams):
//   * rows are transformed dof by dof; the blocks of a finished dof are pushed while the
//     next dof is transformed (the last dof's push is cut into the ky chunks below);
//   * the local ky range is cut into chunks of whole waves of the persistent column kernel;
//     chunk c is transformed as soon as ITS data has landed everywhere (NCCL all-reduce as
//     barrier), and its result is pushed back while chunk c+1 is transformed.
int pipelined_step(gfmd_b200 *h, const double *d_u, double2 *A, double2 *B, double2 *B2)
{
  const GridDesc &g = h->g;
  NcclApi &a = nccl();
  const size_t blk = (size_t) g.d * g.kyb * g.nx_loc;
  const size_t dblk = (size_t) g.kyb * g.nx_loc;                        // one dof of a block
  const size_t pitch = dblk * sizeof(double2);
  const double2 *tw_sub = h->cols_top ? h->fft_sub.desc.core.tw : h->fft_cols.desc.core.tw;
  const int nc = h->nchunks, ck = h->chunk_kl;
  const bool flags = h->sync_flags;
  PeerOut po{};
  if (h->peer_store)
    for (int r = 0; r < g.P; ++r) po.p[r] = (r == g.rank ? B2 : h->peer_recv[1][r]) + g.rank * blk;

  stage_mark(h, 1);
  for (int dof = 0; dof < g.d; ++dof) {
    int rc = fast_rows_fwd(h->fast_rows, d_u, A, g, h->d_tw_ny, h->fft_rows.desc, h->stream, &h->launches, dof, 1);
    if (rc) return fail(h, GFMD_B200_ECUDA, "fast rows_fwd launch failed");
    CU(h, cudaEventRecord(h->ev_row[dof], h->stream));
    for (int k = 1; k < g.P; ++k) {
      const int r = (g.rank + k) % g.P;
      CU(h, cudaStreamWaitEvent(h->copy_stream[r], h->ev_row[dof], 0));
      if (dof < g.d - 1) {
        CU(h, cudaMemcpyAsync(h->peer_recv[0][r] + g.rank * blk + dof * dblk, A + r * blk + dof * dblk,
                              dblk * sizeof(double2), cudaMemcpyDeviceToDevice, h->copy_stream[r]));
      } else {
        for (int c = 0; c < nc; ++c) {
          const int k0 = c * ck, k1 = (c + 1) * ck < g.kyb ? (c + 1) * ck : g.kyb;
          const size_t off = dof * dblk + (size_t) k0 * g.nx_loc;
          CU(h, cudaMemcpyAsync(h->peer_recv[0][r] + g.rank * blk + off, A + r * blk + off,
                                (size_t) (k1 - k0) * g.nx_loc * sizeof(double2), cudaMemcpyDeviceToDevice,
                                h->copy_stream[r]));
          // in stream order behind ALL earlier pushes to r: chunk c of every dof has landed there
          if (flags && (rc = signal_peer(h, h->copy_stream[r], r, kFlagFwd + g.rank * gfmd_b200::kMaxChunks + c)))
            return rc;
          CU(h, cudaEventRecord(h->ev_chunk[c][r], h->copy_stream[r]));
        }
      }
    }
  }
  stage_mark(h, 2);
  stage_mark(h, 3);
  // own block stays on the compute stream
  CU(h, cudaMemcpyAsync(B + g.rank * blk, A + g.rank * blk, blk * sizeof(double2), cudaMemcpyDeviceToDevice,
                        h->stream));
  for (int c = 0; c < nc; ++c) {
    const int k0 = c * ck, k1 = (c + 1) * ck < g.kyb ? (c + 1) * ck : g.kyb;
    const size_t off = (size_t) k0 * g.nx_loc, width = (size_t) (k1 - k0) * g.nx_loc * sizeof(double2);
    // own pushes of this chunk have left A (the column kernel overwrites it) ...
    for (int k = 1; k < g.P; ++k) CU(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[c][(g.rank + k) % g.P], 0));
    // ... and every peer's chunk has arrived in B
    if (flags) {
      int rc = wait_peers(h, kFlagFwd + c, gfmd_b200::kMaxChunks);
      if (rc) return rc;
    } else {
      NC(h, a.AllReduce(h->d_barrier, h->d_barrier, 1, ncclDouble, ncclSum, h->comm, h->stream));
    }
    int rc = fast_cols_fused(h->fast_cols, h->cols_top, B, A, g, tw_sub, h->fft_cols.desc.core.tw, h->d_phi,
                             h->d_linf, h->d_epart, h->d_res, h->num_sms, h->stream, &h->launches, k0, k1,
                             h->peer_store ? &po : nullptr);
    if (rc) return fail(h, GFMD_B200_ECUDA, "fast cols_fused launch failed");
    if (h->peer_store) continue;                 // the kernel has stored into the peers' buffers itself
    CU(h, cudaEventRecord(h->ev_k2[c], h->stream));
    for (int k = 1; k < g.P; ++k) {
      const int r = (g.rank + k) % g.P;
      CU(h, cudaStreamWaitEvent(h->copy_stream[r], h->ev_k2[c], 0));
      CU(h, cudaMemcpy2DAsync(h->peer_recv[1][r] + g.rank * blk + off, pitch, A + r * blk + off, pitch, width,
                              g.d, cudaMemcpyDeviceToDevice, h->copy_stream[r]));
      if (c == nc - 1) {
        if (flags) {      // u0 (written by chunk 0's kernel on rank 0) rides ahead of the return flag
          if ((rc = push_u0(h, h->copy_stream[r], r))) return rc;
          if ((rc = signal_peer(h, h->copy_stream[r], r, kFlagRet + g.rank))) return rc;
        }
        CU(h, cudaEventRecord(h->ev_join[r], h->copy_stream[r]));
      }
    }
  }
  k_finalize<<<1, 256, 0, h->stream>>>(h->d_epart, g.nky_loc << h->cols_top, h->d_res);
  h->launches++;
  stage_mark(h, 4);
  if (!h->peer_store) {
    CU(h, cudaMemcpyAsync(B2 + g.rank * blk, A + g.rank * blk, blk * sizeof(double2), cudaMemcpyDeviceToDevice,
                          h->stream));
    for (int k = 1; k < g.P; ++k) CU(h, cudaStreamWaitEvent(h->stream, h->ev_join[(g.rank + k) % g.P], 0));
  } else if (flags) {
    for (int k = 1; k < g.P; ++k) {
      const int r = (g.rank + k) % g.P;
      int rc = push_u0(h, h->stream, r);
      if (!rc) rc = signal_peer(h, h->stream, r, kFlagRet + g.rank);
      if (rc) return rc;
    }
  }
  if (flags) {
    int rc = wait_peers(h, kFlagRet, 1);
    if (!rc) rc = take_u0(h);
    return rc;
  }
  // u0 all-reduce (gfmd_solver_static.cpp:176) doubles as the barrier of the return pushes
  NC(h, a.AllReduce(h->d_res->u0, h->d_res->u0, (size_t) g.d, ncclDouble, ncclSum, h->comm, h->stream));
  return 0;
}

// Multi-GPU step WITHOUT transposes, overlapped (long columns, nx >= 8192; default for more than two
// ranks).  The exchange is done by the top-digit passes of the column transform themselves: the
// forward pass LOADS each piece of a column from the rank whose row kernels produced it, the backward
// pass STORES each result piece into its owner's return buffer (k_cols_top_pass<.., PEER>, NVLink
// traffic issued by the SMs: 640 GB/s per direction at 8 GPUs against 418 GB/s of copy-engine pushes
// and 564 GB/s of grouped ncclSend/ncclRecv, profiles/r2_stage_times_8gpu_16384x16384.txt).  Per
// ky chunk:   pull stream  top-digit pass <- peers            (NVLink RX bound)
//             main stream  fused column kernel on num_sms - xchg_sms SMs (HBM / FP64 bound)
//             push stream  top-digit pass -> peers            (NVLink TX bound)
// so that chunk c + 1 is pulled and chunk c - 1 pushed while chunk c is transformed.  Ordering across
// ranks: flag words (rows complete -> pull; pushes complete -> backward rows), no collective.
int direct_pipelined_step(gfmd_b200 *h, const double *d_u, double2 *A, double2 *B, double2 *B2)
{
  const GridDesc &g = h->g;
  const size_t blk = (size_t) g.d * g.kyb * g.nx_loc;
  const double2 *tw_sub = h->fft_sub.desc.core.tw;
  const int nc = h->nchunks, ck = h->chunk_kl;
  PeerOut po{}, pin{};
  for (int r = 0; r < g.P; ++r) {
    po.p[r] = (r == g.rank ? B2 : h->peer_recv[1][r]) + g.rank * blk;
    pin.p[r] = (r == g.rank ? A : h->peer_stage[r]) + g.rank * blk;
  }
  stage_mark(h, 1);
  // Rows dof by dof.  As soon as a dof is complete on EVERY rank (flag words), the pull stream's
  // top-digit pass fetches it -- on xchg_sms SMs, while the row kernels of the next dof keep the others
  // busy: NVLink works while the rows are transformed, not only after them.  The last dof is pulled
  // chunk by chunk next to the fused column kernel (below).
  const bool split_rows = h->io_fwd == nullptr && g.d > 1;
  const int nrow_launch = split_rows ? g.d : 1;
  int rc = 0;
  for (int i = 0; i < nrow_launch; ++i) {
    rc = split_rows ? fast_rows_fwd(h->fast_rows, d_u, A, g, h->d_tw_ny, h->fft_rows.desc, h->stream, &h->launches, i, 1)
                    : fast_rows_fwd(h->fast_rows, d_u, A, g, h->d_tw_ny, h->fft_rows.desc, h->stream, &h->launches, 0, -1,
                                    h->io_fwd);
    if (rc) return fail(h, GFMD_B200_ECUDA, "fast rows_fwd launch failed");
    for (int k = 1; k < g.P; ++k)
      if ((rc = signal_peer(h, h->stream, (g.rank + k) % g.P, kFlagRows + g.rank + gfmd_b200::kMaxRanks * i))) return rc;
    CU(h, cudaEventRecord(h->ev_row[i], h->stream));
    CU(h, cudaStreamWaitEvent(h->pull_stream, h->ev_row[i], 0));
    for (int k = 1; k < g.P; ++k) {
      const int p = (g.rank + k) % g.P;
      if (stream_wait32_geq(h->pull_stream, h->flags + kFlagRows + p + gfmd_b200::kMaxRanks * i, h->seq))
        return fail(h, GFMD_B200_ECUDA, "cuStreamWaitValue32 on the rows flag of rank %d failed", p);
    }
    if (split_rows && i < g.d - 1) {       // this dof, all columns, now
      rc = fast_cols_fused(h->fast_cols, h->cols_top, B, B, g, tw_sub, h->fft_cols.desc.core.tw, h->d_phi, h->d_linf,
                           h->d_epart, h->d_res, h->num_sms, h->pull_stream, &h->launches, 0, g.kyb, &po, &pin, nullptr, 1,
                           h->xchg_sms, i, 1);
      if (rc) return fail(h, GFMD_B200_ECUDA, "top-digit pull launch failed");
    }
  }
  stage_mark(h, 2);
  const int pull_dof0 = split_rows ? g.d - 1 : 0, pull_nd = split_rows ? 1 : g.d;   // what the chunk loop still pulls
  stage_mark(h, 3);
  const bool tl = h->timeline && h->profiling;
  if (tl)
    for (int c = 0; c < nc; ++c)
      for (int k = 0; k < 6; ++k)
        if (!h->ev_tl[c][k]) cudaEventCreate(&h->ev_tl[c][k]);
  // the fused kernel needs a whole SM per CTA: leave xchg_sms SMs to the pulling and as many to the pushing pass
  const int cols_sms = h->num_sms - h->xchg_sms - h->push_sms > 8 ? h->num_sms - h->xchg_sms - h->push_sms : h->num_sms;
  for (int c = 0; c < nc; ++c) {
    const int k0 = c * ck, k1 = (c + 1) * ck < g.kyb ? (c + 1) * ck : g.kyb;
    // forward top-digit pass of chunk c, inputs straight from the peers' row outputs, into B
    if (tl) cudaEventRecord(h->ev_tl[c][0], h->pull_stream);
    rc = fast_cols_fused(h->fast_cols, h->cols_top, B, B, g, tw_sub, h->fft_cols.desc.core.tw, h->d_phi, h->d_linf,
                         h->d_epart, h->d_res, h->num_sms, h->pull_stream, &h->launches, k0, k1, &po, &pin, nullptr, 1,
                         h->xchg_sms, pull_dof0, pull_nd);
    if (rc) return fail(h, GFMD_B200_ECUDA, "top-digit pull launch failed");
    if (tl) cudaEventRecord(h->ev_tl[c][1], h->pull_stream);
    CU(h, cudaEventRecord(h->ev_pull[c], h->pull_stream));
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_pull[c], 0));
    if (tl) cudaEventRecord(h->ev_tl[c][2], h->stream);
    rc = fast_cols_fused(h->fast_cols, h->cols_top, B, B, g, tw_sub, h->fft_cols.desc.core.tw, h->d_phi, h->d_linf,
                         h->d_epart, h->d_res, cols_sms, h->stream, &h->launches, k0, k1, &po, &pin, nullptr, 2);
    if (rc) return fail(h, GFMD_B200_ECUDA, "fast cols_fused launch failed");
    if (tl) cudaEventRecord(h->ev_tl[c][3], h->stream);
    CU(h, cudaEventRecord(h->ev_k2[c], h->stream));
    CU(h, cudaStreamWaitEvent(h->push_stream, h->ev_k2[c], 0));
    // backward top-digit pass of chunk c, results straight into the owners' return buffers
    if (tl) cudaEventRecord(h->ev_tl[c][4], h->push_stream);
    rc = fast_cols_fused(h->fast_cols, h->cols_top, B, B, g, tw_sub, h->fft_cols.desc.core.tw, h->d_phi, h->d_linf,
                         h->d_epart, h->d_res, h->num_sms, h->push_stream, &h->launches, k0, k1, &po, &pin, nullptr, 4,
                         h->push_sms);
    if (rc) return fail(h, GFMD_B200_ECUDA, "top-digit push launch failed");
    if (tl) cudaEventRecord(h->ev_tl[c][5], h->push_stream);
  }
  k_finalize<<<1, 256, 0, h->stream>>>(h->d_epart, g.nky_loc << h->cols_top, h->d_res);
  h->launches++;
  stage_mark(h, 4);
  // u0 (rank 0, written by chunk 0's fused kernel) and the return flags ride behind the last push
  CU(h, cudaStreamWaitEvent(h->push_stream, h->ev_k2[nc - 1], 0));
  for (int k = 1; k < g.P; ++k) {
    const int r = (g.rank + k) % g.P;
    rc = push_u0(h, h->push_stream, r);
    if (!rc) rc = signal_peer(h, h->push_stream, r, kFlagRet + g.rank);
    if (rc) return rc;
  }
  CU(h, cudaEventRecord(h->ev_push_done, h->push_stream));
  CU(h, cudaStreamWaitEvent(h->stream, h->ev_push_done, 0));      // B and the peers' buffers are free for the next step
  rc = wait_peers(h, kFlagRet, 1);
  if (!rc) rc = take_u0(h);
  if (tl && !rc && ++h->tl_steps == 5) {          // debugging aid: one timeline (ms after the row kernels' mark)
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->pull_stream);
    cudaStreamSynchronize(h->push_stream);
    for (int c = 0; c < nc; ++c) {
      float t[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], h->ev[2], h->ev_tl[c][k]);
      fprintf(stderr, "gfmd_b200 timeline rank %d chunk %d: pull %.3f-%.3f fused %.3f-%.3f push %.3f-%.3f ms\n", g.rank, c,
              t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    cudaGetLastError();
  }
  return rc;
}

// Host-pipelined solver step (single rank, specialised row kernels): the upload of dof k + 1
// runs on a copy stream while the rows of dof k are transformed; the row results of the way
// back are handed to the download stream dof by dof (events hp_out, consumed by
// gfmd_b200_post_force_host).  Same kernels, same arithmetic as enqueue_solver.
int enqueue_solver_hostpipe(gfmd_b200 *h, const double *u_host)
{
  const GridDesc &g = h->g;
  const size_t nxy = (size_t) g.nx_loc * g.ny;
  double2 *A = h->d_stage;
  if (!h->hp_stream[0]) {
    for (int i = 0; i < 2; ++i) CU(h, cudaStreamCreateWithFlags(&h->hp_stream[i], cudaStreamNonBlocking));
    for (int i = 0; i < g.d; ++i) {
      CU(h, cudaEventCreateWithFlags(&h->hp_in[i], cudaEventDisableTiming));
      CU(h, cudaEventCreateWithFlags(&h->hp_out[i], cudaEventDisableTiming));
    }
    CU(h, cudaEventCreateWithFlags(&h->hp_fork, cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&h->hp_done, cudaEventDisableTiming));
  }
  CU(h, cudaMemsetAsync(&h->d_res->epot, 0, offsetof(StepResults, fsum), h->stream));
  // the uploads may not overtake earlier work of this handle that still reads d_u
  CU(h, cudaEventRecord(h->hp_fork, h->stream));
  CU(h, cudaStreamWaitEvent(h->hp_stream[0], h->hp_fork, 0));
  for (int dof = 0; dof < g.d; ++dof) {
    CU(h, cudaMemcpyAsync(h->d_u + dof * nxy, u_host + dof * nxy, nxy * sizeof(double), cudaMemcpyHostToDevice,
                          h->hp_stream[0]));
    CU(h, cudaEventRecord(h->hp_in[dof], h->hp_stream[0]));
    CU(h, cudaStreamWaitEvent(h->stream, h->hp_in[dof], 0));
    if (fast_rows_fwd(h->fast_rows, h->d_u, A, g, h->d_tw_ny, h->fft_rows.desc, h->stream, &h->launches, dof, 1))
      return fail(h, GFMD_B200_ECUDA, "fast rows_fwd launch failed");
  }
  int nepart = (g.nky_loc << h->cols_top) * (h->fast_cols ? fast_cols_nw(h->fast_cols) : 1);
  if (h->fast_cols) {
    const double2 *tw_sub = h->cols_top ? h->fft_sub.desc.core.tw : h->fft_cols.desc.core.tw;
    if (fast_cols_fused(h->fast_cols, h->cols_top, A, A, g, tw_sub, h->fft_cols.desc.core.tw, h->d_phi, h->d_linf,
                        h->d_epart, h->d_res, h->num_sms, h->stream, &h->launches))
      return fail(h, GFMD_B200_ECUDA, "fast cols_fused launch failed");
  } else {
    launch_generic_cols(h, A, A, &nepart);
  }
  k_finalize<<<1, 256, 0, h->stream>>>(h->d_epart, nepart, h->d_res);
  h->launches++;
  for (int dof = 0; dof < g.d; ++dof) {
    if (fast_rows_inv(h->fast_rows, A, h->d_f, g, h->d_tw_ny, h->fft_rows.desc, h->stream, &h->launches, dof, 1))
      return fail(h, GFMD_B200_ECUDA, "fast rows_inv launch failed");
    CU(h, cudaEventRecord(h->hp_out[dof], h->stream));
  }
  CU(h, cudaGetLastError());
  h->hp_pending = true;
  return 0;
}

bool step_is_direct(const gfmd_b200 *h)
{
  const GridDesc &g = h->g;
  return g.P > 1 && h->ipc_on && h->peer_direct && h->fast_cols == 4096 && h->peer_stage[(g.rank + 1) % g.P];
}

// multi-GPU step with per-dof row launches (pipelined_step): no fused gather there
bool step_is_pipelined(const gfmd_b200 *h)
{
  const GridDesc &g = h->g;
  return g.P > 1 && h->ipc_on && h->fast_cols == 4096 && h->fast_rows && h->nchunks > 1 && !step_is_direct(h);
}

// the kernels of one solver step, enqueued on h->stream
int enqueue_solver(gfmd_b200 *h, const double *d_u, double *d_f)
{
  const GridDesc &g = h->g;
  const int nrow_blocks = g.d * ((g.nx_loc + h->rows_RB - 1) / h->rows_RB);
  double2 *A = h->d_stage;
  double2 *B = g.P > 1 ? h->d_stage2 : h->d_stage;                 // columns arrive here
  double2 *B2 = (g.P > 1 && h->ipc_on) ? h->d_stage3 : B;          // rows arrive here on the way back

  CU(h, cudaMemsetAsync(&h->d_res->epot, 0, offsetof(StepResults, fsum), h->stream));
  if (g.P > 1) ++h->seq;

  // GFMD_B200_PEER_DIRECT: no transposes -- the column stage loads and stores the pieces in the peers' memory
  const bool direct = step_is_direct(h);
  const bool pipelined = step_is_pipelined(h);
  h->top_split = false;
  if (direct && h->cols_top > 0 && h->nchunks > 1 && h->sync_flags && h->fast_rows && h->pull_stream) {
    int rc = direct_pipelined_step(h, d_u, A, B, B2);
    if (rc) return rc;
  } else if (pipelined) {
    int rc = pipelined_step(h, d_u, A, B, B2);
    if (rc) return rc;
  } else {
  stage_mark(h, 1);
  if (h->fast_rows) {
    int rc = fast_rows_fwd(h->fast_rows, d_u, A, g, h->d_tw_ny, h->fft_rows.desc, h->stream, &h->launches, 0, -1,
                           h->io_fwd);
    if (rc) return fail(h, GFMD_B200_ECUDA, "fast rows_fwd launch failed");
  } else {
    if (h->even)
      k_rows_fwd<true><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(
          d_u, A, g, h->fft_rows.desc, h->d_tw_ny, h->rows_RB, h->rows_ld);
    else
      k_rows_fwd<false><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(
          d_u, A, g, h->fft_rows.desc, h->d_tw_ny, h->rows_RB, h->rows_ld);
    h->launches++;
  }
  {
    stage_mark(h, 2);
    if (g.P > 1) {
      if (direct && h->sync_flags) {   // every rank's rows are complete before anyone loads them
        for (int k = 1; k < g.P; ++k) {
          int rc = signal_peer(h, h->stream, (g.rank + k) % g.P, kFlagRows + g.rank);
          if (rc) return rc;
        }
        int rc = wait_peers(h, kFlagRows, 1);
        if (rc) return rc;
      } else if (direct) {
        NC(h, nccl().AllReduce(h->d_barrier, h->d_barrier, 1, ncclDouble, ncclSum, h->comm, h->stream));
      } else {
        int rc = exchange(h, A, B, 0);
        if (rc) return rc;
      }
    }
    stage_mark(h, 3);
    int nepart = (g.nky_loc << h->cols_top) * (h->fast_cols ? fast_cols_nw(h->fast_cols) : 1);
    // GFMD_B200_PEER_STORE: the column stage stores its result pieces straight into the owners' return buffers
    const bool peer_store = g.P > 1 && h->ipc_on && (h->peer_store || direct) && h->fast_cols == 4096;
    PeerOut po{}, pin{};
    if (peer_store) {
      const size_t blk = (size_t) g.d * g.kyb * g.nx_loc;
      for (int r = 0; r < g.P; ++r) po.p[r] = (r == g.rank ? B2 : h->peer_recv[1][r]) + g.rank * blk;
      if (direct)
        for (int r = 0; r < g.P; ++r) pin.p[r] = (r == g.rank ? A : h->peer_stage[r]) + g.rank * blk;
    }
    if (g.nky_loc > 0) {
      if (h->fast_cols) {
        const double2 *tw_sub = h->cols_top ? h->fft_sub.desc.core.tw : h->fft_cols.desc.core.tw;
        // direct: long columns are assembled in B by the pulling top-digit pass and transformed there in place
        h->top_split = h->profiling && h->cols_top > 0;
        int rc = fast_cols_fused(h->fast_cols, h->cols_top, B, direct ? B : A, g, tw_sub, h->fft_cols.desc.core.tw,
                                 h->d_phi, h->d_linf, h->d_epart, h->d_res, h->num_sms, h->stream, &h->launches, 0, -1,
                                 peer_store ? &po : nullptr, direct ? &pin : nullptr,
                                 h->top_split ? h->ev_top : nullptr);
        if (rc) return fail(h, GFMD_B200_ECUDA, "fast cols_fused launch failed");
      } else {
        launch_generic_cols(h, B, A, &nepart);
      }
    }
    k_finalize<<<1, 256, 0, h->stream>>>(h->d_epart, nepart, h->d_res);
    h->launches++;
    stage_mark(h, 4);
    if (g.P > 1) {
      if (!peer_store) {
        int rc = exchange(h, A, B2, 1);      // with flags: carries u0 and waits for the peers' return flags
        if (rc) return rc;
      } else if (h->sync_flags) {            // the kernels stored into the peers' buffers themselves
        for (int k = 1; k < g.P; ++k) {
          const int r = (g.rank + k) % g.P;
          int rc = push_u0(h, h->stream, r);
          if (!rc) rc = signal_peer(h, h->stream, r, kFlagRet + g.rank);
          if (rc) return rc;
        }
        int rc = wait_peers(h, kFlagRet, 1);
        if (!rc) rc = take_u0(h);
        if (rc) return rc;
      }
      if (!(h->sync_flags && h->ipc_on))
        NC(h, nccl().AllReduce(h->d_res->u0, h->d_res->u0, (size_t) g.d, ncclDouble, ncclSum, h->comm,
                               h->stream));
    }
  }
  }   // !pipelined
  stage_mark(h, 5);
  if (h->fast_rows) {
    int rc = fast_rows_inv(h->fast_rows, B2, d_f, g, h->d_tw_ny, h->fft_rows.desc, h->stream, &h->launches, 0, -1,
                           h->io_inv);
    if (rc) return fail(h, GFMD_B200_ECUDA, "fast rows_inv launch failed");
  } else {
    if (h->even)
      k_rows_inv<true><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(
          B2, d_f, g, h->fft_rows.desc, h->d_tw_ny, h->rows_RB, h->rows_ld);
    else
      k_rows_inv<false><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(
          B2, d_f, g, h->fft_rows.desc, h->d_tw_ny, h->rows_RB, h->rows_ld);
    h->launches++;
  }
  stage_mark(h, 6);
  CU(h, cudaGetLastError());
  return 0;
}

// Off-path transforms on the generic kernels (single rank): rows forward, the auxiliary column
// kernel in `mode`, and for AUX_PREC rows backward.  d_in -> (stage, d_spec) or d_out.
int enqueue_aux(gfmd_b200 *h, int mode, const double *d_in, double *d_out, bool want_f, int ncopy)
{
  const GridDesc &g = h->g;
  if (g.P != 1)
    return fail(h, GFMD_B200_EUNSUPPORTED, "q-space dumps and prec_gradient run on a single rank only (the "
                "reference's dumps do too, gfmd_solver_fft.cpp:211-212)");
  if (!h->aux_cols_smem && !h->aux_split_db)
    return fail(h, GFMD_B200_EUNSUPPORTED, "nx = %d: one column (%zu B) does not fit one CTA's shared memory (%zu B); "
                "q-space dumps and prec_gradient are limited to nx <= 8192", g.nx, (size_t) h->cols_ld * sizeof(double2),
                kMaxSmem);
  if (mode == AUX_PREC && !(g.d == 3 || g.d == 6 || g.d == 9 || g.d == 12))
    return fail(h, GFMD_B200_EUNSUPPORTED, "prec_gradient: ndof %d (3, 6, 9, 12 supported)", g.d);
  if (!h->aux_attr_set && !h->aux_cols_smem) {
    CU(h, grow_dyn_smem((const void *) k_cols_split_fft<-1>, h->aux_split_smem));
    CU(h, grow_dyn_smem((const void *) k_cols_split_fft<+1>, h->aux_split_smem));
    h->aux_attr_set = true;
  }
  if (!h->aux_attr_set) {
#define SET_AUX(DT, MODE) CU(h, grow_dyn_smem((const void *) k_cols_aux<DT, MODE>, h->aux_cols_smem))
    switch (g.d) {
      case 3: SET_AUX(3, AUX_SPECTRUM); SET_AUX(3, AUX_PREC); break;
      case 6: SET_AUX(6, AUX_SPECTRUM); SET_AUX(6, AUX_PREC); break;
      case 9: SET_AUX(9, AUX_SPECTRUM); SET_AUX(9, AUX_PREC); break;
      case 12: SET_AUX(12, AUX_SPECTRUM); SET_AUX(12, AUX_PREC); break;
      default: SET_AUX(0, AUX_SPECTRUM); break;
    }
#undef SET_AUX
    h->aux_attr_set = true;
  }
  if (mode == AUX_SPECTRUM && want_f && !h->d_spec)
    CU(h, dmalloc(h, &h->d_spec, (size_t) g.d * g.kyb * g.nx));
  const int nrow_blocks = g.d * ((g.nx_loc + h->rows_RB - 1) / h->rows_RB);
  double2 *A = h->d_stage;
  if (h->even)
    k_rows_fwd<true><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(d_in, A, g, h->fft_rows.desc, h->d_tw_ny,
                                                                          h->rows_RB, h->rows_ld);
  else
    k_rows_fwd<false><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(d_in, A, g, h->fft_rows.desc, h->d_tw_ny,
                                                                           h->rows_RB, h->rows_ld);
  h->launches++;
  const int lognx = ilog2_rt(g.nx);
  if (!h->aux_cols_smem) {
    // column set larger than one CTA: transform phase, per-q kernel on the spectrum in HBM, transform phase
    const int db = h->aux_split_db, ngrp = (g.d + db - 1) / db;
    k_cols_split_fft<-1><<<g.nky_loc * ngrp, h->aux_split_T, h->aux_split_smem, h->stream>>>(A, A, g, h->fft_cols.desc,
                                                                                           h->cols_ld, db);
    const long long nq = (long long) g.nky_loc * g.nx;
    const int pgrid = (int) ((nq + 255) / 256 < (long long) h->num_sms * 8 ? (nq + 255) / 256 : (long long) h->num_sms * 8);
#define LAUNCH_PERQ(DT, MODE)                                                                                          \
  k_aux_perq<DT, MODE><<<pgrid, 256, 0, h->stream>>>(A, want_f ? h->d_spec : nullptr, g, h->d_phi, h->d_cavg,         \
                                                     h->phi_mode(), h->cols_top, lognx, ncopy)
    if (mode == AUX_SPECTRUM) {
      switch (g.d) {
        case 3: LAUNCH_PERQ(3, AUX_SPECTRUM); break;
        case 6: LAUNCH_PERQ(6, AUX_SPECTRUM); break;
        case 9: LAUNCH_PERQ(9, AUX_SPECTRUM); break;
        case 12: LAUNCH_PERQ(12, AUX_SPECTRUM); break;
        default: LAUNCH_PERQ(0, AUX_SPECTRUM); break;
      }
    } else {
      switch (g.d) {
        case 3: LAUNCH_PERQ(3, AUX_PREC); break;
        case 6: LAUNCH_PERQ(6, AUX_PREC); break;
        case 9: LAUNCH_PERQ(9, AUX_PREC); break;
        default: LAUNCH_PERQ(12, AUX_PREC); break;
      }
      k_cols_split_fft<+1><<<g.nky_loc * ngrp, h->aux_split_T, h->aux_split_smem, h->stream>>>(A, A, g, h->fft_cols.desc,
                                                                                             h->cols_ld, db);
      h->launches++;
    }
#undef LAUNCH_PERQ
    h->launches += 2;
  } else {
#define LAUNCH_AUX(DT, MODE)                                                                      \
  k_cols_aux<DT, MODE><<<g.nky_loc, h->cols_T, h->aux_cols_smem, h->stream>>>(                     \
      A, want_f ? h->d_spec : nullptr, g, h->fft_cols.desc, h->d_phi, h->d_cavg, h->phi_mode(), lognx, h->cols_ld, \
      ncopy)
  if (mode == AUX_SPECTRUM) {
    switch (g.d) {
      case 3: LAUNCH_AUX(3, AUX_SPECTRUM); break;
      case 6: LAUNCH_AUX(6, AUX_SPECTRUM); break;
      case 9: LAUNCH_AUX(9, AUX_SPECTRUM); break;
      case 12: LAUNCH_AUX(12, AUX_SPECTRUM); break;
      default: LAUNCH_AUX(0, AUX_SPECTRUM); break;
    }
  } else {
    switch (g.d) {
      case 3: LAUNCH_AUX(3, AUX_PREC); break;
      case 6: LAUNCH_AUX(6, AUX_PREC); break;
      case 9: LAUNCH_AUX(9, AUX_PREC); break;
      default: LAUNCH_AUX(12, AUX_PREC); break;
    }
  }
#undef LAUNCH_AUX
  h->launches++;
  }
  if (mode == AUX_PREC) {
    if (h->even)
      k_rows_inv<true><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(A, d_out, g, h->fft_rows.desc,
                                                                            h->d_tw_ny, h->rows_RB, h->rows_ld);
    else
      k_rows_inv<false><<<nrow_blocks, h->rows_T, h->rows_smem, h->stream>>>(A, d_out, g, h->fft_rows.desc,
                                                                             h->d_tw_ny, h->rows_RB, h->rows_ld);
    h->launches++;
  }
  CU(h, cudaGetLastError());
  return 0;
}

void accumulate_stage_times(gfmd_b200 *h, int first, int last)
{
  // events first..last+1 were recorded; requires a stream sync by the caller
  auto add = [&](int stage, cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) {
      h->stage_ms[stage] += ms;
      h->stage_cnt[stage]++;
    }
  };
  for (int s = first; s <= last; ++s) {
    if (s == 3 && h->top_split) {       // long columns: top-digit passes (stages 7, 8) around the fused kernel
      add(7, h->ev[3], h->ev_top[0]);
      add(3, h->ev_top[0], h->ev_top[1]);
      add(8, h->ev_top[1], h->ev[4]);
      continue;
    }
    add(s, h->ev[s], h->ev[s + 1]);
  }
  cudaGetLastError();
}

int solver_step(gfmd_b200 *h, const double *d_u, double *d_f)
{
  if (!h->phi_set) return fail(h, GFMD_B200_ESTATE, "post_force before set_phi (set_kernel)");
  if (d_u == d_f) return fail(h, GFMD_B200_EINVAL, "u and f must be different buffers");
  const bool graph_ok = h->want_graph && h->g.P == 1 && !h->profiling && !h->io_fwd && !h->io_inv;
  if (!graph_ok) {
    int rc = enqueue_solver(h, d_u, d_f);
    if (rc) return rc;
    if (h->profiling) {
      CU(h, cudaStreamSynchronize(h->stream));
      accumulate_stage_times(h, 1, 5);
    }
    return 0;
  }
  if (!h->graph_exec || h->graph_u != d_u || h->graph_f != d_f) {
    drop_graph(h);
    cudaGraph_t graph = nullptr;
    CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const long long before = h->launches;
    int rc = enqueue_solver(h, d_u, d_f);
    h->graph_launches = h->launches - before;
    h->launches = before;
    cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    CU(h, ce);
    ce = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    CU(h, ce);
    h->graph_u = d_u;
    h->graph_f = d_f;
  }
  CU(h, cudaGraphLaunch(h->graph_exec, h->stream));
  h->launches += h->graph_launches;
  return 0;
}

int fetch_results(gfmd_b200 *h)
{
  CU(h, cudaMemcpyAsync(h->h_res, h->d_res, sizeof(StepResults), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

void try_pin(gfmd_b200 *h, const void *p, size_t bytes)
{
  if (!h->pin_host) return;
  auto it = h->pinned.find(p);
  if (it != h->pinned.end() && it->second >= bytes) return;
  if (it != h->pinned.end()) {
    cudaHostUnregister(const_cast<void *>(p));
    h->pinned.erase(it);
  }
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered)
    return;                 // already page-locked / managed by someone else
  cudaGetLastError();
  if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) == cudaSuccess)
    h->pinned[p] = bytes;
  else
    cudaGetLastError();     // pageable copies still work
}

// Hermitian packing of one q: M = full d x d complex matrix (row-major), scale s.
inline void pack_hermitian(const double *M, const double *Mneg, int d, double s, double *dst,
                           size_t plane_stride, double &amax, double &hdev, double &cdev)
{
  // Phi_h = (Phi(q) + conj Phi(-q)) / 2 ; then its Hermitian part
  auto elem = [&](int i, int j, double &re, double &im) {
    const double *a = M + 2 * ((size_t) i * d + j);
    if (Mneg) {
      const double *b = Mneg + 2 * ((size_t) i * d + j);
      re = 0.5 * (a[0] + b[0]);
      im = 0.5 * (a[1] - b[1]);
      double dr = a[0] - b[0], di = a[1] + b[1];
      double dv = std::sqrt(dr * dr + di * di);
      if (dv > cdev) cdev = dv;
    } else {
      re = a[0];
      im = a[1];
    }
    double av = std::sqrt(a[0] * a[0] + a[1] * a[1]);
    if (av > amax) amax = av;
  };
  int c = d;
  for (int i = 0; i < d; ++i) {
    double re, im;
    elem(i, i, re, im);
    if (std::fabs(im) > hdev) hdev = std::fabs(im);
    dst[(size_t) i * plane_stride] = re * s;
    for (int j = i + 1; j < d; ++j) {
      double r1, i1, r2, i2;
      elem(i, j, r1, i1);
      elem(j, i, r2, i2);
      double dr = r1 - r2, di = i1 + i2;
      double dv = std::sqrt(dr * dr + di * di);
      if (dv > hdev) hdev = dv;
      dst[(size_t) c * plane_stride] = 0.5 * (r1 + r2) * s;
      dst[(size_t) (c + 1) * plane_stride] = 0.5 * (i1 - i2) * s;
      c += 2;
    }
  }
}

}  // namespace

// -------------------------------------------------------------------- ABI ---

extern "C" {

#ifdef GFMD_PHASE_TIMING
int gfmd_b200_debug_phase_cycles(long long *out, int reset)
{
  long long z[16] = {0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(z));
  if (reset) cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
  return 0;
}
#endif

const char *gfmd_b200_version(void) { return "gfmd_b200 0.1 (sm_100a, fp64)"; }

int gfmd_b200_create(gfmd_b200_t **h, int nx, int ny, int ndof, int device)
{
  return create_common(h, nx, ny, ndof, device, 0, 1);
}

int gfmd_b200_create_slab(gfmd_b200_t **h, int nx, int ny, int ndof, int device, int rank, int nranks)
{
  return create_common(h, nx, ny, ndof, device, rank, nranks);
}

int gfmd_b200_get_unique_id(char id[GFMD_B200_UNIQUE_ID_BYTES])
{
  if (!nccl_load()) return fail(nullptr, GFMD_B200_ENCCL, "%s", nccl().err.c_str());
  ncclUniqueId uid;
  static_assert(sizeof(ncclUniqueId) == GFMD_B200_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclResult_t r = nccl().GetUniqueId(&uid);
  if (r != ncclSuccess) return fail(nullptr, GFMD_B200_ENCCL, "ncclGetUniqueId: %s", nccl().GetErrorString(r));
  memcpy(id, &uid, sizeof(uid));
  return 0;
}

int gfmd_b200_comm_init(gfmd_b200_t *h, const char id[GFMD_B200_UNIQUE_ID_BYTES])
{
  if (!h) return GFMD_B200_EINVAL;
  if (h->g.P == 1) return 0;
  if (!nccl_load()) return fail(h, GFMD_B200_ENCCL, "%s", nccl().err.c_str());
  int rc = set_device(h);
  if (rc) return rc;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NC(h, nccl().CommInitRank(&h->comm, h->g.P, uid, h->g.rank));
  return 0;
}

int gfmd_b200_ipc_export(gfmd_b200_t *h, char *handles)
{
  if (!h || !handles) return fail(h, GFMD_B200_EINVAL, "ipc_export: null argument");
  if (h->g.P == 1) return fail(h, GFMD_B200_ESTATE, "ipc_export: single-GPU handle has no exchange");
  int rc = set_device(h);
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == GFMD_B200_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  const size_t nstage = (size_t) h->g.P * h->g.d * h->g.kyb * h->g.nx_loc;
  if (!h->d_stage3) {
    CU(h, dmalloc(h, &h->d_stage3, nstage));
    CU(h, cudaMemset(h->d_stage3, 0, sizeof(double2) * nstage));
    CU(h, cudaDeviceSynchronize());      // before any peer may push into it
  }
  cudaIpcMemHandle_t m[2];
  CU(h, cudaIpcGetMemHandle(&m[0], h->d_stage2));
  CU(h, cudaIpcGetMemHandle(&m[1], h->d_stage3));
  memcpy(handles, m, sizeof(m));
  return 0;
}

int gfmd_b200_ipc_import(gfmd_b200_t *h, const char *all_handles)
{
  if (!h || !all_handles) return fail(h, GFMD_B200_EINVAL, "ipc_import: null argument");
  if (h->g.P == 1) return 0;
  if (!h->d_stage3) return fail(h, GFMD_B200_ESTATE, "ipc_import before ipc_export");
  int rc = set_device(h);
  if (rc) return rc;
  CU(h, cudaStreamSynchronize(h->stream));
  for (int r = 0; r < h->g.P; ++r) {
    if (r == h->g.rank) {
      h->peer_recv[0][r] = h->d_stage2;
      h->peer_recv[1][r] = h->d_stage3;
      continue;
    }
    for (int w = 0; w < 2; ++w) {
      cudaIpcMemHandle_t m;
      memcpy(&m, all_handles + ((size_t) r * 2 + w) * GFMD_B200_IPC_HANDLE_BYTES, sizeof(m));
      void *p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, m, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return fail(h, GFMD_B200_ECUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s (the NCCL exchange stays "
                    "in use)", r, cudaGetErrorString(e));
      h->peer_recv[w][r] = (double2 *) p;
    }
    {
      const size_t nstage = (size_t) h->g.P * h->g.d * h->g.kyb * h->g.nx_loc;
      h->peer_flags[r] = reinterpret_cast<unsigned *>(h->peer_recv[0][r] + nstage);
      h->peer_u0_in[r] = reinterpret_cast<double *>(reinterpret_cast<char *>(h->peer_flags[r]) + kFlagU0Bytes);
    }
    if (!h->copy_stream[r]) {
      CU(h, cudaStreamCreateWithFlags(&h->copy_stream[r], cudaStreamNonBlocking));
      CU(h, cudaEventCreateWithFlags(&h->ev_join[r], cudaEventDisableTiming));
    }
  }
  if (!h->ev_fork) CU(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  if (!h->d_barrier) {
    CU(h, dmalloc(h, &h->d_barrier, (size_t) 1));
    CU(h, cudaMemset(h->d_barrier, 0, sizeof(double)));
    CU(h, cudaDeviceSynchronize());
  }
  // Which exchange: more than two ranks with long columns (nx >= 8192) default to the overlapped step
  // WITHOUT transposes (direct_pipelined_step; needs gfmd_b200_ipc_import_stage), everything else to the
  // copy-engine pushes (pipelined_step).  Measured at 16384 x 16384: 8 GPUs 4.46 ms against 5.14 ms even
  // before the overlap existed (4.09 ms with it), 2 GPUs 13.3 ms against 12.6 ms.  GFMD_B200_PEER_DIRECT=0/1 overrides.
  h->peer_direct = h->g.P >= 4 && h->cols_top > 0 && h->fast_cols == 4096;
  if (const char *e = getenv("GFMD_B200_PEER_STORE")) h->peer_store = atoi(e) != 0;
  if (const char *e = getenv("GFMD_B200_PEER_DIRECT")) h->peer_direct = atoi(e) != 0;
  if (const char *e = getenv("GFMD_B200_XCHG_SMS")) {       // "<pull>" or "<pull>,<push>"
    h->xchg_sms = h->push_sms = atoi(e);
    if (const char *c = strchr(e, ',')) h->push_sms = atoi(c + 1);
  }
  if (h->push_sms < 0) h->push_sms = 0;
  if (const char *e = getenv("GFMD_B200_TIMELINE")) h->timeline = atoi(e) != 0;
  if (h->xchg_sms < 0) h->xchg_sms = 0;
  // chunking of the column stage: whole waves of the persistent column kernel per chunk
  {
    const bool overlapped_direct = h->peer_direct && h->cols_top > 0;
    int want = overlapped_direct ? 8 : 4;
    if (const char *e = getenv("GFMD_B200_CHUNKS")) want = atoi(e);
    if (want < 1) want = 1;
    if (want > gfmd_b200::kMaxChunks) want = gfmd_b200::kMaxChunks;
    int sms = h->num_sms;
    if (overlapped_direct && sms - h->xchg_sms - h->push_sms > 8) sms -= h->xchg_sms + h->push_sms;   // the fused kernel's share
    const int per_wave = sms >> h->cols_top > 0 ? sms >> h->cols_top : 1;   // ky per wave
    int waves = (h->g.kyb + per_wave * want - 1) / (per_wave * want);
    if (waves < 1) waves = 1;
    h->chunk_kl = waves * per_wave;
    h->nchunks = (h->g.kyb + h->chunk_kl - 1) / h->chunk_kl;
    if (h->nchunks > gfmd_b200::kMaxChunks) { h->nchunks = 1; h->chunk_kl = h->g.kyb; }
    for (int i = 0; i < h->g.d; ++i)
      if (!h->ev_row[i]) CU(h, cudaEventCreateWithFlags(&h->ev_row[i], cudaEventDisableTiming));
    for (int c = 0; c < h->nchunks; ++c) {
      if (!h->ev_k2[c]) CU(h, cudaEventCreateWithFlags(&h->ev_k2[c], cudaEventDisableTiming));
      if (!h->ev_pull[c]) CU(h, cudaEventCreateWithFlags(&h->ev_pull[c], cudaEventDisableTiming));
      for (int r = 0; r < h->g.P; ++r)
        if (!h->ev_chunk[c][r]) CU(h, cudaEventCreateWithFlags(&h->ev_chunk[c][r], cudaEventDisableTiming));
    }
  }
  h->ipc_on = true;
  // cross-rank ordering: flag words written and awaited by the streams themselves (default), or
  // GFMD_B200_SYNC=nccl: one-element NCCL all-reduces as barriers (needs gfmd_b200_comm_init)
  {
    const char *e = getenv("GFMD_B200_SYNC");
    const bool want_nccl = e && !strcmp(e, "nccl");
    h->sync_flags = !want_nccl && memops_load();
    if (!h->sync_flags && !h->comm)
      return fail(h, GFMD_B200_ESTATE, "ipc_import: %s and no NCCL communicator (gfmd_b200_comm_init) to fall back on",
                  want_nccl ? "GFMD_B200_SYNC=nccl" : memops().err.c_str());
    if (h->desc.find("sync: ") == std::string::npos)
      h->desc += h->sync_flags ? " | sync: stream flags in peer memory" : " | sync: NCCL all-reduce barriers";
  }
  if (h->peer_store && !h->peer_direct && h->fast_cols == 4096 && h->desc.find("return: ") == std::string::npos)
    h->desc += " | return: in-kernel peer stores";
  return 0;
}

int gfmd_b200_ipc_export_stage(gfmd_b200_t *h, char *handle)
{
  if (!h || !handle) return fail(h, GFMD_B200_EINVAL, "ipc_export_stage: null argument");
  if (h->g.P == 1) return fail(h, GFMD_B200_ESTATE, "ipc_export_stage: single-GPU handle has no exchange");
  int rc = set_device(h);
  if (rc) return rc;
  cudaIpcMemHandle_t m;
  CU(h, cudaIpcGetMemHandle(&m, h->d_stage));
  memcpy(handle, &m, sizeof(m));
  return 0;
}

int gfmd_b200_ipc_import_stage(gfmd_b200_t *h, const char *all_handles)
{
  if (!h || !all_handles) return fail(h, GFMD_B200_EINVAL, "ipc_import_stage: null argument");
  if (h->g.P == 1) return 0;
  if (!h->ipc_on) return fail(h, GFMD_B200_ESTATE, "ipc_import_stage before ipc_import");
  int rc = set_device(h);
  if (rc) return rc;
  CU(h, cudaStreamSynchronize(h->stream));
  for (int r = 0; r < h->g.P; ++r) {
    if (r == h->g.rank) {
      h->peer_stage[r] = h->d_stage;
      continue;
    }
    if (h->peer_stage[r]) continue;
    cudaIpcMemHandle_t m;
    memcpy(&m, all_handles + (size_t) r * GFMD_B200_IPC_HANDLE_BYTES, sizeof(m));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, m, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail(h, GFMD_B200_ECUDA, "cudaIpcOpenMemHandle(rank %d, row output) failed: %s (the peer pushes stay "
                  "in use)", r, cudaGetErrorString(e));
    h->peer_stage[r] = (double2 *) p;
  }
  if (h->peer_direct && h->fast_cols == 4096 && h->cols_top > 0 && !h->pull_stream) {
    int least = 0, greatest = 0;
    CU(h, cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CU(h, cudaStreamCreateWithPriority(&h->pull_stream, cudaStreamNonBlocking, greatest));
    CU(h, cudaStreamCreateWithPriority(&h->push_stream, cudaStreamNonBlocking, greatest));
    CU(h, cudaEventCreateWithFlags(&h->ev_rows_done, cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&h->ev_push_done, cudaEventDisableTiming));
  }
  if (h->peer_direct && h->fast_cols == 4096 && h->desc.find("transposes: ") == std::string::npos)
    h->desc += h->cols_top > 0 && h->nchunks > 1 && h->sync_flags
                   ? " | transposes: none, in-kernel peer loads and stores, overlapped chunk by chunk (" +
                         std::to_string(h->nchunks) + " chunks, " + std::to_string(h->xchg_sms) + " + " +
                         std::to_string(h->push_sms) + " SMs pull + push)"
                   : " | transposes: none, in-kernel peer loads and stores";
  return 0;
}

void gfmd_b200_destroy(gfmd_b200_t *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  drop_graph(h);
  if (h->comm && nccl().lib) nccl().CommDestroy(h->comm);
  for (auto &kv : h->pinned) cudaHostUnregister(const_cast<void *>(kv.first));
  cudaGetLastError();
  for (int w = 0; w < 2; ++w)
    for (int r = 0; r < gfmd_b200::kMaxRanks; ++r)
      if (h->peer_recv[w][r] && r != h->g.rank) cudaIpcCloseMemHandle(h->peer_recv[w][r]);
  for (int r = 0; r < gfmd_b200::kMaxRanks; ++r)
    if (h->peer_stage[r] && r != h->g.rank) cudaIpcCloseMemHandle(h->peer_stage[r]);
  for (int r = 0; r < gfmd_b200::kMaxRanks; ++r) {
    if (h->copy_stream[r]) cudaStreamDestroy(h->copy_stream[r]);
    if (h->ev_join[r]) cudaEventDestroy(h->ev_join[r]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->pull_stream) cudaStreamDestroy(h->pull_stream);
  if (h->push_stream) cudaStreamDestroy(h->push_stream);
  if (h->ev_rows_done) cudaEventDestroy(h->ev_rows_done);
  if (h->ev_push_done) cudaEventDestroy(h->ev_push_done);
  for (int c = 0; c < gfmd_b200::kMaxChunks; ++c) {
    if (h->ev_pull[c]) cudaEventDestroy(h->ev_pull[c]);
    for (int k = 0; k < 6; ++k)
      if (h->ev_tl[c][k]) cudaEventDestroy(h->ev_tl[c][k]);
  }
  for (int i = 0; i < GFMD_B200_MAX_NDOF; ++i)
    if (h->ev_row[i]) cudaEventDestroy(h->ev_row[i]);
  for (int c = 0; c < gfmd_b200::kMaxChunks; ++c) {
    if (h->ev_k2[c]) cudaEventDestroy(h->ev_k2[c]);
    for (int r = 0; r < gfmd_b200::kMaxRanks; ++r)
      if (h->ev_chunk[c][r]) cudaEventDestroy(h->ev_chunk[c][r]);
  }
  cudaFree(h->d_barrier); cudaFree(h->d_stage3);
  cudaFree(h->d_u); cudaFree(h->d_f); cudaFree(h->d_stage); cudaFree(h->d_stage2);
  cudaFree(h->d_phi); cudaFree(h->d_linf); cudaFree(h->d_epart); cudaFree(h->d_fsum_part);
  cudaFree(h->d_res); cudaFree(h->d_tw_ny);
  cudaFree(h->d_spec); cudaFree(h->d_cavg);
  cudaFree(h->d_cmap); cudaFree(h->d_cmap_cnt); cudaFree(h->d_fsum_io);
  if (h->h_res) cudaFreeHost(h->h_res);
  free_fft(h->fft_rows);
  free_fft(h->fft_cols);
  free_fft(h->fft_sub);
  for (int i = 0; i < 8; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 2; ++i)
    if (h->ev_top[i]) cudaEventDestroy(h->ev_top[i]);
  for (int i = 0; i < 2; ++i)
    if (h->hp_stream[i]) cudaStreamDestroy(h->hp_stream[i]);
  for (int i = 0; i < GFMD_B200_MAX_NDOF; ++i) {
    if (h->hp_in[i]) cudaEventDestroy(h->hp_in[i]);
    if (h->hp_out[i]) cudaEventDestroy(h->hp_out[i]);
  }
  if (h->hp_fork) cudaEventDestroy(h->hp_fork);
  if (h->hp_done) cudaEventDestroy(h->hp_done);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char *gfmd_b200_last_error(const gfmd_b200_t *h)
{
  return h ? h->err.c_str() : g_create_error.c_str();
}

int gfmd_b200_get_brick(const gfmd_b200_t *h, int *xlo, int *xhi, int *ylo, int *yhi, int *nxy_loc,
                        int *gammai)
{
  if (!h) return GFMD_B200_EINVAL;
  const GridDesc &g = h->g;
  if (xlo) *xlo = g.x0;
  if (xhi) *xhi = g.x0 + g.nx_loc - 1;
  if (ylo) *ylo = 0;
  if (yhi) *yhi = g.ny - 1;
  if (nxy_loc) *nxy_loc = g.nx_loc * g.ny;
  if (gammai) *gammai = g.x0 == 0 ? 0 : -1;
  return 0;
}

int gfmd_b200_get_q_columns(const gfmd_b200_t *h, int *kylo, int *nky)
{
  if (!h) return GFMD_B200_EINVAL;
  if (kylo) *kylo = h->g.ky0;
  if (nky) *nky = h->g.nky_loc;
  return 0;
}

int gfmd_b200_set_linf(gfmd_b200_t *h, const double *linf)
{
  if (!h) return GFMD_B200_EINVAL;
  int rc = set_device(h);
  if (rc) return rc;
  double tmp[GFMD_B200_MAX_NDOF] = {0};
  if (linf)
    for (int i = 0; i < h->g.d / 3; ++i) tmp[i] = linf[i];
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, h2d_blocking(h->d_linf, tmp, sizeof(tmp)));
  return 0;
}

int gfmd_b200_set_phi(gfmd_b200_t *h, const double *phi, int already_normalised, const double *linf)
{
  if (!h || !phi) return fail(h, GFMD_B200_EINVAL, "set_phi: null argument");
  int rc = set_device(h);
  if (rc) return rc;
  const GridDesc &g = h->g;
  const int d = g.d, nx = g.nx, ny = g.ny;
  const size_t dsq = (size_t) d * d;
  const double s = already_normalised ? 1.0 : 1.0 / ((double) nx * (double) ny);
  const int lognx = ilog2_rt(nx);
  double amax = 0.0, hdev = 0.0, cdev = 0.0;
  CU(h, cudaStreamSynchronize(h->stream));
  // chunk over local ky to bound the host staging buffer
  const int chunk = 64;
  std::vector<double> buf((size_t) chunk * dsq * nx);
  for (int k0 = 0; k0 < g.nky_loc; k0 += chunk) {
    const int nk = g.nky_loc - k0 < chunk ? g.nky_loc - k0 : chunk;
    for (int kl = 0; kl < nk; ++kl) {
      const int ky = g.ky0 + k0 + kl;
      const int kyn = (ny - ky) % ny;
      for (int kx = 0; kx < nx; ++kx) {
        const int kxn = (nx - kx) % nx;
        const double *M = phi + 2 * dsq * ((size_t) kx * ny + ky);
        const double *Mn = phi + 2 * dsq * ((size_t) kxn * ny + kyn);
        size_t off, cstride;
        phi_slot(h->phi_mode(), h->cols_top, lognx, nx, dsq, kx, off, cstride);
        pack_hermitian(M, Mn, d, s, buf.data() + (size_t) kl * dsq * nx + off, cstride, amax, hdev, cdev);
      }
    }
    CU(h, h2d_blocking(h->d_phi + (size_t) k0 * dsq * nx, buf.data(), sizeof(double) * (size_t) nk * dsq * nx));
  }
  h->herm_dev = amax > 0 ? hdev / amax : 0.0;
  h->conj_dev = amax > 0 ? cdev / amax : 0.0;
  if (h->herm_dev > 1e-6)
    return fail(h, GFMD_B200_EPHI, "Phi table is not Hermitian: max |Phi - Phi^H| / max|Phi| = %g", h->herm_dev);
  if (h->conj_dev > 1e-6)
    return fail(h, GFMD_B200_EPHI, "Phi(-q) != conj Phi(q): relative deviation %g (complex u(r)?)", h->conj_dev);
  std::fill(h->phi_cols_set.begin(), h->phi_cols_set.end(), 1);
  h->phi_set = true;
  return gfmd_b200_set_linf(h, linf);
}

int gfmd_b200_set_phi_columns(gfmd_b200_t *h, const double *phi, int ky_first, int nky, int already_normalised)
{
  if (!h || !phi) return fail(h, GFMD_B200_EINVAL, "set_phi_columns: null argument");
  const GridDesc &g = h->g;
  if (nky < 0 || ky_first < g.ky0 || ky_first + nky > g.ky0 + g.nky_loc)
    return fail(h, GFMD_B200_EINVAL, "set_phi_columns: ky range [%d,%d) outside this handle's [%d,%d)",
                ky_first, ky_first + nky, g.ky0, g.ky0 + g.nky_loc);
  int rc = set_device(h);
  if (rc) return rc;
  const int d = g.d, nx = g.nx;
  const size_t dsq = (size_t) d * d;
  const double s = already_normalised ? 1.0 : 1.0 / ((double) nx * (double) g.ny);
  const int lognx = ilog2_rt(nx);
  double amax = 0.0, hdev = 0.0, cdev = 0.0;
  CU(h, cudaStreamSynchronize(h->stream));
  const int chunk = 64;
  std::vector<double> buf((size_t) chunk * dsq * nx);
  for (int k0 = 0; k0 < nky; k0 += chunk) {
    const int nk = nky - k0 < chunk ? nky - k0 : chunk;
    for (int kl = 0; kl < nk; ++kl)
      for (int kx = 0; kx < nx; ++kx) {
        const double *M = phi + 2 * dsq * ((size_t) kx * nky + k0 + kl);
        size_t off, cstride;
        phi_slot(h->phi_mode(), h->cols_top, lognx, nx, dsq, kx, off, cstride);
        pack_hermitian(M, nullptr, d, s, buf.data() + (size_t) kl * dsq * nx + off, cstride, amax, hdev, cdev);
      }
    CU(h, h2d_blocking(h->d_phi + (size_t) (ky_first - g.ky0 + k0) * dsq * nx, buf.data(),
                       sizeof(double) * (size_t) nk * dsq * nx));
  }
  if (amax > 0 && hdev / amax > 1e-6)
    return fail(h, GFMD_B200_EPHI, "Phi table is not Hermitian: relative deviation %g", hdev / amax);
  for (int k = 0; k < nky; ++k) h->phi_cols_set[ky_first - g.ky0 + k] = 1;
  bool all = true;
  for (char c : h->phi_cols_set) all = all && c;
  h->phi_set = all;
  return 0;
}

// uuv_is_device: the (U0, U, V) blocks already live in device memory (gfmd_b200_build_phi_columns_device)
static int build_phi_columns_impl(gfmd_b200_t *h, const double *uuv, int ky_first, int nky, int height,
                                  int normalise, bool uuv_is_device)
{
  if (!h || !uuv) return fail(h, GFMD_B200_EINVAL, "build_phi_columns: null argument");
  const GridDesc &g = h->g;
  if (nky < 0 || ky_first < g.ky0 || ky_first + nky > g.ky0 + g.nky_loc)
    return fail(h, GFMD_B200_EINVAL, "build_phi_columns: ky range [%d,%d) outside this handle's [%d,%d)",
                ky_first, ky_first + nky, g.ky0, g.ky0 + g.nky_loc);
  if (nky == 0) return 0;
  int rc = set_device(h);
  if (rc) return rc;
  const int d = g.d, nx = g.nx;
  const size_t dsq = (size_t) d * d;
  const size_t n = (size_t) nx * nky * 3 * dsq;
  if (!(d == 3 || d == 6 || d == 9 || d == 12))
    return fail(h, GFMD_B200_EUNSUPPORTED, "build_phi_columns: ndof %d (3, 6, 9, 12 supported)", d);
  double2 *d_own = nullptr;
  int *d_flag = nullptr;
  CU(h, cudaStreamSynchronize(h->stream));
  if (uuv_is_device) CU(h, cudaDeviceSynchronize());     // the caller's producer may run on another stream
  CU(h, cudaMalloc((void **) &d_flag, sizeof(int)));     // "out of iterations" flag (height < 0)
  int flag = 0;
  cudaError_t e = h2d_blocking(d_flag, &flag, sizeof(int));
  const double2 *d_uuv = reinterpret_cast<const double2 *>(uuv);
  if (e == cudaSuccess && !uuv_is_device) {
    e = cudaMalloc((void **) &d_own, n * sizeof(double2));
    if (e == cudaSuccess) e = h2d_blocking(d_own, uuv, n * sizeof(double2));
    d_uuv = d_own;
  }
  if (e == cudaSuccess) {
    const double scale = normalise ? 1.0 / ((double) nx * (double) g.ny) : 1.0;
    double *dst = h->d_phi + (size_t) (ky_first - g.ky0) * dsq * nx;
    const long long nq = (long long) nx * nky;
    const int grid = (int) ((nq + 63) / 64);
    const int lognx = ilog2_rt(nx);
    switch (d) {
      case 3: k_build_phi<3><<<grid, 64, 0, h->stream>>>(d_uuv, nx, nky, height, scale, h->phi_mode(), h->cols_top, lognx, dst, d_flag); break;
      case 6: k_build_phi<6><<<grid, 64, 0, h->stream>>>(d_uuv, nx, nky, height, scale, h->phi_mode(), h->cols_top, lognx, dst, d_flag); break;
      case 9: k_build_phi<9><<<grid, 64, 0, h->stream>>>(d_uuv, nx, nky, height, scale, h->phi_mode(), h->cols_top, lognx, dst, d_flag); break;
      default: k_build_phi<12><<<grid, 64, 0, h->stream>>>(d_uuv, nx, nky, height, scale, h->phi_mode(), h->cols_top, lognx, dst, d_flag); break;
    }
    h->launches++;
    e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
  }
  cudaFree(d_own);
  cudaFree(d_flag);
  CU(h, e);
  if (flag)      // the reference aborts here (iterate_Gnn, surface_stiffness.cpp:541-543)
    return fail(h, GFMD_B200_EPHI, "build_phi_columns: out of iterations while evaluating the continued fraction "
                "(height < 0, 100000 iterations, tolerance 1e-8)");
  for (int k = 0; k < nky; ++k) h->phi_cols_set[ky_first - g.ky0 + k] = 1;
  bool all = true;
  for (char c : h->phi_cols_set) all = all && c;
  h->phi_set = all;
  return 0;
}

int gfmd_b200_build_phi_columns(gfmd_b200_t *h, const double *uuv, int ky_first, int nky, int height,
                                int normalise)
{
  return build_phi_columns_impl(h, uuv, ky_first, nky, height, normalise, false);
}

int gfmd_b200_build_phi_columns_device(gfmd_b200_t *h, const double *d_uuv, int ky_first, int nky, int height,
                                       int normalise)
{
  return build_phi_columns_impl(h, d_uuv, ky_first, nky, height, normalise, true);
}

int gfmd_b200_phi_deviation(const gfmd_b200_t *h, double *herm_dev, double *conj_dev)
{
  if (!h) return GFMD_B200_EINVAL;
  if (herm_dev) *herm_dev = h->herm_dev;
  if (conj_dev) *conj_dev = h->conj_dev;
  return 0;
}

int gfmd_b200_post_force_device(gfmd_b200_t *h, const double *d_u, double *d_f)
{
  if (!h) return GFMD_B200_EINVAL;
  int rc = set_device(h);
  if (rc) return rc;
  return solver_step(h, d_u ? d_u : h->d_u, d_f ? d_f : h->d_f);
}

int gfmd_b200_pre_force_async_host(gfmd_b200_t *h, const double *u)
{
  if (!h || !u) return fail(h, GFMD_B200_EINVAL, "pre_force_async_host: null argument");
  int rc = set_device(h);
  if (rc) return rc;
  const size_t bytes = sizeof(double) * (size_t) h->g.d * h->g.nx_loc * h->g.ny;
  try_pin(h, u, bytes);
  h->hp_pending = false;
  if (h->hp_enabled && h->g.P == 1 && h->fast_rows && !h->want_graph && !h->profiling) {
    if (!h->phi_set) return fail(h, GFMD_B200_ESTATE, "post_force before set_phi (set_kernel)");
    rc = enqueue_solver_hostpipe(h, u);
  } else {
    CU(h, cudaMemcpyAsync(h->d_u, u, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = solver_step(h, h->d_u, h->d_f);
  }
  if (rc) return rc;
  h->pending_u = u;
  return 0;
}

int gfmd_b200_post_force_host(gfmd_b200_t *h, const double *u, double *f, double *epot, double *u0)
{
  if (!h || !u || !f) return fail(h, GFMD_B200_EINVAL, "post_force_host: null argument");
  int rc = set_device(h);
  if (rc) return rc;
  const size_t bytes = sizeof(double) * (size_t) h->g.d * h->g.nx_loc * h->g.ny;
  if (h->pending_u != u) {
    rc = gfmd_b200_pre_force_async_host(h, u);
    if (rc) return rc;
  }
  h->pending_u = nullptr;
  try_pin(h, f, bytes);
  if (h->hp_pending) {
    // download each dof as soon as its rows are back in real space
    const size_t nxy = (size_t) h->g.nx_loc * h->g.ny;
    for (int dof = 0; dof < h->g.d; ++dof) {
      CU(h, cudaStreamWaitEvent(h->hp_stream[1], h->hp_out[dof], 0));
      CU(h, cudaMemcpyAsync(f + dof * nxy, h->d_f + dof * nxy, nxy * sizeof(double), cudaMemcpyDeviceToHost,
                            h->hp_stream[1]));
    }
    CU(h, cudaEventRecord(h->hp_done, h->hp_stream[1]));
    CU(h, cudaStreamWaitEvent(h->stream, h->hp_done, 0));
    h->hp_pending = false;
  } else {
    CU(h, cudaMemcpyAsync(f, h->d_f, bytes, cudaMemcpyDeviceToHost, h->stream));
  }
  rc = fetch_results(h);
  if (rc) return rc;
  if (epot) *epot = h->h_res->epot;
  if (u0)
    for (int i = 0; i < h->g.d; ++i) u0[i] = h->h_res->u0[i];
  return 0;
}

int gfmd_b200_spectrum_host(gfmd_b200_t *h, const double *u, double *uq, double *fq)
{
  if (!h || !u || !uq) return fail(h, GFMD_B200_EINVAL, "spectrum_host: null argument");
  if (!h->phi_set && fq) return fail(h, GFMD_B200_ESTATE, "spectrum_host before set_phi (set_kernel)");
  if (h->pending_u) return fail(h, GFMD_B200_ESTATE, "spectrum_host between pre_force and post_force");
  int rc = set_device(h);
  if (rc) return rc;
  const GridDesc &g = h->g;
  const int d = g.d, nx = g.nx, ny = g.ny, nyh = g.nyh;
  const size_t bytes = sizeof(double) * (size_t) d * nx * ny;
  CU(h, cudaMemcpyAsync(h->d_u, u, bytes, cudaMemcpyHostToDevice, h->stream));
  rc = enqueue_aux(h, AUX_SPECTRUM, h->d_u, nullptr, fq != nullptr, 0);
  if (rc) return rc;
  const size_t nhalf = (size_t) d * nyh * nx;
  std::vector<double2> half(nhalf);
  for (int which = 0; which < (fq ? 2 : 1); ++which) {
    CU(h, cudaMemcpyAsync(half.data(), which ? h->d_spec : h->d_stage, nhalf * sizeof(double2),
                          cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    // half spectrum [dof][ky][kx] -> the reference's q_buffer [ix*ny + iy][dof]; the other half
    // by X(-q) = conj X(q) (real field; Phi(-q) = conj Phi(q))
    double2 *out = reinterpret_cast<double2 *>(which ? fq : uq);
    for (int ix = 0; ix < nx; ++ix)
      for (int iy = 0; iy < ny; ++iy) {
        const bool lower = iy < nyh;
        const int ky = lower ? iy : ny - iy, kx = lower ? ix : (nx - ix) % nx;
        for (int dof = 0; dof < d; ++dof) {
          double2 v = half[((size_t) dof * nyh + ky) * nx + kx];
          if (!lower) v.y = -v.y;
          out[((size_t) ix * ny + iy) * d + dof] = v;
        }
      }
  }
  return 0;
}

int gfmd_b200_prec_gradient_host(gfmd_b200_t *h, const double *cavg, const double *grad, double *gP,
                                 int first3_only)
{
  if (!h || !cavg || !grad || !gP) return fail(h, GFMD_B200_EINVAL, "prec_gradient_host: null argument");
  if (!h->phi_set) return fail(h, GFMD_B200_ESTATE, "prec_gradient before set_phi (set_kernel)");
  if (h->pending_u) return fail(h, GFMD_B200_ESTATE, "prec_gradient between pre_force and post_force");
  int rc = set_device(h);
  if (rc) return rc;
  const GridDesc &g = h->g;
  const size_t bytes = sizeof(double) * (size_t) g.d * g.nx_loc * g.ny;
  if (!h->d_cavg) CU(h, dmalloc(h, &h->d_cavg, (size_t) GFMD_B200_MAX_NDOF * GFMD_B200_MAX_NDOF));
  CU(h, cudaMemcpyAsync(h->d_cavg, cavg, sizeof(double) * g.d * g.d, cudaMemcpyHostToDevice, h->stream));
  // g / gP belong to the minimiser, which may reallocate them: never page-lock them here
  CU(h, cudaMemcpyAsync(h->d_u, grad, bytes, cudaMemcpyHostToDevice, h->stream));
  rc = enqueue_aux(h, AUX_PREC, h->d_u, h->d_f, false, first3_only ? 3 : g.d);
  if (rc) return rc;
  CU(h, cudaMemcpyAsync(gP, h->d_f, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int gfmd_b200_gather(gfmd_b200_t *h, const double *d_x, const double *d_xeq, int *d_gid, const int *d_mask,
                     int groupbit, int nall, double xprd, double yprd, int dxshift, int dyshift, double *d_u)
{
  if (!h || !d_x || !d_xeq || !d_gid || !d_mask) return fail(h, GFMD_B200_EINVAL, "gather: null argument");
  int rc = set_device(h);
  if (rc) return rc;
  CU(h, cudaMemsetAsync(&h->d_res->natoms_gathered, 0, sizeof(int), h->stream));
  CU(h, cudaMemsetAsync(&h->d_res->n_out_of_range, 0, sizeof(int), h->stream));
  stage_mark(h, 0);
  if (nall > 0) {
    k_gather<<<atom_blocks(h, nall), kAtomTile, 0, h->stream>>>(d_x, d_xeq, d_gid, d_mask, groupbit, nall, h->g, xprd,
                                                        yprd, dxshift, dyshift, d_u ? d_u : h->d_u, h->d_res);
    h->launches++;
  }
  if (h->profiling) {
    cudaEventRecord(h->ev[1], h->stream);
    CU(h, cudaStreamSynchronize(h->stream));
    accumulate_stage_times(h, 0, 0);
  }
  CU(h, cudaGetLastError());
  return 0;
}

int gfmd_b200_scatter(gfmd_b200_t *h, const double *d_fgrid, const int *d_gid, const int *d_mask, int groupbit,
                      int nall, int nlocal, double *d_f)
{
  if (!h || !d_gid || !d_mask || !d_f) return fail(h, GFMD_B200_EINVAL, "scatter: null argument");
  int rc = set_device(h);
  if (rc) return rc;
  const int nblk = atom_blocks(h, nall);
  if (nblk > h->fsum_part_cap) {
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_fsum_part);
    h->d_fsum_part = nullptr;
    CU(h, dmalloc(h, &h->d_fsum_part, (size_t) 3 * nblk));
    h->fsum_part_cap = nblk;
  }
  CU(h, cudaMemsetAsync(&h->d_res->natoms_scattered, 0, sizeof(int), h->stream));
  if (h->profiling) cudaEventRecord(h->ev[6], h->stream);
  if (nall > 0) {
    k_scatter<<<nblk, kAtomTile, 0, h->stream>>>(d_fgrid ? d_fgrid : h->d_f, d_gid, d_mask, groupbit, nall, nlocal,
                                           h->g, d_f, h->d_fsum_part, h->d_res);
    k_sum_partials<<<1, 256, 0, h->stream>>>(h->d_fsum_part, nblk, 3, h->d_res->fsum);
    h->launches += 2;
  } else {
    CU(h, cudaMemsetAsync(h->d_res->fsum, 0, 3 * sizeof(double), h->stream));
  }
  if (h->profiling) {
    cudaEventRecord(h->ev[7], h->stream);
    CU(h, cudaStreamSynchronize(h->stream));
    accumulate_stage_times(h, 6, 6);
  }
  CU(h, cudaGetLastError());
  return 0;
}

int gfmd_b200_build_cell_map(gfmd_b200_t *h, int *d_gid, const int *d_mask, int groupbit, int nall, int nlocal,
                             int dxshift, int dyshift, int *usable)
{
  if (!h || !d_gid || !d_mask) return fail(h, GFMD_B200_EINVAL, "build_cell_map: null argument");
  if (usable) *usable = 0;
  int rc = set_device(h);
  if (rc) return rc;
  const GridDesc &g = h->g;
  const size_t ncell = (size_t) (g.d / 3) * g.nx_loc * g.ny;
  h->cmap_valid = false;
  if (!h->d_cmap) {
    CU(h, dmalloc(h, &h->d_cmap, ncell));
    CU(h, dmalloc(h, &h->d_cmap_cnt, (size_t) 4));
  }
  CU(h, cudaMemsetAsync(h->d_cmap, 0xff, sizeof(int) * ncell, h->stream));          // -1: empty cell
  CU(h, cudaMemsetAsync(h->d_cmap_cnt, 0, sizeof(int) * 4, h->stream));
  if (nall > 0) {
    k_build_cellmap<<<atom_blocks(h, nall), kAtomTile, 0, h->stream>>>(d_gid, d_mask, groupbit, nall, g, dxshift,
                                                                       dyshift, h->d_cmap, h->d_cmap_cnt);
    h->launches++;
  }
  CU(h, cudaMemcpyAsync(h->cmap_cnt, h->d_cmap_cnt, sizeof(int) * 3, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaGetLastError());
  // usable: every cell of the brick holds exactly one atom of the group (then gather and scatter
  // through the map are the reference's loops exactly), and this handle's row kernels have the fused form
  const bool bijective = h->cmap_cnt[1] == 0 && h->cmap_cnt[2] == 0 && (size_t) h->cmap_cnt[0] == ncell;
  h->cmap_valid = bijective && fast_rows_has_atomio(h->fast_rows);
  h->cmap_gid = d_gid; h->cmap_mask = d_mask;
  h->cmap_groupbit = groupbit; h->cmap_nall = nall; h->cmap_nlocal = nlocal;
  if (h->cmap_valid && !h->d_fsum_io) {
    FastRowsCfg frc;
    fast_rows_cfg(h->fast_rows, frc);
    CU(h, dmalloc(h, &h->d_fsum_io, (size_t) (g.nx_loc / frc.rb) * g.d));
  }
  if (usable) *usable = h->cmap_valid ? 1 : 0;
  return 0;
}

int gfmd_b200_drop_cell_map(gfmd_b200_t *h)
{
  if (!h) return GFMD_B200_EINVAL;
  h->cmap_valid = false;
  return 0;
}

int gfmd_b200_full_step(gfmd_b200_t *h, const double *d_x, const double *d_xeq, int *d_gid, const int *d_mask,
                        int groupbit, int nall, int nlocal, double xprd, double yprd, double *d_f)
{
  if (!h) return GFMD_B200_EINVAL;
  // fused path: a valid cell map built from exactly these arrays (gfmd_b200_build_cell_map)
  const bool mapped = h->cmap_valid && h->cmap_gid == d_gid && h->cmap_mask == d_mask && h->cmap_groupbit == groupbit &&
                      h->cmap_nall == nall && h->cmap_nlocal == nlocal && d_x && d_xeq && d_f && !h->want_graph;
  if (!mapped) {
    int rc = gfmd_b200_gather(h, d_x, d_xeq, d_gid, d_mask, groupbit, nall, xprd, yprd, 0, 0, nullptr);
    if (rc) return rc;
    rc = solver_step(h, h->d_u, h->d_f);
    if (rc) return rc;
    return gfmd_b200_scatter(h, nullptr, d_gid, d_mask, groupbit, nall, nlocal, d_f);
  }
  int rc = set_device(h);
  if (rc) return rc;
  AtomIO io{};
  io.x = d_x; io.xeq = d_xeq; io.fat = d_f; io.cmap = h->d_cmap; io.fsum_part = h->d_fsum_io;
  io.xprd = xprd; io.yprd = yprd; io.nlocal = nlocal;
  // the multi-GPU pipeline launches its forward rows dof by dof (each launch would re-read the atoms):
  // separate gather there, fused scatter everywhere
  const bool fuse_gather = !step_is_pipelined(h);
  if (!fuse_gather) {
    rc = gfmd_b200_gather(h, d_x, d_xeq, d_gid, d_mask, groupbit, nall, xprd, yprd, 0, 0, nullptr);
    if (rc) return rc;
  } else if (h->profiling) {
    cudaEventRecord(h->ev[0], h->stream);          // stage 0 (gather) is empty: it lives inside rows_fwd
    cudaEventRecord(h->ev[1], h->stream);
  }
  h->io_fwd = fuse_gather ? &io : nullptr;
  h->io_inv = &io;
  rc = solver_step(h, h->d_u, h->d_f);
  h->io_fwd = h->io_inv = nullptr;
  if (rc) return rc;
  FastRowsCfg frc;
  fast_rows_cfg(h->fast_rows, frc);
  k_sum_fsum_io<<<1, 256, 0, h->stream>>>(h->d_fsum_io, h->g.nx_loc / frc.rb, h->g.d, h->d_res->fsum);
  h->launches++;
  // counters as the separate kernels report them: every atom of the map is gathered and scattered
  if (fuse_gather) {
    CU(h, cudaMemcpyAsync(&h->d_res->natoms_gathered, h->d_cmap_cnt, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    CU(h, cudaMemsetAsync(&h->d_res->n_out_of_range, 0, sizeof(int), h->stream));
  }
  CU(h, cudaMemcpyAsync(&h->d_res->natoms_scattered, h->d_cmap_cnt, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  CU(h, cudaGetLastError());
  return 0;
}

int gfmd_b200_get_results(gfmd_b200_t *h, double *epot, double *u0, double fsum[3], int counters[3])
{
  if (!h) return GFMD_B200_EINVAL;
  int rc = set_device(h);
  if (rc) return rc;
  rc = fetch_results(h);
  if (rc) return rc;
  if (epot) *epot = h->h_res->epot;
  if (u0)
    for (int i = 0; i < h->g.d; ++i) u0[i] = h->h_res->u0[i];
  if (fsum)
    for (int i = 0; i < 3; ++i) fsum[i] = h->h_res->fsum[i];
  if (counters) {
    counters[0] = h->h_res->natoms_gathered;
    counters[1] = h->h_res->natoms_scattered;
    counters[2] = h->h_res->n_out_of_range;
  }
  return 0;
}

double *gfmd_b200_device_u(gfmd_b200_t *h) { return h ? h->d_u : nullptr; }
double *gfmd_b200_device_f(gfmd_b200_t *h) { return h ? h->d_f : nullptr; }
void *gfmd_b200_stream(gfmd_b200_t *h) { return h ? (void *) h->stream : nullptr; }

int gfmd_b200_set_stream(gfmd_b200_t *h, void *s)
{
  if (!h) return GFMD_B200_EINVAL;
  int rc = set_device(h);
  if (rc) return rc;
  CU(h, cudaStreamSynchronize(h->stream));
  drop_graph(h);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t) s;
  h->own_stream = false;
  return 0;
}

int gfmd_b200_synchronize(gfmd_b200_t *h)
{
  if (!h) return GFMD_B200_EINVAL;
  int rc = set_device(h);
  if (rc) return rc;
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int gfmd_b200_pin_host_buffers(gfmd_b200_t *h, int on)
{
  if (!h) return GFMD_B200_EINVAL;
  h->pin_host = on != 0;
  if (!on) {
    for (auto &kv : h->pinned) cudaHostUnregister(const_cast<void *>(kv.first));
    cudaGetLastError();
    h->pinned.clear();
  }
  return 0;
}

int gfmd_b200_host_pipeline(gfmd_b200_t *h, int on)
{
  if (!h) return GFMD_B200_EINVAL;
  if (h->pending_u) return fail(h, GFMD_B200_ESTATE, "host_pipeline between pre_force and post_force");
  if (on >= 0) h->hp_enabled = on != 0;
  return (h->hp_enabled && h->g.P == 1 && h->fast_rows) ? 1 : 0;     // whether it takes effect
}

int gfmd_b200_use_graph(gfmd_b200_t *h, int on)
{
  if (!h) return GFMD_B200_EINVAL;
  h->want_graph = on != 0;
  if (!on) drop_graph(h);
  return 0;
}

long long gfmd_b200_launch_count(const gfmd_b200_t *h) { return h ? h->launches : 0; }

int gfmd_b200_profile(gfmd_b200_t *h, int on)
{
  if (!h) return GFMD_B200_EINVAL;
  h->profiling = on != 0;
  if (on) {
    for (int i = 0; i < GFMD_B200_NSTAGES; ++i) {
      h->stage_ms[i] = 0.0;
      h->stage_cnt[i] = 0;
    }
  }
  return 0;
}

int gfmd_b200_get_stage_times(gfmd_b200_t *h, double ms[GFMD_B200_NSTAGES], long long counts[GFMD_B200_NSTAGES])
{
  if (!h) return GFMD_B200_EINVAL;
  for (int i = 0; i < GFMD_B200_NSTAGES; ++i) {
    if (ms) ms[i] = h->stage_ms[i];
    if (counts) counts[i] = h->stage_cnt[i];
  }
  return 0;
}

const char *gfmd_b200_describe(gfmd_b200_t *h) { return h ? h->desc.c_str() : ""; }

double gfmd_b200_memory_usage(const gfmd_b200_t *h) { return h ? h->bytes : 0.0; }

}  // extern "C"
