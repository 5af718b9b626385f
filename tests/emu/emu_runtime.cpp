/* TEST INFRASTRUCTURE ONLY -- runtime of the CUDA-on-CPU emulation shim (see
 * tests/emu/include/cuda_runtime.h).  One OS thread; the threads of a block are cooperative
 * fibers (ucontext) scheduled round robin, blocks run one after the other:
 *   __syncthreads / __syncwarp   barriers over the live fibers of the block / warp
 *   __shfl_down_sync             exchange through a per-warp slot array between two warp barriers
 * Execution is deterministic; a fiber runs until it reaches a barrier or returns, so a
 * missing barrier shows up as a wrong (often NaN) result rather than as a rare race.
 */
#include "cuda_runtime.h"

#include <sys/mman.h>
#include <sched.h>
#include <ucontext.h>

#include <chrono>
#include <cstdio>
#include <mutex>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace {

constexpr size_t kStack = 512 * 1024;

struct Fiber {
  ucontext_t ctx;
  void *stack = nullptr;
  bool done = true;
  uint3 tidx{};
};

struct Bar {
  int count = 0;
  unsigned gen = 0;
};

std::vector<Fiber> g_fib;
ucontext_t g_main;
int g_cur = 0, g_n = 0, g_live = 0;
const std::function<void()> *g_body = nullptr;
Bar g_bar_block, g_bar_warp[32];
int g_warp_live[32];
uint64_t g_shfl[32][32];
void *g_dyn = nullptr;

void switch_to(int next)
{
  const int prev = g_cur;
  g_cur = next;
  threadIdx = g_fib[next].tidx;
  swapcontext(&g_fib[prev].ctx, &g_fib[next].ctx);
  threadIdx = g_fib[g_cur].tidx;
}

// Scheduling order of the fibers of a block: g_order is a permutation of the thread indices, the
// scheduler walks it round robin.  GFMD_EMU_SCHED = "reverse" or "random[:seed]" (default: thread
// order) -- CUDA promises no order between barriers, so every order must give the same results;
// running the tests under several orders exposes missing barriers.
std::vector<int> g_order, g_pos;
int g_sched_mode = -1;        // 0 thread order, 1 reverse, 2 random
unsigned long long g_sched_state = 0x9E3779B97F4A7C15ull;

void make_order(int n)
{
  if (g_sched_mode < 0) {
    const char *e = getenv("GFMD_EMU_SCHED");
    g_sched_mode = 0;
    if (e && !strncmp(e, "reverse", 7)) g_sched_mode = 1;
    if (e && !strncmp(e, "random", 6)) {
      g_sched_mode = 2;
      if (e[6] == ':') g_sched_state ^= strtoull(e + 7, nullptr, 10) * 0xD1342543DE82EF95ull;
    }
  }
  g_order.resize(n);
  g_pos.resize(n);
  for (int i = 0; i < n; ++i) g_order[i] = g_sched_mode == 1 ? n - 1 - i : i;
  if (g_sched_mode == 2)
    for (int i = n - 1; i > 0; --i) {                      // Fisher-Yates with xorshift64*
      g_sched_state ^= g_sched_state >> 12;
      g_sched_state ^= g_sched_state << 25;
      g_sched_state ^= g_sched_state >> 27;
      const int j = (int) ((g_sched_state * 0x2545F4914F6CDD1Dull >> 33) % (unsigned) (i + 1));
      const int t = g_order[i]; g_order[i] = g_order[j]; g_order[j] = t;
    }
  for (int i = 0; i < n; ++i) g_pos[g_order[i]] = i;
}

int next_live(int from)
{
  const int p = g_pos[from];
  for (int k = 1; k <= g_n; ++k) {
    const int c = g_order[(p + k) % g_n];
    if (!g_fib[c].done) return c;
  }
  return -1;
}

void yield()
{
  const int nx = next_live(g_cur);
  if (nx >= 0 && nx != g_cur) switch_to(nx);
}

void release_if_complete(Bar &b, int live)
{
  if (b.count > 0 && b.count >= live) {
    b.count = 0;
    b.gen++;
  }
}

void bar_wait(Bar &b, const int &live, const char *what)
{
  const unsigned gen = b.gen;
  if (++b.count >= live) {
    b.count = 0;
    b.gen++;
    return;
  }
  long spins = 0;
  while (b.gen == gen) {
    yield();
    if (++spins > 50000000L) {
      fprintf(stderr, "cuda emu: deadlock in %s (block %u, thread %d): not every live thread reaches it\n", what,
              blockIdx.x, g_cur);
      abort();
    }
  }
}

void fiber_main()
{
  (*g_body)();
  Fiber &f = g_fib[g_cur];
  f.done = true;
  --g_live;
  const int w = g_cur / 32;
  --g_warp_live[w];
  release_if_complete(g_bar_block, g_live);
  release_if_complete(g_bar_warp[w], g_warp_live[w]);
  const int nx = next_live(g_cur);
  if (nx < 0) {
    setcontext(&g_main);
  } else {
    g_cur = nx;
    threadIdx = g_fib[nx].tidx;
    setcontext(&g_fib[nx].ctx);
  }
}

}  // namespace

namespace emu {

void sync_block() { bar_wait(g_bar_block, g_live, "__syncthreads"); }

void sync_warp() { bar_wait(g_bar_warp[g_cur / 32], g_warp_live[g_cur / 32], "__syncwarp"); }

uint64_t shfl_down_raw(uint64_t v, int delta)
{
  const int w = g_cur / 32, lane = g_cur % 32;
  g_shfl[w][lane] = v;
  sync_warp();
  const int src = lane + delta;
  const uint64_t r = (src < 32 && w * 32 + src < g_n) ? g_shfl[w][src] : v;
  sync_warp();
  return r;
}

void *dyn_smem() { return g_dyn; }

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
  // one launch at a time: several host threads (the ranks of an emulated multi-GPU run, see
  // tests/test_emulated_multi_rank.py) share the one fiber scheduler
  static std::mutex launch_mutex;
  std::lock_guard<std::mutex> lock(launch_mutex);
  const int n = (int) (block.x * block.y * block.z);
  if (n < 1 || n > 1024 || smem_bytes > 232448) {
    fprintf(stderr, "cuda emu: block of %d threads, %zu B of dynamic shared memory\n", n, smem_bytes);
    abort();
  }
  if ((int) g_fib.size() < n) g_fib.resize(n);
  for (int t = 0; t < n; ++t)
    if (!g_fib[t].stack) {
      g_fib[t].stack = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
      if (g_fib[t].stack == MAP_FAILED) abort();
    }
  void *dyn = nullptr;
  if (posix_memalign(&dyn, 128, smem_bytes ? smem_bytes : 16)) abort();
  g_body = &body;
  g_n = n;
  blockDim = block;
  gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        memset(dyn, 0xff, smem_bytes ? smem_bytes : 16);      /* NaN: reads of unwritten smem show */
        g_dyn = dyn;
        g_bar_block = Bar();
        for (int w = 0; w < 32; ++w) {
          g_bar_warp[w] = Bar();
          g_warp_live[w] = 0;
        }
        for (int t = 0; t < n; ++t) {
          Fiber &f = g_fib[t];
          f.done = false;
          f.tidx.x = t % block.x;
          f.tidx.y = (t / block.x) % block.y;
          f.tidx.z = t / (block.x * block.y);
          g_warp_live[t / 32]++;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_main, 0);
        }
        g_live = n;
        make_order(n);
        g_cur = g_order[0];
        threadIdx = g_fib[g_cur].tidx;
        swapcontext(&g_main, &g_fib[g_cur].ctx);
      }
  g_dyn = nullptr;
  free(dyn);
}

}  // namespace emu

/* ---- runtime API -------------------------------------------------------------------------- */

struct emu_stream { int unused; };
struct emu_event { double t_ms; };
static emu_stream g_stream_obj;

static double now_ms()
{
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

const char *cudaGetErrorString(cudaError_t e)
{
  switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorNotSupported: return "operation not supported (CPU emulation)";
    case cudaErrorMemoryAllocation: return "out of memory";
    default: return "error (CPU emulation)";
  }
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 3; return cudaSuccess; }   /* "SMs" */
cudaError_t cudaMalloc(void **p, size_t bytes)
{
  if (posix_memalign(p, 256, bytes ? bytes : 256)) return cudaErrorMemoryAllocation;
  memset(*p, 0xff, bytes);                                  /* NaN: reads of unwritten memory show */
  return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t bytes) { *p = malloc(bytes); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
  for (size_t r = 0; r < h; ++r) memmove((char *) d + r * dp, (const char *) s + r * sp, w);
  return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = &g_stream_obj; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = &g_stream_obj; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = &g_stream_obj; return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *a, int *b) { *a = 0; *b = -1; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0.0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float) (b->t_ms - a->t_ms); return cudaSuccess; }
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) { *g = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
/* "inter-process" handles inside one process: the handle carries the pointer */
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned)
{
  memcpy(p, h.reserved, sizeof(*p));
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
cudaError_t emuStreamWriteValue32(cudaStream_t, unsigned *addr, unsigned value)
{
  __atomic_store_n(addr, value, __ATOMIC_RELEASE);
  return cudaSuccess;
}
cudaError_t emuStreamWaitValue32Geq(cudaStream_t, unsigned *addr, unsigned value)
{
  const auto t0 = std::chrono::steady_clock::now();
  while ((int) (__atomic_load_n(addr, __ATOMIC_ACQUIRE) - value) < 0) {
    sched_yield();
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(600)) {
      fprintf(stderr, "cuda emu: stream wait on a flag timed out (value %u never arrived)\n", value);
      return cudaErrorInvalidValue;
    }
  }
  return cudaSuccess;
}
