/* TEST INFRASTRUCTURE ONLY -- a CUDA-on-CPU emulation shim, never part of the product.
 *
 * Lets `tests/emu/build.py` compile the UNCHANGED kernel sources of user-gfmd_b200/csrc
 * (after a mechanical rewrite of the `<<< >>>` launches, tests/emu/preprocess.py) with g++ into
 * tests/emu/_build/libgfmd_b200_emu.so, so that the logic of every kernel -- indexing,
 * barriers, shuffles, twiddles, table layouts -- is exercised by `pytest -m "not gpu"` on
 * machines without a GPU.  Threads of a block run as cooperative fibers (emu_runtime.cpp);
 * dynamic shared memory and "device" allocations start filled with NaN so that reads of
 * never-written memory show.  It proves nothing about performance or about real data races,
 * and the product library neither contains nor looks for it: libgfmd_b200.so still fails
 * with GFMD_B200_ENOGPU when there is no device.
 */
#pragma once
#define GFMD_CUDA_EMU 1

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static

struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct alignas(8) int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

/* ---- device-side intrinsics ------------------------------------------------------------ */
namespace emu {
void sync_block();
void sync_warp();
uint64_t shfl_down_raw(uint64_t v, int delta);
void *dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
}

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int delta)
{
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  raw = emu::shfl_down_raw(raw, delta);
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }       /* fibers: one OS thread */
static inline double atomicAdd(double *p, double v) { double o = *p; *p += v; return o; }
static inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }
static inline long long clock64() { return 0; }
using std::fma;
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

/* ---- runtime API (synchronous: every "async" call completes before it returns) ---------- */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
typedef struct emu_graph *cudaGraph_t;
typedef struct emu_graph_exec *cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };

const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDevice(int *);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaDeviceGetAttribute(int *, cudaDeviceAttr, int);
cudaError_t cudaMalloc(void **, size_t);
cudaError_t cudaFree(void *);
cudaError_t cudaMallocHost(void **, size_t);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemcpy2DAsync(void *, size_t, const void *, size_t, size_t, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemset(void *, int, size_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t *);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *, unsigned, int);
cudaError_t cudaDeviceGetStreamPriorityRange(int *, int *);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
cudaError_t cudaHostRegister(void *, size_t, unsigned);
cudaError_t cudaHostUnregister(void *);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *, const void *);
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode);
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long);
cudaError_t cudaGraphDestroy(cudaGraph_t);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t);
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *);
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void *);
/* stream memory operations (the product uses the driver's cuStreamWriteValue32 / cuStreamWaitValue32):
 * streams are synchronous here, so the write happens at once and the wait spins (ranks = host threads) */
cudaError_t emuStreamWriteValue32(cudaStream_t, unsigned *addr, unsigned value);
cudaError_t emuStreamWaitValue32Geq(cudaStream_t, unsigned *addr, unsigned value);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F *, cudaFuncAttribute, int) { return cudaSuccess; }
