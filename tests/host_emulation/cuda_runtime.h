// TEST INFRASTRUCTURE ONLY -- host emulation of the small part of the CUDA runtime and of the
// CUDA execution model that ryujin_b200/csrc uses, so that the SAME kernel and host sources can
// be compiled with g++ into tests/host_emulation/libryujin_b200_hostemu.so and the GPU parity
// suite can run in CI on a box without a GPU (tests/test_host_emulation.py).
//
// This file shadows <cuda_runtime.h> on the include path of that build and nowhere else.  The
// product library (ryujin_b200/libryujin_b200.so) is built by nvcc for sm_100a, contains no host
// path for the kernels and fails with RJ_ERR_NO_DEVICE without a GPU.
//
// Execution model: a kernel launch runs its blocks one after the other; the threads of a block
// are ucontext coroutines scheduled round-robin, so __syncthreads() and __shfl_xor_sync() have
// their CUDA meaning (a thread that has exited counts as arrived).  "Device memory" is host
// memory, streams are synchronous, events are time stamps.  Stream capture records the kernel
// launches, async copies and (emulated) NCCL calls issued between Begin/EndCapture as closures
// without executing them; cudaGraphLaunch replays them in order -- the semantics the CUDA-graph
// stepping of the TimeIntegrator relies on (arguments frozen at capture, data read at replay).
// All engine state is thread_local: several ranks can run as host threads of one process, with
// the in-process NCCL emulation of nccl.h carrying the halo exchanges between them.
#pragma once

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ucontext.h>
#include <vector>

#define RJ_HOST_EMULATION 1

/* ---- language extensions ---- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local

struct rj_emu_dim3 {
  unsigned int x = 1, y = 1, z = 1;
};
extern thread_local rj_emu_dim3 threadIdx, blockIdx, blockDim, gridDim; /* one emulated device per host thread */

/* ---- error type and memory ---- */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "not supported by the host emulation"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
/* fresh "device" memory is poisoned (all bits set: NaN as double, 4294967295 as index) so that a kernel
 * that consumes memory nobody initialised fails the parity tests instead of reading convenient zeros */
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n)
{
  *p = static_cast<T *>(std::malloc(n ? n : 1));
  std::memset(*p, 0xff, n ? n : 1);
  return cudaSuccess;
}
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n)
{
  *p = static_cast<T *>(std::malloc(n ? n : 1));
  std::memset(*p, 0xff, n ? n : 1);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }

/* ---- streams, events, graphs ---- */
typedef struct rj_emu_stream *cudaStream_t;
typedef struct rj_emu_event *cudaEvent_t;
struct rj_emu_graph;
struct rj_emu_graph_exec;
typedef rj_emu_graph *cudaGraph_t;
typedef rj_emu_graph_exec *cudaGraphExec_t;
struct rj_emu_stream { int id; };
struct rj_emu_event { std::chrono::steady_clock::time_point t; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaStreamCaptureModeRelaxed = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new rj_emu_stream{0}; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = new rj_emu_stream{1}; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
/* stream capture: closures recorded instead of executed */
struct rj_emu_graph { std::vector<std::function<void()>> nodes; };
struct rj_emu_graph_exec { std::vector<std::function<void()>> nodes; };
extern thread_local rj_emu_graph *rj_emu_capture; /* non-null between Begin/EndCapture on this host thread */
template <typename F> inline void rj_emu_enqueue(F f)
{
  if (rj_emu_capture)
    rj_emu_capture->nodes.emplace_back(f);
  else
    f();
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t)
{
  rj_emu_enqueue([=]() { std::memcpy(d, s, n); });
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t)
{
  rj_emu_enqueue([=]() { std::memset(d, v, n); });
  return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new rj_emu_event(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new rj_emu_event(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; } /* everything runs eagerly */
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int)
{
  if (rj_emu_capture)
    return cudaErrorNotSupported;
  rj_emu_capture = new rj_emu_graph();
  return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g)
{
  *g = rj_emu_capture;
  rj_emu_capture = nullptr;
  return *g ? cudaSuccess : cudaErrorNotSupported;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long)
{
  *e = new rj_emu_graph_exec{g->nodes};
  return cudaSuccess;
}
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t)
{
  for (auto &node : e->nodes)
    node();
  return cudaSuccess;
}
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }

/* CUDA puts the C math classification functions into the global namespace */
using std::isinf;
using std::isnan;

/* ---- device intrinsics (the build uses -ffp-contract=off: IEEE mul / add like -fmad=false) ---- */
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline long long __double_as_longlong(double x) { long long u; std::memcpy(&u, &x, 8); return u; }
inline double __longlong_as_double(long long u) { double x; std::memcpy(&x, &u, 8); return x; }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const auto old = *p; if (v < old) *p = v; return old; }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const auto old = *p; if (v > old) *p = v; return old; }
inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { const unsigned int old = *p; *p = old + v; return old; }
inline void __threadfence() {}
inline int atomicOr(int *p, int v) { const int old = *p; *p = old | v; return old; }

/* ---- block execution engine (host_emulation.cc) ---- */
void rj_emu_syncthreads();
unsigned long long rj_emu_shfl_xor(unsigned long long value, int lane_mask);
inline void __syncthreads() { rj_emu_syncthreads(); }
inline unsigned long long __shfl_xor_sync(unsigned mask, unsigned long long value, int lane_mask)
{
  (void)mask;
  return rj_emu_shfl_xor(value, lane_mask);
}

struct rj_emu_thunk {
  virtual void run() = 0;
  virtual ~rj_emu_thunk() = default;
};
void rj_emu_run_grid(unsigned int grid, unsigned int block, rj_emu_thunk &kernel_call, size_t dynamic_smem = 0);
unsigned char *rj_emu_dynamic_smem(); /* the running block's dynamic shared memory (poisoned per block) */
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

template <typename F> struct rj_emu_call : rj_emu_thunk {
  F f;
  explicit rj_emu_call(F g) : f(g) {}
  void run() override { f(); }
};

/* the launch sites of the product (module.cu: launch(), launch_tile()) go through these macros */
#define RJ_LAUNCH_KERNEL_SMEM(kernel, grid, block, smem, stream, ...)                                        \
  do {                                                                                                       \
    const unsigned int rj_emu_grid = (grid), rj_emu_block = (block);                                         \
    const size_t rj_emu_smem = (smem);                                                                       \
    rj_emu_enqueue([=]() {                                                                                   \
      auto rj_emu_body = [&]() { kernel(__VA_ARGS__); };                                                     \
      rj_emu_call<decltype(rj_emu_body)> rj_emu_thunk_(rj_emu_body);                                         \
      rj_emu_run_grid(rj_emu_grid, rj_emu_block, rj_emu_thunk_, rj_emu_smem);                                \
    });                                                                                                      \
  } while (0)
#define RJ_LAUNCH_KERNEL(kernel, grid, block, stream, ...) RJ_LAUNCH_KERNEL_SMEM(kernel, grid, block, 0, stream, __VA_ARGS__)
