// Shared device/host helpers for libisdqn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/isdqn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libisdqn_b200 targets sm_100a (B200) only"
#endif

namespace isdqn {

constexpr int kNumSMs = 148;  // B200

void set_last_cuda_error(cudaError_t e, const char* where);

#define ISDQN_CUDA_CHECK(expr)                              \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) {                                \
      ::isdqn::set_last_cuda_error(_e, #expr);              \
      return ISDQN_E_CUDA;                                  \
    }                                                       \
  } while (0)

#define ISDQN_LAUNCH_CHECK() ISDQN_CUDA_CHECK(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Optional per-kernel timing (isdqn_profile_begin/end): every launch site marks its name; a mark records a CUDA
// event on the launching stream, so consecutive marks bracket one kernel.  Off (one predictable branch) by default.
extern bool g_profile_on;
void profile_mark(cudaStream_t s, const char* name);
#define ISDQN_PROF(stream, name)                              \
  do {                                                        \
    if (::isdqn::g_profile_on) ::isdqn::profile_mark(stream, name); \
  } while (0)

// two-stream backward pass (runtime.cu)
constexpr int kSideEvents = 16, kSideStreams = 2;
bool fork_enabled();
cudaStream_t side_stream(int i);
cudaEvent_t side_event(int i);

// Programmatic dependent launch (PDL).  The kernels of one learner step form a chain of short dependent launches; with
// the launch attribute below the NEXT kernel's CTAs are scheduled (and run their prologue: barrier init, TMEM
// allocation) while the current kernel still executes, and block in pdl_wait() until it has completed and its memory
// is visible.  Every kernel launched through launch_pdl calls pdl_sync() (or trigger + wait) before it touches global
// memory, so the chain keeps full stream-order semantics (completion is transitive: a grid cannot complete before
// the grids it waited on).  A kernel that allocates TMEM triggers only AFTER its allocation (a dependent that grabbed
// the columns first would wait on a grid that waits on it).  Opt-in with ISDQN_PDL=1 (measured slower at batch 32:
// early-resident dependents compete with the running kernel); the instructions are no-ops without the attribute.
bool pdl_enabled();
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_trigger();
  pdl_wait();
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Exclusive prefix sum of one int per thread over the whole CTA (THREADS <= 1024, multiple of 32).
// warp_tot: 32 ints of shared memory, total: 1 int of shared memory.  Contains two __syncthreads().
template <int THREADS>
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot, int* total) {
  // v: per-thread count; returns the exclusive prefix over the block; *total = block sum.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < THREADS / 32 ? warp_tot[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += y;
    }
    if (lane < THREADS / 32) warp_tot[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  const int res = warp_tot[warp] + inc - v;
  return res;
}

}  // namespace isdqn
