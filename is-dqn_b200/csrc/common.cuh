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

// Function attributes (opt-in shared-memory sizes), occupancy results and side streams belong to ONE device context: caches
// of them are kept per device, so that a process driving several GPUs (two agents behind torch.cuda.set_device) gets the
// opt-in on each of them.  `first()` is true once per device for each PerDevice object.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
  bool first() {
    const int d = current_device();
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
template <typename T>
struct PerDevice {
  T v[kMaxDevices] = {};
  T& get() { return v[current_device()]; }
};
// SM count of the current device (cached); the grid-sizing heuristics use the B200 constant kNumSMs, the kernels whose
// CORRECTNESS depends on co-residency (acting: grid barrier) use this
int num_sms();

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
int fork_mode();
cudaStream_t side_stream(int i);
cudaEvent_t side_event(int i);

// Programmatic dependent launch (PDL).  The kernels of one learner step form a chain of short dependent launches; with
// the launch attribute below the NEXT kernel's CTAs are scheduled (and run their prologue: barrier init, TMEM
// allocation) while the current kernel still executes, and block in pdl_wait() until it has completed and its memory
// is visible.  Every kernel launched through launch_pdl calls pdl_sync() (or trigger + wait) before it touches global
// memory, so the chain keeps full stream-order semantics (completion is transitive: a grid cannot complete before
// the grids it waited on).  A kernel that allocates TMEM triggers only AFTER its allocation (a dependent that grabbed
// the columns first would wait on a grid that waits on it).  On by default since the TMA-fed kernels (ISDQN_PDL=0
// switches it off; with the LDGSTS-era kernels it had measured 5 % slower); the instructions are no-ops without the
// attribute.
bool pdl_enabled();

// Optional device-side timeline (isdqn_trace_set): CTA (0,0,0) of every step kernel appends its start time (globaltimer,
// ns) to a buffer — [0] = count, [1..] = times — which gives kernel-to-kernel intervals INSIDE a graph replay, where
// neither events nor a profiler can look without perturbing it.  One (static) pointer per translation unit.
static __device__ unsigned long long* g_trace_buf = nullptr;
static inline cudaError_t trace_set_local(unsigned long long* p) { return cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)); }
// entry = (tag << 56) | time; tag 0 = kernel start, 1.. = phase marks inside tc_gemm_kernel (any one thread of CTA 0)
__device__ __forceinline__ void trace_mark(unsigned tag) {
  if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
    unsigned long long* t = g_trace_buf;
    if (t) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      const unsigned long long i = atomicAdd(t, 1ull);
      if (i < 4000) t[1 + i] = ((unsigned long long)tag << 56) | (now & 0x00ffffffffffffffull);
    }
  }
}
__device__ __forceinline__ void trace_kernel_start() {
  if (threadIdx.x == 0) trace_mark(0);
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ISDQN_PDL_LATE = 1 (default): trigger AFTER the wait, so that only the direct successor of the running kernel is
// resident early instead of the whole chain cascading onto the SMs; 0 = trigger first (the earlier placement).  Both
// measured 141.5-142.1 us per batch-32 step against 149.5-156 us without PDL (profiles/r01_summary.md).
#ifndef ISDQN_PDL_LATE
#define ISDQN_PDL_LATE 1
#endif
__device__ __forceinline__ void pdl_wait_then_trigger() {
  pdl_wait();
#if ISDQN_PDL_LATE
  pdl_trigger();
#endif
}
__device__ __forceinline__ void pdl_sync() {
  trace_kernel_start();
#if ISDQN_PDL_LATE
  pdl_wait();
  pdl_trigger();
#else
  pdl_trigger();
  pdl_wait();
#endif
}

// Experiment (ISDQN_CARVEOUT=1): give every kernel of the step the same (maximum) shared-memory carve-out so that the
// SMs never reconfigure the L1/shared split between a tensor-core kernel and a small streaming one.  Measured slower.
void prefer_max_shared_once(const void* kernel);
void prefer_max_shared(const void* kernel);  // unconditional (kernels that must co-reside with the tensor-core kernels)
template <typename... KArgs>
static inline void co_resident_with_tc(void (*kernel)(KArgs...)) {
  prefer_max_shared(reinterpret_cast<const void*>(kernel));
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
  prefer_max_shared_once(reinterpret_cast<const void*>(kernel));
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

// L2 eviction-priority hints (createpolicy + ld/st .L2::cache_hint).  The optimiser state of the Atari network
// (parameters, both Adam moments, the bf16 shadow: 57 MB) fits the 126 MB L2 next to everything else a batch-32 step
// touches; marking it evict_last keeps it resident from one step to the next, so Adam streams from L2, not HBM.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_f4_hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_f4_hint(float4* ptr, const float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_u2_hint(uint2* ptr, const uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(ptr), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
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
