// Library plumbing: error strings, CUDA-graph helpers, NCCL (dlopen'ed) for the data-parallel mode.
#include <cstdlib>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace isdqn {
static thread_local char g_last_error[512] = "";
void set_last_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------- profiler
bool g_profile_on = false;

// Kernels whose shared-memory carve-outs differ cannot be resident on one SM at the same time (measured: a default-
// carve-out Adam on a side stream serialised against the tensor-core kernels).  The streaming kernels that run between
// the fork and the join of the backward pass therefore ask for the tensor-core kernels' carve-out (maximum shared);
// ISDQN_CARVEOUT=1 extends that to every kernel launched through launch_pdl (measured slower: the rest lose L1).
void prefer_max_shared(const void* kernel) {
  static const void* seen[kMaxDevices][256];
  static int n_seen[kMaxDevices] = {};
  const int d = current_device();
  for (int i = 0; i < n_seen[d]; ++i)
    if (seen[d][i] == kernel) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  if (n_seen[d] < 256) seen[d][n_seen[d]++] = kernel;
}
void prefer_max_shared_once(const void* kernel) {
  static const bool all = [] {
    const char* e = getenv("ISDQN_CARVEOUT");
    return e && e[0] == '1';
  }();
  if (all) prefer_max_shared(kernel);
}

bool pdl_enabled() {
  static const bool on = [] {
    // on by default: 141.7 vs 150.9 us per batch-32 step on the same box (profiles/r01_summary.md); ISDQN_PDL=0 disables
    const char* e = getenv("ISDQN_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
// Side stream + events of the two-stream backward pass (tc_learner.cu).  One set per process: the learner entry points
// are not re-entrant across host threads.  ISDQN_FORK=1 enables it (measured slower at batch 32).
int fork_mode() {
  static const int mode = [] {
    // 0 (default) one stream | 1 weight gradients + early Adam on side streams | 2 early Adam only.  Measured at batch
    // 32 (profiles/r01_summary.md): neither beats the single stream once every kernel of the backward pass shares one
    // carve-out — the cross-stream graph edges cost what the overlap gains.
    const char* e = getenv("ISDQN_FORK");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0;
  }();
  return mode;
}
// (streams and events belong to one device: one set per device)
cudaStream_t side_stream(int i) {
  static cudaStream_t s[kMaxDevices][kSideStreams] = {};
  static bool tried[kMaxDevices][kSideStreams] = {};
  if (i < 0 || i >= kSideStreams) return nullptr;
  const int d = current_device();
  if (!tried[d][i]) {
    tried[d][i] = true;
    if (cudaStreamCreateWithFlags(&s[d][i], cudaStreamNonBlocking) != cudaSuccess) s[d][i] = nullptr;
  }
  return s[d][i];
}
cudaEvent_t side_event(int i) {
  static cudaEvent_t ev[kMaxDevices][kSideEvents] = {};
  if (i < 0 || i >= kSideEvents) return nullptr;
  const int d = current_device();
  if (!ev[d][i] && cudaEventCreateWithFlags(&ev[d][i], cudaEventDisableTiming) != cudaSuccess) ev[d][i] = nullptr;
  return ev[d][i];
}

namespace {
constexpr int kMaxMarks = 512;
cudaEvent_t g_marks[kMaxMarks];
const char* g_mark_names[kMaxMarks];
int g_n_marks = 0;
}  // namespace
void profile_mark(cudaStream_t s, const char* name) {
  if (g_n_marks >= kMaxMarks) return;
  if (!g_marks[g_n_marks]) cudaEventCreate(&g_marks[g_n_marks]);
  cudaEventRecord(g_marks[g_n_marks], s);
  g_mark_names[g_n_marks] = name;
  ++g_n_marks;
}
}  // namespace isdqn

using namespace isdqn;

extern "C" int isdqn_profile_begin(void) {
  g_n_marks = 0;
  g_profile_on = true;
  return ISDQN_OK;
}

// Closes the capture with a final mark on `stream`, synchronises, and writes up to `max_entries` (name, ms)
// pairs: entry i is the time between mark i and mark i+1.  Returns the number of entries (or a negative error).
extern "C" int isdqn_profile_end(void* stream, int32_t max_entries, char* names, int32_t name_stride, float* ms) {
  if (!g_profile_on) return ISDQN_E_INVALID;
  profile_mark(as_stream(stream), "end");
  g_profile_on = false;
  if (g_n_marks < 1) return 0;
  ISDQN_CUDA_CHECK(cudaEventSynchronize(g_marks[g_n_marks - 1]));
  int n = g_n_marks - 1;
  if (n > max_entries) n = max_entries;
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    ISDQN_CUDA_CHECK(cudaEventElapsedTime(&t, g_marks[i], g_marks[i + 1]));
    if (ms) ms[i] = t;
    if (names && name_stride > 0) {
      strncpy(names + (size_t)i * name_stride, g_mark_names[i], name_stride - 1);
      names[(size_t)i * name_stride + name_stride - 1] = 0;
    }
  }
  return n;
}

// Keeps the GPU busy for ~`micros` so that the host can queue the launches that follow without gaps.
namespace isdqn {
__global__ void spin_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}
}  // namespace isdqn
extern "C" int isdqn_spin(void* stream, int32_t micros) {
  if (micros < 0 || micros > 100000) return ISDQN_E_INVALID;
  spin_kernel<<<1, 1, 0, as_stream(stream)>>>((long long)micros * 1900);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

namespace isdqn {
int num_sms() {
  static PerDevice<int> sms;
  int& n = sms.get();
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n > 0 ? n : kNumSMs;
}
}  // namespace isdqn

// ---------------------------------------------------------------------------------------------- head draw (host)
// iSDQN.best_action draws its online head with `jax.random.randint(key, (), 0, K)` (slimdqn/networks/isdqn.py:129).  JAX is
// not in this image; the draw is restated from jax 0.4.30 (jax/_src/prng.py threefry_2x32 / threefry_split /
// threefry_random_bits with jax_threefry_partitionable = False, jax/_src/random.py _randint), so that the SAME JAX key picks
// the SAME head.  The block function is pinned by the Random123 known-answer vectors (tests/test_threefry.py); the
// composition around it is a restatement (unpinned).
namespace {
inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  static const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int g = 0; g < 5; ++g) {
    const int* R = (g & 1) ? R1 : R0;
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl32(x1, R[i]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
}
}  // namespace

extern "C" void isdqn_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out2) {
  threefry2x32(k0, k1, x0, x1);
  out2[0] = x0;
  out2[1] = x1;
}

// jax.random.randint(key, (), minval, maxval) for an int32 result and a raw threefry key (k0, k1); maxval > minval.
extern "C" int32_t isdqn_threefry_randint(uint32_t k0, uint32_t k1, int32_t minval, int32_t maxval) {
  if (maxval <= minval) return minval;
  // k1, k2 = split(key): counts [0, 1, 2, 3] -> lanes (0, 2) and (1, 3); outputs concatenated [y0(0), y0(1), y1(0), y1(1)]
  uint32_t a0 = 0, a1 = 2, b0 = 1, b1 = 3;
  threefry2x32(k0, k1, a0, a1);
  threefry2x32(k0, k1, b0, b1);
  const uint32_t ka[2] = {a0, b0}, kb[2] = {a1, b1};
  // random_bits(key, 32, ()): one count, padded to the pair (0, 0); the first output word
  uint32_t h0 = 0, h1 = 0, l0 = 0, l1 = 0;
  threefry2x32(ka[0], ka[1], h0, h1);
  threefry2x32(kb[0], kb[1], l0, l1);
  const uint32_t span = (uint32_t)((int64_t)maxval - (int64_t)minval);
  uint32_t mult = 65536u % span;  // (2^16 mod span)^2 mod span == 2^32 mod span, in uint32 arithmetic as _randint does
  mult = (uint32_t)(mult * mult) % span;
  const uint32_t off = ((h0 % span) * mult + (l0 % span)) % span;
  return (int32_t)((int64_t)minval + (int64_t)off);
}

// jax.random.split(key, num): out_keys[2 * num] (raw key words).  threefry_2x32(key, iota(2 * num)): the counts are split in
// two halves that form the two lanes, the outputs of the lanes are concatenated and reshaped (num, 2).
extern "C" void isdqn_threefry_split(uint32_t k0, uint32_t k1, int32_t num, uint32_t* out_keys) {
  if (num < 1 || !out_keys) return;
  const int half = num;  // 2 * num counts -> lanes x0 = [0, num), x1 = [num, 2 num)
  for (int i = 0; i < half; ++i) {
    uint32_t a = (uint32_t)i, b = (uint32_t)(half + i);
    threefry2x32(k0, k1, a, b);
    out_keys[i] = a;          // first half of the flat output
    out_keys[half + i] = b;   // second half
  }
}

// jax.random.uniform(key) (float32 in [0, 1)): 32 random bits -> mantissa of a float in [1, 2) minus 1
extern "C" float isdqn_threefry_uniform(uint32_t k0, uint32_t k1) {
  uint32_t x0 = 0, x1 = 0;
  threefry2x32(k0, k1, x0, x1);
  const uint32_t bits = (x0 >> 9) | 0x3F800000u;
  float f;
  memcpy(&f, &bits, sizeof(f));
  return f - 1.0f;
}

extern "C" int isdqn_abi_version(void) { return ISDQN_ABI_VERSION; }

extern "C" const char* isdqn_strerror(int code) {
  switch (code) {
    case ISDQN_OK: return "ok";
    case ISDQN_E_INVALID: return "invalid argument";
    case ISDQN_E_TOO_LARGE: return "request exceeds a kernel limit";
    case ISDQN_E_CUDA: return "CUDA runtime error";
    case ISDQN_E_UNSUPPORTED: return "unsupported configuration";
    case ISDQN_E_NCCL: return "NCCL error";
    default: return "unknown error";
  }
}

extern "C" const char* isdqn_last_cuda_error(void) { return g_last_error; }

// ------------------------------------------------------------------------------------------------- graphs
extern "C" int isdqn_graph_begin(void* stream) {
  ISDQN_CUDA_CHECK(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
  return ISDQN_OK;
}

extern "C" int isdqn_graph_end(void* stream, void** out_graph_exec) {
  if (!out_graph_exec) return ISDQN_E_INVALID;
  cudaGraph_t graph = nullptr;
  ISDQN_CUDA_CHECK(cudaStreamEndCapture(as_stream(stream), &graph));
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "cudaGraphInstantiate");
    return ISDQN_E_CUDA;
  }
  *out_graph_exec = exec;
  return ISDQN_OK;
}

extern "C" int isdqn_graph_launch(void* graph_exec, void* stream) {
  if (!graph_exec) return ISDQN_E_INVALID;
  ISDQN_CUDA_CHECK(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)));
  return ISDQN_OK;
}

extern "C" int isdqn_graph_destroy(void* graph_exec) {
  if (!graph_exec) return ISDQN_OK;
  ISDQN_CUDA_CHECK(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
  return ISDQN_OK;
}

// --------------------------------------------------------------------------------------------------- NCCL
// Minimal, version-stable slice of nccl.h (ncclUniqueId is 128 bytes; enums as in NCCL 2.x).
namespace {
struct NcclUid {
  char b[128];
};
struct NcclApi {
  typedef NcclUid Uid;
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid /* ncclUniqueId by value */, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.handle) return ISDQN_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    snprintf(g_last_error, sizeof(g_last_error), "dlopen(libnccl.so.2) failed: %s", dlerror());
    return ISDQN_E_NCCL;
  }
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(dlsym(h, "ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    snprintf(g_last_error, sizeof(g_last_error), "libnccl is missing a required symbol");
    return ISDQN_E_NCCL;
  }
  g_nccl.handle = h;
  return ISDQN_OK;
}

int nccl_check(int rc, const char* where) {
  if (rc == 0) return ISDQN_OK;
  snprintf(g_last_error, sizeof(g_last_error), "%s: NCCL error %d (%s)", where, rc,
           g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return ISDQN_E_NCCL;
}
}  // namespace

extern "C" int isdqn_dp_unique_id(uint8_t* h_unique_id_128) {
  if (!h_unique_id_128) return ISDQN_E_INVALID;
  int rc = nccl_load();
  if (rc) return rc;
  return nccl_check(g_nccl.GetUniqueId(h_unique_id_128), "ncclGetUniqueId");
}

extern "C" int isdqn_dp_init(const uint8_t* h_unique_id_128, int32_t rank, int32_t world, void** out_comm) {
  if (!h_unique_id_128 || !out_comm || world < 1 || rank < 0 || rank >= world) return ISDQN_E_INVALID;
  int rc = nccl_load();
  if (rc) return rc;
  NcclApi::Uid uid;
  memcpy(uid.b, h_unique_id_128, 128);
  void* comm = nullptr;
  rc = nccl_check(g_nccl.CommInitRank(&comm, world, uid, rank), "ncclCommInitRank");
  if (rc) return rc;
  *out_comm = comm;
  return ISDQN_OK;
}

extern "C" int isdqn_dp_allreduce_f32(void* comm, float* d_buf, int64_t n, void* stream) {
  if (!comm || !d_buf || n < 0) return ISDQN_E_INVALID;
  if (!g_nccl.handle) return ISDQN_E_NCCL;
  // ncclFloat32 = 7, ncclSum = 0
  return nccl_check(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, 7, 0, comm, as_stream(stream)), "ncclAllReduce");
}

// The gradient minus one range [skip_off, skip_off + skip_n) that was all-reduced earlier (the hidden Dense kernel, started on
// the communication stream while the convolution backward was still running): the two remaining ranges as one NCCL group.
int isdqn_dp_allreduce_rest(void* comm, float* d_buf, int64_t n, int64_t skip_off, int64_t skip_n, void* stream) {
  if (!comm || !d_buf || n < 0 || skip_off < 0 || skip_n < 0 || skip_off + skip_n > n) return ISDQN_E_INVALID;
  if (!g_nccl.handle) return ISDQN_E_NCCL;
  const int64_t tail = n - skip_off - skip_n;
  const bool group = g_nccl.GroupStart && g_nccl.GroupEnd && skip_off > 0 && tail > 0;
  int rc = ISDQN_OK;
  if (group) rc = nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
  if (!rc && skip_off > 0)
    rc = nccl_check(g_nccl.AllReduce(d_buf, d_buf, (size_t)skip_off, 7, 0, comm, as_stream(stream)), "ncclAllReduce");
  if (!rc && tail > 0)
    rc = nccl_check(g_nccl.AllReduce(d_buf + skip_off + skip_n, d_buf + skip_off + skip_n, (size_t)tail, 7, 0, comm, as_stream(stream)),
                    "ncclAllReduce");
  if (group) {
    const int rc2 = nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
    if (!rc) rc = rc2;
  }
  return rc;
}

extern "C" int isdqn_dp_destroy(void* comm) {
  if (!comm) return ISDQN_OK;
  if (!g_nccl.handle) return ISDQN_E_NCCL;
  return nccl_check(g_nccl.CommDestroy(comm), "ncclCommDestroy");
}

// Device-side kernel-start timeline (common.cuh: trace_kernel_start).  d_buf: uint64[4001] zero-initialised device memory
// (NULL switches it off).  Diagnostic only.
int isdqn_trace_set_learner(unsigned long long* buf);
int isdqn_trace_set_tc(unsigned long long* buf);
int isdqn_trace_set_acting(unsigned long long* buf);
int isdqn_trace_set_adam_stream(unsigned long long* buf);
extern "C" int isdqn_trace_set(void* d_buf) {
  int rc = isdqn_trace_set_learner(reinterpret_cast<unsigned long long*>(d_buf));
  if (rc) return rc;
  rc = isdqn_trace_set_acting(reinterpret_cast<unsigned long long*>(d_buf));
  if (rc) return rc;
  rc = isdqn_trace_set_adam_stream(reinterpret_cast<unsigned long long*>(d_buf));
  if (rc) return rc;
  return isdqn_trace_set_tc(reinterpret_cast<unsigned long long*>(d_buf));
}

// ---------------------------------------------------------------------------------------------- host-batch staging
// Plumbing for the reference-facing call with HOST batches (iSDQN.learn_on_batch): events and the two asynchronous copy
// sequences, so that the Python side issues one call per step instead of a dozen framework calls.
extern "C" int isdqn_event_create(void** out_event) {
  if (!out_event) return ISDQN_E_INVALID;
  cudaEvent_t e = nullptr;
  ISDQN_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out_event = e;
  return ISDQN_OK;
}
extern "C" int isdqn_event_destroy(void* event) {
  if (event) ISDQN_CUDA_CHECK(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
  return ISDQN_OK;
}
extern "C" int isdqn_event_record(void* event, void* stream) {
  if (!event) return ISDQN_E_INVALID;
  ISDQN_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), as_stream(stream)));
  return ISDQN_OK;
}

extern "C" int isdqn_event_synchronize(void* event) {
  if (!event) return ISDQN_E_INVALID;
  ISDQN_CUDA_CHECK(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
  return ISDQN_OK;
}

// h_src (pinned) -> d_stage on copy_stream, then d_stage -> d_dst on step_stream:
//   copy_stream waits ev_stage_free (the step-stream copy that last read d_stage), copies, records ev_h2d_done;
//   step_stream waits ev_h2d_done, copies d_stage -> d_dst, re-records ev_stage_free.
extern "C" int isdqn_stage_batch(const void* h_src, void* d_stage, void* d_dst, int64_t bytes, void* copy_stream,
                                 void* step_stream, void* ev_h2d_done, void* ev_stage_free) {
  if (!h_src || !d_stage || bytes < 1 || !ev_h2d_done || !ev_stage_free) return ISDQN_E_INVALID;
  cudaStream_t cs = as_stream(copy_stream), ss = as_stream(step_stream);
  cudaEvent_t e_h2d = reinterpret_cast<cudaEvent_t>(ev_h2d_done), e_free = reinterpret_cast<cudaEvent_t>(ev_stage_free);
  ISDQN_CUDA_CHECK(cudaStreamWaitEvent(cs, e_free, 0));
  ISDQN_CUDA_CHECK(cudaMemcpyAsync(d_stage, h_src, (size_t)bytes, cudaMemcpyHostToDevice, cs));
  ISDQN_CUDA_CHECK(cudaEventRecord(e_h2d, cs));
  ISDQN_CUDA_CHECK(cudaStreamWaitEvent(ss, e_h2d, 0));
  if (d_dst) {
    ISDQN_CUDA_CHECK(cudaMemcpyAsync(d_dst, d_stage, (size_t)bytes, cudaMemcpyDeviceToDevice, ss));
    ISDQN_CUDA_CHECK(cudaEventRecord(e_free, ss));
  }
  // d_dst == NULL: the step reads d_stage in place; the caller records ev_stage_free (isdqn_event_record) on the step
  // stream once the work that reads the slot has been enqueued
  return ISDQN_OK;
}

extern "C" int isdqn_write_async(void* d_dst, const void* h_src, int64_t bytes, void* stream) {
  if (!d_dst || !h_src || bytes < 1) return ISDQN_E_INVALID;
  ISDQN_CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return ISDQN_OK;
}

// d_src -> h_dst (pinned) on `stream`, then record `event` (isdqn_event_synchronize(event) makes h_dst readable)
extern "C" int isdqn_read_async(void* h_dst, const void* d_src, int64_t bytes, void* stream, void* event) {
  if (!h_dst || !d_src || bytes < 1 || !event) return ISDQN_E_INVALID;
  ISDQN_CUDA_CHECK(cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  ISDQN_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), as_stream(stream)));
  return ISDQN_OK;
}
