// Replay samplers on the device, bit-exact with numpy's Generator(PCG64) as used by
// slimdqn/sample_collection/samplers.py:
//   uniform      samplers.py:39-49    Generator.integers(n, size)  -> buffered 32-bit stream + Lemire rejection
//   prioritized  samplers.py:105-116  Generator.uniform(0, root, size) + SumTree.query descent
// Every thread jumps the 128-bit LCG straight to its own stream position (O(log) multiplies), so the draws are
// produced in parallel yet are exactly the sequential ones; rejections are resolved with a block scan.
#include "common.cuh"
#include "pcg64.cuh"

namespace isdqn {

// THREADS candidates are examined per round; a batch of 32 uses the 128-thread instantiation (one short scan)
template <int kUniformThreads>
__global__ void __launch_bounds__(kUniformThreads)
sample_uniform_kernel(uint64_t* rng, uint32_t n_valid, int size, const int32_t* __restrict__ index_to_key,
                      int capacity, int32_t* __restrict__ out_index, int32_t* __restrict__ out_key,
                      int32_t* __restrict__ out_slot, const int* only_if = nullptr, const int32_t* n_valid_dev = nullptr) {
  if (only_if && !*only_if) return;  // fallback launch of the many-CTA path: normally nothing to do
  if (n_valid_dev) n_valid = (uint32_t)*n_valid_dev;  // captured in a CUDA graph: the live count is read at replay time
  if (n_valid == 0u) return;
  __shared__ int warp_tot[32];
  __shared__ int total_sh;
  __shared__ unsigned long long consumed_final;
  const int tid = threadIdx.x;
  const PcgMirror m0 = pcg_load(rng);

  auto emit = [&](int pos, int32_t index) {
    if (out_index) out_index[pos] = index;
    if (index_to_key) {
      const int32_t key = index_to_key[index];
      if (out_key) out_key[pos] = key;
      if (out_slot) out_slot[pos] = key % capacity;
    }
  };

  if (n_valid == 1u) {  // rng == 0: numpy fills with `off` and consumes no randomness
    for (int i = tid; i < size; i += kUniformThreads) emit(i, 0);
    return;
  }
  // Lemire: accept iff low32(v * n) >= (2^32 - n) % n   (the reference only evaluates the threshold when
  // low32 < n, which is the same predicate because threshold < n)
  const uint32_t threshold = (uint32_t)((0x100000000ull - n_valid) % n_valid);
  int written = 0;
  unsigned long long consumed = 0;  // 32-bit values consumed so far
  while (written < size) {
    const unsigned long long c = consumed + tid;
    const uint32_t v = pcg_next32_at(m0, c);
    const unsigned long long prod = (unsigned long long)v * n_valid;
    const int accept = ((uint32_t)prod >= threshold) ? 1 : 0;
    const int rank = block_exclusive_scan<kUniformThreads>(accept, warp_tot, &total_sh);
    const int total = total_sh;
    const int remaining = size - written;
    if (accept && rank < remaining) {
      emit(written + rank, (int32_t)(prod >> 32));
      if (rank == remaining - 1) consumed_final = c + 1;
    }
    __syncthreads();
    if (total >= remaining) {
      consumed = consumed_final;
      written = size;
    } else {
      consumed += kUniformThreads;
      written += total;
    }
    __syncthreads();
  }
  if (tid == 0) pcg_store(rng, pcg_after_next32(m0, consumed));
}

// ---- many-CTA variant for large draws (the throughput shape: 65,536 per launch).  The accept/reject test of draw
// position p depends only on p, and a draw lands at output index p - (#rejected positions before p).  Pass 1 counts the
// rejections of every CTA's slice of positions, pass 2 re-evaluates the slice and emits; rejections are rare
// (probability n_valid / 2^32 each), so the slices cover size + kUniformMargin positions and the single-CTA kernel is the
// fallback for the astronomically unlikely case that this is not enough (flag in the workspace).
constexpr int kUniformMargin = 2048;
constexpr int kUniformWide = 1024;  // threads per CTA
struct UniformWs {                  // device workspace (isdqn_sample_uniform_workspace_bytes)
  uint64_t new_rng[6];
  int fallback;                     // 1: pass 2 gave up, the single-CTA kernel must run
  int pad;
  int rej[1024];                    // rejections per CTA slice
};

__global__ void __launch_bounds__(kUniformWide)
uniform_count_kernel(const uint64_t* __restrict__ rng, uint32_t n_valid, int slice, UniformWs* ws) {
  __shared__ int warp_tot[32];
  __shared__ int total_sh;
  const PcgMirror m0 = pcg_load(rng);
  const uint32_t threshold = (uint32_t)((0x100000000ull - n_valid) % n_valid);
  int rejected = 0;
  const unsigned long long p0 = (unsigned long long)blockIdx.x * slice;
  for (int r = threadIdx.x; r < slice; r += kUniformWide) {
    const uint32_t v = pcg_next32_at(m0, p0 + r);
    rejected += ((uint32_t)((unsigned long long)v * n_valid) >= threshold) ? 0 : 1;
  }
  block_exclusive_scan<kUniformWide>(rejected, warp_tot, &total_sh);
  if (threadIdx.x == 0) {
    ws->rej[blockIdx.x] = total_sh;
    if (blockIdx.x == 0) ws->fallback = 0;
  }
}

__global__ void __launch_bounds__(kUniformWide)
uniform_emit_kernel(const uint64_t* __restrict__ rng, uint32_t n_valid, int size, int slice,
                    const int32_t* __restrict__ index_to_key, int capacity, int32_t* __restrict__ out_index,
                    int32_t* __restrict__ out_key, int32_t* __restrict__ out_slot, UniformWs* ws) {
  __shared__ int warp_tot[32];
  __shared__ int total_sh;
  const int tid = threadIdx.x;
  const PcgMirror m0 = pcg_load(rng);
  const uint32_t threshold = (uint32_t)((0x100000000ull - n_valid) % n_valid);
  long long rej_before = 0, rej_all = 0;
  for (int c = 0; c < (int)gridDim.x; ++c) {
    const int r = ws->rej[c];
    if (c < (int)blockIdx.x) rej_before += r;
    rej_all += r;
  }
  if ((long long)gridDim.x * slice - rej_all < size) {  // not enough accepted draws in the covered positions
    if (blockIdx.x == 0 && tid == 0) ws->fallback = 1;
    return;
  }
  const unsigned long long p0 = (unsigned long long)blockIdx.x * slice;
  long long out_base = (long long)p0 - rej_before;  // output index of this CTA's first accepted draw
  for (int r0 = 0; r0 < slice && out_base < size; r0 += kUniformWide) {
    const unsigned long long p = p0 + r0 + tid;
    const uint32_t v = pcg_next32_at(m0, p);
    const unsigned long long prod = (unsigned long long)v * n_valid;
    const int accept = ((uint32_t)prod >= threshold) ? 1 : 0;
    const int rank = block_exclusive_scan<kUniformWide>(accept, warp_tot, &total_sh);
    const long long o = out_base + rank;
    if (accept && o < size) {
      const int32_t index = (int32_t)(prod >> 32);
      if (out_index) out_index[o] = index;
      if (index_to_key) {
        const int32_t key = index_to_key[index];
        if (out_key) out_key[o] = key;
        if (out_slot) out_slot[o] = key % capacity;
      }
      if (o == size - 1) pcg_store(ws->new_rng, pcg_after_next32(m0, p + 1));  // exactly the values numpy consumed
    }
    out_base += total_sh;
    __syncthreads();
  }
}

// runs after pass 2: installs the new generator state, or (fallback) hands over to the single-CTA kernel's logic
__global__ void uniform_finish_kernel(uint64_t* rng, const UniformWs* ws) {
  if (ws->fallback) return;
  if (threadIdx.x < 6) rng[threadIdx.x] = ws->new_rng[threadIdx.x];
}

constexpr int kPrioThreads = 128;

__global__ void __launch_bounds__(kPrioThreads)
sample_prioritized_kernel(const uint64_t* __restrict__ rng, const double* __restrict__ nodes, int depth, int size,
                          int n_valid, const int32_t* __restrict__ index_to_key, int capacity,
                          int32_t* __restrict__ out_index, int32_t* __restrict__ out_key, int32_t* __restrict__ out_slot,
                          double* __restrict__ out_target, double* __restrict__ out_prob, uint32_t* status) {
  const PcgMirror m0 = pcg_load(rng);
  const double root = nodes[0];
  const int first_leaf = (1 << (depth - 1)) - 1;
  uint32_t st = 0;
  if (root == 0.0) st |= ISDQN_ST_EMPTY_TREE;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += gridDim.x * blockDim.x) {
    // Generator.uniform: low + (high - low) * ((next64 >> 11) * 2^-53), one next64 per element
    const uint64_t o = pcg_out64_at(m0.state, m0.inc, (uint64_t)i);
    const double u = (double)(o >> 11) * (1.0 / 9007199254740992.0);
    double t = __dadd_rn(0.0, __dmul_rn(root - 0.0, u));
    if (out_target) out_target[i] = t;
    if (!(t >= 0.0 && t < root)) st |= ISDQN_ST_TARGET_RANGE;
    int node = 0;
    for (int level = 0; level < depth - 1; ++level) {
      const int left = 2 * node + 1;
      const double ls = __ldg(nodes + left);
      if (t < ls) {
        node = left;
      } else {
        node = left + 1;
        t = t - ls;
        if (level + 1 < depth - 1 && !(t < __ldg(nodes + left + 1))) st |= ISDQN_ST_DESCENT_ASSERT;
      }
    }
    const int32_t index = node - first_leaf;
    if (out_index) out_index[i] = index;
    // probability of the drawn leaf (importance weights of a prioritized training loop): leaf / root
    if (out_prob) out_prob[i] = root > 0.0 ? __ldg(nodes + node) / root : 0.0;
    if (index_to_key) {
      // samplers.py:112 `self._index_to_key[index]` raises IndexError for an index past the live keys (an empty tree or a
      // rounding that lands on a zero leaf): flag it and look a live entry up instead, so that nothing downstream (key %
      // capacity, the gather) ever sees a slot outside the tables
      int32_t safe = index;
      if (n_valid >= 0 && index >= n_valid) {
        st |= ISDQN_ST_INDEX_RANGE;
        safe = n_valid > 0 ? n_valid - 1 : 0;
      }
      const int32_t key = n_valid == 0 ? 0 : index_to_key[safe];
      if (out_key) out_key[i] = key;
      if (out_slot) out_slot[i] = (int32_t)(((int64_t)key % capacity + capacity) % capacity);
    }
  }
  if (st && status) atomicOr(status, st);
}

// One training batch (size <= 1024, one CTA, thread i = draw i) for a captured step: the same draw as
// sample_prioritized_kernel with the live count read from the device, plus the importance weights of the batch
// (n_valid * P(i)) ** -beta / max_i(.) as float32 (Schaul et al. 2016) and the generator advance, in one launch.
__global__ void __launch_bounds__(1024)
sample_prioritized_train_kernel(uint64_t* __restrict__ rng, const double* __restrict__ nodes, int depth, int size,
                                const int32_t* __restrict__ n_valid_dev, const int32_t* __restrict__ index_to_key, int capacity,
                                int32_t* __restrict__ out_key, int32_t* __restrict__ out_slot, double* __restrict__ out_prob,
                                const double* __restrict__ beta_dev, float* __restrict__ out_weight, uint32_t* status) {
  __shared__ double red[32];
  // the top kTopLevels levels of the heap, loaded in one cooperative round trip: a draw is a chain of (depth - 1) dependent
  // 8-byte reads, and with one CTA there is nothing to hide their latency behind — 11 of the 20 come out of shared memory
  constexpr int kTopLevels = 11;
  __shared__ double top[(1 << kTopLevels) - 1];
  const int cached_levels = depth - 1 < kTopLevels ? depth - 1 : kTopLevels;
  const int n_top = (1 << cached_levels) - 1;
  for (int j = threadIdx.x; j < n_top; j += blockDim.x) top[j] = __ldg(nodes + j);
  const double beta = beta_dev ? *beta_dev : 0.0;
  const PcgMirror m0 = pcg_load(rng);
  const int n_valid = *n_valid_dev;
  __syncthreads();
  const double root = n_top > 0 ? top[0] : nodes[0];
  const int first_leaf = (1 << (depth - 1)) - 1;
  const int i = threadIdx.x;
  uint32_t st = 0;
  double w = 0.0;
  if (i == 0 && root == 0.0) st |= ISDQN_ST_EMPTY_TREE;
  if (i < size) {
    const uint64_t o = pcg_out64_at(m0.state, m0.inc, (uint64_t)i);
    const double u = (double)(o >> 11) * (1.0 / 9007199254740992.0);
    double t = __dadd_rn(0.0, __dmul_rn(root - 0.0, u));
    if (!(t >= 0.0 && t < root)) st |= ISDQN_ST_TARGET_RANGE;
    int node = 0;
    for (int level = 0; level < depth - 1; ++level) {
      const int left = 2 * node + 1;
      const bool in_top = left + 1 < n_top;
      const double ls = in_top ? top[left] : __ldg(nodes + left);
      if (t < ls) {
        node = left;
      } else {
        node = left + 1;
        t = t - ls;
        if (level + 1 < depth - 1 && !(t < (in_top ? top[left + 1] : __ldg(nodes + left + 1)))) st |= ISDQN_ST_DESCENT_ASSERT;
      }
    }
    const int32_t index = node - first_leaf;
    const double prob = root > 0.0 ? __ldg(nodes + node) / root : 0.0;
    if (out_prob) out_prob[i] = prob;
    int32_t safe = index;
    if (index >= n_valid) {
      st |= ISDQN_ST_INDEX_RANGE;
      safe = n_valid > 0 ? n_valid - 1 : 0;
    }
    const int32_t key = n_valid == 0 ? 0 : index_to_key[safe];
    if (out_key) out_key[i] = key;
    if (out_slot) out_slot[i] = (int32_t)(((int64_t)key % capacity + capacity) % capacity);
    // (a zero-probability draw was flagged above — the reference raises there; its weight is 0, not inf)
    w = (out_weight && prob > 0.0) ? pow(prob * (double)n_valid, -beta) : 0.0;
  }
  if (out_weight) {  // block maximum
    double mx = w;
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
      mx = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (threadIdx.x == 0) red[0] = mx;
    }
    __syncthreads();
    if (i < size) out_weight[i] = red[0] > 0.0 ? (float)(w / red[0]) : 0.f;
  }
  if (st && status) atomicOr(status, st);
  __syncthreads();  // every thread has read the generator state
  if (threadIdx.x == 0) {
    PcgMirror m = m0;
    m.state = pcg_advance(m.state, m.inc, (uint64_t)size);
    pcg_store(rng, m);
  }
}

__global__ void pcg_advance64_kernel(uint64_t* rng, uint64_t steps) {
  PcgMirror m = pcg_load(rng);
  m.state = pcg_advance(m.state, m.inc, steps);
  pcg_store(rng, m);
}

__global__ void scatter_rows_i32_kernel(int32_t* __restrict__ table, int width, const int32_t* __restrict__ pidx,
                                        const int32_t* __restrict__ pval, int n) {
  // rows in one patch list are distinct (the host keeps only the last write per row)
  const int total = n * width;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int i = t / width, j = t - i * width;
    table[(int64_t)pidx[i] * width + j] = pval[t];
  }
}

__global__ void scatter_rows_f64_kernel(double* __restrict__ table, const int32_t* __restrict__ pidx,
                                        const double* __restrict__ pval, int n) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) table[pidx[t]] = pval[t];
}

}  // namespace isdqn

using namespace isdqn;

extern "C" int isdqn_sample_uniform(uint64_t* d_rng, int32_t n_valid, int32_t size, const int32_t* d_index_to_key,
                                    int32_t capacity, int32_t* d_out_index, int32_t* d_out_key, int32_t* d_out_slot,
                                    void* stream) {
  if (!d_rng || n_valid < 1 || size < 0 || (d_index_to_key && capacity < 1)) return ISDQN_E_INVALID;
  if (size > (1 << 20)) return ISDQN_E_TOO_LARGE;
  if (size == 0) return ISDQN_OK;
  ISDQN_PROF(as_stream(stream), "sample_uniform");
  if (size <= 96)
    sample_uniform_kernel<128><<<1, 128, 0, as_stream(stream)>>>(d_rng, (uint32_t)n_valid, size, d_index_to_key, capacity,
                                                                 d_out_index, d_out_key, d_out_slot);
  else
    sample_uniform_kernel<1024><<<1, 1024, 0, as_stream(stream)>>>(d_rng, (uint32_t)n_valid, size, d_index_to_key, capacity,
                                                                   d_out_index, d_out_key, d_out_slot);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

// isdqn_sample_uniform with the number of live indices read from device memory when the kernel RUNS: the launch can be
// captured in a CUDA graph together with the gather and the learner step and replayed while the buffer fills.
extern "C" int isdqn_sample_uniform_dev(uint64_t* d_rng, const int32_t* d_n_valid, int32_t size, const int32_t* d_index_to_key,
                                        int32_t capacity, int32_t* d_out_index, int32_t* d_out_key, int32_t* d_out_slot,
                                        void* stream) {
  if (!d_rng || !d_n_valid || size < 0 || (d_index_to_key && capacity < 1)) return ISDQN_E_INVALID;
  if (size > 8192) return ISDQN_E_TOO_LARGE;
  if (size == 0) return ISDQN_OK;
  ISDQN_PROF(as_stream(stream), "sample_uniform");
  if (size <= 96)
    sample_uniform_kernel<128><<<1, 128, 0, as_stream(stream)>>>(d_rng, 1u, size, d_index_to_key, capacity, d_out_index, d_out_key,
                                                                 d_out_slot, nullptr, d_n_valid);
  else
    sample_uniform_kernel<1024><<<1, 1024, 0, as_stream(stream)>>>(d_rng, 1u, size, d_index_to_key, capacity, d_out_index,
                                                                   d_out_key, d_out_slot, nullptr, d_n_valid);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int64_t isdqn_sample_uniform_workspace_bytes(void) { return (int64_t)sizeof(UniformWs); }

extern "C" int isdqn_sample_uniform_ws(uint64_t* d_rng, int32_t n_valid, int32_t size, const int32_t* d_index_to_key,
                                       int32_t capacity, int32_t* d_out_index, int32_t* d_out_key, int32_t* d_out_slot,
                                       void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (size < 2048 || n_valid == 1 || !d_workspace || workspace_bytes < (int64_t)sizeof(UniformWs))
    return isdqn_sample_uniform(d_rng, n_valid, size, d_index_to_key, capacity, d_out_index, d_out_key, d_out_slot, stream);
  if (!d_rng || n_valid < 1 || (d_index_to_key && capacity < 1)) return ISDQN_E_INVALID;
  if (size > (1 << 20)) return ISDQN_E_TOO_LARGE;
  cudaStream_t s = as_stream(stream);
  UniformWs* ws = reinterpret_cast<UniformWs*>(d_workspace);
  int grid = ceil_div(size + kUniformMargin, kUniformWide);
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
  const int slice = ceil_div(ceil_div(size + kUniformMargin, grid), kUniformWide) * kUniformWide;
  ISDQN_PROF(s, "sample_uniform");
  uniform_count_kernel<<<grid, kUniformWide, 0, s>>>(d_rng, (uint32_t)n_valid, slice, ws);
  ISDQN_LAUNCH_CHECK();
  uniform_emit_kernel<<<grid, kUniformWide, 0, s>>>(d_rng, (uint32_t)n_valid, size, slice, d_index_to_key, capacity, d_out_index,
                                                    d_out_key, d_out_slot, ws);
  ISDQN_LAUNCH_CHECK();
  uniform_finish_kernel<<<1, 32, 0, s>>>(d_rng, ws);
  ISDQN_LAUNCH_CHECK();
  sample_uniform_kernel<1024><<<1, 1024, 0, s>>>(d_rng, (uint32_t)n_valid, size, d_index_to_key, capacity, d_out_index, d_out_key,
                                                 d_out_slot, &ws->fallback);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_sample_prioritized(uint64_t* d_rng, const double* d_nodes, int depth, int32_t size, int32_t n_valid,
                                        const int32_t* d_index_to_key, int32_t capacity, int32_t* d_out_index,
                                        int32_t* d_out_key, int32_t* d_out_slot, double* d_out_target,
                                        double* d_out_prob, uint32_t* d_status, void* stream) {
  if (!d_rng || !d_nodes || depth < 1 || depth > 31 || size < 0 || (d_index_to_key && capacity < 1))
    return ISDQN_E_INVALID;
  if (size == 0) return ISDQN_OK;
  int grid = ceil_div(size, kPrioThreads);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  ISDQN_PROF(as_stream(stream), "sample_prioritized");
  sample_prioritized_kernel<<<grid, kPrioThreads, 0, as_stream(stream)>>>(d_rng, d_nodes, depth, size, n_valid, d_index_to_key,
                                                                          capacity, d_out_index, d_out_key, d_out_slot,
                                                                          d_out_target, d_out_prob, d_status);
  ISDQN_LAUNCH_CHECK();
  ISDQN_PROF(as_stream(stream), "pcg_advance");
  pcg_advance64_kernel<<<1, 1, 0, as_stream(stream)>>>(d_rng, (uint64_t)size);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_sample_prioritized_train(uint64_t* d_rng, const double* d_nodes, int depth, int32_t size,
                                              const int32_t* d_n_valid, const int32_t* d_index_to_key, int32_t capacity,
                                              int32_t* d_out_key, int32_t* d_out_slot, double* d_out_prob,
                                              const double* d_beta, float* d_out_weight, uint32_t* d_status, void* stream) {
  if (!d_rng || !d_nodes || !d_n_valid || !d_index_to_key || depth < 1 || depth > 31 || size < 1 || capacity < 1 ||
      (d_out_weight && !d_beta))
    return ISDQN_E_INVALID;
  if (size > 1024) return ISDQN_E_TOO_LARGE;
  ISDQN_PROF(as_stream(stream), "sample_prioritized");
  const int threads = size <= 256 ? 256 : ceil_div(size, 32) * 32;  // (idle threads still help load the top of the heap)
  sample_prioritized_train_kernel<<<1, threads, 0, as_stream(stream)>>>(d_rng, d_nodes, depth, size, d_n_valid, d_index_to_key,
                                                                        capacity, d_out_key, d_out_slot, d_out_prob, d_beta,
                                                                        d_out_weight, d_status);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_scatter_rows_i32(int32_t* d_table, int32_t width, const int32_t* d_patch_index,
                                      const int32_t* d_patch_value, int32_t n_patches, void* stream) {
  if (!d_table || !d_patch_index || !d_patch_value || width < 1 || n_patches < 0) return ISDQN_E_INVALID;
  if (n_patches == 0) return ISDQN_OK;
  const int total = n_patches * width;
  int grid = ceil_div(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  scatter_rows_i32_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_table, width, d_patch_index, d_patch_value, n_patches);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_scatter_rows_f64(double* d_table, const int32_t* d_patch_index, const double* d_patch_value,
                                      int32_t n_patches, void* stream) {
  if (!d_table || !d_patch_index || !d_patch_value || n_patches < 0) return ISDQN_E_INVALID;
  if (n_patches == 0) return ISDQN_OK;
  int grid = ceil_div(n_patches, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  scatter_rows_f64_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_table, d_patch_index, d_patch_value, n_patches);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}
