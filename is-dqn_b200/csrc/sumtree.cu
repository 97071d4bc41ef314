// Sum-tree kernels (float64 array heap), bit-exact with slimdqn/sample_collection/sum_tree.py.
//
//   query  sum_tree.py:58-102   one thread walks QPT independent descents in lock step (QPT loads in flight per
//                               thread); the top levels of the heap are staged in shared memory when a CTA has
//                               enough queries to amortise the staging (throughput shapes).
//   set    sum_tree.py:20-47    one CTA: bitonic sort of (leaf, position) in shared memory, first-duplicate-wins
//                               compaction, delta = value - old leaf, then for every level the runs of equal
//                               ancestors are folded sequentially in ascending-leaf order — exactly the order of
//                               np.add.at — so the float64 heap stays bit-identical to the reference's.
#include <cstdlib>

#include "common.cuh"

namespace isdqn {

// ------------------------------------------------------------------------------------------------- query
constexpr int kQueryThreads = 256;
constexpr int kQueryPerThread = 4;
constexpr int kQueryMaxTopLevels = 12;  // 4095 nodes = 32 KB of shared memory

template <int QPT>
__global__ void __launch_bounds__(kQueryThreads)
sumtree_query_kernel(const double* __restrict__ nodes, int depth, const double* __restrict__ targets, int64_t n,
                     int32_t* __restrict__ out, uint32_t* status, int top_levels) {
  extern __shared__ double top[];
  const int n_top = top_levels > 0 ? (1 << top_levels) - 1 : 0;
  for (int i = threadIdx.x; i < n_top; i += blockDim.x) top[i] = nodes[i];
  if (n_top) __syncthreads();

  const int first_leaf = (1 << (depth - 1)) - 1;
  const double root = nodes[0];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  uint32_t st = 0;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * QPT) {
    double t[QPT];
    int node[QPT];
    bool live[QPT];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t q = base + (int64_t)j * stride;
      live[j] = q < n;
      t[j] = live[j] ? targets[q] : 0.0;
      node[j] = 0;
      if (live[j] && !(t[j] >= 0.0 && t[j] < root)) st |= ISDQN_ST_TARGET_RANGE;
    }
    for (int level = 0; level < depth - 1; ++level) {
      const bool child_in_top = (level + 1) < top_levels;
      const bool child_internal = (level + 1) < (depth - 1);
      double ls[QPT];
#pragma unroll
      for (int j = 0; j < QPT; ++j) {
        const int left = 2 * node[j] + 1;
        ls[j] = child_in_top ? top[left] : __ldg(nodes + left);
      }
#pragma unroll
      for (int j = 0; j < QPT; ++j) {
        const int left = 2 * node[j] + 1;
        if (t[j] < ls[j]) {
          node[j] = left;
        } else {
          node[j] = left + 1;
          t[j] = t[j] - ls[j];
          if (child_internal && live[j]) {
            // sum_tree.py:82: the next iteration asserts target < nodes[node] (delta-propagated sums can
            // violate it by one rounding); going left satisfies it by construction.
            const double rs = child_in_top ? top[left + 1] : __ldg(nodes + left + 1);
            if (!(t[j] < rs)) st |= ISDQN_ST_DESCENT_ASSERT;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t q = base + (int64_t)j * stride;
      if (live[j]) out[q] = node[j] - first_leaf;
    }
  }
  if (st && status) atomicOr(status, st);
}

// --------------------------------------------------------------------------------------------------- set
struct SetSmem {
  unsigned long long* keys;  // [MAXM]  (leaf << 32 | position), sorted ascending
  int* uleaf;                // [MAXM]  unique leaves, ascending
  double* udelta;            // [MAXM]
  int* warp_tot;             // [32]
  int* misc;                 // [4]: total, bad flag
  double* red;               // [32]
};

// With TAGS a value -(1+j) stands for "the current value of leaf j" (read before this op modifies anything):
// PrioritizedSamplingDistribution.remove moves the last leaf's priority into the hole (samplers.py:99-102)
// and would otherwise need a device->host read per eviction.  A value in (-1, 0) (ISDQN_SUMTREE_TAG_MAX = -0.5)
// stands for "max_recorded_priority as it is when the op starts": a prioritized training loop inserts new
// transitions at that priority without reading it back.
template <bool TAGS>
__device__ __forceinline__ double resolve_value(double v, const double* nodes, int first_leaf, int n_leaves, double max_old,
                                                int* bad) {
  if (TAGS && v < 0.0) {
    if (v > -1.0) return max_old;
    const int src = (int)(-v) - 1;
    if (src < 0 || src >= n_leaves) {
      *bad |= 2;
      return 0.0;
    }
    return nodes[first_leaf + src];
  }
  return v;
}

// One `set` executed by the whole CTA.  Ends with a __syncthreads(), so it can be called back to back.
template <int MAXM, int THREADS, bool TAGS>
__device__ void sumtree_set_block(double* nodes, int depth, const int32_t* __restrict__ idx,
                                  const double* __restrict__ val, int m, double* max_prio, uint32_t* status,
                                  const SetSmem& sm) {
  const int tid = threadIdx.x;
  const int n_leaves = 1 << (depth - 1);
  const int first_leaf = n_leaves - 1;

  // (0) sum_tree.py:31 — nothing is modified when any value is negative (or NaN); indices must be leaves.
  int bad = 0;
  double vmax = 0.0;
  // (max_prio is only written by thread 0 in step (1), after the barrier below: every thread reads the same old value)
  const double max_old = (TAGS && max_prio) ? *max_prio : 0.0;
  for (int i = tid; i < m; i += THREADS) {
    const double v = resolve_value<TAGS>(val[i], nodes, first_leaf, n_leaves, max_old, &bad);
    const int l = idx[i];
    if (!(v >= 0.0)) bad |= 1;
    if (l < 0 || l >= n_leaves) bad |= 2;
    vmax = fmax(vmax, v);
  }
  bad = __syncthreads_or(bad);
  if (bad) {
    // which of the two: recompute cheaply on one thread's worth of data is not possible; flag both classes
    int mine = 0;
    for (int i = tid; i < m; i += THREADS) {
      int b2 = 0;
      if (!(resolve_value<TAGS>(val[i], nodes, first_leaf, n_leaves, max_old, &b2) >= 0.0)) mine |= ISDQN_ST_NEGATIVE_VALUE;
      if (b2 || idx[i] < 0 || idx[i] >= n_leaves) mine |= ISDQN_ST_INDEX_RANGE;
    }
    if (mine && status) atomicOr(status, (uint32_t)mine);
    __syncthreads();
    return;
  }
  // (1) sum_tree.py:32 — max_recorded_priority
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if ((tid & 31) == 0) sm.red[tid >> 5] = vmax;
  __syncthreads();
  if (tid == 0) {
    double r = sm.red[0];
    for (int w = 1; w < THREADS / 32; ++w) r = fmax(r, sm.red[w]);
    if (max_prio) *max_prio = fmax(*max_prio, r);
  }

  // (2) sort (leaf, position) ascending: np.unique(node_indices, return_index=True) keeps the first position
  int P = 1;
  while (P < m) P <<= 1;
  for (int i = tid; i < P; i += THREADS)
    sm.keys[i] = i < m ? (((unsigned long long)(uint32_t)idx[i]) << 32) | (uint32_t)i : ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = sm.keys[i], b = sm.keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            sm.keys[i] = b;
            sm.keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }

  // (3) compaction of first occurrences + delta against the OLD leaf value (sum_tree.py:34,39-40)
  const int per = (P + THREADS - 1) / THREADS;
  const int lo = tid * per;
  int cnt = 0;
  for (int i = lo; i < lo + per && i < m; ++i) {
    const uint32_t leaf = (uint32_t)(sm.keys[i] >> 32);
    cnt += (i == 0 || leaf != (uint32_t)(sm.keys[i - 1] >> 32)) ? 1 : 0;
  }
  int pos = block_exclusive_scan<THREADS>(cnt, sm.warp_tot, &sm.misc[0]);
  const int u = sm.misc[0];
  for (int i = lo; i < lo + per && i < m; ++i) {
    const unsigned long long key = sm.keys[i];
    const uint32_t leaf = (uint32_t)(key >> 32);
    if (i == 0 || leaf != (uint32_t)(sm.keys[i - 1] >> 32)) {
      sm.uleaf[pos] = (int)leaf;
      int b2 = 0;
      sm.udelta[pos] = resolve_value<TAGS>(val[(uint32_t)key], nodes, first_leaf, n_leaves, max_old, &b2) - nodes[first_leaf + (int)leaf];
      ++pos;
    }
  }
  __syncthreads();

  // (4) sum_tree.py:41-47 — every level: np.add.at(nodes, ancestors, deltas) == sequential fold per run of
  // equal ancestors, ascending-leaf order.  Levels touch disjoint nodes, runs are disjoint: no races.
  const int tasks = u * depth;
  for (int task = tid; task < tasks; task += THREADS) {
    const int s = task / u;  // how many levels above the leaves
    const int i = task - s * u;
    const int node = ((first_leaf + sm.uleaf[i] + 1) >> s) - 1;
    if (i > 0 && (((first_leaf + sm.uleaf[i - 1] + 1) >> s) - 1) == node) continue;  // not a run head
    double acc = nodes[node];
    int j = i;
    do {
      acc = __dadd_rn(acc, sm.udelta[j]);
      ++j;
    } while (j < u && (((first_leaf + sm.uleaf[j] + 1) >> s) - 1) == node);
    nodes[node] = acc;
  }
  __syncthreads();
}

template <int MAXM, int THREADS>
__device__ __forceinline__ SetSmem carve_set_smem(unsigned char* raw) {
  SetSmem sm;
  sm.keys = reinterpret_cast<unsigned long long*>(raw);
  sm.udelta = reinterpret_cast<double*>(raw + sizeof(unsigned long long) * MAXM);
  sm.red = reinterpret_cast<double*>(raw + 16 * MAXM);
  sm.uleaf = reinterpret_cast<int*>(raw + 16 * MAXM + 32 * sizeof(double));
  sm.warp_tot = sm.uleaf + MAXM;
  sm.misc = sm.warp_tot + 32;
  return sm;
}
template <int MAXM>
constexpr size_t set_smem_bytes() {
  return 16 * (size_t)MAXM + 32 * sizeof(double) + sizeof(int) * ((size_t)MAXM + 32 + 4);
}

template <int MAXM, int THREADS>
__global__ void __launch_bounds__(THREADS)
sumtree_set_kernel(double* nodes, int depth, const int32_t* __restrict__ idx, const double* __restrict__ val, int m,
                   double* max_prio, uint32_t* status, const int32_t* abort_flag = nullptr) {
  extern __shared__ __align__(16) unsigned char set_raw[];
  if (abort_flag && *abort_flag) return;  // isdqn_sumtree_set_keys: a key was not live, nothing is modified
  SetSmem sm = carve_set_smem<MAXM, THREADS>(set_raw);
  sumtree_set_block<MAXM, THREADS, false>(nodes, depth, idx, val, m, max_prio, status, sm);
}

// One `set` of m <= 2 entries executed by ONE warp, no sort network and no block barrier: lane s owns level s above the
// leaves (depth <= 31 <= warp size).  These are the ops of the add / evict traffic (PrioritizedSamplingDistribution.add:
// m = 1, .remove: m = 2, samplers.py:67-103).  Same arithmetic as sumtree_set_block: first duplicate wins, deltas against
// the leaf values before the op, ascending-leaf fold order where the two chains share an ancestor.
__device__ __forceinline__ void sumtree_set_small(double* nodes, int depth, int m, int i0, double v0, int i1, double v1,
                                                  double* max_prio, uint32_t* status) {
  const int lane = threadIdx.x & 31;
  const int n_leaves = 1 << (depth - 1);
  const int first_leaf = n_leaves - 1;
  const double max_old = max_prio ? *max_prio : 0.0;
  int bad = 0;
  v0 = resolve_value<true>(v0, nodes, first_leaf, n_leaves, max_old, &bad);
  if (m == 2) v1 = resolve_value<true>(v1, nodes, first_leaf, n_leaves, max_old, &bad);
  int st = 0;
  if (bad || i0 < 0 || i0 >= n_leaves || (m == 2 && (i1 < 0 || i1 >= n_leaves))) st |= ISDQN_ST_INDEX_RANGE;
  if (!(v0 >= 0.0) || (m == 2 && !(v1 >= 0.0))) st |= ISDQN_ST_NEGATIVE_VALUE;
  if (st) {  // sum_tree.py:31 — nothing is modified
    if (lane == 0 && status) atomicOr(status, (uint32_t)st);
    return;
  }
  if (lane == 0 && max_prio) *max_prio = fmax(max_old, m == 2 ? fmax(v0, v1) : v0);
  if (m == 2 && i0 == i1) m = 1;  // np.unique(..., return_index=True): the first occurrence wins
  if (m == 2 && i1 < i0) {        // ascending leaves
    const int ti = i0; i0 = i1; i1 = ti;
    const double tv = v0; v0 = v1; v1 = tv;
  }
  // this lane's nodes (its loads are issued together with the leaf loads below: one round trip for both)
  const bool act = lane < depth;
  const int n0 = ((first_leaf + i0 + 1) >> lane) - 1;
  const int n1 = m == 2 ? ((first_leaf + i1 + 1) >> lane) - 1 : n0;
  const double c0 = act ? nodes[n0] : 0.0;
  const double c1 = (act && n1 != n0) ? nodes[n1] : 0.0;
  const double d0 = v0 - nodes[first_leaf + i0];
  const double d1 = m == 2 ? v1 - nodes[first_leaf + i1] : 0.0;
  if (act) {
    if (m == 1) {
      nodes[n0] = __dadd_rn(c0, d0);
    } else if (n0 != n1) {
      nodes[n0] = __dadd_rn(c0, d0);
      nodes[n1] = __dadd_rn(c1, d1);
    } else {
      nodes[n0] = __dadd_rn(__dadd_rn(c0, d0), d1);
    }
  }
  __syncwarp();  // the next op of this warp reads what other lanes wrote
}

// One `set` of m <= 32 entries executed by ONE warp with no shared memory and no block barrier (the batch-32 priority
// update of a training step, PrioritizedSamplingDistribution.update, samplers.py:76-88).  Lane i holds entry i.
//   sort (leaf, position) ascending with a shuffle bitonic network; the first of equal leaves wins (np.unique(...,
//   return_index=True)); delta = value - old leaf; then level by level every run of equal ancestors is folded
//   sequentially in ascending-leaf order by its first lane — np.add.at's order, so the float64 heap stays bit-identical.
// All node loads of a lane (its ancestor on every level) are issued before the first fold: one round trip, not depth.
template <bool TAGS>
__device__ __forceinline__ void sumtree_set_warp(double* nodes, int depth, int m, int leaf_in, double v_in, double* max_prio,
                                                 uint32_t* status) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int n_leaves = 1 << (depth - 1);
  const int first_leaf = n_leaves - 1;
  const bool have = lane < m;
  const double max_old = max_prio ? *max_prio : 0.0;
  int bad = 0;
  double v = have ? resolve_value<TAGS>(v_in, nodes, first_leaf, n_leaves, max_old, &bad) : 0.0;
  int st = 0;
  if (have && (bad || leaf_in < 0 || leaf_in >= n_leaves)) st |= ISDQN_ST_INDEX_RANGE;
  if (have && !(v >= 0.0)) st |= ISDQN_ST_NEGATIVE_VALUE;
  st = __reduce_or_sync(FULL, (unsigned)st);
  if (st) {  // sum_tree.py:31 — nothing is modified
    if (lane == 0 && status) atomicOr(status, (uint32_t)st);
    return;
  }
  {  // sum_tree.py:32
    double vm = have ? v : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vm = fmax(vm, __shfl_xor_sync(FULL, vm, o));
    if (lane == 0 && max_prio) *max_prio = fmax(max_old, vm);
  }
  // ---- sort by (leaf, position); absent lanes carry the largest key
  unsigned long long key = have ? (((unsigned long long)(uint32_t)leaf_in) << 32) | (uint32_t)lane : ~0ull;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(FULL, key, j);
      const double ov = __shfl_xor_sync(FULL, v, j);
      const bool up = (lane & k) == 0;        // ascending block
      const bool lower = (lane & j) == 0;     // this lane keeps the smaller key of the pair in an ascending block
      const bool take_min = up == lower;
      const bool swap = take_min ? (ok < key) : (ok > key);
      if (swap) {
        key = ok;
        v = ov;
      }
    }
  }
  const int leaf = (int)(key >> 32);
  const bool present = key != ~0ull;
  const int prev_leaf = __shfl_up_sync(FULL, leaf, 1);
  const bool prev_present = __shfl_up_sync(FULL, (int)present, 1) != 0;
  const bool active = present && (lane == 0 || !prev_present || prev_leaf != leaf);  // first occurrence of its leaf
  // ---- this lane's chain of nodes: loads first, then the folds
  double cur[31];
#pragma unroll
  for (int s = 0; s < 31; ++s) cur[s] = (active && s < depth) ? nodes[((first_leaf + leaf + 1) >> s) - 1] : 0.0;
  const double delta = active ? __dadd_rn(v, -cur[0]) : 0.0;  // against the OLD leaf value (sum_tree.py:34)
  const unsigned amask = __ballot_sync(FULL, active);
#pragma unroll 1
  for (int s = 0; s < depth; ++s) {
    const int node = active ? ((first_leaf + leaf + 1) >> s) - 1 : -1 - lane;
    // previous ACTIVE lane's node (ascending leaves => equal ancestors are contiguous among the active lanes)
    const unsigned below = amask & ((1u << lane) - 1u);
    const int prev_lane = below ? 31 - __clz(below) : lane;
    const int prev_node = __shfl_sync(FULL, node, prev_lane);
    const bool head = active && (!below || prev_node != node);
    double acc = cur[s < 31 ? s : 30];
    // every head folds the deltas of the active lanes j >= its own lane that share its node, in lane order
    for (int j = 0; j < 32; ++j) {
      const int nj = __shfl_sync(FULL, node, j);
      const double dj = __shfl_sync(FULL, delta, j);
      if (head && j >= lane && ((amask >> j) & 1u) && nj == node) acc = __dadd_rn(acc, dj);
    }
    if (head) nodes[node] = acc;
  }
  __syncwarp();
}

template <bool TAGS>
__global__ void __launch_bounds__(32)
sumtree_set_warp_kernel(double* nodes, int depth, const int32_t* __restrict__ idx, const double* __restrict__ val, int m,
                        double* max_prio, uint32_t* status, const int32_t* abort_flag) {
  if (abort_flag && *abort_flag) return;
  const int lane = threadIdx.x;
  sumtree_set_warp<TAGS>(nodes, depth, m, lane < m ? idx[lane] : 0, lane < m ? val[lane] : 0.0, max_prio, status);
}

constexpr int kOpsChunk = 256;       // ops whose descriptors are staged in shared memory at a time
constexpr int kOpsChunkEntries = 1024;

template <int MAXM, int THREADS>
__global__ void __launch_bounds__(THREADS)
sumtree_set_ops_kernel(double* nodes, int depth, const int32_t* __restrict__ op_offset, int n_ops,
                       const int32_t* __restrict__ idx, const double* __restrict__ val, double* max_prio,
                       uint32_t* status, int warp_ops) {
  extern __shared__ __align__(16) unsigned char set_raw[];
  SetSmem sm = carve_set_smem<MAXM, THREADS>(set_raw);
  // Descriptors of the next kOpsChunk ops (offsets, and their entries when they fit) are staged with coalesced loads:
  // a small op then costs the round trips of its tree accesses only, not three more for its own description.
  __shared__ int s_off[kOpsChunk + 1];
  __shared__ int s_idx[kOpsChunkEntries];
  __shared__ double s_val[kOpsChunkEntries];
  const int warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < n_ops; c0 += kOpsChunk) {
    const int c1 = min(n_ops, c0 + kOpsChunk);
    for (int i = threadIdx.x; i <= c1 - c0; i += THREADS) s_off[i] = op_offset[c0 + i];
    __syncthreads();
    const int e0 = s_off[0], e1 = s_off[c1 - c0];
    const bool staged = e1 - e0 <= kOpsChunkEntries && e1 >= e0;
    if (staged) {
      for (int i = threadIdx.x; i < e1 - e0; i += THREADS) {
        s_idx[i] = idx[e0 + i];
        s_val[i] = val[e0 + i];
      }
    }
    __syncthreads();
    bool small_pending = false;  // warp 0 may still be working on small ops the other warps have skipped
    for (int op = c0; op < c1; ++op) {
      const int b = s_off[op - c0], e = s_off[op - c0 + 1];
      const int m = e - b;
      if (m > MAXM || m < 0) {
        if (threadIdx.x == 0 && status) atomicOr(status, ISDQN_ST_OP_TOO_LARGE);
        continue;
      }
      if (m == 0) continue;
      if (m <= 2) {
        if (warp == 0) {
          const int i0 = staged ? s_idx[b - e0] : idx[b];
          const double v0 = staged ? s_val[b - e0] : val[b];
          const int i1 = m == 2 ? (staged ? s_idx[b + 1 - e0] : idx[b + 1]) : 0;
          const double v1 = m == 2 ? (staged ? s_val[b + 1 - e0] : val[b + 1]) : 0.0;
          sumtree_set_small(nodes, depth, m, i0, v0, i1, v1, max_prio, status);
        }
        small_pending = true;
        continue;
      }
      if (m <= 32 && warp_ops) {  // a batch-sized update inside the queue: still one warp, no block barrier
        if (warp == 0) {
          const int lane = threadIdx.x & 31;
          sumtree_set_warp<true>(nodes, depth, m, lane < m ? (staged ? s_idx[b + lane - e0] : idx[b + lane]) : 0,
                                 lane < m ? (staged ? s_val[b + lane - e0] : val[b + lane]) : 0.0, max_prio, status);
        }
        small_pending = true;
        continue;
      }
      if (small_pending) {  // sets are strictly ordered: this op reads the nodes the small ops wrote
        __syncthreads();
        small_pending = false;
      }
      sumtree_set_block<MAXM, THREADS, true>(nodes, depth, idx + b, val + b, m, max_prio, status, sm);
    }
    __syncthreads();  // (also: the staged descriptors are no longer read)
  }
}

// PrioritizedSamplingDistribution.update (samplers.py:76-88) with keys and priorities that live on the device: key ->
// dense index through the device mirror of `_key_to_index` (slot = key mod n_slots), priority -> priority ** alpha (0 stays 0).
// PRIO: 0 = float64 [n], 1 = float32 [n], 2 = float32 [prio_rows][n] averaged over the rows in ascending order (the per-head
// |TD| matrix the loss kernel writes).  A key that is not live sets ISDQN_ST_KEY_MISSING and the abort flag (the
// reference raises KeyError before it touches the tree).
// one entry of that translation: (dense index, priority ** alpha) of keys[i]; returns false for a key that is not live
__device__ __forceinline__ bool key_to_entry(int i, const int32_t* __restrict__ keys, const void* __restrict__ prio, int prio_kind,
                                             int prio_rows, int n, double prio_offset, double alpha,
                                             const int32_t* __restrict__ key_slot_to_index, int n_slots,
                                             const int32_t* __restrict__ index_to_key, int n_valid, int32_t* idx_out, double* p_out) {
  const int32_t key = keys[i];
  const int slot = (int)((((int64_t)key % n_slots) + n_slots) % n_slots);
  const int32_t idx = key_slot_to_index[slot];
  const bool live = idx >= 0 && idx < n_valid && index_to_key[idx] == key;
  double p;
  if (prio_kind == 0) {
    p = reinterpret_cast<const double*>(prio)[i];
  } else if (prio_kind == 1) {
    p = (double)reinterpret_cast<const float*>(prio)[i];
  } else {
    float acc = 0.f;
    for (int r = 0; r < prio_rows; ++r) acc += reinterpret_cast<const float*>(prio)[(int64_t)r * n + i];
    p = (double)(acc / (float)prio_rows);
  }
  p += prio_offset;
  if (!(alpha == 1.0)) p = p == 0.0 ? 0.0 : pow(p, alpha);
  *idx_out = live ? idx : 0;
  *p_out = p;
  return live;
}

__global__ void __launch_bounds__(256)
sumtree_keys_to_set_kernel(const int32_t* __restrict__ keys, const void* __restrict__ prio, int prio_kind, int prio_rows, int n,
                           double prio_offset, double alpha, const int32_t* __restrict__ key_slot_to_index, int n_slots,
                           const int32_t* __restrict__ index_to_key, int n_valid, int32_t* __restrict__ out_idx,
                           double* __restrict__ out_val, int32_t* abort_flag, uint32_t* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t idx;
  double p;
  if (!key_to_entry(i, keys, prio, prio_kind, prio_rows, n, prio_offset, alpha, key_slot_to_index, n_slots, index_to_key, n_valid,
                    &idx, &p)) {
    if (status) atomicOr(status, ISDQN_ST_KEY_MISSING);
    atomicExch(abort_flag, 1);
  }
  out_idx[i] = idx;
  out_val[i] = p;
}

// n <= 32 (the priority update of one training batch): translation and set in ONE warp-sized launch — lane = entry
__global__ void __launch_bounds__(32)
sumtree_keys_set_warp_kernel(double* nodes, int depth, const int32_t* __restrict__ keys, const void* __restrict__ prio,
                             int prio_kind, int prio_rows, int n, double prio_offset, double alpha,
                             const int32_t* __restrict__ key_slot_to_index, int n_slots,
                             const int32_t* __restrict__ index_to_key, int n_valid, double* max_prio, uint32_t* status) {
  const int lane = threadIdx.x;
  int32_t idx = 0;
  double p = 0.0;
  bool dead = false;
  if (lane < n)
    dead = !key_to_entry(lane, keys, prio, prio_kind, prio_rows, n, prio_offset, alpha, key_slot_to_index, n_slots, index_to_key,
                         n_valid, &idx, &p);
  if (dead && status) atomicOr(status, ISDQN_ST_KEY_MISSING);
  if (__any_sync(0xffffffffu, dead)) return;  // the reference raises KeyError before it touches the tree
  sumtree_set_warp<false>(nodes, depth, n, idx, p, max_prio, status);
}

__global__ void clear_flag_kernel(int32_t* flag) { *flag = 0; }

}  // namespace isdqn

using namespace isdqn;

static bool warp_set_enabled() {
  static const bool on = [] {
    const char* e = getenv("ISDQN_WARP_SET");
    return !(e && e[0] == '0');
  }();
  return on;
}

extern "C" int isdqn_sumtree_query(const double* d_nodes, int depth, const double* d_targets, int64_t n,
                                   int32_t* d_out_index, uint32_t* d_status, void* stream) {
  if (!d_nodes || !d_targets || !d_out_index || depth < 1 || depth > 31 || n < 0) return ISDQN_E_INVALID;
  if (n == 0) return ISDQN_OK;
  // A descent is a chain of dependent L2 reads: what counts is the number of descents in flight.  Up to one query per
  // thread while that still fits the resident threads (the 65,536-query launch ran 31.7 us with 4 per thread on 64 CTAs
  // against 20.4 us for the fused sampler's one per thread, profiles/r02_replay_ncu.md); 4 per thread beyond.
  const int64_t cap = (int64_t)kNumSMs * 8;
  const bool one = n <= cap * kQueryThreads;
  const int64_t per_cta = (int64_t)kQueryThreads * (one ? 1 : kQueryPerThread);
  int64_t grid = ceil_div<int64_t>(n, per_cta);
  if (grid > cap) grid = cap;
  // staging pays once a CTA serves many queries: each query needs <= 12 of the staged nodes
  int top_levels = 0;
  if (n / grid >= 4096) top_levels = depth - 1 < kQueryMaxTopLevels ? depth - 1 : kQueryMaxTopLevels;
  const size_t smem = top_levels > 0 ? sizeof(double) * ((1u << top_levels) - 1) : 0;
  ISDQN_PROF(as_stream(stream), "sumtree_query");
  if (one)
    sumtree_query_kernel<1><<<(unsigned)grid, kQueryThreads, smem, as_stream(stream)>>>(d_nodes, depth, d_targets, n, d_out_index,
                                                                                        d_status, top_levels);
  else
    sumtree_query_kernel<kQueryPerThread><<<(unsigned)grid, kQueryThreads, smem, as_stream(stream)>>>(
        d_nodes, depth, d_targets, n, d_out_index, d_status, top_levels);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_sumtree_set(double* d_nodes, int depth, const int32_t* d_index, const double* d_value,
                                 int32_t n, double* d_max_priority, uint32_t* d_status, void* stream) {
  if (!d_nodes || !d_index || !d_value || depth < 1 || depth > 31 || n < 0) return ISDQN_E_INVALID;
  if (n == 0) return ISDQN_OK;
  if (n > ISDQN_SUMTREE_SET_MAX) return ISDQN_E_TOO_LARGE;
  ISDQN_PROF(as_stream(stream), "sumtree_set");
  if (n <= 32 && warp_set_enabled()) {
    sumtree_set_warp_kernel<false><<<1, 32, 0, as_stream(stream)>>>(d_nodes, depth, d_index, d_value, n, d_max_priority, d_status,
                                                                   nullptr);
  } else if (n <= 1024) {
    sumtree_set_kernel<1024, 256><<<1, 256, set_smem_bytes<1024>(), as_stream(stream)>>>(
        d_nodes, depth, d_index, d_value, n, d_max_priority, d_status);
  } else {
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
      ISDQN_CUDA_CHECK(cudaFuncSetAttribute(sumtree_set_kernel<ISDQN_SUMTREE_SET_MAX, 1024>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)set_smem_bytes<ISDQN_SUMTREE_SET_MAX>()));
      }
    sumtree_set_kernel<ISDQN_SUMTREE_SET_MAX, 1024>
        <<<1, 1024, set_smem_bytes<ISDQN_SUMTREE_SET_MAX>(), as_stream(stream)>>>(d_nodes, depth, d_index, d_value, n,
                                                                                  d_max_priority, d_status);
  }
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int64_t isdqn_sumtree_set_keys_workspace_bytes(int32_t n) {
  return n < 0 ? -1 : 16 + (int64_t)(((int64_t)n * 4 + 15) / 16 * 16) + (int64_t)n * 8;
}

extern "C" int isdqn_sumtree_set_keys(double* d_nodes, int depth, const int32_t* d_keys, const void* d_priorities,
                                      int32_t prio_kind, int32_t prio_rows, int32_t n, double prio_offset, double alpha,
                                      const int32_t* d_key_slot_to_index, int32_t n_slots, const int32_t* d_index_to_key,
                                      int32_t n_valid, double* d_max_priority, uint32_t* d_status, void* d_workspace,
                                      int64_t workspace_bytes, void* stream) {
  if (!d_nodes || !d_keys || !d_priorities || !d_key_slot_to_index || !d_index_to_key || !d_workspace || depth < 1 || depth > 31 ||
      n < 0 || n_slots < 1 || prio_kind < 0 || prio_kind > 2 || (prio_kind == 2 && prio_rows < 1))
    return ISDQN_E_INVALID;
  if (n == 0) return ISDQN_OK;
  if (n > 1024) return ISDQN_E_TOO_LARGE;
  if (workspace_bytes < isdqn_sumtree_set_keys_workspace_bytes(n)) return ISDQN_E_INVALID;
  cudaStream_t s = as_stream(stream);
  if (n <= 32 && warp_set_enabled()) {
    ISDQN_PROF(s, "sumtree_keys_set");
    sumtree_keys_set_warp_kernel<<<1, 32, 0, s>>>(d_nodes, depth, d_keys, d_priorities, prio_kind, prio_rows, n, prio_offset, alpha,
                                                  d_key_slot_to_index, n_slots, d_index_to_key, n_valid, d_max_priority, d_status);
    ISDQN_LAUNCH_CHECK();
    return ISDQN_OK;
  }
  int32_t* flag = reinterpret_cast<int32_t*>(d_workspace);
  int32_t* idx = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(d_workspace) + 16);
  double* val = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(d_workspace) + 16 + ((int64_t)n * 4 + 15) / 16 * 16);
  ISDQN_PROF(s, "sumtree_keys_to_set");
  clear_flag_kernel<<<1, 1, 0, s>>>(flag);
  sumtree_keys_to_set_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d_keys, d_priorities, prio_kind, prio_rows, n, prio_offset, alpha, d_key_slot_to_index,
                                                             n_slots, d_index_to_key, n_valid, idx, val, flag, d_status);
  ISDQN_LAUNCH_CHECK();
  ISDQN_PROF(s, "sumtree_set");
  if (n <= 32 && warp_set_enabled())
    sumtree_set_warp_kernel<false><<<1, 32, 0, s>>>(d_nodes, depth, idx, val, n, d_max_priority, d_status, flag);
  else
    sumtree_set_kernel<1024, 256><<<1, 256, set_smem_bytes<1024>(), s>>>(d_nodes, depth, idx, val, n, d_max_priority, d_status, flag);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_sumtree_set_ops(double* d_nodes, int depth, const int32_t* d_op_offset, int32_t n_ops,
                                     const int32_t* d_index, const double* d_value, double* d_max_priority,
                                     uint32_t* d_status, void* stream) {
  if (!d_nodes || !d_op_offset || !d_index || !d_value || depth < 1 || depth > 31 || n_ops < 0)
    return ISDQN_E_INVALID;
  if (n_ops == 0) return ISDQN_OK;
  ISDQN_PROF(as_stream(stream), "sumtree_set_ops");
  sumtree_set_ops_kernel<ISDQN_SUMTREE_OP_MAX, 256><<<1, 256, set_smem_bytes<ISDQN_SUMTREE_OP_MAX>(), as_stream(stream)>>>(
      d_nodes, depth, d_op_offset, n_ops, d_index, d_value, d_max_priority, d_status, warp_set_enabled() ? 1 : 0);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}
