/*
 * isdqn_b200.h — C-ABI of libisdqn_b200.so: the B200 (sm_100a) learner hot path of iS-DQN.
 *
 * The reference (theovincent/iS-DQN, package `slimdqn`) has no native boundary: the hot path sits behind a
 * Python object API.  Every entry point below replaces the body of one reference method;
 * the `replaces:` line cites it (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * stub a maintainer adds on the reference side, and the XLA-FFI wrapper shape for a JAX host.
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is a DEVICE pointer owned by the caller and borrowed for
 *     the duration of the call; h_* are host pointers.  No entry point allocates device memory.
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; all work is enqueued on it, nothing synchronises
 *     unless the function name ends in `_sync`.
 *   - return value: 0 = ISDQN_OK, negative = ISDQN_E_* (see isdqn_strerror).  Nothing throws.
 *   - data-dependent failures that the reference reports with exceptions (negative priority, target out of
 *     range, ...) are reported through a caller-provided device status word `d_status` (bit mask ISDQN_ST_*),
 *     OR-ed by the kernels; the host wrapper reads it when it needs the result anyway.
 *   - thread-compatible: no global mutable state except the lazily dlopen'ed NCCL handle.
 */
#ifndef ISDQN_B200_H_
#define ISDQN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISDQN_ABI_VERSION 3

/* return codes */
#define ISDQN_OK 0
#define ISDQN_E_INVALID (-1)     /* bad argument (null pointer, non-positive size, unsupported shape) */
#define ISDQN_E_TOO_LARGE (-2)   /* request exceeds a documented kernel limit */
#define ISDQN_E_CUDA (-3)        /* a CUDA runtime call failed; see isdqn_last_cuda_error() */
#define ISDQN_E_UNSUPPORTED (-4) /* feature of the reference that is out of scope (batch_norm, impala) */
#define ISDQN_E_NCCL (-5)        /* NCCL missing or a NCCL call failed */

/* device status bits */
#define ISDQN_ST_NEGATIVE_VALUE 1u   /* sum_tree.py:31 `assert (values >= 0.0).all()` would fire */
#define ISDQN_ST_TARGET_RANGE 2u     /* sum_tree.py:74 ValueError: target outside [0, root) */
#define ISDQN_ST_DESCENT_ASSERT 4u   /* sum_tree.py:82 `assert (targets < nodes[node]).all()` would fire */
#define ISDQN_ST_EMPTY_TREE 8u       /* samplers.py:106 root == 0.0 */
#define ISDQN_ST_INDEX_RANGE 16u      /* leaf index outside the heap (NumPy would raise IndexError / wrap) */
#define ISDQN_ST_OP_TOO_LARGE 32u     /* an op of isdqn_sumtree_set_ops exceeds ISDQN_SUMTREE_OP_MAX */
#define ISDQN_ST_KEY_MISSING 64u      /* samplers.py:84 `self._key_to_index[key]` would raise KeyError */

int isdqn_abi_version(void);
const char* isdqn_strerror(int code);
const char* isdqn_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------------ sum tree
 * Array heap of float64, `depth` levels, (1<<depth)-1 nodes, first leaf at (1<<(depth-1))-1
 * (slimdqn/sample_collection/sum_tree.py:11-18). */

/* replaces: SumTree.query  sum_tree.py:58-102.  Batched inverse-CDF descent, strict `<` go-left rule.
 * d_out_index[i] = leaf index (int32).  Sets ISDQN_ST_TARGET_RANGE / ISDQN_ST_DESCENT_ASSERT. */
int isdqn_sumtree_query(const double* d_nodes, int depth, const double* d_targets, int64_t n,
                        int32_t* d_out_index, uint32_t* d_status, void* stream);

/* replaces: SumTree.set  sum_tree.py:20-47.  De-duplicates (first occurrence wins), then applies
 * delta = value - old_leaf to the leaf and every ancestor as a sequential left fold in ascending-leaf order
 * (the np.add.at order), so the float64 node array is bit-identical to the reference's.
 * *d_max_priority = max(*d_max_priority, max(values)) (sum_tree.py:32).  n <= ISDQN_SUMTREE_SET_MAX.
 * If any value is negative/NaN nothing is modified and ISDQN_ST_NEGATIVE_VALUE is set. */
#define ISDQN_SUMTREE_SET_MAX 8192
int isdqn_sumtree_set(double* d_nodes, int depth, const int32_t* d_index, const double* d_value, int32_t n,
                      double* d_max_priority, uint32_t* d_status, void* stream);

/* A queue of `n_ops` sets applied strictly in order by one launch: op j covers entries
 * [d_op_offset[j], d_op_offset[j+1]) of d_index/d_value, each op at most ISDQN_SUMTREE_OP_MAX entries.
 * In this entry point only, a value -(1+j) means "the value leaf j holds when the op starts" (the swap-remove
 * of samplers.py:99-102 without a device->host read) and ISDQN_SUMTREE_TAG_MAX means "max_recorded_priority as it
 * is when the op starts" (a prioritized training loop inserts new transitions at that priority).
 * replaces: the per-transition SumTree.set calls of PrioritizedSamplingDistribution.add / .remove
 * (samplers.py:67-74, 90-103), which the reference issues one NumPy call at a time. */
#define ISDQN_SUMTREE_OP_MAX 1024
#define ISDQN_SUMTREE_TAG_MAX (-0.5)
int isdqn_sumtree_set_ops(double* d_nodes, int depth, const int32_t* d_op_offset, int32_t n_ops,
                          const int32_t* d_index, const double* d_value, double* d_max_priority,
                          uint32_t* d_status, void* stream);

/* replaces: PrioritizedSamplingDistribution.update  samplers.py:76-88 (+ ReplayBuffer.update, replay_buffer.py:215-220)
 * for keys and priorities that already live on the device (the |TD| of the step that just ran): index =
 * d_key_slot_to_index[key mod n_slots] (device mirror of `_key_to_index`, verified against d_index_to_key), value =
 * (priority + prio_offset) ** alpha (0 stays 0; alpha == 1 is exact, otherwise CUDA's pow; prio_offset is the usual
 * small constant that keeps a zero TD error drawable, 0 for the reference's semantics), then one isdqn_sumtree_set.
 * prio_kind: 0 = float64 [n]; 1 = float32 [n]; 2 = float32 [prio_rows][n], averaged over the rows (the per-head |TD| matrix
 * of isdqn_train.d_td_abs).  A key that is not live sets ISDQN_ST_KEY_MISSING and leaves the tree untouched.  n <= 1024.
 * d_workspace: isdqn_sumtree_set_keys_workspace_bytes(n) bytes of device scratch. */
int64_t isdqn_sumtree_set_keys_workspace_bytes(int32_t n);
int isdqn_sumtree_set_keys(double* d_nodes, int depth, const int32_t* d_keys, const void* d_priorities, int32_t prio_kind,
                           int32_t prio_rows, int32_t n, double prio_offset, double alpha, const int32_t* d_key_slot_to_index, int32_t n_slots,
                           const int32_t* d_index_to_key, int32_t n_valid, double* d_max_priority, uint32_t* d_status,
                           void* d_workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------ samplers
 * d_rng: 6 x uint64 mirror of numpy's PCG64 state:
 *   [0]=state low 64, [1]=state high 64, [2]=inc low, [3]=inc high, [4]=has_uint32, [5]=uinteger. */

/* replaces: UniformSamplingDistribution.sample  samplers.py:39-49.
 * Draws exactly what `Generator.integers(n_valid, size=size)` draws (PCG64 next32 stream + Lemire rejection),
 * advances d_rng accordingly, d_out_index[i] = dense index, d_out_key[i] = d_index_to_key[index],
 * d_out_slot[i] = key % capacity (may be NULL, as may d_out_key with d_index_to_key). size <= 1<<20. */
int isdqn_sample_uniform(uint64_t* d_rng, int32_t n_valid, int32_t size, const int32_t* d_index_to_key,
                         int32_t capacity, int32_t* d_out_index, int32_t* d_out_key, int32_t* d_out_slot,
                         void* stream);

/* isdqn_sample_uniform with n_valid read from device memory when the kernel runs (graph-capturable while the buffer
 * fills); size <= 8192. */
int isdqn_sample_uniform_dev(uint64_t* d_rng, const int32_t* d_n_valid, int32_t size, const int32_t* d_index_to_key,
                             int32_t capacity, int32_t* d_out_index, int32_t* d_out_key, int32_t* d_out_slot,
                             void* stream);

/* isdqn_sample_uniform for large draws: same results and generator state, spread over the whole GPU (two passes over the
 * draw positions + a finish kernel).  d_workspace: isdqn_sample_uniform_workspace_bytes() of device scratch owned by the
 * caller; draws below 2048 (or a NULL workspace) take the single-CTA kernel. */
int64_t isdqn_sample_uniform_workspace_bytes(void);
int isdqn_sample_uniform_ws(uint64_t* d_rng, int32_t n_valid, int32_t size, const int32_t* d_index_to_key,
                            int32_t capacity, int32_t* d_out_index, int32_t* d_out_key, int32_t* d_out_slot,
                            void* d_workspace, int64_t workspace_bytes, void* stream);

/* replaces: PrioritizedSamplingDistribution.sample  samplers.py:105-116 (+ SumTree.query).
 * targets = Generator.uniform(0.0, root, size) with root read on the device, then the descent of
 * isdqn_sumtree_query.  d_out_target / d_out_prob may be NULL; d_out_prob[i] = leaf priority / root, the probability
 * the draw had (importance weights).  Sets ISDQN_ST_EMPTY_TREE when root == 0.  n_valid = number of live dense
 * indices: a descent that ends at or beyond it (empty tree, or a rounding that lands on a zero leaf — the reference's
 * `self._index_to_key[index]` raises IndexError there) sets ISDQN_ST_INDEX_RANGE and yields a live key instead, so the
 * slots handed to the gather are always valid; n_valid < 0 skips the check. */
int isdqn_sample_prioritized(uint64_t* d_rng, const double* d_nodes, int depth, int32_t size, int32_t n_valid,
                             const int32_t* d_index_to_key, int32_t capacity, int32_t* d_out_index,
                             int32_t* d_out_key, int32_t* d_out_slot, double* d_out_target, double* d_out_prob,
                             uint32_t* d_status, void* stream);

/* One training batch for a CAPTURED step (size <= 1024): the draw of isdqn_sample_prioritized with the live count read on
 * the device (*d_n_valid), the importance weights (n_valid * P(i))^-beta / max over the batch as float32 (beta read from
 * *d_beta, so that an annealed beta needs no re-capture; d_out_weight may be NULL) and the generator advance in one launch; nothing depends on host state, so the launch can live in the same CUDA
 * graph as the learner step and isdqn_sumtree_set_keys (new functionality: the reference never trains from its
 * prioritized sampler; draw = samplers.py:105-116). */
int isdqn_sample_prioritized_train(uint64_t* d_rng, const double* d_nodes, int depth, int32_t size, const int32_t* d_n_valid,
                                   const int32_t* d_index_to_key, int32_t capacity, int32_t* d_out_key, int32_t* d_out_slot,
                                   double* d_out_prob, const double* d_beta, float* d_out_weight, uint32_t* d_status, void* stream);

/* Scatter of host-accumulated patches into a device int32 table (index_to_key, element metadata):
 * d_table[d_patch_index[i]*width + j] = d_patch_value[i*width + j].  Later patches win. */
int isdqn_scatter_rows_i32(int32_t* d_table, int32_t width, const int32_t* d_patch_index,
                           const int32_t* d_patch_value, int32_t n_patches, void* stream);
int isdqn_scatter_rows_f64(double* d_table, const int32_t* d_patch_index, const double* d_patch_value,
                           int32_t n_patches, void* stream);

/* ------------------------------------------------------------------------------------------------- gather
 * Frame ring: `n_slots` frames of `frame_stride` bytes each (frame_stride % 16 == 0), slot `zero_slot` is all
 * zeros and stands for the reference's zero padding (replay_buffer.py:131-147).  Element metadata is indexed by
 * element slot = key % capacity: d_elem_frames[slot][2*stack] = ring slots of the `stack` state frames then the
 * `stack` next_state frames; d_elem_action int64, d_elem_reward float64, d_elem_terminal uint8.
 *
 * replaces: ReplayBuffer.sample  replay_buffer.py:198-213 (itemgetter + unpack + np.stack) for the drawn slots.
 * Outputs (batched on axis 0, layout = np.stack of ReplayElement fields):
 *   d_out_state / d_out_next : [n][frame_elems][stack] of out_dtype, observation axis order preserved, stack last
 *   d_out_action int64[n], d_out_reward float64[n], d_out_terminal uint8[n]   (any may be NULL) */
#define ISDQN_OUT_RAW 0  /* same bytes as stored (bit-exact parity path) */
#define ISDQN_OUT_F32 1  /* uint8 -> float32 x/255 (architectures/dqn.py:51), elem_size must be 1 */
#define ISDQN_OUT_BF16 2 /* uint8 -> bfloat16(x/255), elem_size must be 1 */
int isdqn_gather_stacks(const uint8_t* d_frames, int64_t frame_stride, int32_t frame_elems, int32_t elem_size,
                        int32_t stack, const int32_t* d_elem_frames, const int64_t* d_elem_action,
                        const double* d_elem_reward, const uint8_t* d_elem_terminal, const int32_t* d_slots,
                        int32_t n, int32_t out_dtype, void* d_out_state, void* d_out_next,
                        int64_t* d_out_action, double* d_out_reward, uint8_t* d_out_terminal, void* stream);

/* ------------------------------------------------------------------------------------------------ learner
 * Flat parameter vector: leaves packed in execution order, each leaf start aligned to 8 floats:
 *   cnn: Conv_i.kernel (HWIO), Conv_i.bias, [LayerNorm_i.scale, LayerNorm_i.bias] for i=0..2, then the Dense
 *        tail; fc: Dense tail only.  Dense kernels are (in, out) row-major; activations are NHWC.
 *   impala (dqn.py:7-36, 77-86): for each Stack_s, s = 0..2 (features[s] channels):
 *        Conv_0.kernel (3x3 HWIO), Conv_0.bias, then per residual block j = 0, 1:
 *        [LayerNorm_j.scale, LayerNorm_j.bias,] Conv_{1+2j}.kernel, .bias, Conv_{2+2j}.kernel, .bias;
 *        then [LayerNorm_0.scale, .bias] of the DQNNet itself, then the Dense tail (Dense_0 [, LayerNorm_1], ..., head).
 * (slimdqn/networks/architectures/dqn.py:47-103; names are flax's auto-names.) */
#define ISDQN_ARCH_CNN 0
#define ISDQN_ARCH_FC 1
#define ISDQN_ARCH_IMPALA 2
#define ISDQN_MAX_FEATURES 8
#define ISDQN_MAX_LEAVES 64

typedef struct isdqn_net {
  int32_t arch;        /* ISDQN_ARCH_* */
  int32_t layer_norm;  /* 0/1 */
  int32_t obs_h, obs_w, obs_c; /* cnn: (84,84,4); fc: obs_c = flattened observation size, h = w = 1 */
  int32_t n_features;
  int32_t features[ISDQN_MAX_FEATURES];
  int32_t n_heads;     /* K = n_bellman_iterations (the net has 1+K heads) */
  int32_t n_actions;   /* A */
} isdqn_net;

typedef struct isdqn_layout {
  int32_t n_leaves;
  int64_t offset[ISDQN_MAX_LEAVES]; /* in floats */
  int64_t size[ISDQN_MAX_LEAVES];   /* in floats */
  int64_t total;                    /* padded total, multiple of 8 */
} isdqn_layout;

int isdqn_net_layout(const isdqn_net* net, isdqn_layout* out);
/* bytes of scratch the forward (n_rows images) / the training step (batch B) need */
int64_t isdqn_forward_workspace_bytes(const isdqn_net* net, int32_t n_rows);
int64_t isdqn_learn_workspace_bytes(const isdqn_net* net, int32_t batch);

/* replaces: DQNNet.apply  architectures/dqn.py:47-103 (+ the reshape of isdqn.py:39-41, which is a view).
 * d_input: cnn uint8 [n_rows][H][W][C] (normalised /255 inside) or, with input_is_float, float32 of the same
 * shape (tests/utils.py feeds float32 in [0,1) which the reference also divides by 255); fc float32 [n_rows][D].
 * d_q: float32 [n_rows][(1+K)*A]. */
int isdqn_forward(const isdqn_net* net, const float* d_params, const void* d_input, int32_t input_is_float,
                  int32_t n_rows, float* d_q, void* d_workspace, int64_t workspace_bytes, void* stream);

/* replaces: the loss tail of iSDQN.loss_on_batch + compute_target  isdqn.py:97-109 and its gradient.
 * d_q_all float32 [2B][(1+K)*A] (rows [0,B) = s, rows [B,2B) = s').  d_losses[k] = mean_b td^2 (k = 0..K-1);
 * d_dq (may be NULL) float32 [B][(1+K)*A] = d(sum_k losses)/d q_all[:B], scaled with 1/batch_global.
 * reward float64 -> float32, action int64, terminal uint8 (the dtypes rb.sample() yields). */
int isdqn_heads_td_loss(const float* d_q_all, const int64_t* d_action, const double* d_reward,
                        const uint8_t* d_terminal, float gamma_n, int32_t batch, int32_t batch_global,
                        int32_t n_heads, int32_t n_actions, float* d_losses, float* d_dq, void* stream);
/* the same with the prioritized-replay extras of isdqn_train (either may be NULL) */
int isdqn_heads_td_loss_weighted(const float* d_q_all, const int64_t* d_action, const double* d_reward,
                                 const uint8_t* d_terminal, float gamma_n, int32_t batch, int32_t batch_global,
                                 int32_t n_heads, int32_t n_actions, const float* d_is_weights, float* d_losses,
                                 float* d_dq, float* d_td_abs, void* stream);

/* replaces: optax.adam(lr, eps).update + optax.apply_updates  isdqn.py:46,85-86 (optax 0.2.4 semantics:
 * eps outside the sqrt, bias correction with count+1).  *d_count is incremented on the device. */
int isdqn_adam_step(float* d_params, const float* d_grads, float* d_mu, float* d_nu, int32_t* d_count, float lr,
                    float b1, float b2, float eps, int64_t n, void* stream);

/* Same update for a counter the caller already advanced (*d_count = t >= 1); used inside the fused step, where the
 * loss kernel advances the counter so that the whole step stays one capturable launch sequence. */
int isdqn_adam_step_nocount(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count,
                            float lr, float b1, float b2, float eps, int64_t n, void* stream);

/* replaces: iSDQN.shift_params  isdqn.py:111-125: kernel[:, :-A] = kernel[:, A:], bias[:-A] = bias[A:]. */
int isdqn_shift_heads(float* d_kernel, float* d_bias, int32_t n_in, int32_t n_heads, int32_t n_actions,
                      void* stream);

typedef struct isdqn_batch {
  const void* d_state;      /* cnn: uint8 [B][H][W][C]; fc: float32 [B][D] */
  const void* d_next_state;
  const int64_t* d_action;
  const double* d_reward;
  const uint8_t* d_terminal;
} isdqn_batch;

typedef struct isdqn_train {
  float gamma_n;            /* gamma ** update_horizon */
  float lr, b1, b2, eps;
  int32_t batch;            /* rows on this device */
  int32_t batch_global;     /* = batch unless data parallel */
  float* d_params; float* d_grads; float* d_mu; float* d_nu; int32_t* d_count;
  float* d_losses;          /* [K] per-head mean TD^2 of THIS step (local share under DP) */
  void* d_workspace; int64_t workspace_bytes;
  void* nccl_comm;          /* NULL, or the handle from isdqn_dp_init: gradients are all-reduced before Adam */
  int32_t compute_dtype;    /* ISDQN_COMPUTE_F32 (CUDA-core fp32, 1e-5 parity) or ISDQN_COMPUTE_BF16 (tcgen05, 2e-2) */
  void* d_workspace_tc; int64_t workspace_tc_bytes; /* bf16 path only: isdqn_learn_workspace_tc_bytes() */
  void* d_params_bf16;      /* bf16 path only: bf16 shadow of d_params (layout.total elements).  Adam keeps it current;  */
  int32_t refresh_shadow;   /* set to 1 to rebuild it from d_params at the start of the call (parameters were changed     */
                            /* outside isdqn_learn_on_batch since the last call)                                          */
  double* d_cumulated;      /* NULL, or [K]: isdqn_learn_on_batch also does d_cumulated[k] += d_losses[k] (the                */
                            /* `self.cumulated_losses += losses` of isdqn.py:62, kept on the device)                        */
  /* prioritized training driver (new functionality: the reference never wires its prioritized sampler to an agent) */
  const float* d_is_weights; /* NULL, or [B] importance weights: losses[k] = mean_b w_b td^2 (gradient scaled alike)        */
  float* d_td_abs;           /* NULL, or [K][B]: |td| of every online head and sample (isdqn_sumtree_set_keys, prio_kind 2)    */
} isdqn_train;
#define ISDQN_COMPUTE_F32 0
#define ISDQN_COMPUTE_BF16 1
/* bytes of bf16 scratch the tensor-core path needs (0 if the network is not eligible: cnn with 32/64/128/256
 * channel convs and hidden Dense widths that are multiples of 64) */
int64_t isdqn_learn_workspace_tc_bytes(const isdqn_net* net, int32_t batch);

/* fp32 -> bf16 (round to nearest even) of n elements, n % 4 == 0: builds the parameter shadow */
int isdqn_cast_f32_to_bf16(const float* d_src, void* d_dst_bf16, int64_t n, void* stream);

/* Building block / test entry of the tcgen05 tile engine: D[M][N] (fp32) = A B^T, bf16 operands.
 *   a_mn_major = 0: A is [M][lda] (K contiguous); 1: A is [K][lda] (M contiguous); likewise B with N.
 * splits > 1 writes the partial products of `splits` K ranges at d_c + z*M*N. lda, ldb, N multiples of 8. */
int isdqn_tc_gemm_bf16(const void* d_a, int64_t lda, int32_t a_mn_major, const void* d_b, int64_t ldb, int32_t b_mn_major,
                       float* d_c, int32_t M, int32_t N, int32_t K, int32_t splits, void* stream);

/* replaces: iSDQN.loss_on_batch  isdqn.py:92-103 (forward on concat(s, s') + loss, no gradient). */
int isdqn_loss_on_batch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* batch, float* d_q_all,
                        void* stream);
/* replaces: iSDQN.learn_on_batch  isdqn.py:82-90: forward, loss, backward, [allreduce], Adam — one call,
 * graph-capturable (no host synchronisation, no allocation). */
int isdqn_learn_on_batch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* batch, void* stream);
/* backward only up to the gradient (used by parity tests and by the DP path) */
int isdqn_grad_on_batch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* batch, void* stream);

/* replaces: iSDQN.best_action  isdqn.py:127-135 for a given head index: argmax_a q[1+idx_network][a]. */
int isdqn_best_action(const isdqn_net* net, const float* d_params, const void* d_state, int32_t input_is_float,
                      int32_t idx_network, int32_t* d_action, void* d_workspace, int64_t workspace_bytes,
                      void* stream);

/* ---------------------------------------------------------------------------------------- CUDA graph helpers
 * The step above is ~20 launches; at batch 32 it is launch-latency bound, so the host captures it once. */
int isdqn_graph_begin(void* stream);
int isdqn_graph_end(void* stream, void** out_graph_exec);
int isdqn_graph_launch(void* graph_exec, void* stream);
int isdqn_graph_destroy(void* graph_exec);

/* ------------------------------------------------------------------------------------------------- profiler
 * Measurement support for bench.py (no counterpart in the reference): between begin and end every kernel launch
 * of this library records a CUDA event on its stream; end returns (name, milliseconds) per launch. */
int isdqn_profile_begin(void);
int isdqn_profile_end(void* stream, int32_t max_entries, char* names, int32_t name_stride, float* ms);
/* a ~micros busy-wait kernel: lets the host queue the following launches back to back (no launch gaps) */
int isdqn_spin(void* stream, int32_t micros);

/* -------------------------------------------------------------------------------------------- data parallel
 * New functionality (the reference has no collective): gradient all-reduce over NVLink. NCCL is
 * dlopen'ed on first use.  h_unique_id is the 128-byte ncclUniqueId produced on rank 0. */
int isdqn_dp_unique_id(uint8_t* h_unique_id_128);
int isdqn_dp_init(const uint8_t* h_unique_id_128, int32_t rank, int32_t world, void** out_comm);
int isdqn_dp_allreduce_f32(void* comm, float* d_buf, int64_t n, void* stream);
int isdqn_dp_destroy(void* comm);

/* Host-batch staging for the reference-facing call `learn_on_batch(params, opt_state, host batch)` (isdqn.py:82; the
 * reference's implicit device_put at the jit boundary).  Events are created without timing.
 * isdqn_stage_batch: h_src (pinned) -> d_stage on copy_stream [after ev_stage_free], record ev_h2d_done; step_stream waits
 * it, copies d_stage -> d_dst and re-records ev_stage_free (d_dst == NULL: no second copy — the step reads d_stage in place
 * and the caller records ev_stage_free behind it with isdqn_event_record).  isdqn_read_async: d_src -> h_dst (pinned) on stream, then
 * records event.  Nothing synchronises except isdqn_event_synchronize. */
int isdqn_event_create(void** out_event);
int isdqn_event_destroy(void* event);
int isdqn_event_record(void* event, void* stream);
int isdqn_event_synchronize(void* event);
int isdqn_stage_batch(const void* h_src, void* d_stage, void* d_dst, int64_t bytes, void* copy_stream, void* step_stream,
                      void* ev_h2d_done, void* ev_stage_free);
int isdqn_read_async(void* h_dst, const void* d_src, int64_t bytes, void* stream, void* event);
int isdqn_write_async(void* d_dst, const void* h_src, int64_t bytes, void* stream); /* pinned host -> device */

/* Acting path (isdqn.py:127-135): d_out[h] = argmax_a d_q[h * n_actions + a] for every head h of ONE row of Q-values
 * (first maximum, like jnp.argmax); the caller picks head 1 + idx on the host.  Together with isdqn_loss_on_batch on a
 * batch of one (s' aliased to s) this is captured in a CUDA graph by the Python host (iSDQN.best_action). */
int isdqn_argmax_heads(const float* d_q, int32_t n_heads_total, int32_t n_actions, int32_t* d_out, void* stream);

/* Acting path as ONE kernel (replaces the jitted body of iSDQN.best_action, isdqn.py:127-135, and the
 * DQNNet forward it runs, architectures/dqn.py:47-103, for the `cnn` architecture on ONE uint8 observation stack
 * [H][W][C]): conv torso + LayerNorm + hidden Dense + head layer + the greedy action of EVERY head (d_actions[1 + K],
 * first maximum like jnp.argmax; the caller picks head 1 + idx).  fp32 master weights whatever the learner's compute
 * dtype.  d_q (optional) receives the (1 + K) * A Q-values.  d_workspace: isdqn_act_workspace_bytes(net) bytes, ZEROED
 * once by the caller (its first 16 bytes are the grid barrier, re-armed by every launch); 0 bytes = the network is
 * outside what this path covers (ISDQN_E_UNSUPPORTED from isdqn_act; use isdqn_best_action).
 * isdqn_act_host: the whole env-step call — pinned observation -> d_obs, the kernel, actions -> pinned host, and the
 * one synchronisation the reference's `.item()` performs (on `event`). */
int64_t isdqn_act_workspace_bytes(const isdqn_net* net);
int isdqn_act(const isdqn_net* net, const float* d_params, const uint8_t* d_obs, float* d_q, int32_t* d_actions,
              void* d_workspace, int64_t workspace_bytes, void* stream);
int isdqn_act_host(const isdqn_net* net, const float* d_params, const uint8_t* h_obs_pinned, uint8_t* d_obs,
                   int64_t obs_bytes, float* d_q, int32_t* d_actions, int32_t* h_actions_pinned, void* d_workspace,
                   int64_t workspace_bytes, void* stream, void* event);
/* isdqn_act_mapped: the same call with no copies and no event — the kernel reads the observation out of pinned host
 * memory (mapped into the device's address space) and writes the 1 + K actions and then `seq` into *h_flag_pinned straight
 * back; the host spins on the flag (isdqn_act_wait, or inside the call when timeout_us > 0).  ISDQN_E_CUDA on timeout. */
int isdqn_act_mapped(const isdqn_net* net, const float* d_params, const uint8_t* h_obs_pinned, float* d_q,
                     int32_t* h_actions_pinned, int32_t* h_flag_pinned, int32_t seq, void* d_workspace,
                     int64_t workspace_bytes, void* stream, int64_t timeout_us);
int isdqn_act_wait(const int32_t* h_flag_pinned, int32_t seq, int64_t timeout_us);

/* Host helpers (no device work).  replaces: the head draw of iSDQN.best_action, `jax.random.randint(key, (), 0, K)`
 * (isdqn.py:129), restated from jax 0.4.30's threefry PRNG so that the same raw JAX key (uint32[2]) picks the same head.
 * isdqn_threefry2x32: the Threefry-2x32 (20 rounds) block function, out2 = E_key(x0, x1). */
void isdqn_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* out2);
int32_t isdqn_threefry_randint(uint32_t k0, uint32_t k1, int32_t minval, int32_t maxval);
/* jax.random.split(key, num) -> out_keys[2 * num] raw key words (key i = out_keys[2i], out_keys[2i + 1]) and
 * jax.random.uniform(key) (float32 in [0, 1)): the epsilon-greedy draws of select_action
 * (slimdqn/sample_collection/utils.py:8-15), restated like isdqn_threefry_randint. */
void isdqn_threefry_split(uint32_t k0, uint32_t k1, int32_t num, uint32_t* out_keys);
float isdqn_threefry_uniform(uint32_t k0, uint32_t k1);

/* Diagnostic: device-side timeline.  While d_buf (uint64[4001], zero-initialised device memory) is set, CTA (0,0,0) of
 * every learner-step kernel appends its start time in ns (%globaltimer) at d_buf[1 + d_buf[0]++].  NULL switches it off. */
int isdqn_trace_set(void* d_buf);

#ifdef __cplusplus
}
#endif
#endif /* ISDQN_B200_H_ */
