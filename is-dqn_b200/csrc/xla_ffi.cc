// XLA FFI registration shim (north_star: "a thin C-ABI exposed as a JAX/XLA FFI custom call").
//
// Compiled ONLY when the XLA FFI headers are on the include path (`make ffi XLA_FFI_INCLUDE=<dir holding xla/ffi/api/ffi.h>`,
// e.g. `python -c "import jaxlib; print(jaxlib.__path__[0] + '/include')"`).  This image ships neither JAX nor the
// headers, so the translation unit is not part of the default build and is untested here; it is mechanical:
// every entry point of include/isdqn_b200.h already has the FFI calling convention (device pointers + sizes + a stream, no
// allocation, status return).  Python side (jax >= 0.4.31):
//     jax.ffi.register_ffi_target("isdqn_sumtree_query", jax.ffi.pycapsule(lib.IsdqnSumTreeQuery), platform="CUDA")
//     idx, st = jax.ffi.ffi_call("isdqn_sumtree_query", (ShapeDtypeStruct((n,), int32), ShapeDtypeStruct((1,), uint32)))(
//                   nodes, targets, depth=21)
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define ISDQN_HAVE_XLA_FFI 1
#endif
#endif

#ifdef ISDQN_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "../../include/isdqn_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(isdqn_strerror(rc)); }

// replaces: SumTree.query  slimdqn/sample_collection/sum_tree.py:58-102
static ffi::Error SumTreeQuery(cudaStream_t stream, ffi::Buffer<ffi::F64> nodes, ffi::Buffer<ffi::F64> targets, int32_t depth,
                               ffi::ResultBuffer<ffi::S32> out, ffi::ResultBuffer<ffi::U32> status) {
  return Status(isdqn_sumtree_query(nodes.typed_data(), depth, targets.typed_data(), (int64_t)targets.element_count(),
                                    out->typed_data(), status->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IsdqnSumTreeQuery, SumTreeQuery,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("depth")
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>());

// replaces: SumTree.set  sum_tree.py:20-47 (nodes / max_priority are updated in place: declare them as input-output aliases)
static ffi::Error SumTreeSet(cudaStream_t stream, ffi::Buffer<ffi::F64> nodes, ffi::Buffer<ffi::S32> index,
                             ffi::Buffer<ffi::F64> value, ffi::Buffer<ffi::F64> max_priority, int32_t depth,
                             ffi::ResultBuffer<ffi::U32> status) {
  return Status(isdqn_sumtree_set(nodes.typed_data(), depth, index.typed_data(), value.typed_data(), (int32_t)index.element_count(),
                                  max_priority.typed_data(), status->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IsdqnSumTreeSet, SumTreeSet,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<int32_t>("depth")
                                  .Ret<ffi::Buffer<ffi::U32>>());

// replaces: ReplayBuffer.sample  replay_buffer.py:198-213 for drawn element slots (uint8 stack-4 frames, raw output)
static ffi::Error GatherStacks(cudaStream_t stream, ffi::Buffer<ffi::U8> frames, ffi::Buffer<ffi::S32> elem_frames,
                               ffi::Buffer<ffi::S64> elem_action, ffi::Buffer<ffi::F64> elem_reward,
                               ffi::Buffer<ffi::U8> elem_terminal, ffi::Buffer<ffi::S32> slots, int64_t frame_stride,
                               int32_t frame_elems, int32_t stack, ffi::ResultBuffer<ffi::U8> state, ffi::ResultBuffer<ffi::U8> next,
                               ffi::ResultBuffer<ffi::S64> action, ffi::ResultBuffer<ffi::F64> reward,
                               ffi::ResultBuffer<ffi::U8> terminal) {
  return Status(isdqn_gather_stacks(frames.typed_data(), frame_stride, frame_elems, 1, stack, elem_frames.typed_data(),
                                    elem_action.typed_data(), elem_reward.typed_data(), elem_terminal.typed_data(),
                                    slots.typed_data(), (int32_t)slots.element_count(), ISDQN_OUT_RAW, state->typed_data(),
                                    next->typed_data(), action->typed_data(), reward->typed_data(), terminal->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IsdqnGatherStacks, GatherStacks,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Attr<int64_t>("frame_stride")
                                  .Attr<int32_t>("frame_elems")
                                  .Attr<int32_t>("stack")
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// replaces: the loss tail of iSDQN.loss_on_batch + compute_target  slimdqn/networks/isdqn.py:97-109 (+ its gradient)
static ffi::Error HeadsTdLoss(cudaStream_t stream, ffi::Buffer<ffi::F32> q_all, ffi::Buffer<ffi::S64> action,
                              ffi::Buffer<ffi::F64> reward, ffi::Buffer<ffi::U8> terminal, float gamma_n, int32_t n_heads,
                              int32_t n_actions, ffi::ResultBuffer<ffi::F32> losses, ffi::ResultBuffer<ffi::F32> dq) {
  const int32_t batch = (int32_t)action.element_count();
  return Status(isdqn_heads_td_loss(q_all.typed_data(), action.typed_data(), reward.typed_data(), terminal.typed_data(), gamma_n,
                                    batch, batch, n_heads, n_actions, losses->typed_data(), dq->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IsdqnHeadsTdLoss, HeadsTdLoss,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Attr<float>("gamma_n")
                                  .Attr<int32_t>("n_heads")
                                  .Attr<int32_t>("n_actions")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// replaces: optax.adam(...).update + optax.apply_updates  isdqn.py:46,85-86 (params / mu / nu / count aliased in place)
static ffi::Error AdamStep(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> grads, ffi::Buffer<ffi::F32> mu,
                           ffi::Buffer<ffi::F32> nu, ffi::Buffer<ffi::S32> count, float lr, float b1, float b2, float eps) {
  return Status(isdqn_adam_step(params.typed_data(), grads.typed_data(), mu.typed_data(), nu.typed_data(), count.typed_data(), lr,
                                b1, b2, eps, (int64_t)params.element_count(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(IsdqnAdamStep, AdamStep,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Attr<float>("lr")
                                  .Attr<float>("b1")
                                  .Attr<float>("b2")
                                  .Attr<float>("eps"));
#else
// XLA FFI headers not found: nothing to compile (the tested caller in this image is ctypes).
extern "C" int isdqn_xla_ffi_available(void) { return 0; }
#endif
