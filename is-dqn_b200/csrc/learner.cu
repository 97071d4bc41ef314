// Learner entry points (fp32 path): forward, K-head TD loss, backward, Adam — host orchestration of the kernels
// in learner_kernels.cuh.  Replaces the jitted body of iSDQN.learn_on_batch / loss_on_batch / best_action
// (slimdqn/networks/isdqn.py:82-135) and DQNNet.__call__ (slimdqn/networks/architectures/dqn.py:47-103).
#include "learner_kernels.cuh"
#include "impala_kernels.cuh"
#include "dense_small.cuh"
#include "plan.cuh"

using namespace isdqn;

extern "C" int isdqn_adam_step_nocount(float* d_params, const float* d_grads, float* d_mu, float* d_nu,
                                       const int32_t* d_count, float lr, float b1, float b2, float eps, int64_t n,
                                       void* stream);

int isdqn_tc_train_dispatch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, bool backward, bool update,
                            float* q_out, void* stream);
// building blocks of the tile engine (tc_learner.cu)
bool isdqn_tc_conv_ok(const isdqn::Layer& L);
int isdqn_tc_conv_fwd(const isdqn::Layer& L, const void* x16, int rows, const void* w16, const float* params, void* out16,
                      cudaStream_t s, float in_scale = 1.0f);
int isdqn_tc_conv_wgrad(const isdqn::Layer& L, const void* x16, const void* dz16, float* part, int rows_l, int splits,
                        int* real_splits, cudaStream_t s, float in_scale = 1.0f);
int isdqn_tc_conv_dgrad(const isdqn::Layer& L, const void* dz16, const void* w16, float* dx, int B, cudaStream_t s);
int isdqn_cast_bf16_launch(const float* src, void* dst16, int64_t n, cudaStream_t s);
int isdqn_adam_launch(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count, float lr,
                      float b1, float b2, float eps, int64_t n, void* d_shadow_bf16, void* stream, int64_t skip_begin,
                      int64_t skip_len, int max_ctas);

namespace {

inline float* wsp(void* ws, int64_t off) { return off < 0 ? nullptr : reinterpret_cast<float*>(ws) + off; }

template <int KIND>
int launch_conv_fwd_kind(const ConvArgs& a, cudaStream_t s) {
  ISDQN_PROF(s, "conv_fwd");
  if (a.Cout <= 32) {
    conv_fwd_kernel<64, 32, 4, KIND><<<ceil_div(a.M, 64), kGemmThreads, 0, s>>>(a);
  } else if (a.Cout <= 64) {
    conv_fwd_kernel<32, 64, 4, KIND><<<ceil_div(a.M, 32), kGemmThreads, 0, s>>>(a);
  } else if (a.Cout <= 128) {
    conv_fwd_kernel<32, 128, 8, KIND><<<ceil_div(a.M, 32), kGemmThreads, 0, s>>>(a);
  } else if (a.Cout <= 256) {
    conv_fwd_kernel<16, 256, 8, KIND><<<ceil_div(a.M, 16), kGemmThreads, 0, s>>>(a);
  } else {
    return ISDQN_E_TOO_LARGE;
  }
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

int launch_conv_fwd(const ConvArgs& a, int kind, cudaStream_t s) {
  switch (kind) {
    case IN_U8_255: return launch_conv_fwd_kind<IN_U8_255>(a, s);
    case IN_F32: return launch_conv_fwd_kind<IN_F32>(a, s);
    default: return launch_conv_fwd_kind<IN_F32_255>(a, s);
  }
}

int launch_gemm(const GemmArgs& g, int splits, cudaStream_t s, const char* tag) {
  ISDQN_PROF(s, tag);
  dim3 grid(ceil_div(g.M, 64), ceil_div(g.N, 64), splits);
  const bool a_kfast = g.sak == 1;
  const bool b_nfast = g.sbn == 1;
  if (a_kfast && b_nfast) gemm_strided_kernel<true, true><<<grid, kGemmThreads, 0, s>>>(g);
  else if (a_kfast) gemm_strided_kernel<true, false><<<grid, kGemmThreads, 0, s>>>(g);
  else if (b_nfast) gemm_strided_kernel<false, true><<<grid, kGemmThreads, 0, s>>>(g);
  else gemm_strided_kernel<false, false><<<grid, kGemmThreads, 0, s>>>(g);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

void fill_conv_geom(const Layer& L, ConvArgs* a) {
  a->H = L.H; a->W = L.W; a->Cin = L.Cin; a->OH = L.OH; a->OW = L.OW; a->Cout = L.out_dim;
  a->ksz = L.ksz; a->stride = L.stride; a->pad_y = L.pad_y; a->pad_x = L.pad_x;
  a->K = L.in_dim;
}

// Hidden Dense layers of an `fc` plan on the tile engine (the Dense tail of the impala network in tensor-core mode):
// bf16 copies of the layer inputs, bf16 kernels out of the parameter shadow, fp32 partials / LayerNorm / head as always.
struct DenseTc {
  const __nv_bfloat16* in16;                        // bf16 copy of the plan's input [rows][in_dim of layer 0]
  __nv_bfloat16* act16[ISDQN_MAX_FEATURES + 1];     // bf16 copy of the output of hidden layer l [rows][out_dim]
  __nv_bfloat16* dz16;                              // [B][widest hidden layer]: gradient of the pre-activation output
  const __nv_bfloat16* shadow;                      // bf16 parameters, same offsets as the plan's
};
bool dense_small_on() {  // A/B: ISDQN_DENSE_SMALL=0 keeps the tiled GEMM for the small-batch Dense layers
  static const bool on = [] {
    const char* e = getenv("ISDQN_DENSE_SMALL");
    return !(e && e[0] == '0');
  }();
  return on;
}
bool dense_tc_ok(const Layer& L) { return L.type == 1 && L.in_dim % 8 == 0 && L.out_dim % 8 == 0; }

// Forward through every layer.  in0/in1: the two halves of concat(s, s') (in1 may be null, n0 = rows then).
int run_forward(const Plan& p, const Workspace& w, void* ws, const float* params, const void* in0, const void* in1,
                int n0, int rows, int rows_train, int in_kind, cudaStream_t s, const DenseTc* dtc = nullptr) {
  for (int l = 0; l < p.n_layers; ++l) {
    const Layer& L = p.L[l];
    const bool first = l == 0;
    float* out = wsp(ws, w.act[l]);
    const float* ln_g = L.has_ln ? params + L.g_off : nullptr;
    const float* ln_b = L.has_ln ? params + L.beta_off : nullptr;
    if (L.type == 0) {
      ConvArgs a;
      fill_conv_geom(L, &a);
      a.in0 = first ? in0 : wsp(ws, w.act[l - 1]);
      a.in1 = first ? in1 : nullptr;
      a.n_img0 = first ? n0 : rows;
      a.M = rows * L.pix;
      a.w = params + L.w_off; a.bias = params + L.b_off; a.ln_g = ln_g; a.ln_b = ln_b; a.relu = L.relu;
      a.out = out;
      a.xhat = rows_train > 0 ? wsp(ws, w.xhat[l]) : nullptr;
      a.rstd = rows_train > 0 ? wsp(ws, w.rstd[l]) : nullptr;
      a.m_train = rows_train * L.pix;
      int rc = launch_conv_fwd(a, first ? in_kind : IN_F32, s);
      if (rc) return rc;
    } else {
      if (L.out_dim > kRowThreads * kRowMaxPerThread) return ISDQN_E_TOO_LARGE;
      // dense input: previous activation viewed as [rows][in_dim]; a first dense layer (fc) reads float input
      // from the two halves, which must then be contiguous rows: handled by two GEMM launches.
      const int splits = dense_fwd_splits(rows, L.out_dim, L.in_dim);
      const int kps = ceil_div(ceil_div(L.in_dim, splits), kBK) * kBK;
      const int real_splits = ceil_div(L.in_dim, kps);
      if (!first && !L.has_ln && !L.relu && L.out_dim <= 128 && L.in_dim <= kHeadMaxK) {  // the head layer
        ISDQN_PROF(s, "head_fwd");
        ISDQN_CUDA_CHECK(launch_head_fwd(s, wsp(ws, w.act[l - 1]), params + L.w_off, params + L.b_off, L.in_dim, L.out_dim, out, rows));
        continue;
      }
      const bool on_tc = dtc != nullptr && l + 1 < p.n_layers && dense_tc_ok(L) && (!first || in1 == nullptr);
      const bool direct = !on_tc && real_splits == 1 && !L.has_ln && !L.relu;
      float* part = direct ? out : wsp(ws, w.fwd_part);
      const int64_t split_stride = (int64_t)rows * L.out_dim;
      int n_parts = real_splits;
      if (on_tc) {
        // the tile engine splits the K axis in 64-element chunks: ask for a split count that is a fixed point of its rule
        // (chunks per split = ceil(chunks / splits), splits = ceil(chunks / chunks per split)), so that exactly n_parts
        // partials are written
        const int chunks = ceil_div(L.in_dim, 64);
        n_parts = real_splits < chunks ? real_splits : chunks;
        for (int it = 0; it < 8; ++it) n_parts = ceil_div(chunks, ceil_div(chunks, n_parts));
        int rc = isdqn_tc_gemm_bf16(first ? dtc->in16 : dtc->act16[l - 1], L.in_dim, 0, dtc->shadow + L.w_off, L.out_dim, 1, part,
                                    rows, L.out_dim, L.in_dim, n_parts, s);
        if (rc) return rc;
      }
      // few rows: the memory-shaped kernel (dense_small.cuh) instead of the tiled GEMM; same partial layout
      const bool small_fwd = !on_tc && !direct && rows <= 64 && L.out_dim % 4 == 0 && dense_small_on();
      if (small_fwd) {
        ISDQN_PROF(s, "dense_fwd_small");
        dense_fwd_small_kernel<<<dim3(ceil_div(L.out_dim, 128), real_splits), 256, 0, s>>>(
            reinterpret_cast<const float*>(first ? in0 : wsp(ws, w.act[l - 1])),
            reinterpret_cast<const float*>(first ? in1 : nullptr), first ? n0 : rows, rows, L.in_dim, L.out_dim, params + L.w_off,
            part, split_stride, kps);
        ISDQN_LAUNCH_CHECK();
      }
      for (int half = 0; half < 2 && !on_tc && !small_fwd; ++half) {
        GemmArgs g;
        int m_begin, m_count;
        if (first) {
          if (half == 0) { g.A = reinterpret_cast<const float*>(in0); m_begin = 0; m_count = n0; }
          else { g.A = reinterpret_cast<const float*>(in1); m_begin = n0; m_count = rows - n0; }
        } else {
          if (half == 1) break;
          g.A = wsp(ws, w.act[l - 1]); m_begin = 0; m_count = rows;
        }
        if (m_count <= 0) continue;
        g.sam = L.in_dim; g.sak = 1;
        g.B = params + L.w_off; g.sbk = L.out_dim; g.sbn = 1;
        g.C = part + (int64_t)m_begin * L.out_dim; g.ldc = L.out_dim; g.split_stride = split_stride;
        g.M = m_count; g.N = L.out_dim; g.K = L.in_dim; g.k_per_split = kps;
        g.bias = direct ? params + L.b_off : nullptr;
        int rc = launch_gemm(g, real_splits, s, "dense_fwd_gemm");
        if (rc) return rc;
      }
      if (!direct) {
        ISDQN_PROF(s, "dense_finalize");
        dense_finalize_kernel<<<rows, kRowThreads, 0, s>>>(
            part, n_parts, split_stride, rows, L.out_dim, params + L.b_off, ln_g, ln_b, L.relu, out,
            rows_train > 0 ? wsp(ws, w.xhat[l]) : nullptr, rows_train > 0 ? wsp(ws, w.rstd[l]) : nullptr, rows_train,
            on_tc ? dtc->act16[l] : nullptr);
        ISDQN_LAUNCH_CHECK();
      }
    }
  }
  return ISDQN_OK;
}

// weight gradient of a convolution: deterministic partials over `splits` row ranges; returns the number of partials
// written to `part` ([splits][K][Cout]) or -(error code)
int launch_conv_wgrad_f32(const Layer& L, const void* in, int in_kind, int rows_l, const float* dz, float* part, int splits,
                          cudaStream_t s) {
  ConvWgradArgs a;
  a.in = in;
  a.H = L.H; a.W = L.W; a.Cin = L.Cin; a.OH = L.OH; a.OW = L.OW; a.Cout = L.out_dim;
  a.ksz = L.ksz; a.stride = L.stride; a.pad_y = L.pad_y; a.pad_x = L.pad_x;
  a.M = rows_l; a.K = L.in_dim; a.dz = dz; a.part = part;
  a.rows_per_split = ceil_div(ceil_div(rows_l, splits), kBK) * kBK;
  const int real_splits = ceil_div(rows_l, a.rows_per_split);
  dim3 grid(ceil_div(L.in_dim, 64), ceil_div(L.out_dim, 64), real_splits);
  ISDQN_PROF(s, "conv_wgrad");
  if (in_kind == IN_U8_255) conv_wgrad_kernel<IN_U8_255><<<grid, kGemmThreads, 0, s>>>(a);
  else if (in_kind == IN_F32_255) conv_wgrad_kernel<IN_F32_255><<<grid, kGemmThreads, 0, s>>>(a);
  else conv_wgrad_kernel<IN_F32><<<grid, kGemmThreads, 0, s>>>(a);
  if (cudaGetLastError() != cudaSuccess) return -ISDQN_E_CUDA;
  return real_splits;
}

// input gradient of a convolution: dx [n_img*H*W][Cin] from dz [n_img*OH*OW][Cout]
int launch_conv_dgrad_f32(const Layer& L, int n_img, const float* dz, const float* w, float* dx, cudaStream_t s) {
  ConvDgradArgs a;
  a.H = L.H; a.W = L.W; a.Cin = L.Cin; a.OH = L.OH; a.OW = L.OW; a.Cout = L.out_dim;
  a.ksz = L.ksz; a.stride = L.stride; a.pad_y = L.pad_y; a.pad_x = L.pad_x;
  a.n_img = n_img;
  a.taps = ceil_div(L.ksz, L.stride);
  a.Kd = a.taps * a.taps * L.out_dim;
  a.dz = dz; a.w = w; a.dx = dx;
  const int rows_max = n_img * ceil_div(L.H, L.stride) * ceil_div(L.W, L.stride);
  const int cy = ceil_div(L.Cin, 64), cz = L.stride * L.stride;
  ISDQN_PROF(s, "conv_dgrad");
  if (ceil_div(rows_max, 64) * cy * cz < 2 * kNumSMs) {  // small grid: 32-row tiles
    conv_dgrad_kernel<32, 4><<<dim3(ceil_div(rows_max, 32), cy, cz), kGemmThreads, 0, s>>>(a);
  } else {
    conv_dgrad_kernel<64, 8><<<dim3(ceil_div(rows_max, 64), cy, cz), kGemmThreads, 0, s>>>(a);
  }
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

int check_common(const isdqn_net* net, Plan* p) {
  if (!net) return ISDQN_E_INVALID;
  if (net->n_heads > kMaxHeads || net->n_actions > kMaxActions) return ISDQN_E_TOO_LARGE;
  return build_plan(net, p);
}

int run_loss(const Plan& p, const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, const float* q_all,
             float* dq, float* dbias, int32_t* count, cudaStream_t s) {
  ISDQN_PROF(s, "heads_td_loss");
  heads_td_loss_kernel<<<net->n_heads, kLossThreads, 0, s>>>(q_all, b->d_action, b->d_reward, b->d_terminal, tr->gamma_n,
                                                  tr->batch, tr->batch_global, net->n_heads, net->n_actions,
                                                  tr->d_losses, dq, dbias, count, count ? tr->d_cumulated : nullptr,
                                                  tr->d_is_weights, tr->d_td_abs);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

// d_input (optional): receives dL/d(input of layer 0) [B][in_dim] (the impala torso continues the backward pass from it)
int run_backward(const Plan& p, const Workspace& w, void* ws, const isdqn_train* tr, const isdqn_batch* b, int in_kind,
                 cudaStream_t s, float* d_input = nullptr, const DenseTc* dtc = nullptr) {
  const int B = tr->batch;
  const float* params = tr->d_params;
  float* grads = tr->d_grads;
  SegmentList segs;
  segs.count = 0;
  auto add_seg = [&](const float* src, float* dst, int64_t stride, int n, int parts) {
    Segment& sg = segs.s[segs.count++];
    sg.src = src; sg.dst = dst; sg.stride = stride; sg.n = n; sg.parts = parts; sg.s2d_cout = 0;
  };
  float* dz = wsp(ws, w.dq);  // gradient w.r.t. the pre-activation output of layer l
  for (int l = p.n_layers - 1; l >= 0; --l) {
    const Layer& L = p.L[l];
    const bool first = l == 0;
    const int rows_l = B * L.pix;
    const bool on_tc = dtc != nullptr && l + 1 < p.n_layers && dense_tc_ok(L);  // (its dz16 was written by the step below)
    // ---- weight gradient
    if (on_tc) {
      // dW[in][out] = X^T dz: both operands MN-major ([k = b][in], [k = b][out])
      int rc = isdqn_tc_gemm_bf16(first ? dtc->in16 : dtc->act16[l - 1], L.in_dim, 1, dtc->dz16, L.out_dim, 1, grads + L.w_off,
                                  L.in_dim, L.out_dim, B, 1, s);
      if (rc) return rc;
    } else if (L.type == 1 && B <= 64 && L.out_dim % 4 == 0 && dense_small_on()) {
      ISDQN_PROF(s, "dense_wgrad_small");
      dense_wgrad_small_kernel<<<dim3(ceil_div(L.in_dim, 16), ceil_div(L.out_dim, 512)), 128, 0, s>>>(
          first ? reinterpret_cast<const float*>(b->d_state) : wsp(ws, w.act[l - 1]), dz, grads + L.w_off, B, L.in_dim, L.out_dim);
      ISDQN_LAUNCH_CHECK();
    } else if (L.type == 1) {
      GemmArgs g;  // dW[in][out] = X^T dz : A(m'=in, k'=b) = X[b*in_dim + in]
      g.A = first ? reinterpret_cast<const float*>(b->d_state) : wsp(ws, w.act[l - 1]);
      g.sam = 1; g.sak = L.in_dim;
      g.B = dz; g.sbk = L.out_dim; g.sbn = 1;
      g.C = grads + L.w_off; g.ldc = L.out_dim; g.split_stride = 0;
      g.M = L.in_dim; g.N = L.out_dim; g.K = B; g.k_per_split = ceil_div(B, kBK) * kBK;
      g.bias = nullptr;
      int rc = launch_gemm(g, 1, s, "dense_wgrad_gemm");
      if (rc) return rc;
    } else {
      const int real_splits = launch_conv_wgrad_f32(L, first ? b->d_state : wsp(ws, w.act[l - 1]), first ? in_kind : IN_F32, rows_l,
                                                    dz, wsp(ws, w.wpart[l]), w.wsplits[l], s);
      if (real_splits < 0) return -real_splits;
      float* part_l = wsp(ws, w.wpart[l]);
      add_seg(part_l, grads + L.w_off, (int64_t)L.in_dim * L.out_dim, L.in_dim * L.out_dim, real_splits);
    }
    // ---- bias / LayerNorm parameter gradients come from the column partials of the kernel that produced dz
    if (L.relu) {
      const float* cp = wsp(ws, w.colpart[l]);
      const int64_t st = 3 * (int64_t)L.out_dim;
      add_seg(cp, grads + L.b_off, st, L.out_dim, w.col_ctas[l]);
      if (L.has_ln) {
        add_seg(cp + L.out_dim, grads + L.g_off, st, L.out_dim, w.col_ctas[l]);
        add_seg(cp + 2 * L.out_dim, grads + L.beta_off, st, L.out_dim, w.col_ctas[l]);
      }
    }  // the head layer's bias gradient was written by the loss kernel
    if (first && !d_input) break;
    // ---- input gradient = dL/d(out of layer l-1), then through that layer's ReLU + LayerNorm
    const Layer& P = p.L[first ? 0 : l - 1];
    float* dprev = first ? d_input : wsp(ws, w.dbuf[l & 1]);
    if (on_tc) {
      // dX[b][in] = dz[b][out] W^T: A = dz16 [b][out] and B = W16 [in][out] are both K-major
      int rc = isdqn_tc_gemm_bf16(dtc->dz16, L.out_dim, 0, dtc->shadow + L.w_off, L.out_dim, 0, dprev, B, L.in_dim, L.out_dim, 1, s);
      if (rc) return rc;
    } else if (L.type == 1 && B <= 32 && dense_small_on()) {
      ISDQN_PROF(s, "dense_dgrad_small");
      dense_dgrad_small_kernel<<<ceil_div(L.in_dim, 32), 256, 0, s>>>(dz, params + L.w_off, dprev, B, L.in_dim, L.out_dim);
      ISDQN_LAUNCH_CHECK();
    } else if (L.type == 1) {
      GemmArgs g;  // dX[b][in] = dz[b][out] W^T : B(k=out, n=in) = W[in*out_dim + out]
      g.A = dz; g.sam = L.out_dim; g.sak = 1;
      g.B = params + L.w_off; g.sbk = 1; g.sbn = L.out_dim;
      g.C = dprev; g.ldc = L.in_dim; g.split_stride = 0;
      g.M = B; g.N = L.in_dim; g.K = L.out_dim; g.k_per_split = ceil_div(L.out_dim, kBK) * kBK;
      g.bias = nullptr;
      int rc = launch_gemm(g, 1, s, "dense_dgrad_gemm");
      if (rc) return rc;
    } else {
      int rc = launch_conv_dgrad_f32(L, B, dz, params + L.w_off, dprev, s);
      if (rc) return rc;
    }
    if (first) break;
    {
      const int rows_p = B * P.pix;
      const float* g_ = P.has_ln ? params + P.g_off : nullptr;
      const float* b_ = P.has_ln ? params + P.beta_off : nullptr;
      ISDQN_PROF(s, "ln_relu_bwd");
      __nv_bfloat16* dz16_p = (dtc != nullptr && dense_tc_ok(P)) ? dtc->dz16 : nullptr;
      if (ln_bwd_use_warp(P.out_dim)) {
        ISDQN_CUDA_CHECK(launch_ln_relu_bwd_warp(w.col_ctas[l - 1], s, dprev, wsp(ws, w.xhat[l - 1]), wsp(ws, w.rstd[l - 1]), g_, b_,
                                wsp(ws, w.act[l - 1]), rows_p, P.out_dim, wsp(ws, w.colpart[l - 1]), dz16_p, nullptr));
      } else {
        ln_relu_bwd_block_kernel<<<w.col_ctas[l - 1], kRowThreads, 0, s>>>(
            dprev, wsp(ws, w.xhat[l - 1]), wsp(ws, w.rstd[l - 1]), g_, b_, wsp(ws, w.act[l - 1]), rows_p, P.out_dim,
            wsp(ws, w.colpart[l - 1]), dz16_p);
      }
      ISDQN_LAUNCH_CHECK();
    }
    dz = dprev;
  }
  if (segs.count > 0) {
    ISDQN_PROF(s, "reduce_segments");
    ISDQN_CUDA_CHECK(launch_reduce_segments(segs, s));
  }
  return ISDQN_OK;
}

// ======================================================================================================= impala
// architecture_type == "impala" (slimdqn/networks/architectures/dqn.py:7-36, 77-86), fp32 only:
//   3 x Stack(features[s]):  x = Conv3x3(x); x = max_pool 3x3 / 2 SAME; 2 x { r = x; x = relu(LN(x)); x = relu(Conv3x3(x));
//   x = Conv3x3(x) + r };  then relu(LN(x)), flatten (h, w, c), Dense tail as in the other architectures.
// The torso is walked explicitly below; the Dense tail is an `fc` Plan over the flattened torso output, run by the
// same run_forward / run_loss / run_backward as every other network.
struct ImpalaStack {
  int Cin, C, Hin, Win, H, W, pool_pad_y, pool_pad_x;
  int64_t w[5], b[5], g[2], beta[2];  // Conv_0..4, LayerNorm_0..1 of the Stack (offsets into the flat vector)
};
struct ImpalaPlan {
  ImpalaStack st[3];
  int has_ln;
  int64_t fin_g, fin_beta;  // DQNNet's own LayerNorm_0 behind the stacks
  int64_t tail_base;        // first float of the Dense tail's leaves
  int flat;                 // H*W*C of the last stack = input width of the tail
  isdqn_net tail_net;
  Plan tail;
  isdqn_layout layout;
};

Layer impala_conv_layer(int H, int W, int Cin, int Cout) {
  Layer L = {};
  L.type = 0;
  L.H = H; L.W = W; L.Cin = Cin; L.OH = H; L.OW = W; L.ksz = 3; L.stride = 1; L.pad_y = L.pad_x = 1;  // 'SAME'
  L.in_dim = 9 * Cin; L.out_dim = Cout; L.pix = H * W;
  return L;
}

int build_impala_plan(const isdqn_net* net, ImpalaPlan* ip) {
  if (!net || net->arch != ISDQN_ARCH_IMPALA) return ISDQN_E_INVALID;
  if (net->n_heads < 1 || net->n_actions < 1 || net->n_heads > kMaxHeads || net->n_actions > kMaxActions)
    return net->n_heads < 1 || net->n_actions < 1 ? ISDQN_E_INVALID : ISDQN_E_TOO_LARGE;
  if (net->n_features < 3 || net->n_features > ISDQN_MAX_FEATURES) return ISDQN_E_INVALID;
  if (net->obs_h < 1 || net->obs_w < 1 || net->obs_c < 1) return ISDQN_E_INVALID;
  for (int i = 0; i < net->n_features; ++i)
    if (net->features[i] < 1) return ISDQN_E_INVALID;
  isdqn_layout& lay = ip->layout;
  lay.n_leaves = 0;
  int64_t off = 0;
  bool overflow = false;
  auto leaf = [&](int64_t size) {
    const int64_t o = off;
    if (lay.n_leaves >= ISDQN_MAX_LEAVES) { overflow = true; return o; }
    lay.offset[lay.n_leaves] = o;
    lay.size[lay.n_leaves] = size;
    lay.n_leaves++;
    off = (off + size + 7) & ~(int64_t)7;
    return o;
  };
  ip->has_ln = net->layer_norm ? 1 : 0;
  int h = net->obs_h, w = net->obs_w, c = net->obs_c;
  for (int s = 0; s < 3; ++s) {
    ImpalaStack& S = ip->st[s];
    S.Cin = c; S.C = net->features[s]; S.Hin = h; S.Win = w;
    if (S.C > 256) return ISDQN_E_TOO_LARGE;  // one CTA covers all output channels of a convolution
    same_pad(h, 3, 2, &S.H, &S.pool_pad_y);
    same_pad(w, 3, 2, &S.W, &S.pool_pad_x);
    S.w[0] = leaf((int64_t)9 * c * S.C);
    S.b[0] = leaf(S.C);
    for (int j = 0; j < 2; ++j) {
      S.g[j] = ip->has_ln ? leaf(S.C) : -1;
      S.beta[j] = ip->has_ln ? leaf(S.C) : -1;
      for (int q = 1; q <= 2; ++q) {
        S.w[q + 2 * j] = leaf((int64_t)9 * S.C * S.C);
        S.b[q + 2 * j] = leaf(S.C);
      }
    }
    h = S.H; w = S.W; c = S.C;
  }
  ip->fin_g = ip->has_ln ? leaf(c) : -1;
  ip->fin_beta = ip->has_ln ? leaf(c) : -1;
  ip->tail_base = off;
  ip->flat = h * w * c;
  isdqn_net& t = ip->tail_net;
  t = *net;
  t.arch = ISDQN_ARCH_FC;
  t.obs_h = t.obs_w = 1;
  t.obs_c = ip->flat;
  t.n_features = net->n_features - 3;
  for (int i = 0; i < ISDQN_MAX_FEATURES; ++i) t.features[i] = i < t.n_features ? net->features[i + 3] : 0;
  int rc = build_plan(&t, &ip->tail);
  if (rc) return rc;
  for (int i = 0; i < ip->tail.layout.n_leaves; ++i) {
    if (lay.n_leaves >= ISDQN_MAX_LEAVES) return ISDQN_E_TOO_LARGE;
    lay.offset[lay.n_leaves] = ip->tail_base + ip->tail.layout.offset[i];
    lay.size[lay.n_leaves] = ip->tail.layout.size[i];
    lay.n_leaves++;
  }
  if (overflow) return ISDQN_E_TOO_LARGE;
  lay.total = ip->tail_base + ip->tail.layout.total;
  return ISDQN_OK;
}

// partial-sum scratch of one Stack's backward pass (offsets in floats); re-used stack after stack
struct ImpalaPartials {
  int64_t wpart[5];
  int wsplits[5];
  int64_t colsum[3];  // bias gradients of the convolutions without an activation: Conv_0, Conv_2, Conv_4
  int colsum_ctas[3];
  int64_t colpart[4];  // [ctas][3][C] of: relu behind Conv_1, LayerNorm_0, relu behind Conv_3, LayerNorm_1
  int col_ctas;
  int64_t total;
};
int colsum_ctas_for(int rows) {
  int c = ceil_div(rows, 256);
  if (c > 4 * kNumSMs) c = 4 * kNumSMs;
  return c < 1 ? 1 : c;
}
void impala_partials(const ImpalaStack& S, int B, ImpalaPartials* o) {
  int64_t off = 0;
  auto take = [&](int64_t n) {
    const int64_t r = off;
    off = align4(off + n);
    return r;
  };
  const int rows_a = B * S.Hin * S.Win, rows = B * S.H * S.W;
  for (int i = 0; i < 5; ++i) {
    const int K = 9 * (i == 0 ? S.Cin : S.C);
    o->wsplits[i] = conv_wgrad_splits(i == 0 ? rows_a : rows, K, S.C);
    o->wpart[i] = take((int64_t)o->wsplits[i] * K * S.C);
  }
  for (int i = 0; i < 3; ++i) {
    o->colsum_ctas[i] = colsum_ctas_for(i == 0 ? rows_a : rows);
    o->colsum[i] = take((int64_t)o->colsum_ctas[i] * S.C);
  }
  o->col_ctas = ln_bwd_ctas(rows, S.C);
  for (int i = 0; i < 4; ++i) o->colpart[i] = take((int64_t)o->col_ctas * 3 * S.C);
  o->total = off;
}

struct ImpalaWs {  // offsets in floats; -1 = not allocated
  int64_t x0;  // Conv_0 output of the stack being computed (dead after the pooling)
  struct St {
    int64_t widx;              // uint8 window index of the pooling maximum (training rows)
    int64_t x[3];              // block inputs / outputs: x[0] = pooled, x[1] = after block 0, x[2] = stack output
    int64_t t[2], u[2];        // relu(LN(x[j])), relu(Conv_{1+2j}(t[j]))
    int64_t xhat[2], rstd[2];  // LayerNorm_j statistics of the training rows
  } st[3];
  int64_t tf, xhf, rsf;  // relu(LN_0(stack 2 output)) = tail input, and that LayerNorm's statistics
  int64_t finpart;
  int fin_ctas;
  int64_t tail;  // the tail Plan's Workspace starts here
  Workspace tailw;
  int64_t G, T1, T2, G0, part;
  int64_t total;
};

void carve_impala(const ImpalaPlan& ip, int rows, int B, ImpalaWs* w) {
  int64_t off = 0;
  auto take = [&](int64_t n) {
    const int64_t r = off;
    off = align4(off + n);
    return r;
  };
  int64_t max_x0 = 0, max_g = 0, max_g0 = 0, max_part = 0;
  for (int s = 0; s < 3; ++s) {
    const ImpalaStack& S = ip.st[s];
    const int64_t a = (int64_t)S.Hin * S.Win * S.C, n = (int64_t)S.H * S.W * S.C;
    if (rows * a > max_x0) max_x0 = rows * a;
    if (B * n > max_g) max_g = B * n;
    if (B * a > max_g0) max_g0 = B * a;
    if (B > 0) {
      ImpalaPartials P;
      impala_partials(S, B, &P);
      if (P.total > max_part) max_part = P.total;
    }
  }
  w->x0 = take(max_x0);
  for (int s = 0; s < 3; ++s) {
    const ImpalaStack& S = ip.st[s];
    const int64_t n = (int64_t)S.H * S.W * S.C;
    ImpalaWs::St& O = w->st[s];
    O.widx = B > 0 ? take((B * n + 3) / 4) : -1;
    for (int j = 0; j < 3; ++j) O.x[j] = take(rows * n);
    for (int j = 0; j < 2; ++j) {
      O.t[j] = take(rows * n);
      O.u[j] = take(rows * n);
      O.xhat[j] = (B > 0 && ip.has_ln) ? take(B * n) : -1;
      O.rstd[j] = (B > 0 && ip.has_ln) ? take((int64_t)B * S.H * S.W) : -1;
    }
  }
  const ImpalaStack& L = ip.st[2];
  w->tf = take((int64_t)rows * ip.flat);
  w->xhf = (B > 0 && ip.has_ln) ? take((int64_t)B * ip.flat) : -1;
  w->rsf = (B > 0 && ip.has_ln) ? take((int64_t)B * L.H * L.W) : -1;
  carve_workspace(ip.tail, rows, B, &w->tailw);
  w->tail = take(w->tailw.total);
  if (B > 0) {
    w->G = take(max_g);
    w->T1 = take(max_g);
    w->T2 = take(max_g);
    w->G0 = take(max_g0);
    w->part = take(max_part);
    w->fin_ctas = ln_bwd_ctas(B * L.H * L.W, L.C);
    w->finpart = take((int64_t)w->fin_ctas * 3 * L.C);
  } else {
    w->G = w->T1 = w->T2 = w->G0 = w->part = w->finpart = -1;
    w->fin_ctas = 0;
  }
  w->total = off;
}

// bf16 buffers of the tensor-core mode (offsets in BYTES into isdqn_train.d_workspace_tc)
struct ImpalaTc {
  int64_t x0;  // Conv_0 output of stacks 1, 2 (scratch)
  int64_t c;   // output of a block's second convolution (scratch)
  struct St {
    int64_t t[2], u[2];  // relu(LN(x[j])), relu(Conv_{1+2j}(t[j]))
    int64_t x3;          // bf16 copy of the stack output (input of the next stack's Conv_0); -1 for the last stack
  } st[3];
  int64_t G, T, G0;  // bf16 copies of the gradients the weight / input gradient GEMMs consume
  // first convolution (3 x 3 x obs_c <= 8 input channels) on the tile engine: frames as bf16 integers 0..255 with the
  // channel axis zero-padded to 8 (one 16-byte chunk per pixel), the kernel padded alike, its gradient un-padded at the end
  int first_tc;
  int64_t xpad, wpad, gpad;  // bf16 [rows][H][W][8]; bf16 [3][3][8][C]; fp32 [3][3][8][C]
  // Dense tail: bf16 copies of its input and of the hidden activations, bf16 dz of a hidden layer
  int64_t tf, tact[ISDQN_MAX_FEATURES + 1], tdz;
  int64_t total;
};
Layer impala_first_layer_padded(const ImpalaStack& S) { return impala_conv_layer(S.Hin, S.Win, 8, S.C); }
bool first_conv_tc_on() {  // A/B: ISDQN_IMPALA_FIRST_TC=0 keeps Stack_0/Conv_0 on the CUDA-core kernels
  static const bool on = [] {
    const char* e = getenv("ISDQN_IMPALA_FIRST_TC");
    return !(e && e[0] == '0');
  }();
  return on;
}
void carve_impala_tc(const ImpalaPlan& ip, int rows, int B, ImpalaTc* t) {
  int64_t off = 0;
  auto take = [&](int64_t n_elems) {
    const int64_t r = off;
    off = (off + 2 * n_elems + 255) & ~(int64_t)255;
    return r;
  };
  int64_t max_x0 = 0, max_c = 0, max_g = 0, max_g0 = 0;
  t->first_tc = (ip.st[0].Cin <= 8 && isdqn_tc_conv_ok(impala_first_layer_padded(ip.st[0])) && first_conv_tc_on()) ? 1 : 0;
  for (int s = 0; s < 3; ++s) {
    const ImpalaStack& S = ip.st[s];
    const int64_t a = (int64_t)S.Hin * S.Win * S.C, n = (int64_t)S.H * S.W * S.C;
    const bool a_tc = s > 0 || t->first_tc;
    if (a_tc && rows * a > max_x0) max_x0 = rows * a;
    if (rows * n > max_c) max_c = rows * n;
    if (B * n > max_g) max_g = B * n;
    if (a_tc && B * a > max_g0) max_g0 = B * a;
  }
  t->x0 = take(max_x0);
  t->c = take(max_c);
  for (int s = 0; s < 3; ++s) {
    const ImpalaStack& S = ip.st[s];
    const int64_t n = (int64_t)S.H * S.W * S.C;
    for (int j = 0; j < 2; ++j) {
      t->st[s].t[j] = take(rows * n);
      t->st[s].u[j] = take(rows * n);
    }
    t->st[s].x3 = s < 2 ? take(rows * n) : -1;
  }
  if (B > 0) {
    t->G = take(max_g);
    t->T = take(max_g);
    t->G0 = take(max_g0);
  } else {
    t->G = t->T = t->G0 = -1;
  }
  {
    t->tf = take((int64_t)rows * ip.flat);
    int widest = 8;
    for (int l = 0; l < ISDQN_MAX_FEATURES + 1; ++l) t->tact[l] = -1;
    for (int l = 0; l + 1 < ip.tail.n_layers; ++l) {
      t->tact[l] = take((int64_t)rows * ip.tail.L[l].out_dim);
      if (ip.tail.L[l].out_dim > widest) widest = ip.tail.L[l].out_dim;
    }
    t->tdz = B > 0 ? take((int64_t)B * widest) : -1;
  }
  if (t->first_tc) {
    const ImpalaStack& S = ip.st[0];
    t->xpad = take((int64_t)rows * S.Hin * S.Win * 8);
    t->wpad = take((int64_t)72 * S.C);
    t->gpad = B > 0 ? take((int64_t)2 * 72 * S.C) : -1;  // fp32: twice the bf16 element size
  } else {
    t->xpad = t->wpad = t->gpad = -1;
  }
  t->total = off;
}
inline __nv_bfloat16* w16(void* wt, int64_t off) {
  return off < 0 ? nullptr : reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(wt) + off);
}
// every convolution but Conv_0 of the first stack (3 x 3 x 4 input channels: K = 36) goes to the tile engine
bool impala_tc_eligible(const ImpalaPlan& ip) {
  for (int s = 0; s < 3; ++s) {
    const ImpalaStack& S = ip.st[s];
    if (s > 0 && !isdqn_tc_conv_ok(impala_conv_layer(S.Hin, S.Win, S.Cin, S.C))) return false;
    if (!isdqn_tc_conv_ok(impala_conv_layer(S.H, S.W, S.C, S.C))) return false;
  }
  return true;
}

int impala_conv(const Layer& L, const void* in0, const void* in1, int n0, int rows, int kind, const float* w, const float* bias,
                int relu, const float* residual, float* out, cudaStream_t s) {
  ConvArgs a;
  fill_conv_geom(L, &a);
  a.in0 = in0; a.in1 = in1; a.n_img0 = n0;
  a.M = rows * L.pix;
  a.w = w; a.bias = bias; a.ln_g = nullptr; a.ln_b = nullptr; a.relu = relu;
  a.out = out; a.xhat = nullptr; a.rstd = nullptr; a.m_train = 0;
  a.residual = residual;
  return launch_conv_fwd(a, kind, s);
}

int grid_for(int64_t n) {
  int64_t g = ceil_div<int64_t>(n, 256);
  if (g > 16 * kNumSMs) g = 16 * kNumSMs;
  return g < 1 ? 1 : (int)g;
}

bool tail_tc_on() {  // A/B: ISDQN_IMPALA_TAIL_TC=0 keeps the Dense tail on the CUDA-core GEMMs
  static const bool on = [] {
    const char* e = getenv("ISDQN_IMPALA_TAIL_TC");
    return !(e && e[0] == '0');
  }();
  return on;
}
void fill_dense_tc(const ImpalaPlan& ip, const ImpalaTc& t, void* wt, const __nv_bfloat16* shadow, DenseTc* d) {
  d->in16 = w16(wt, t.tf);
  for (int l = 0; l < ISDQN_MAX_FEATURES + 1; ++l) d->act16[l] = w16(wt, t.tact[l]);
  d->dz16 = w16(wt, t.tdz);
  d->shadow = shadow + ip.tail_base;
}

// t / wt / shadow: tensor-core mode (bf16 activations in wt, bf16 kernels out of the parameter shadow); null = fp32 mode
int impala_forward(const ImpalaPlan& ip, const ImpalaWs& w, void* ws, const float* params, const void* in0, const void* in1,
                   int n0, int rows, int rows_train, int in_kind, cudaStream_t s, const ImpalaTc* t = nullptr, void* wt = nullptr,
                   const __nv_bfloat16* shadow = nullptr) {
  const bool tc = t != nullptr;
  const float* prev = nullptr;
  for (int si = 0; si < 3; ++si) {
    const ImpalaStack& S = ip.st[si];
    const ImpalaWs::St& O = w.st[si];
    Layer La = impala_conv_layer(S.Hin, S.Win, S.Cin, S.C);
    Layer Lb = impala_conv_layer(S.H, S.W, S.C, S.C);
    float* x0 = wsp(ws, w.x0);
    const int64_t n_out = (int64_t)rows * S.H * S.W * S.C;
    uint8_t* widx = reinterpret_cast<uint8_t*>(wsp(ws, O.widx));
    const int n_train = rows_train > 0 ? rows_train : 0;
    int rc;
    if (tc && (si > 0 || t->first_tc)) {
      La.relu = 0; La.has_ln = 0; La.b_off = S.b[0];
      if (si == 0) {
        Layer Lp = impala_first_layer_padded(S);
        Lp.relu = 0; Lp.has_ln = 0; Lp.b_off = S.b[0];
        const int64_t n_pix = (int64_t)rows * S.Hin * S.Win;
        ISDQN_PROF(s, "frames_pad8_bf16");
        ISDQN_CUDA_CHECK(launch_pdl(u8_frames_pad8_bf16_kernel, dim3(grid_for(n_pix)), dim3(256), 0, s,
                                    reinterpret_cast<const uint8_t*>(in0), reinterpret_cast<const uint8_t*>(in1),
                                    (int64_t)n0 * S.Hin * S.Win, n_pix, S.Cin, w16(wt, t->xpad)));
        ISDQN_CUDA_CHECK(launch_pdl(pad_first_kernel_bf16, dim3(ceil_div(72 * S.C, 256)), dim3(256), 0, s, shadow + S.w[0], S.Cin, S.C,
                                    w16(wt, t->wpad)));
        rc = isdqn_tc_conv_fwd(Lp, w16(wt, t->xpad), rows, w16(wt, t->wpad), params, w16(wt, t->x0), s,
                               in_kind == IN_F32 ? 1.0f : 1.0f / 255.0f);
      } else {
        rc = isdqn_tc_conv_fwd(La, w16(wt, t->st[si - 1].x3), rows, shadow + S.w[0], params, w16(wt, t->x0), s);
      }
      if (rc) return rc;
      ISDQN_PROF(s, "maxpool_fwd");
      if (S.C % 8 == 0) {
        ISDQN_CUDA_CHECK(launch_pdl(maxpool3s2_fwd_bf16x8_kernel, dim3(grid_for(n_out / 8)), dim3(256), 0, s, w16(wt, t->x0), rows,
                                    S.Hin, S.Win, S.C, S.H, S.W, S.pool_pad_y, S.pool_pad_x, wsp(ws, O.x[0]), widx, n_train));
      } else {
        ISDQN_CUDA_CHECK(launch_pdl(maxpool3s2_fwd_kernel<__nv_bfloat16>, dim3(grid_for(n_out)), dim3(256), 0, s, w16(wt, t->x0),
                                    rows, S.Hin, S.Win, S.C, S.H, S.W, S.pool_pad_y, S.pool_pad_x, wsp(ws, O.x[0]), widx, n_train));
      }
    } else {
      rc = impala_conv(La, si == 0 ? in0 : prev, si == 0 ? in1 : nullptr, si == 0 ? n0 : rows, rows, si == 0 ? in_kind : IN_F32,
                       params + S.w[0], params + S.b[0], 0, nullptr, x0, s);
      if (rc) return rc;
      ISDQN_PROF(s, "maxpool_fwd");
      ISDQN_CUDA_CHECK(launch_pdl(maxpool3s2_fwd_kernel<float>, dim3(grid_for(n_out)), dim3(256), 0, s, x0, rows, S.Hin, S.Win, S.C,
                                  S.H, S.W, S.pool_pad_y, S.pool_pad_x, wsp(ws, O.x[0]), widx, n_train));
    }
    ISDQN_LAUNCH_CHECK();
    const int prow = rows * S.H * S.W;
    for (int j = 0; j < 2; ++j) {
      ISDQN_PROF(s, "ln_relu_fwd");
      // tensor-core mode, j == 1: the skip connection of block 0 (x[1] = x[0] + bf16 output of Conv_2) is folded into this
      // launch, which reads x[0] and the convolution output and writes x[1] next to its normalised form
      const bool fold = tc && j == 1;
      ISDQN_CUDA_CHECK(launch_ln_relu_fwd_warp(s, wsp(ws, O.x[fold ? 0 : j]), prow, S.C, ip.has_ln ? params + S.g[j] : nullptr,
                                               ip.has_ln ? params + S.beta[j] : nullptr, tc ? nullptr : wsp(ws, O.t[j]),
                                               tc ? w16(wt, t->st[si].t[j]) : nullptr, wsp(ws, O.xhat[j]), wsp(ws, O.rstd[j]),
                                               rows_train * S.H * S.W, fold ? w16(wt, t->c) : nullptr,
                                               fold ? wsp(ws, O.x[1]) : nullptr));
      if (tc) {
        Lb.has_ln = 0;
        Lb.relu = 1; Lb.b_off = S.b[1 + 2 * j];
        rc = isdqn_tc_conv_fwd(Lb, w16(wt, t->st[si].t[j]), rows, shadow + S.w[1 + 2 * j], params, w16(wt, t->st[si].u[j]), s);
        if (rc) return rc;
        Lb.relu = 0; Lb.b_off = S.b[2 + 2 * j];
        rc = isdqn_tc_conv_fwd(Lb, w16(wt, t->st[si].u[j]), rows, shadow + S.w[2 + 2 * j], params, w16(wt, t->c), s);
        if (rc) return rc;
        if (j == 1 && si < 2) {  // stack output: the next stack's first convolution reads the bf16 copy
          ISDQN_PROF(s, "residual_add");
          ISDQN_CUDA_CHECK(launch_pdl(residual_add_fwd_kernel, dim3(grid_for(n_out)), dim3(256), 0, s, wsp(ws, O.x[1]),
                                      w16(wt, t->c), wsp(ws, O.x[2]), w16(wt, t->st[si].x3), n_out));
        }
      } else {
        rc = impala_conv(Lb, wsp(ws, O.t[j]), nullptr, rows, rows, IN_F32, params + S.w[1 + 2 * j], params + S.b[1 + 2 * j], 1,
                         nullptr, wsp(ws, O.u[j]), s);
        if (rc) return rc;
        rc = impala_conv(Lb, wsp(ws, O.u[j]), nullptr, rows, rows, IN_F32, params + S.w[2 + 2 * j], params + S.b[2 + 2 * j], 0,
                         wsp(ws, O.x[j]), wsp(ws, O.x[j + 1]), s);
        if (rc) return rc;
      }
    }
    prev = wsp(ws, O.x[2]);
  }
  const ImpalaStack& L = ip.st[2];
  ISDQN_PROF(s, "ln_relu_fwd");
  ISDQN_CUDA_CHECK(launch_ln_relu_fwd_warp(s, tc ? wsp(ws, w.st[2].x[1]) : prev, rows * L.H * L.W, L.C,
                                           ip.has_ln ? params + ip.fin_g : nullptr, ip.has_ln ? params + ip.fin_beta : nullptr,
                                           wsp(ws, w.tf), tc ? w16(wt, t->tf) : nullptr, wsp(ws, w.xhf), wsp(ws, w.rsf),
                                           rows_train * L.H * L.W, tc ? w16(wt, t->c) : nullptr,
                                           tc ? wsp(ws, w.st[2].x[2]) : nullptr));
  DenseTc dtc;
  if (tc) fill_dense_tc(ip, *t, wt, shadow, &dtc);
  return run_forward(ip.tail, w.tailw, wsp(ws, w.tail), params + ip.tail_base, wsp(ws, w.tf), nullptr, rows, rows, rows_train,
                     IN_F32, s, (tc && tail_tc_on()) ? &dtc : nullptr);
}

int impala_backward(const ImpalaPlan& ip, const ImpalaWs& w, void* ws, const isdqn_train* tr, const isdqn_batch* b, int in_kind,
                    cudaStream_t s, const ImpalaTc* t = nullptr, void* wt = nullptr, const __nv_bfloat16* shadow = nullptr) {
  const bool tc = t != nullptr;
  const int B = tr->batch;
  const float* params = tr->d_params;
  float* grads = tr->d_grads;
  float *G = wsp(ws, w.G), *T1 = wsp(ws, w.T1), *T2 = wsp(ws, w.T2), *G0 = wsp(ws, w.G0), *pb = wsp(ws, w.part);
  __nv_bfloat16 *G16 = tc ? w16(wt, t->G) : nullptr, *T16 = tc ? w16(wt, t->T) : nullptr, *G016 = tc ? w16(wt, t->G0) : nullptr;
  {  // Dense tail; its input gradient is dL/d relu(LN_0(x)) of the last stack
    isdqn_train trt = *tr;
    trt.d_params = tr->d_params + ip.tail_base;
    trt.d_grads = tr->d_grads + ip.tail_base;
    isdqn_batch bt = *b;
    bt.d_state = wsp(ws, w.tf);
    DenseTc dtc;
    if (tc) fill_dense_tc(ip, *t, wt, shadow, &dtc);
    int rc = run_backward(ip.tail, w.tailw, wsp(ws, w.tail), &trt, &bt, IN_F32, s, G, (tc && tail_tc_on()) ? &dtc : nullptr);
    if (rc) return rc;
  }
  SegmentList segs;
  segs.count = 0;
  auto add_seg = [&](const float* src, float* dst, int64_t stride, int n, int parts) {
    Segment& sg = segs.s[segs.count++];
    sg.src = src; sg.dst = dst; sg.stride = stride; sg.n = n; sg.parts = parts; sg.s2d_cout = 0;
  };
  // d (post-ReLU gradient) -> gradient of the LayerNorm input, in place (and / or as bf16); [1] / [2] of the column
  // partials are the LayerNorm scale / bias gradients, [0] (sum of the result) the bias gradient of a convolution right
  // below a ReLU
  auto act_bwd = [&](float* d, const float* xhat, const float* rstd, int64_t g_off, int64_t beta_off, const float* act,
                     const __nv_bfloat16* act16, int rows, int C, float* colpart, int ctas, __nv_bfloat16* dz16,
                     bool store_d) -> int {
    ISDQN_PROF(s, "ln_relu_bwd");
    const bool ln = g_off >= 0;
    ISDQN_CUDA_CHECK(launch_ln_relu_bwd_warp(ctas, s, d, ln ? xhat : nullptr, ln ? rstd : nullptr, ln ? params + g_off : nullptr,
                                             ln ? params + beta_off : nullptr, act, rows, C, colpart, dz16, act16, store_d));
    if (ln) {
      add_seg(colpart + C, grads + g_off, 3 * (int64_t)C, C, ctas);
      add_seg(colpart + 2 * C, grads + beta_off, 3 * (int64_t)C, C, ctas);
    }
    return ISDQN_OK;
  };
  auto colsum = [&](const float* x, int rows, int C, float* part, int ctas, float* dst) -> int {
    ISDQN_PROF(s, "colsum");
    ISDQN_CUDA_CHECK(launch_pdl(colsum_partials_kernel, dim3(ctas), dim3(256), 0, s, x, rows, C, part));
    add_seg(part, dst, C, C, ctas);
    return ISDQN_OK;
  };
  // weight gradient partials + input gradient of one convolution (x / x16: its input, dz / dz16: gradient of its output)
  auto conv_bwd = [&](const Layer& L, const void* x, int x_kind, const __nv_bfloat16* x16, const float* dz,
                      const __nv_bfloat16* dz16, int rows_l, float* part, int splits, int64_t w_off, float* dx, bool on_tc) -> int {
    int sp = 0;
    if (on_tc) {
      int rc = isdqn_tc_conv_wgrad(L, x16, dz16, part, rows_l, splits, &sp, s);
      if (rc) return rc;
    } else {
      sp = launch_conv_wgrad_f32(L, x, x_kind, rows_l, dz, part, splits, s);
      if (sp < 0) return -sp;
    }
    add_seg(part, grads + w_off, (int64_t)L.in_dim * L.out_dim, L.in_dim * L.out_dim, sp);
    if (!dx) return ISDQN_OK;
    if (on_tc) return isdqn_tc_conv_dgrad(L, dz16, shadow + w_off, dx, B, s);
    return launch_conv_dgrad_f32(L, B, dz, params + w_off, dx, s);
  };
  {
    const ImpalaStack& L = ip.st[2];
    int rc = act_bwd(G, wsp(ws, w.xhf), wsp(ws, w.rsf), ip.fin_g, ip.fin_beta, wsp(ws, w.tf), nullptr, B * L.H * L.W, L.C,
                     wsp(ws, w.finpart), w.fin_ctas, G16, true);
    if (rc) return rc;
  }
  for (int si = 2; si >= 0; --si) {
    const ImpalaStack& S = ip.st[si];
    const ImpalaWs::St& O = w.st[si];
    ImpalaPartials P;
    impala_partials(S, B, &P);
    const int rows = B * S.H * S.W, C = S.C;
    const int64_t n = (int64_t)rows * C;
    const Layer Lb = impala_conv_layer(S.H, S.W, C, C);
    for (int j = 1; j >= 0; --j) {
      const int cb = 1 + 2 * j, cc = 2 + 2 * j;
      const __nv_bfloat16* t16 = tc ? w16(wt, t->st[si].t[j]) : nullptr;
      const __nv_bfloat16* u16 = tc ? w16(wt, t->st[si].u[j]) : nullptr;
      // G = dL/d(block output) = dL/d(Conv_cc output) (no activation) and, through the skip, part of dL/d(block input)
      int rc;
      if (si == 2 && j == 1) {  // (the LayerNorm backward that produced this G left its column sums in finpart[.][0])
        add_seg(wsp(ws, w.finpart), grads + S.b[cc], 3 * (int64_t)C, C, w.fin_ctas);
      } else {
        rc = colsum(G, rows, C, pb + P.colsum[1 + j], P.colsum_ctas[1 + j], grads + S.b[cc]);
        if (rc) return rc;
      }
      rc = conv_bwd(Lb, wsp(ws, O.u[j]), IN_F32, u16, G, G16, rows, pb + P.wpart[cc], P.wsplits[cc], S.w[cc], T1, tc);
      if (rc) return rc;
      // through relu(Conv_cb(.)): T1 -> dL/d(Conv_cb output)
      float* cp = pb + P.colpart[2 * j];
      rc = act_bwd(T1, nullptr, nullptr, -1, -1, wsp(ws, O.u[j]), u16, rows, C, cp, P.col_ctas, T16, !tc);
      if (rc) return rc;
      add_seg(cp, grads + S.b[cb], 3 * (int64_t)C, C, P.col_ctas);
      rc = conv_bwd(Lb, wsp(ws, O.t[j]), IN_F32, t16, T1, T16, rows, pb + P.wpart[cb], P.wsplits[cb], S.w[cb], T2, tc);
      if (rc) return rc;
      // through relu(LN_j(.)): T2 -> dL/d(block input) along the convolution branch; the skip branch adds G
      rc = act_bwd(T2, wsp(ws, O.xhat[j]), wsp(ws, O.rstd[j]), S.g[j], S.beta[j], wsp(ws, O.t[j]), t16, rows, C,
                   pb + P.colpart[2 * j + 1], P.col_ctas, nullptr, true);
      if (rc) return rc;
      ISDQN_PROF(s, "residual_add");
      ISDQN_CUDA_CHECK(launch_pdl(add_inplace_kernel, dim3(grid_for(n)), dim3(256), 0, s, G, T2, n, G16));
    }
    const bool a_tc = tc && (si > 0 || t->first_tc);
    ISDQN_PROF(s, "maxpool_bwd");
    {
      const int64_t n_in = (int64_t)B * S.Hin * S.Win * C;
      const uint8_t* widx = reinterpret_cast<const uint8_t*>(wsp(ws, O.widx));
      if (C % 4 == 0) {
        ISDQN_CUDA_CHECK(launch_pdl(maxpool3s2_bwd_x4_kernel, dim3(grid_for(n_in / 4)), dim3(256), 0, s, G, widx, B, S.Hin, S.Win, C,
                                    S.H, S.W, S.pool_pad_y, S.pool_pad_x, G0, a_tc ? G016 : nullptr));
      } else {
        ISDQN_CUDA_CHECK(launch_pdl(maxpool3s2_bwd_kernel, dim3(grid_for(n_in)), dim3(256), 0, s, G, widx, B, S.Hin, S.Win, C, S.H,
                                    S.W, S.pool_pad_y, S.pool_pad_x, G0, a_tc ? G016 : nullptr));
      }
    }
    const Layer La = impala_conv_layer(S.Hin, S.Win, S.Cin, C);
    const int rows_a = B * S.Hin * S.Win;
    int rc = colsum(G0, rows_a, C, pb + P.colsum[0], P.colsum_ctas[0], grads + S.b[0]);
    if (rc) return rc;
    const void* in = si == 0 ? b->d_state : static_cast<const void*>(wsp(ws, w.st[si - 1].x[2]));
    float* gpad = nullptr;
    if (a_tc && si == 0) {
      // padded first convolution: partials [splits][3][3][8][C] -> gpad, un-padded into the gradient after the reduction.
      // (The partial scratch was sized for 9 * Cin rows per split: the padded kernel has 72, so fewer splits fit.)
      const Layer Lp = impala_first_layer_padded(S);
      gpad = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(wt) + t->gpad);
      int splits = (int)(((int64_t)P.wsplits[0] * 9 * S.Cin) / 72);
      if (splits < 1) splits = 1;
      int sp = 0;
      rc = isdqn_tc_conv_wgrad(Lp, w16(wt, t->xpad), G016, pb + P.wpart[0], rows_a, splits, &sp, s,
                               in_kind == IN_F32 ? 1.0f : 1.0f / 255.0f);
      if (rc) return rc;
      add_seg(pb + P.wpart[0], gpad, (int64_t)72 * C, 72 * C, sp);
    } else {
      rc = conv_bwd(La, in, si == 0 ? in_kind : IN_F32, (a_tc && si > 0) ? w16(wt, t->st[si - 1].x3) : nullptr, G0, G016, rows_a,
                    pb + P.wpart[0], P.wsplits[0], S.w[0], si > 0 ? G : nullptr, a_tc);
      if (rc) return rc;
    }
    if (a_tc && si > 0) {  // the input gradient came out in fp32: the next stack's GEMMs want it in bf16 as well
      const ImpalaStack& Pn = ip.st[si - 1];
      rc = isdqn_cast_bf16_launch(G, G16, (int64_t)B * Pn.H * Pn.W * Pn.C, s);
      if (rc) return rc;
    }
    // this stack's partials are folded before the next stack re-uses the scratch
    ISDQN_PROF(s, "reduce_segments");
    ISDQN_CUDA_CHECK(launch_reduce_segments(segs, s));
    segs.count = 0;
    if (gpad) {
      ISDQN_CUDA_CHECK(launch_pdl(unpad_first_kernel_grad, dim3(ceil_div(9 * S.Cin * C, 256)), dim3(256), 0, s, gpad, S.Cin, C,
                                  grads + S.w[0]));
    }
  }
  return ISDQN_OK;
}

int impala_train(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, bool backward, bool update, float* q_out,
                 void* stream) {
  if (!tr || !b) return ISDQN_E_INVALID;
  ImpalaPlan ip;
  int rc = build_impala_plan(net, &ip);
  if (rc) return rc;
  const bool tc = tr->compute_dtype == ISDQN_COMPUTE_BF16;
  if (tc && !impala_tc_eligible(ip)) return ISDQN_E_UNSUPPORTED;
  if (!tr->d_params || !tr->d_losses || !tr->d_workspace || tr->batch < 1 || tr->batch_global < tr->batch) return ISDQN_E_INVALID;
  if (tc && (!tr->d_workspace_tc || !tr->d_params_bf16)) return ISDQN_E_INVALID;
  if (!b->d_state || !b->d_next_state || !b->d_action || !b->d_reward || !b->d_terminal) return ISDQN_E_INVALID;
  if (backward && !tr->d_grads) return ISDQN_E_INVALID;
  if (update && (!tr->d_mu || !tr->d_nu || !tr->d_count)) return ISDQN_E_INVALID;
  const int B = tr->batch;
  ImpalaWs w;
  carve_impala(ip, 2 * B, B, &w);
  if (w.total * (int64_t)sizeof(float) > tr->workspace_bytes) return ISDQN_E_INVALID;
  ImpalaTc t;
  if (tc) {
    carve_impala_tc(ip, 2 * B, B, &t);
    if (t.total > tr->workspace_tc_bytes) return ISDQN_E_INVALID;
  }
  cudaStream_t s = as_stream(stream);
  void* ws = tr->d_workspace;
  void* wt = tc ? tr->d_workspace_tc : nullptr;
  __nv_bfloat16* shadow = tc ? reinterpret_cast<__nv_bfloat16*>(tr->d_params_bf16) : nullptr;
  if (tc && tr->refresh_shadow) {
    rc = isdqn_cast_bf16_launch(tr->d_params, shadow, ip.layout.total, s);
    if (rc) return rc;
  }
  rc = impala_forward(ip, w, ws, tr->d_params, b->d_state, b->d_next_state, B, 2 * B, backward ? B : 0, IN_U8_255, s,
                      tc ? &t : nullptr, wt, shadow);
  if (rc) return rc;
  void* tws = wsp(ws, w.tail);
  const Layer& last = ip.tail.L[ip.tail.n_layers - 1];
  const float* q_all = wsp(tws, w.tailw.act[ip.tail.n_layers - 1]);
  rc = run_loss(ip.tail, net, tr, b, q_all, backward ? wsp(tws, w.tailw.dq) : nullptr,
                backward ? tr->d_grads + ip.tail_base + last.b_off : nullptr, update ? tr->d_count : nullptr, s);
  if (rc) return rc;
  if (q_out)
    ISDQN_CUDA_CHECK(cudaMemcpyAsync(q_out, q_all, sizeof(float) * 2 * (size_t)B * ip.tail.n_out, cudaMemcpyDeviceToDevice, s));
  if (!backward) return ISDQN_OK;
  rc = impala_backward(ip, w, ws, tr, b, IN_U8_255, s, tc ? &t : nullptr, wt, shadow);
  if (rc) return rc;
  if (!update) return ISDQN_OK;
  if (tr->nccl_comm) {
    ISDQN_PROF(s, "nccl_allreduce");
    rc = isdqn_dp_allreduce_f32(tr->nccl_comm, tr->d_grads, ip.layout.total, stream);
    if (rc) return rc;
  }
  return isdqn_adam_launch(tr->d_params, tr->d_grads, tr->d_mu, tr->d_nu, tr->d_count, tr->lr, tr->b1, tr->b2, tr->eps,
                           ip.layout.total, shadow, stream, 0, 0, 0);
}

// forward only (isdqn_forward / isdqn_best_action); returns the device pointer of q [rows][n_out] through q_dev
int impala_infer(const isdqn_net* net, const float* d_params, const void* d_input, int input_is_float, int rows, void* d_workspace,
                 int64_t workspace_bytes, const float** q_dev, int* n_out, cudaStream_t s) {
  ImpalaPlan ip;
  int rc = build_impala_plan(net, &ip);
  if (rc) return rc;
  if (!d_params || !d_input || !d_workspace || rows < 1) return ISDQN_E_INVALID;
  ImpalaWs w;
  carve_impala(ip, rows, 0, &w);
  if (w.total * (int64_t)sizeof(float) > workspace_bytes) return ISDQN_E_INVALID;
  rc = impala_forward(ip, w, d_workspace, d_params, d_input, nullptr, rows, rows, 0, input_is_float ? IN_F32_255 : IN_U8_255, s);
  if (rc) return rc;
  *q_dev = wsp(wsp(d_workspace, w.tail), w.tailw.act[ip.tail.n_layers - 1]);
  *n_out = ip.tail.n_out;
  return ISDQN_OK;
}

int train_common(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, bool backward, bool update,
                 float* q_out, void* stream) {
  if (net && net->arch == ISDQN_ARCH_IMPALA) return impala_train(net, tr, b, backward, update, q_out, stream);
  if (tr && tr->compute_dtype == ISDQN_COMPUTE_BF16) return isdqn_tc_train_dispatch(net, tr, b, backward, update, q_out, stream);
  Plan p;
  int rc = check_common(net, &p);
  if (rc) return rc;
  if (!tr || !b || !tr->d_params || !tr->d_losses || !tr->d_workspace || tr->batch < 1 || tr->batch_global < tr->batch)
    return ISDQN_E_INVALID;
  if (!b->d_state || !b->d_next_state || !b->d_action || !b->d_reward || !b->d_terminal) return ISDQN_E_INVALID;
  if (backward && !tr->d_grads) return ISDQN_E_INVALID;
  if (update && (!tr->d_mu || !tr->d_nu || !tr->d_count)) return ISDQN_E_INVALID;
  const int B = tr->batch;
  Workspace w;
  carve_workspace(p, 2 * B, B, &w);
  if (w.total * (int64_t)sizeof(float) > tr->workspace_bytes) return ISDQN_E_INVALID;
  cudaStream_t s = as_stream(stream);
  const int in_kind = net->arch == ISDQN_ARCH_CNN ? IN_U8_255 : IN_F32;
  rc = run_forward(p, w, tr->d_workspace, tr->d_params, b->d_state, b->d_next_state, B, 2 * B, backward ? B : 0, in_kind, s);
  if (rc) return rc;
  const float* q_all = wsp(tr->d_workspace, w.act[p.n_layers - 1]);
  const Layer& last = p.L[p.n_layers - 1];
  rc = run_loss(p, net, tr, b, q_all, backward ? wsp(tr->d_workspace, w.dq) : nullptr,
                backward ? tr->d_grads + last.b_off : nullptr, update ? tr->d_count : nullptr, s);
  if (rc) return rc;
  if (q_out)
    ISDQN_CUDA_CHECK(cudaMemcpyAsync(q_out, q_all, sizeof(float) * 2 * (size_t)B * p.n_out, cudaMemcpyDeviceToDevice, s));
  if (!backward) return ISDQN_OK;
  rc = run_backward(p, w, tr->d_workspace, tr, b, in_kind, s);
  if (rc) return rc;
  if (!update) return ISDQN_OK;
  if (tr->nccl_comm) {
    ISDQN_PROF(s, "nccl_allreduce");
    rc = isdqn_dp_allreduce_f32(tr->nccl_comm, tr->d_grads, p.layout.total, stream);
    if (rc) return rc;
  }
  return isdqn_adam_step_nocount(tr->d_params, tr->d_grads, tr->d_mu, tr->d_nu, tr->d_count, tr->lr, tr->b1, tr->b2,
                                 tr->eps, p.layout.total, stream);
}

}  // namespace

// bytes of bf16 scratch of the impala torso in tensor-core mode (0: the network is not eligible)
int64_t isdqn_impala_tc_bytes(const isdqn_net* net, int batch) {
  ImpalaPlan ip;
  if (build_impala_plan(net, &ip) || batch < 1) return -1;
  if (!impala_tc_eligible(ip)) return 0;
  ImpalaTc t;
  carve_impala_tc(ip, 2 * batch, batch, &t);
  return t.total;
}

extern "C" int isdqn_net_layout(const isdqn_net* net, isdqn_layout* out) {
  if (!out) return ISDQN_E_INVALID;
  if (net && net->arch == ISDQN_ARCH_IMPALA) {
    ImpalaPlan ip;
    const int rci = build_impala_plan(net, &ip);
    if (rci) return rci;
    *out = ip.layout;
    return ISDQN_OK;
  }
  Plan p;
  int rc = build_plan(net, &p);
  if (rc) return rc;
  *out = p.layout;
  return ISDQN_OK;
}

extern "C" int64_t isdqn_forward_workspace_bytes(const isdqn_net* net, int32_t n_rows) {
  if (net && net->arch == ISDQN_ARCH_IMPALA) {
    ImpalaPlan ip;
    if (build_impala_plan(net, &ip) || n_rows < 1) return -1;
    ImpalaWs w;
    carve_impala(ip, n_rows, 0, &w);
    return w.total * (int64_t)sizeof(float);
  }
  Plan p;
  if (build_plan(net, &p) || n_rows < 1) return -1;
  Workspace w;
  carve_workspace(p, n_rows, 0, &w);
  return w.total * (int64_t)sizeof(float);
}

extern "C" int64_t isdqn_learn_workspace_bytes(const isdqn_net* net, int32_t batch) {
  if (net && net->arch == ISDQN_ARCH_IMPALA) {
    ImpalaPlan ip;
    if (build_impala_plan(net, &ip) || batch < 1) return -1;
    ImpalaWs w;
    carve_impala(ip, 2 * batch, batch, &w);
    return w.total * (int64_t)sizeof(float);
  }
  Plan p;
  if (build_plan(net, &p) || batch < 1) return -1;
  Workspace w;
  carve_workspace(p, 2 * batch, batch, &w);
  return w.total * (int64_t)sizeof(float);
}

extern "C" int isdqn_forward(const isdqn_net* net, const float* d_params, const void* d_input, int32_t input_is_float,
                             int32_t n_rows, float* d_q, void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (net && net->arch == ISDQN_ARCH_IMPALA) {
    if (!d_q) return ISDQN_E_INVALID;
    const float* q = nullptr;
    int n_out = 0;
    const int rci = impala_infer(net, d_params, d_input, input_is_float, n_rows, d_workspace, workspace_bytes, &q, &n_out,
                                 as_stream(stream));
    if (rci) return rci;
    ISDQN_CUDA_CHECK(cudaMemcpyAsync(d_q, q, sizeof(float) * (size_t)n_rows * n_out, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return ISDQN_OK;
  }
  Plan p;
  int rc = check_common(net, &p);
  if (rc) return rc;
  if (!d_params || !d_input || !d_q || !d_workspace || n_rows < 1) return ISDQN_E_INVALID;
  Workspace w;
  carve_workspace(p, n_rows, 0, &w);
  if (w.total * (int64_t)sizeof(float) > workspace_bytes) return ISDQN_E_INVALID;
  cudaStream_t s = as_stream(stream);
  const int in_kind = net->arch == ISDQN_ARCH_CNN ? (input_is_float ? IN_F32_255 : IN_U8_255) : IN_F32;
  rc = run_forward(p, w, d_workspace, d_params, d_input, nullptr, n_rows, n_rows, 0, in_kind, s);
  if (rc) return rc;
  ISDQN_CUDA_CHECK(cudaMemcpyAsync(d_q, wsp(d_workspace, w.act[p.n_layers - 1]), sizeof(float) * (size_t)n_rows * p.n_out,
                                   cudaMemcpyDeviceToDevice, s));
  return ISDQN_OK;
}

extern "C" int isdqn_heads_td_loss_weighted(const float* d_q_all, const int64_t* d_action, const double* d_reward,
                                            const uint8_t* d_terminal, float gamma_n, int32_t batch, int32_t batch_global,
                                            int32_t n_heads, int32_t n_actions, const float* d_is_weights, float* d_losses,
                                            float* d_dq, float* d_td_abs, void* stream) {
  if (!d_q_all || !d_action || !d_reward || !d_terminal || !d_losses || batch < 1 || batch_global < batch ||
      n_heads < 1 || n_actions < 1)
    return ISDQN_E_INVALID;
  if (n_heads > kMaxHeads || n_actions > kMaxActions) return ISDQN_E_TOO_LARGE;
  heads_td_loss_kernel<<<n_heads, kLossThreads, 0, as_stream(stream)>>>(d_q_all, d_action, d_reward, d_terminal, gamma_n, batch,
                                                                  batch_global, n_heads, n_actions, d_losses, d_dq,
                                                                  nullptr, nullptr, nullptr, d_is_weights, d_td_abs);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_heads_td_loss(const float* d_q_all, const int64_t* d_action, const double* d_reward,
                                   const uint8_t* d_terminal, float gamma_n, int32_t batch, int32_t batch_global,
                                   int32_t n_heads, int32_t n_actions, float* d_losses, float* d_dq, void* stream) {
  return isdqn_heads_td_loss_weighted(d_q_all, d_action, d_reward, d_terminal, gamma_n, batch, batch_global, n_heads, n_actions,
                                      nullptr, d_losses, d_dq, nullptr, stream);
}

// Adam over the flat vector; *d_count must already hold the step number t >= 1.  `shadow` (optional): bf16 copy of
// the updated parameters, written in the same pass.
int isdqn_adam_launch(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count, float lr,
                      float b1, float b2, float eps, int64_t n, void* d_shadow_bf16, void* stream, int64_t skip_begin,
                      int64_t skip_len, int max_ctas) {
  if (!d_params || !d_grads || !d_mu || !d_nu || !d_count || n < 0 || (n & 3)) return ISDQN_E_INVALID;
  if (skip_begin < 0 || skip_len < 0 || (skip_begin & 3) || (skip_len & 3) || skip_begin + skip_len > n) return ISDQN_E_INVALID;
  const int64_t n4 = (n - skip_len) / 4;
  if (n4 == 0) return ISDQN_OK;
  int64_t grid = ceil_div<int64_t>(n4, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  ISDQN_PROF(as_stream(stream), "adam");
  co_resident_with_tc(adam_kernel);
  ISDQN_CUDA_CHECK(launch_pdl(adam_kernel, dim3((unsigned)grid), dim3(256), 0, as_stream(stream), d_params, d_grads, d_mu,
                              d_nu, d_count, lr, b1, b2, eps, n4, reinterpret_cast<__nv_bfloat16*>(d_shadow_bf16), skip_begin / 4,
                              skip_len / 4));
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

// Rank-B weight gradient of a Dense kernel [Kin][N] recomputed on the fly + its Adam update (dense_wgrad_adam_kernel).
// act: bf16 [B][lda] (first Kin columns), dz: bf16 [B][N].  The caller checks isdqn_dense_wgrad_adam_ok first and keeps
// the separate weight gradient + Adam when the shape does not qualify.
bool isdqn_dense_wgrad_adam_ok(int B, int Kin, int N, int64_t w_off) {
  static const bool on = [] {
    const char* e = getenv("ISDQN_FUSED_DENSE_ADAM");
    return !(e && e[0] == '0');
  }();
  return on && B >= 1 && B <= kDwaMaxB && Kin % 8 == 0 && N % 8 == 0 && w_off % 8 == 0;
}
int isdqn_dense_wgrad_adam_launch(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count,
                                  float lr, float b1, float b2, float eps, void* d_shadow_bf16, int64_t n_total, int64_t w_off,
                                  const void* d_act_bf16, int64_t lda, const void* d_dz_bf16, int B, int Kin, int N,
                                  void* stream) {
  // the whole flat vector [0, n_total): the Dense kernel at w_off from its recomputed gradient, every other leaf from d_grads
  if (!d_params || !d_grads || !d_mu || !d_nu || !d_count || !d_act_bf16 || !d_dz_bf16 || (n_total & 3)) return ISDQN_E_INVALID;
  if (!isdqn_dense_wgrad_adam_ok(B, Kin, N, w_off) || w_off + (int64_t)Kin * N > n_total) return ISDQN_E_UNSUPPORTED;
  static const int want_ctas = [] {
    const char* e = getenv("ISDQN_DWA_CTAS");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 3 * kNumSMs;
  }();
  const int gy = ceil_div(N, kDwaCols);
  const int n_tiles = ceil_div(Kin, kDwaTileRows);
  int gx = want_ctas / gy;
  if (gx < 1) gx = 1;
  const int tpc = ceil_div(n_tiles, gx);
  gx = ceil_div(n_tiles, tpc);
  const int64_t rest4 = (n_total - (int64_t)Kin * N) / 4;
  int rest_ctas = (int)ceil_div<int64_t>(rest4, 2 * kDwaThreads);
  if (rest_ctas > kNumSMs) rest_ctas = kNumSMs;
  const size_t smem = dwa_smem_bytes(B, tpc);
  if (smem > 100 * 1024) return ISDQN_E_UNSUPPORTED;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    ISDQN_CUDA_CHECK(cudaFuncSetAttribute(dense_wgrad_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  ISDQN_PROF(as_stream(stream), "dense_wgrad_adam");
  co_resident_with_tc(dense_wgrad_adam_kernel);
  ISDQN_CUDA_CHECK(launch_pdl(dense_wgrad_adam_kernel, dim3((unsigned)(rest_ctas + gx), (unsigned)gy), dim3(kDwaThreads), smem,
                              as_stream(stream), d_params, d_grads, d_mu, d_nu, d_count, lr, b1, b2, eps,
                              reinterpret_cast<__nv_bfloat16*>(d_shadow_bf16), w_off, n_total / 4,
                              reinterpret_cast<const __nv_bfloat16*>(d_act_bf16), lda,
                              reinterpret_cast<const __nv_bfloat16*>(d_dz_bf16), B, Kin, N, tpc, rest_ctas));
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_adam_step_nocount(float* d_params, const float* d_grads, float* d_mu, float* d_nu,
                                       const int32_t* d_count, float lr, float b1, float b2, float eps, int64_t n,
                                       void* stream) {
  return isdqn_adam_launch(d_params, d_grads, d_mu, d_nu, d_count, lr, b1, b2, eps, n, nullptr, stream, 0, 0, 0);
}

extern "C" int isdqn_adam_step(float* d_params, const float* d_grads, float* d_mu, float* d_nu, int32_t* d_count,
                               float lr, float b1, float b2, float eps, int64_t n, void* stream) {
  if (!d_count) return ISDQN_E_INVALID;
  count_inc_kernel<<<1, 1, 0, as_stream(stream)>>>(d_count);
  ISDQN_LAUNCH_CHECK();
  return isdqn_adam_step_nocount(d_params, d_grads, d_mu, d_nu, d_count, lr, b1, b2, eps, n, stream);
}

extern "C" int isdqn_shift_heads(float* d_kernel, float* d_bias, int32_t n_in, int32_t n_heads, int32_t n_actions,
                                 void* stream) {
  if (!d_kernel || !d_bias || n_in < 1 || n_heads < 1 || n_actions < 1) return ISDQN_E_INVALID;
  if ((int64_t)n_heads * n_actions > 1024) return ISDQN_E_TOO_LARGE;
  shift_heads_kernel<<<n_in + 1, 256, 0, as_stream(stream)>>>(d_kernel, d_bias, n_in, n_heads, n_actions);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_loss_on_batch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* batch, float* d_q_all,
                                   void* stream) {
  return train_common(net, tr, batch, false, false, d_q_all, stream);
}

extern "C" int isdqn_grad_on_batch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* batch, void* stream) {
  return train_common(net, tr, batch, true, false, nullptr, stream);
}

extern "C" int isdqn_learn_on_batch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* batch, void* stream) {
  return train_common(net, tr, batch, true, true, nullptr, stream);
}

extern "C" int isdqn_best_action(const isdqn_net* net, const float* d_params, const void* d_state, int32_t input_is_float,
                                 int32_t idx_network, int32_t* d_action, void* d_workspace, int64_t workspace_bytes,
                                 void* stream) {
  if (net && net->arch == ISDQN_ARCH_IMPALA) {
    if (!d_action || idx_network < 0 || idx_network >= net->n_heads) return ISDQN_E_INVALID;
    const float* q = nullptr;
    int n_out = 0;
    const int rci = impala_infer(net, d_params, d_state, input_is_float, 1, d_workspace, workspace_bytes, &q, &n_out, as_stream(stream));
    if (rci) return rci;
    argmax_head_kernel<<<1, 32, 0, as_stream(stream)>>>(q, net->n_actions, 1 + idx_network, d_action);
    ISDQN_LAUNCH_CHECK();
    return ISDQN_OK;
  }
  Plan p;
  int rc = check_common(net, &p);
  if (rc) return rc;
  if (!d_params || !d_state || !d_action || !d_workspace || idx_network < 0 || idx_network >= net->n_heads)
    return ISDQN_E_INVALID;
  Workspace w;
  carve_workspace(p, 1, 0, &w);
  if (w.total * (int64_t)sizeof(float) > workspace_bytes) return ISDQN_E_INVALID;
  cudaStream_t s = as_stream(stream);
  const int in_kind = net->arch == ISDQN_ARCH_CNN ? (input_is_float ? IN_F32_255 : IN_U8_255) : IN_F32;
  rc = run_forward(p, w, d_workspace, d_params, d_state, nullptr, 1, 1, 0, in_kind, s);
  if (rc) return rc;
  argmax_head_kernel<<<1, 32, 0, s>>>(wsp(d_workspace, w.act[p.n_layers - 1]), net->n_actions, 1 + idx_network, d_action);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_argmax_heads(const float* d_q, int32_t n_heads_total, int32_t n_actions, int32_t* d_out, void* stream) {
  if (!d_q || !d_out || n_heads_total < 1 || n_actions < 1) return ISDQN_E_INVALID;
  argmax_heads_kernel<<<ceil_div(n_heads_total, 64), 64, 0, as_stream(stream)>>>(d_q, n_heads_total, n_actions, d_out);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

int isdqn_trace_set_learner(unsigned long long* buf) { return isdqn::trace_set_local(buf) == cudaSuccess ? ISDQN_OK : ISDQN_E_CUDA; }
