// Learner entry points (fp32 path): forward, K-head TD loss, backward, Adam — host orchestration of the kernels
// in learner_kernels.cuh.  Replaces the jitted body of iSDQN.learn_on_batch / loss_on_batch / best_action
// (slimdqn/networks/isdqn.py:82-135) and DQNNet.__call__ (slimdqn/networks/architectures/dqn.py:47-103).
#include "learner_kernels.cuh"
#include "plan.cuh"

using namespace isdqn;

extern "C" int isdqn_adam_step_nocount(float* d_params, const float* d_grads, float* d_mu, float* d_nu,
                                       const int32_t* d_count, float lr, float b1, float b2, float eps, int64_t n,
                                       void* stream);

int isdqn_tc_train_dispatch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, bool backward, bool update,
                            float* q_out, void* stream);

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

// Forward through every layer.  in0/in1: the two halves of concat(s, s') (in1 may be null, n0 = rows then).
int run_forward(const Plan& p, const Workspace& w, void* ws, const float* params, const void* in0, const void* in1,
                int n0, int rows, int rows_train, int in_kind, cudaStream_t s) {
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
      const bool direct = real_splits == 1 && !L.has_ln && !L.relu;
      float* part = direct ? out : wsp(ws, w.fwd_part);
      const int64_t split_stride = (int64_t)rows * L.out_dim;
      for (int half = 0; half < 2; ++half) {
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
            part, real_splits, split_stride, rows, L.out_dim, params + L.b_off, ln_g, ln_b, L.relu, out,
            rows_train > 0 ? wsp(ws, w.xhat[l]) : nullptr, rows_train > 0 ? wsp(ws, w.rstd[l]) : nullptr, rows_train);
        ISDQN_LAUNCH_CHECK();
      }
    }
  }
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

int run_backward(const Plan& p, const Workspace& w, void* ws, const isdqn_train* tr, const isdqn_batch* b, int in_kind,
                 cudaStream_t s) {
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
    // ---- weight gradient
    if (L.type == 1) {
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
      ConvWgradArgs a;
      a.in = first ? b->d_state : wsp(ws, w.act[l - 1]);
      a.H = L.H; a.W = L.W; a.Cin = L.Cin; a.OH = L.OH; a.OW = L.OW; a.Cout = L.out_dim;
      a.ksz = L.ksz; a.stride = L.stride; a.pad_y = L.pad_y; a.pad_x = L.pad_x;
      a.M = rows_l; a.K = L.in_dim; a.dz = dz; a.part = wsp(ws, w.wpart[l]);
      const int splits = w.wsplits[l];
      a.rows_per_split = ceil_div(ceil_div(rows_l, splits), kBK) * kBK;
      const int real_splits = ceil_div(rows_l, a.rows_per_split);
      dim3 grid(ceil_div(L.in_dim, 64), ceil_div(L.out_dim, 64), real_splits);
      ISDQN_PROF(s, "conv_wgrad");
      if (first) {
        if (in_kind == IN_U8_255) conv_wgrad_kernel<IN_U8_255><<<grid, kGemmThreads, 0, s>>>(a);
        else if (in_kind == IN_F32_255) conv_wgrad_kernel<IN_F32_255><<<grid, kGemmThreads, 0, s>>>(a);
        else conv_wgrad_kernel<IN_F32><<<grid, kGemmThreads, 0, s>>>(a);
      } else {
        conv_wgrad_kernel<IN_F32><<<grid, kGemmThreads, 0, s>>>(a);
      }
      ISDQN_LAUNCH_CHECK();
      add_seg(a.part, grads + L.w_off, (int64_t)L.in_dim * L.out_dim, L.in_dim * L.out_dim, real_splits);
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
    if (first) break;
    // ---- input gradient = dL/d(out of layer l-1), then through that layer's ReLU + LayerNorm
    const Layer& P = p.L[l - 1];
    float* dprev = wsp(ws, w.dbuf[l & 1]);
    if (L.type == 1) {
      GemmArgs g;  // dX[b][in] = dz[b][out] W^T : B(k=out, n=in) = W[in*out_dim + out]
      g.A = dz; g.sam = L.out_dim; g.sak = 1;
      g.B = params + L.w_off; g.sbk = 1; g.sbn = L.out_dim;
      g.C = dprev; g.ldc = L.in_dim; g.split_stride = 0;
      g.M = B; g.N = L.in_dim; g.K = L.out_dim; g.k_per_split = ceil_div(L.out_dim, kBK) * kBK;
      g.bias = nullptr;
      int rc = launch_gemm(g, 1, s, "dense_dgrad_gemm");
      if (rc) return rc;
    } else {
      ConvDgradArgs a;
      a.H = L.H; a.W = L.W; a.Cin = L.Cin; a.OH = L.OH; a.OW = L.OW; a.Cout = L.out_dim;
      a.ksz = L.ksz; a.stride = L.stride; a.pad_y = L.pad_y; a.pad_x = L.pad_x;
      a.n_img = B;
      a.taps = ceil_div(L.ksz, L.stride);
      a.Kd = a.taps * a.taps * L.out_dim;
      a.dz = dz; a.w = params + L.w_off; a.dx = dprev;
      const int rows_max = B * ceil_div(L.H, L.stride) * ceil_div(L.W, L.stride);
      dim3 grid(ceil_div(rows_max, 64), ceil_div(L.Cin, 64), L.stride * L.stride);
      ISDQN_PROF(s, "conv_dgrad");
      conv_dgrad_kernel<<<grid, kGemmThreads, 0, s>>>(a);
      ISDQN_LAUNCH_CHECK();
    }
    {
      const int rows_p = B * P.pix;
      const float* g_ = P.has_ln ? params + P.g_off : nullptr;
      const float* b_ = P.has_ln ? params + P.beta_off : nullptr;
      ISDQN_PROF(s, "ln_relu_bwd");
      if (ln_bwd_use_warp(P.out_dim)) {
        ISDQN_CUDA_CHECK(launch_ln_relu_bwd_warp(w.col_ctas[l - 1], s, dprev, wsp(ws, w.xhat[l - 1]), wsp(ws, w.rstd[l - 1]), g_, b_,
                                wsp(ws, w.act[l - 1]), rows_p, P.out_dim, wsp(ws, w.colpart[l - 1]), nullptr, nullptr));
      } else {
        ln_relu_bwd_block_kernel<<<w.col_ctas[l - 1], kRowThreads, 0, s>>>(
            dprev, wsp(ws, w.xhat[l - 1]), wsp(ws, w.rstd[l - 1]), g_, b_, wsp(ws, w.act[l - 1]), rows_p, P.out_dim,
            wsp(ws, w.colpart[l - 1]));
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

int train_common(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, bool backward, bool update,
                 float* q_out, void* stream) {
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

extern "C" int isdqn_net_layout(const isdqn_net* net, isdqn_layout* out) {
  if (!out) return ISDQN_E_INVALID;
  Plan p;
  int rc = build_plan(net, &p);
  if (rc) return rc;
  *out = p.layout;
  return ISDQN_OK;
}

extern "C" int64_t isdqn_forward_workspace_bytes(const isdqn_net* net, int32_t n_rows) {
  Plan p;
  if (build_plan(net, &p) || n_rows < 1) return -1;
  Workspace w;
  carve_workspace(p, n_rows, 0, &w);
  return w.total * (int64_t)sizeof(float);
}

extern "C" int64_t isdqn_learn_workspace_bytes(const isdqn_net* net, int32_t batch) {
  Plan p;
  if (build_plan(net, &p) || batch < 1) return -1;
  Workspace w;
  carve_workspace(p, 2 * batch, batch, &w);
  return w.total * (int64_t)sizeof(float);
}

extern "C" int isdqn_forward(const isdqn_net* net, const float* d_params, const void* d_input, int32_t input_is_float,
                             int32_t n_rows, float* d_q, void* d_workspace, int64_t workspace_bytes, void* stream) {
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
