// Host-side description of the network (layers, parameter layout, workspace layout) shared by every learner
// entry point.  Mirrors DQNNet's construction order (slimdqn/networks/architectures/dqn.py:47-103).
#pragma once
#include "common.cuh"

namespace isdqn {

struct Layer {
  int type;  // 0 = conv, 1 = dense
  // conv geometry (flax padding='SAME': out = ceil(in/s), pad_total = max((out-1)s + k - in, 0), lo = total/2)
  int H, W, Cin, OH, OW, ksz, stride, pad_y, pad_x;
  int in_dim;   // dense: input features; conv: ksz*ksz*Cin (the implicit-GEMM K)
  int out_dim;  // dense: output features; conv: Cout
  int pix;      // rows per sample: OH*OW for conv, 1 for dense
  int has_ln, relu;
  int64_t w_off, b_off, g_off, beta_off;  // offsets into the flat parameter vector (floats); g/beta -1 without LN
};

struct Plan {
  int n_layers;
  Layer L[ISDQN_MAX_FEATURES + 1];
  isdqn_layout layout;
  int n_out;  // (1+K)*A
};

static inline int64_t align4(int64_t x) { return (x + 3) & ~(int64_t)3; }

static inline void same_pad(int in, int k, int s, int* out, int* lo) {
  *out = (in + s - 1) / s;
  int total = (*out - 1) * s + k - in;
  if (total < 0) total = 0;
  *lo = total / 2;
}

// returns ISDQN_OK or an error code
static inline int build_plan(const isdqn_net* net, Plan* p) {
  if (!net || !p) return ISDQN_E_INVALID;
  if (net->n_heads < 1 || net->n_actions < 1) return ISDQN_E_INVALID;
  if (net->n_features < 0 || net->n_features > ISDQN_MAX_FEATURES) return ISDQN_E_INVALID;
  for (int i = 0; i < net->n_features; ++i)
    if (net->features[i] < 1) return ISDQN_E_INVALID;
  p->n_layers = 0;
  p->n_out = (1 + net->n_heads) * net->n_actions;
  isdqn_layout& lay = p->layout;
  lay.n_leaves = 0;
  int64_t off = 0;
  auto leaf = [&](int64_t size) {  // leaves start on 8-element boundaries: 16-byte aligned in fp32 AND in the bf16 shadow
    const int64_t o = off;
    lay.offset[lay.n_leaves] = o;
    lay.size[lay.n_leaves] = size;
    lay.n_leaves++;
    off = (off + size + 7) & ~(int64_t)7;
    return o;
  };
  int start = 0;
  int fan_in = 0;
  if (net->arch == ISDQN_ARCH_CNN) {
    if (net->n_features < 3) return ISDQN_E_INVALID;
    static const int ks[3] = {8, 4, 3}, ss[3] = {4, 2, 1};
    int h = net->obs_h, w = net->obs_w, c = net->obs_c;
    if (h < 1 || w < 1 || c < 1) return ISDQN_E_INVALID;
    for (int i = 0; i < 3; ++i) {
      Layer& L = p->L[p->n_layers++];
      L.type = 0;
      L.H = h; L.W = w; L.Cin = c; L.ksz = ks[i]; L.stride = ss[i];
      same_pad(h, ks[i], ss[i], &L.OH, &L.pad_y);
      same_pad(w, ks[i], ss[i], &L.OW, &L.pad_x);
      L.in_dim = ks[i] * ks[i] * c;
      L.out_dim = net->features[i];
      if (L.out_dim > 256) return ISDQN_E_TOO_LARGE;  // the fused LayerNorm epilogue covers <= 256 channels per CTA
      L.pix = L.OH * L.OW;
      L.has_ln = net->layer_norm ? 1 : 0;
      L.relu = 1;
      L.w_off = leaf((int64_t)L.in_dim * L.out_dim);
      L.b_off = leaf(L.out_dim);
      L.g_off = L.has_ln ? leaf(L.out_dim) : -1;
      L.beta_off = L.has_ln ? leaf(L.out_dim) : -1;
      h = L.OH; w = L.OW; c = L.out_dim;
    }
    fan_in = h * w * c;
    start = 3;
  } else if (net->arch == ISDQN_ARCH_FC) {
    fan_in = net->obs_c;
    if (fan_in < 1) return ISDQN_E_INVALID;
    start = 0;
  } else {
    return ISDQN_E_UNSUPPORTED;  // impala: out of scope
  }
  for (int i = start; i <= net->n_features; ++i) {
    const bool last = i == net->n_features;
    Layer& L = p->L[p->n_layers++];
    L.type = 1;
    L.H = L.W = L.OH = L.OW = 1; L.Cin = fan_in; L.ksz = L.stride = 1; L.pad_y = L.pad_x = 0;
    L.in_dim = fan_in;
    L.out_dim = last ? p->n_out : net->features[i];
    L.pix = 1;
    L.has_ln = (!last && net->layer_norm) ? 1 : 0;
    L.relu = last ? 0 : 1;
    L.w_off = leaf((int64_t)L.in_dim * L.out_dim);
    L.b_off = leaf(L.out_dim);
    L.g_off = L.has_ln ? leaf(L.out_dim) : -1;
    L.beta_off = L.has_ln ? leaf(L.out_dim) : -1;
    fan_in = L.out_dim;
  }
  lay.total = off;
  return ISDQN_OK;
}

// Workspace carving (floats).  `rows` = samples through the forward pass, `B` = samples with a backward pass.
struct Workspace {
  int64_t act[ISDQN_MAX_FEATURES + 1];   // [rows*pix][out_dim] post-activation output of layer l (last = q)
  int64_t xhat[ISDQN_MAX_FEATURES + 1];  // [B*pix][out_dim] normalised pre-affine value (LN layers)
  int64_t rstd[ISDQN_MAX_FEATURES + 1];  // [B*pix]
  int64_t fwd_part;                      // split-K partial sums of the dense forward (reused per layer)
  int64_t dq;                            // [B][n_out]
  int64_t dbuf[2];                       // ping-pong gradient buffers (d_out / dz of the current layer)
  int64_t wpart[ISDQN_MAX_FEATURES + 1]; // split partial sums of the conv weight gradients
  int wsplits_tc[ISDQN_MAX_FEATURES + 1];
  int wsplits[ISDQN_MAX_FEATURES + 1];
  int64_t colpart[ISDQN_MAX_FEATURES + 1];  // [ctas][3][out_dim] column partials of the LN/ReLU backward
  int col_ctas[ISDQN_MAX_FEATURES + 1];
  int64_t total;                         // floats
};

static inline int dense_fwd_splits(int rows, int N, int K) {
  const int tiles = ceil_div(rows, 64) * ceil_div(N, 64);
  int s = (2 * kNumSMs) / tiles;
  const int max_by_k = K / 64 > 0 ? K / 64 : 1;  // at least 64 of K per split
  if (s > max_by_k) s = max_by_k;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

static inline int conv_wgrad_splits(int M, int K, int N) {
  const int tiles = ceil_div(K, 64) * ceil_div(N, 64);
  int s = (2 * kNumSMs) / tiles;
  const int max_by_m = M / 128 > 0 ? M / 128 : 1;
  if (s > max_by_m) s = max_by_m;
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return s;
}
// tensor-core path: 128-row tiles of the K axis, all output channels in one tile => fewer tiles, more splits to fill
// two CTAs per SM (the partial buffer is sized for the larger of the two counts)
static inline int conv_wgrad_splits_tc(int M, int K) {
  int s = (2 * kNumSMs) / ceil_div(K, 128);
  const int max_by_m = M / 128 > 0 ? M / 128 : 1;
  if (s > max_by_m) s = max_by_m;
  if (s > 160) s = 160;
  if (s < 1) s = 1;
  return s;
}

// weight gradient of the (fp32) head layer on the tensor-core path: split the batch axis once it is long
static inline int head_wgrad_splits(int B, int K, int N) {
  const int tiles = ceil_div(K, 64) * ceil_div(N, 64);
  int s = (2 * kNumSMs) / tiles;
  const int max_by_b = B / 128 > 0 ? B / 128 : 1;
  if (s > max_by_b) s = max_by_b;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

constexpr int kColpartFusedCap = 4 * kNumSMs;
// LN/ReLU backward: <= 256 channels -> one warp per row (8 rows per CTA pass); wider -> one CTA per row
static inline bool ln_bwd_use_warp(int C) { return C <= 256; }
static inline int ln_bwd_ctas(int rows, int C) {
  int c = ln_bwd_use_warp(C) ? ceil_div(rows, 64) : rows;
  const int cap = ln_bwd_use_warp(C) ? 8 * kNumSMs : 2 * kNumSMs;  // streaming kernel: many resident warps
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  return c;
}

static inline void carve_workspace(const Plan& p, int rows, int B, Workspace* w) {
  int64_t off = 0;
  auto take = [&](int64_t n) {
    const int64_t o = off;
    off = align4(off + n);
    return o;
  };
  int64_t max_part = 0, max_d = 0;
  for (int l = 0; l < p.n_layers; ++l) {
    const Layer& L = p.L[l];
    w->act[l] = take((int64_t)rows * L.pix * L.out_dim);
    w->xhat[l] = (B > 0 && L.has_ln) ? take((int64_t)B * L.pix * L.out_dim) : -1;
    w->rstd[l] = (B > 0 && L.has_ln) ? take((int64_t)B * L.pix) : -1;
    if (L.type == 1) {
      const int64_t part = (int64_t)dense_fwd_splits(rows, L.out_dim, L.in_dim) * rows * L.out_dim;
      if (part > max_part) max_part = part;
    }
    const int64_t d = (int64_t)B * L.pix * L.out_dim;
    if (d > max_d) max_d = d;
  }
  w->fwd_part = take(max_part);
  if (B > 0) {
    w->dq = take((int64_t)B * p.n_out);
    w->dbuf[0] = take(max_d);
    w->dbuf[1] = take(max_d);
    for (int l = 0; l < p.n_layers; ++l) {
      const Layer& L = p.L[l];
      if (L.type == 0) {
        w->wsplits[l] = conv_wgrad_splits(B * L.pix, L.in_dim, L.out_dim);
        w->wsplits_tc[l] = conv_wgrad_splits_tc(B * L.pix, L.in_dim);
        const int most = w->wsplits[l] > w->wsplits_tc[l] ? w->wsplits[l] : w->wsplits_tc[l];
        w->wpart[l] = take((int64_t)most * L.in_dim * L.out_dim);
      } else if (l == p.n_layers - 1 && head_wgrad_splits(B, L.in_dim, L.out_dim) > 1) {
        w->wsplits[l] = head_wgrad_splits(B, L.in_dim, L.out_dim);  // (used by the tensor-core path only)
        w->wpart[l] = take((int64_t)w->wsplits[l] * L.in_dim * L.out_dim);
      } else {
        w->wsplits[l] = 1;
        w->wpart[l] = -1;
      }
      if (L.relu) {
        w->col_ctas[l] = ln_bwd_ctas(B * L.pix, L.out_dim);
        // (conv layers of <= 64 channels: room for one partial per CTA of the input-gradient launch that fuses this
        // layer's LayerNorm backward, tc_learner.cu)
        const int cap = (L.type == 0 && L.out_dim <= 64 && w->col_ctas[l] < kColpartFusedCap) ? kColpartFusedCap : w->col_ctas[l];
        w->colpart[l] = take((int64_t)cap * 3 * L.out_dim);
      } else {
        w->col_ctas[l] = 0;
        w->colpart[l] = -1;
      }
    }
  } else {
    w->dq = w->dbuf[0] = w->dbuf[1] = -1;
  }
  w->total = off;
}

}  // namespace isdqn
