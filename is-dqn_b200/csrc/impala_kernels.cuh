// Kernels only the `impala` torso needs (slimdqn/networks/architectures/dqn.py:7-36, 77-86): 3x3 stride-2 max pooling
// forward / backward, a stand-alone LayerNorm + ReLU forward (the Stack normalises BEFORE its convolutions), column
// sums (bias gradients of the convolutions that have no activation behind them) and the residual accumulate.
// fp32, NHWC, deterministic (no atomics).  The convolutions, the LayerNorm / ReLU backward and the Dense tail are the
// kernels of learner_kernels.cuh.
#pragma once
#include "learner_kernels.cuh"

namespace isdqn {

// flax nn.max_pool(x, (3, 3), strides=(2, 2), padding="SAME") (dqn.py:23): -inf padding, out = ceil(in / 2).
// widx[img][oy][ox][c] = ky*3 + kx of the FIRST maximum in row-major window order (XLA's select-and-scatter with a
// `>=` select keeps the earlier element on ties), written for the images that get a backward pass.
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TIn>
__global__ void __launch_bounds__(256)
maxpool3s2_fwd_kernel(const TIn* __restrict__ x, int n_img, int H, int W, int C, int OH, int OW, int pad_y, int pad_x,
                      float* __restrict__ y, uint8_t* __restrict__ widx, int n_img_train) {
  pdl_sync();
  const int64_t total = (int64_t)n_img * OH * OW * C;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int img = (int)(p / OH);
    float best = 0.f;
    int bi = -1;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - pad_y + ky;
      if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - pad_x + kx;
        if ((unsigned)ix >= (unsigned)W) continue;
        const float v = to_f32(x[(((int64_t)img * H + iy) * W + ix) * C + c]);
        if (bi < 0 || v > best) {
          best = v;
          bi = ky * 3 + kx;
        }
      }
    }
    y[i] = best;
    if (img < n_img_train) widx[i] = (uint8_t)bi;
  }
}

// gradient of the pooling in gather form: input pixel (iy, ix) collects gy of the (at most 2 x 2) windows whose
// recorded maximum is this pixel, in a fixed order.
static __global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ widx, int n_img, int H, int W, int C, int OH,
                      int OW, int pad_y, int pad_x, float* __restrict__ gx, __nv_bfloat16* __restrict__ gx16) {
  pdl_sync();
  const int64_t total = (int64_t)n_img * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int ix = (int)(p % W);
    p /= W;
    const int iy = (int)(p % H);
    const int img = (int)(p / H);
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = iy + pad_y - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= OH) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = ix + pad_x - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= OW) continue;
        const int64_t o = (((int64_t)img * OH + oy) * OW + ox) * C + c;
        if (widx[o] == ky * 3 + kx) acc += gy[o];
      }
    }
    gx[i] = acc;
    if (gx16) gx16[i] = __float2bfloat16_rn(acc);
  }
}

// The same pooling for bf16 input with C % 8 == 0: one thread per (output pixel, 8 channels), 16-byte loads
static __global__ void __launch_bounds__(256)
maxpool3s2_fwd_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, int n_img, int H, int W, int C, int OH, int OW, int pad_y,
                             int pad_x, float* __restrict__ y, uint8_t* __restrict__ widx, int n_img_train) {
  pdl_sync();
  const int C8 = C >> 3;
  const int64_t total = (int64_t)n_img * OH * OW * C8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c8 = (int)(i % C8);
    int64_t p = i / C8;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const int img = (int)(p / OH);
    float best[8];
    int bi[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { best[q] = 0.f; bi[q] = -1; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - pad_y + ky;
      if ((unsigned)iy >= (unsigned)H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - pad_x + kx;
        if ((unsigned)ix >= (unsigned)W) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(x + (((int64_t)img * H + iy) * W + ix) * C + c8 * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(h[q]);
          if (bi[2 * q] < 0 || f.x > best[2 * q]) { best[2 * q] = f.x; bi[2 * q] = ky * 3 + kx; }
          if (bi[2 * q + 1] < 0 || f.y > best[2 * q + 1]) { best[2 * q + 1] = f.y; bi[2 * q + 1] = ky * 3 + kx; }
        }
      }
    }
    float4* yo = reinterpret_cast<float4*>(y + i * 8);
    yo[0] = make_float4(best[0], best[1], best[2], best[3]);
    yo[1] = make_float4(best[4], best[5], best[6], best[7]);
    if (img < n_img_train) {
      uint2 wv;
      wv.x = (unsigned)bi[0] | ((unsigned)bi[1] << 8) | ((unsigned)bi[2] << 16) | ((unsigned)bi[3] << 24);
      wv.y = (unsigned)bi[4] | ((unsigned)bi[5] << 8) | ((unsigned)bi[6] << 16) | ((unsigned)bi[7] << 24);
      *reinterpret_cast<uint2*>(widx + i * 8) = wv;
    }
  }
}

// gradient of the pooling, C % 4 == 0: one thread per (input pixel, 4 channels)
static __global__ void __launch_bounds__(256)
maxpool3s2_bwd_x4_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ widx, int n_img, int H, int W, int C, int OH,
                         int OW, int pad_y, int pad_x, float* __restrict__ gx, __nv_bfloat16* __restrict__ gx16) {
  pdl_sync();
  const int C4 = C >> 2;
  const int64_t total = (int64_t)n_img * H * W * C4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c4 = (int)(i % C4);
    int64_t p = i / C4;
    const int ix = (int)(p % W);
    p /= W;
    const int iy = (int)(p % H);
    const int img = (int)(p / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = iy + pad_y - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= OH) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = ix + pad_x - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= OW) continue;
        const int64_t o = (((int64_t)img * OH + oy) * OW + ox) * C + c4 * 4;
        const unsigned wv = *reinterpret_cast<const unsigned*>(widx + o);
        const unsigned k = (unsigned)(ky * 3 + kx);
        if (((wv & 0xffu) == k) | (((wv >> 8) & 0xffu) == k) | (((wv >> 16) & 0xffu) == k) | ((wv >> 24) == k)) {
          const float4 g = *reinterpret_cast<const float4*>(gy + o);
          if ((wv & 0xffu) == k) acc.x += g.x;
          if (((wv >> 8) & 0xffu) == k) acc.y += g.y;
          if (((wv >> 16) & 0xffu) == k) acc.z += g.z;
          if ((wv >> 24) == k) acc.w += g.w;
        }
      }
    }
    *reinterpret_cast<float4*>(gx + i * 4) = acc;
    if (gx16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&lo);
      pk.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(gx16 + i * 4) = pk;
    }
  }
}

// t[r][:] = relu(LayerNorm(x[r][:]))  (flax LayerNorm over the last axis, eps 1e-6; ln_g == null: relu only).
// One warp per row, C <= 32 * MAXJ.  The normalised values / reciprocal deviations of the first rows_train rows are kept
// for the backward pass.
template <int MAXJ>
__global__ void __launch_bounds__(256)
ln_relu_fwd_warp_kernel(const float* __restrict__ x, int rows, int C, const float* __restrict__ ln_g,
                        const float* __restrict__ ln_b, float* __restrict__ t, __nv_bfloat16* __restrict__ t16,
                        float* __restrict__ xhat, float* __restrict__ rstd, int rows_train,
                        const __nv_bfloat16* __restrict__ res16, float* __restrict__ x_out) {
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gam[MAXJ], bet[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const int n = lane + 32 * j;
    gam[j] = (ln_g && n < C) ? ln_g[n] : 0.f;
    bet[j] = (ln_g && n < C) ? ln_b[n] : 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    float v[MAXJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int n = lane + 32 * j;
      v[j] = n < C ? x[(int64_t)r * C + n] : 0.f;
      if (res16 && n < C) {  // the block's skip connection: x + (bf16 output of its second convolution), kept as x_out
        v[j] += __bfloat162float(res16[(int64_t)r * C + n]);
        x_out[(int64_t)r * C + n] = v[j];
      }
      s += v[j];
    }
    float rs = 0.f;
    if (ln_g) {
      const float mean = warp_sum(s) * inv_c;
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int n = lane + 32 * j;
        v[j] = n < C ? v[j] - mean : 0.f;
        s2 += v[j] * v[j];
      }
      rs = rsqrtf(warp_sum(s2) * inv_c + kLnEps);
    }
    const bool save = ln_g && xhat && r < rows_train;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int n = lane + 32 * j;
      if (n < C) {
        float y = v[j];
        if (ln_g) {
          const float xh = v[j] * rs;
          if (save) xhat[(int64_t)r * C + n] = xh;
          y = xh * gam[j] + bet[j];
        }
        y = fmaxf(y, 0.f);
        if (t) t[(int64_t)r * C + n] = y;
        if (t16) t16[(int64_t)r * C + n] = __float2bfloat16_rn(y);
      }
    }
    if (save && lane == 0) rstd[r] = rs;
  }
}

static inline cudaError_t launch_ln_relu_fwd_warp(cudaStream_t s, const float* x, int rows, int C, const float* ln_g,
                                                  const float* ln_b, float* t, __nv_bfloat16* t16, float* xhat, float* rstd,
                                                  int rows_train, const __nv_bfloat16* res16 = nullptr, float* x_out = nullptr) {
  int ctas = ceil_div(rows, 8);
  if (ctas > 16 * kNumSMs) ctas = 16 * kNumSMs;
#define ISDQN_LN_FWD(MAXJ)                                                                                             \
  return launch_pdl(ln_relu_fwd_warp_kernel<MAXJ>, dim3(ctas), dim3(256), 0, s, x, rows, C, ln_g, ln_b, t, t16, xhat, rstd, \
                    rows_train, res16, x_out)
  if (C <= 32) ISDQN_LN_FWD(1);
  if (C <= 64) ISDQN_LN_FWD(2);
  if (C <= 128) ISDQN_LN_FWD(4);
  ISDQN_LN_FWD(8);
#undef ISDQN_LN_FWD
}

// part[cta][n] = sum over the CTA's rows of x[r][n]  (C <= 256; rows dealt round-robin, fixed order => deterministic)
static __global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ part) {
  pdl_sync();
  __shared__ float sm[256];
  int cw = 32;  // columns per row group: C rounded up to a power of two (<= 256)
  while (cw < C) cw <<= 1;
  const int groups = 256 / cw;
  const int col = threadIdx.x % cw, grp = threadIdx.x / cw;
  float acc = 0.f;
  if (col < C) {
    // four rows in flight per thread (the loop is a chain of L2 / HBM round trips otherwise); the order of the additions is
    // fixed by (grid, rows) alone
    const int step = gridDim.x * groups;
    int r = blockIdx.x * groups + grp;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (; r + 3 * step < rows; r += 4 * step) {
      const float v0 = x[(int64_t)r * C + col], v1 = x[(int64_t)(r + step) * C + col];
      const float v2 = x[(int64_t)(r + 2 * step) * C + col], v3 = x[(int64_t)(r + 3 * step) * C + col];
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; r < rows; r += step) a0 += x[(int64_t)r * C + col];
    acc = (a0 + a1) + (a2 + a3);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += sm[g * cw + threadIdx.x];
    part[(int64_t)blockIdx.x * C + threadIdx.x] = t;
  }
}

// First convolution of the first Stack on the tile engine: uint8 frames [n][H][W][C <= 8] of the two batch halves -> bf16
// [n][H][W][8] holding the exact integers 0..255 (the 1/255 is applied to the fp32 accumulator), channels C..7 zero, so
// that one pixel is one 16-byte chunk of the implicit-GEMM gather.
static __global__ void __launch_bounds__(256)
u8_frames_pad8_bf16_kernel(const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1, int64_t n_pix0, int64_t n_pix, int C,
                           __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_pix; i += (int64_t)gridDim.x * 256) {
    const uint8_t* src = i < n_pix0 ? in0 + i * C : in1 + (i - n_pix0) * C;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __float2bfloat16_rn(c < C ? (float)src[c] : 0.f);
    *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(v);
  }
}
// bf16 kernel [3][3][C][Cout] -> [3][3][8][Cout] with zero rows for the padded channels
static __global__ void pad_first_kernel_bf16(const __nv_bfloat16* __restrict__ w, int C, int Cout, __nv_bfloat16* __restrict__ wp) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 72 * Cout) return;
  const int co = i % Cout, r = i / Cout, c = r % 8, tap = r / 8;
  wp[i] = c < C ? w[(tap * C + c) * Cout + co] : __float2bfloat16_rn(0.f);
}
// gradient of the padded kernel [3][3][8][Cout] (fp32) -> the real kernel's gradient [3][3][C][Cout]
static __global__ void unpad_first_kernel_grad(const float* __restrict__ gp, int C, int Cout, float* __restrict__ g) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C * Cout) return;
  const int co = i % Cout, r = i / Cout, c = r % C, tap = r / C;
  g[i] = gp[(tap * 8 + c) * Cout + co];
}

// dst += src  (the two branches of a residual connection meet here in the backward pass); optional bf16 copy of the sum
static __global__ void __launch_bounds__(256)
add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n, __nv_bfloat16* __restrict__ dst16) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float v = dst[i] + src[i];
    dst[i] = v;
    if (dst16) dst16[i] = __float2bfloat16_rn(v);
  }
}

// tensor-core path, forward: out = x + c (c = bf16 output of the block's second convolution); optional bf16 copy
static __global__ void __launch_bounds__(256)
residual_add_fwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ c, float* __restrict__ out,
                        __nv_bfloat16* __restrict__ out16, int64_t n) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float v = x[i] + __bfloat162float(c[i]);
    out[i] = v;
    if (out16) out16[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace isdqn
