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

// t[r][:] = relu(LayerNorm(x[r][:]))  (flax LayerNorm over the last axis, eps 1e-6; ln_g == null: relu only).
// One warp per row, C <= 32 * MAXJ.  The normalised values / reciprocal deviations of the first rows_train rows are kept
// for the backward pass.
template <int MAXJ>
__global__ void __launch_bounds__(256)
ln_relu_fwd_warp_kernel(const float* __restrict__ x, int rows, int C, const float* __restrict__ ln_g,
                        const float* __restrict__ ln_b, float* __restrict__ t, __nv_bfloat16* __restrict__ t16,
                        float* __restrict__ xhat, float* __restrict__ rstd, int rows_train) {
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
                                                  int rows_train) {
  int ctas = ceil_div(rows, 8);
  if (ctas > 16 * kNumSMs) ctas = 16 * kNumSMs;
  if (C <= 32) ln_relu_fwd_warp_kernel<1><<<ctas, 256, 0, s>>>(x, rows, C, ln_g, ln_b, t, t16, xhat, rstd, rows_train);
  else if (C <= 64) ln_relu_fwd_warp_kernel<2><<<ctas, 256, 0, s>>>(x, rows, C, ln_g, ln_b, t, t16, xhat, rstd, rows_train);
  else if (C <= 128) ln_relu_fwd_warp_kernel<4><<<ctas, 256, 0, s>>>(x, rows, C, ln_g, ln_b, t, t16, xhat, rstd, rows_train);
  else ln_relu_fwd_warp_kernel<8><<<ctas, 256, 0, s>>>(x, rows, C, ln_g, ln_b, t, t16, xhat, rstd, rows_train);
  return cudaGetLastError();
}

// part[cta][n] = sum over the CTA's rows of x[r][n]  (C <= 256; rows dealt round-robin, fixed order => deterministic)
static __global__ void __launch_bounds__(256)
colsum_partials_kernel(const float* __restrict__ x, int rows, int C, float* __restrict__ part) {
  __shared__ float sm[256];
  int cw = 32;  // columns per row group: C rounded up to a power of two (<= 256)
  while (cw < C) cw <<= 1;
  const int groups = 256 / cw;
  const int col = threadIdx.x % cw, grp = threadIdx.x / cw;
  float acc = 0.f;
  if (col < C)
    for (int r = blockIdx.x * groups + grp; r < rows; r += gridDim.x * groups) acc += x[(int64_t)r * C + col];
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += sm[g * cw + threadIdx.x];
    part[(int64_t)blockIdx.x * C + threadIdx.x] = t;
  }
}

// dst += src  (the two branches of a residual connection meet here in the backward pass); optional bf16 copy of the sum
static __global__ void __launch_bounds__(256)
add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n, __nv_bfloat16* __restrict__ dst16) {
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
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float v = x[i] + __bfloat162float(c[i]);
    out[i] = v;
    if (out16) out16[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace isdqn
