// fp32 learner kernels (CUDA-core FFMA, fp32 accumulate): the 1e-5 parity path of the iS-DQN learner.
//
//   conv_fwd_kernel    implicit GEMM (im2col gathered on the fly), one CTA covers ALL output channels of its
//                      rows, so bias + LayerNorm-over-channels + ReLU (architectures/dqn.py:55-72) run in the
//                      epilogue out of registers; also emits the normalised value and 1/std for the rows that
//                      get a backward pass.
//   conv_wgrad_kernel  dW[K][Cout] = im2col(X)^T dZ, split over the rows, deterministic partial sums.
//   conv_dgrad_kernel  dX = dZ (*) W^T as a gather (no atomics), stride-s convs split into s*s parity classes so
//                      only the taps that really hit an input pixel are multiplied.
//   gemm_strided_kernel  Dense forward / wgrad / dgrad with arbitrary operand strides and split-K.
//   dense_finalize_kernel, ln_relu_bwd_*_kernel, reduce_segments_kernel, heads_td_loss_kernel, adam_kernel.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace isdqn {

constexpr int kGemmThreads = 128;
constexpr int kTM = 4;
constexpr int kBK = 16;
constexpr float kLnEps = 1e-6f;  // flax nn.LayerNorm default

template <int BM, int BN, int TN>
struct TileCfg {
  static constexpr int LANES_N = BN / TN;
  static constexpr int ROWS_T = kGemmThreads / LANES_N;
  static_assert(ROWS_T * kTM == BM, "tile shape must use exactly 128 threads");
  static_assert(TN % 4 == 0, "TN must be a multiple of 4");
};

template <int BM, int BN, int TN>
__device__ __forceinline__ void tile_fma(const float (*As)[BM + 4], const float (*Bs)[BN + 4], float (&acc)[kTM][TN],
                                         int ty, int tx) {
#pragma unroll
  for (int kk = 0; kk < kBK; ++kk) {
    const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * kTM]);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
    float b[TN];
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
      b[j] = b4.x; b[j + 1] = b4.y; b[j + 2] = b4.z; b[j + 3] = b4.w;
    }
#pragma unroll
    for (int i = 0; i < kTM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------------------------------ input readers
enum { IN_U8_255 = 0, IN_F32 = 1, IN_F32_255 = 2 };

template <int KIND>
__device__ __forceinline__ float read_in(const void* p, int64_t idx) {
  if (KIND == IN_U8_255) return __fdiv_rn((float)reinterpret_cast<const uint8_t*>(p)[idx], 255.0f);
  if (KIND == IN_F32_255) return __fdiv_rn(reinterpret_cast<const float*>(p)[idx], 255.0f);
  return reinterpret_cast<const float*>(p)[idx];
}

struct ConvArgs {
  const void* in0;  // images [0, n_img0)
  const void* in1;  // images [n_img0, ...) (the s' half of concat(s, s'), isdqn.py:95); may be null
  int n_img0;
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x;
  int M, K;  // M = n_img*OH*OW rows, K = ksz*ksz*Cin
  const float* w;     // [K][Cout]  (HWIO flattened)
  const float* bias;  // [Cout]
  const float* ln_g;  // [Cout] or null
  const float* ln_b;
  int relu;
  float* out;   // [M][Cout]
  float* xhat;  // [m_train][Cout] or null
  float* rstd;  // [m_train] or null
  int m_train;
  const float* residual = nullptr;  // [M][Cout] added to conv + bias (impala block output, dqn.py:34); only without LayerNorm
};

// ----------------------------------------------------------------------------------------------- conv fwd
template <int BM, int BN, int TN, int KIND>
__global__ void __launch_bounds__(kGemmThreads) conv_fwd_kernel(const ConvArgs a) {
  typedef TileCfg<BM, BN, TN> Cfg;
  __shared__ __align__(16) float As[kBK][BM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  __shared__ int ri_img[BM], ri_iy0[BM], ri_ix0[BM];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int tx = tid % Cfg::LANES_N, ty = tid / Cfg::LANES_N;

  for (int r = tid; r < BM; r += kGemmThreads) {
    const int m = m0 + r;
    if (m < a.M) {
      const int img = m / (a.OH * a.OW);
      const int rem = m - img * (a.OH * a.OW);
      const int oy = rem / a.OW, ox = rem - oy * a.OW;
      ri_img[r] = img;
      ri_iy0[r] = oy * a.stride - a.pad_y;
      ri_ix0[r] = ox * a.stride - a.pad_x;
    } else {
      ri_img[r] = -1;
      ri_iy0[r] = ri_ix0[r] = 0;
    }
  }
  __syncthreads();

  float acc[kTM][TN];
#pragma unroll
  for (int i = 0; i < kTM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int a_kk = tid % kBK, a_r0 = tid / kBK;
  constexpr int A_ROWS_PER_PASS = kGemmThreads / kBK;
  constexpr int NA = BM / A_ROWS_PER_PASS, NB = kBK * BN / kGemmThreads;
  // software pipeline: the global loads of chunk k0 + kBK are in registers while chunk k0 is multiplied (at batch 32 the
  // grids are 1-3 CTAs per SM: nothing else hides the load latency).  Same loads, same order of additions.
  float ra[NA], rb[NB];
  auto fetch = [&](int k0) {
    {  // A tile: im2col gather, k (= ky,kx,c with c fastest: contiguous in NHWC) fastest across lanes
      const int k = k0 + a_kk;
      const bool kvalid = k < a.K;
      const int c = k % a.Cin;
      const int t = k / a.Cin;
      const int kx = t % a.ksz, ky = t / a.ksz;
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const int r = a_r0 + q * A_ROWS_PER_PASS;
        float v = 0.f;
        const int img = ri_img[r];
        const int iy = ri_iy0[r] + ky, ix = ri_ix0[r] + kx;
        if (kvalid && img >= 0 && (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W) {
          const bool second = img >= a.n_img0;
          const void* src = second ? a.in1 : a.in0;
          const int li = second ? img - a.n_img0 : img;
          v = read_in<KIND>(src, (((int64_t)li * a.H + iy) * a.W + ix) * a.Cin + c);
        }
        ra[q] = v;
      }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {  // B tile: weights, n fastest
      const int i = tid + q * kGemmThreads;
      const int nn = i % BN, kk = i / BN;
      const int k = k0 + kk;
      rb[q] = (k < a.K && nn < a.Cout) ? __ldg(a.w + (int64_t)k * a.Cout + nn) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < a.K; k0 += kBK) {
#pragma unroll
    for (int q = 0; q < NA; ++q) As[a_kk][a_r0 + q * A_ROWS_PER_PASS] = ra[q];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int i = tid + q * kGemmThreads;
      Bs[i / BN][i % BN] = rb[q];
    }
    __syncthreads();
    if (k0 + kBK < a.K) fetch(k0 + kBK);
    tile_fma<BM, BN, TN>(As, Bs, acc, ty, tx);
    __syncthreads();
  }

  // epilogue: bias -> LayerNorm over the channel axis -> ReLU, all out of registers
  float bias[TN], g[TN], be[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = tx * TN + j;
    const bool nv = n < a.Cout;
    bias[j] = nv ? a.bias[n] : 0.f;
    g[j] = (nv && a.ln_g) ? a.ln_g[n] : 0.f;
    be[j] = (nv && a.ln_g) ? a.ln_b[n] : 0.f;
  }
  const float inv_c = 1.0f / (float)a.Cout;
#pragma unroll
  for (int i = 0; i < kTM; ++i) {
    const int m = m0 + ty * kTM + i;
    float z[TN];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      z[j] = acc[i][j] + bias[j];
      if (tx * TN + j < a.Cout) s += z[j];
    }
    float y[TN], xh[TN], r = 0.f;
    if (a.ln_g) {
#pragma unroll
      for (int o = Cfg::LANES_N / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const float d = z[j] - mean;
        if (tx * TN + j < a.Cout) s2 += d * d;
      }
#pragma unroll
      for (int o = Cfg::LANES_N / 2; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      r = rsqrtf(s2 * inv_c + kLnEps);
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        xh[j] = (z[j] - mean) * r;
        y[j] = xh[j] * g[j] + be[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) y[j] = z[j];
      if (a.residual != nullptr && m < a.M) {
#pragma unroll
        for (int j = 0; j < TN; ++j)
          if (tx * TN + j < a.Cout) y[j] += a.residual[(int64_t)m * a.Cout + tx * TN + j];
      }
    }
    if (m < a.M) {
      const bool save = a.xhat != nullptr && a.ln_g != nullptr && m < a.m_train;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = tx * TN + j;
        if (n < a.Cout) {
          a.out[(int64_t)m * a.Cout + n] = a.relu ? fmaxf(y[j], 0.f) : y[j];
          if (save) a.xhat[(int64_t)m * a.Cout + n] = xh[j];
        }
      }
      if (save && tx == 0) a.rstd[m] = r;
    }
  }
}

// --------------------------------------------------------------------------------------------- conv wgrad
struct ConvWgradArgs {
  const void* in;  // layer input for the rows with a backward pass: [n_img][H][W][Cin]
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x;
  int M, K;
  const float* dz;  // [M][Cout]
  float* part;      // [splits][K][Cout]
  int rows_per_split;
};

template <int KIND>
__global__ void __launch_bounds__(kGemmThreads) conv_wgrad_kernel(const ConvWgradArgs a) {
  constexpr int BM = 64, BN = 64, TN = 8;  // BM tiles the K (= ky,kx,c) axis of dW, BN its Cout axis
  typedef TileCfg<BM, BN, TN> Cfg;
  __shared__ __align__(16) float As[kBK][BM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  const int tid = threadIdx.x;
  const int kc0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid % Cfg::LANES_N, ty = tid / Cfg::LANES_N;
  const int m_begin = blockIdx.z * a.rows_per_split;
  const int m_end = min(a.M, m_begin + a.rows_per_split);

  // this thread always gathers the same column kc of the im2col matrix
  const int a_mm = tid % BM, a_kk0 = tid / BM;
  const int kc = kc0 + a_mm;
  const bool kcvalid = kc < a.K;
  const int c = kc % a.Cin;
  const int t = kc / a.Cin;
  const int kx = t % a.ksz, ky = t / a.ksz;
  const int b_nn = tid % BN, b_kk0 = tid / BN;

  float acc[kTM][TN];
#pragma unroll
  for (int i = 0; i < kTM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // software pipeline (as in conv_fwd_kernel): the next chunk's global loads are in registers while this one is multiplied.
  // Every thread tracks the output pixel (image, oy, ox) of the NA im2col rows it gathers and steps them by kBK per chunk.
  constexpr int PA = kGemmThreads / BM, NA = kBK / PA, PB = kGemmThreads / BN, NB = kBK / PB;
  int p_img[NA], p_oy[NA], p_ox[NA];
#pragma unroll
  for (int q = 0; q < NA; ++q) {
    const int m = m_begin + a_kk0 + q * PA;
    p_img[q] = m / (a.OH * a.OW);
    const int rem = m - p_img[q] * (a.OH * a.OW);
    p_oy[q] = rem / a.OW;
    p_ox[q] = rem - p_oy[q] * a.OW;
  }
  float ra[NA], rb[NB];
  auto fetch = [&](int mb) {
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int m = mb + a_kk0 + q * PA;
      float v = 0.f;
      const int iy = p_oy[q] * a.stride - a.pad_y + ky, ix = p_ox[q] * a.stride - a.pad_x + kx;
      if (kcvalid && m < m_end && (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W)
        v = read_in<KIND>(a.in, (((int64_t)p_img[q] * a.H + iy) * a.W + ix) * a.Cin + c);
      ra[q] = v;
      p_ox[q] += kBK;  // the row this slot gathers in the next chunk
      while (p_ox[q] >= a.OW) {
        p_ox[q] -= a.OW;
        if (++p_oy[q] == a.OH) {
          p_oy[q] = 0;
          ++p_img[q];
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int m = mb + b_kk0 + q * PB;
      const int n = n0 + b_nn;
      rb[q] = (m < m_end && n < a.Cout) ? a.dz[(int64_t)m * a.Cout + n] : 0.f;
    }
  };
  if (m_begin < m_end) fetch(m_begin);
  for (int mb = m_begin; mb < m_end; mb += kBK) {
#pragma unroll
    for (int q = 0; q < NA; ++q) As[a_kk0 + q * PA][a_mm] = ra[q];
#pragma unroll
    for (int q = 0; q < NB; ++q) Bs[b_kk0 + q * PB][b_nn] = rb[q];
    __syncthreads();
    if (mb + kBK < m_end) fetch(mb + kBK);
    tile_fma<BM, BN, TN>(As, Bs, acc, ty, tx);
    __syncthreads();
  }
  float* dst = a.part + (int64_t)blockIdx.z * a.K * a.Cout;
#pragma unroll
  for (int i = 0; i < kTM; ++i) {
    const int k = kc0 + ty * kTM + i;
    if (k >= a.K) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < a.Cout) dst[(int64_t)k * a.Cout + n] = acc[i][j];
    }
  }
}

// --------------------------------------------------------------------------------------------- conv dgrad
struct ConvDgradArgs {
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x;
  int n_img;
  int taps;  // ceil(ksz / stride): taps per axis that can hit one input pixel
  int Kd;    // taps*taps*Cout
  const float* dz;  // [n_img*OH*OW][Cout]
  const float* w;   // [ksz][ksz][Cin][Cout]
  float* dx;        // [n_img*H*W][Cin]
};

// <64, 8>: 64 x 64 tiles; <32, 4>: 32 x 64 tiles — twice the CTAs for the small grids of a batch-32 step
template <int BM, int TN>
__global__ void __launch_bounds__(kGemmThreads) conv_dgrad_kernel(const ConvDgradArgs a) {
  constexpr int BN = 64;
  typedef TileCfg<BM, BN, TN> Cfg;
  __shared__ __align__(16) float As[kBK][BM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  __shared__ int ri_img[BM], ri_oy[BM], ri_ox[BM], ri_pix[BM];
  const int tid = threadIdx.x;
  const int tx = tid % Cfg::LANES_N, ty = tid / Cfg::LANES_N;
  const int s = a.stride;
  // parity class of this CTA: input pixels with (iy + pad_y) % s == ry, (ix + pad_x) % s == rx
  const int ry = blockIdx.z / s, rx = blockIdx.z % s;
  const int iy_first = ((ry - a.pad_y) % s + s) % s, ix_first = ((rx - a.pad_x) % s + s) % s;
  const int ny = iy_first < a.H ? (a.H - iy_first + s - 1) / s : 0;
  const int nx = ix_first < a.W ? (a.W - ix_first + s - 1) / s : 0;
  const int rows = a.n_img * ny * nx;
  const int m0 = blockIdx.x * BM;
  if (m0 >= rows) return;
  const int n0 = blockIdx.y * BN;

  for (int r = tid; r < BM; r += kGemmThreads) {
    const int m = m0 + r;
    if (m < rows) {
      const int img = m / (ny * nx);
      const int rem = m - img * (ny * nx);
      const int iyc = rem / nx, ixc = rem - iyc * nx;
      const int iy = iy_first + s * iyc, ix = ix_first + s * ixc;
      ri_img[r] = img;
      ri_oy[r] = (iy + a.pad_y - ry) / s;  // output row hit by tap ty_=0; tap ty_ hits oy - ty_
      ri_ox[r] = (ix + a.pad_x - rx) / s;
      ri_pix[r] = (img * a.H + iy) * a.W + ix;
    } else {
      ri_img[r] = -1;
      ri_oy[r] = ri_ox[r] = ri_pix[r] = 0;
    }
  }
  __syncthreads();

  float acc[kTM][TN];
#pragma unroll
  for (int i = 0; i < kTM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int a_kk = tid % kBK, a_r0 = tid / kBK;
  constexpr int PASS = kGemmThreads / kBK, NA = BM / PASS, NB = BN / PASS;
  // software pipeline (as in conv_fwd_kernel): the next chunk's global loads are in registers while this one is multiplied
  float ra[NA], rb[NB];
  auto fetch = [&](int k0) {
    const int k = k0 + a_kk;
    const int co = k % a.Cout;
    const int t = k / a.Cout;
    const int tx_ = t % a.taps, ty_ = t / a.taps;
    const int ky = ry + s * ty_, kx = rx + s * tx_;
    const bool tapvalid = k < a.Kd && ky < a.ksz && kx < a.ksz;
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const int r = a_r0 + q * PASS;
      float v = 0.f;
      const int img = ri_img[r];
      const int oy = ri_oy[r] - ty_, ox = ri_ox[r] - tx_;
      if (tapvalid && img >= 0 && (unsigned)oy < (unsigned)a.OH && (unsigned)ox < (unsigned)a.OW)
        v = a.dz[(((int64_t)img * a.OH + oy) * a.OW + ox) * a.Cout + co];
      ra[q] = v;
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {  // B tile: W[ky][kx][c][co], co (= k) fastest
      const int cc = n0 + a_r0 + q * PASS;
      rb[q] = (tapvalid && cc < a.Cin) ? __ldg(a.w + (((int64_t)ky * a.ksz + kx) * a.Cin + cc) * a.Cout + co) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < a.Kd; k0 += kBK) {
#pragma unroll
    for (int q = 0; q < NA; ++q) As[a_kk][a_r0 + q * PASS] = ra[q];
#pragma unroll
    for (int q = 0; q < NB; ++q) Bs[a_kk][a_r0 + q * PASS] = rb[q];
    __syncthreads();
    if (k0 + kBK < a.Kd) fetch(k0 + kBK);
    tile_fma<BM, BN, TN>(As, Bs, acc, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < kTM; ++i) {
    const int r = ty * kTM + i;
    if (ri_img[r] < 0) continue;
    float* dst = a.dx + (int64_t)ri_pix[r] * a.Cin;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int cc = n0 + tx * TN + j;
      if (cc < a.Cin) dst[cc] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------ strided GEMM
struct GemmArgs {
  const float* A; int64_t sam, sak;  // A(m,k) = A[m*sam + k*sak]
  const float* B; int64_t sbk, sbn;  // B(k,n) = B[k*sbk + n*sbn]
  float* C; int64_t ldc;             // split z writes C + z*split_stride
  int64_t split_stride;
  int M, N, K, k_per_split;
  const float* bias;                 // optional (splits == 1): C = A B + bias[n]
};

template <bool A_KFAST, bool B_NFAST>
__global__ void __launch_bounds__(kGemmThreads) gemm_strided_kernel(const GemmArgs g) {
  pdl_sync();
  constexpr int BM = 64, BN = 64, TN = 8;
  typedef TileCfg<BM, BN, TN> Cfg;
  __shared__ __align__(16) float As[kBK][BM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid % Cfg::LANES_N, ty = tid / Cfg::LANES_N;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);

  float acc[kTM][TN];
#pragma unroll
  for (int i = 0; i < kTM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += kBK) {
#pragma unroll
    for (int i = tid; i < BM * kBK; i += kGemmThreads) {
      const int kk = A_KFAST ? i % kBK : i / BM;
      const int mm = A_KFAST ? i / kBK : i % BM;
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < k_end) ? __ldg(g.A + (int64_t)m * g.sam + (int64_t)k * g.sak) : 0.f;
    }
#pragma unroll
    for (int i = tid; i < BN * kBK; i += kGemmThreads) {
      const int nn = B_NFAST ? i % BN : i / kBK;
      const int kk = B_NFAST ? i / BN : i % kBK;
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < k_end) ? __ldg(g.B + (int64_t)k * g.sbk + (int64_t)n * g.sbn) : 0.f;
    }
    __syncthreads();
    tile_fma<BM, BN, TN>(As, Bs, acc, ty, tx);
    __syncthreads();
  }
  float* C = g.C + (int64_t)blockIdx.z * g.split_stride;
#pragma unroll
  for (int i = 0; i < kTM; ++i) {
    const int m = m0 + ty * kTM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.N) C[(int64_t)m * g.ldc + n] = acc[i][j] + (g.bias ? g.bias[n] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------- dense finalize (bias + LN + ReLU)
constexpr int kRowThreads = 256;
constexpr int kRowMaxPerThread = 8;  // dense widths up to 2048

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kRowThreads / 32; ++w) t += red[w];
  __syncthreads();
  return t;
}

// out[r][n] = act( LN( sum_s part[s][r][n] + bias[n] ) ); one CTA per row
static __global__ void __launch_bounds__(kRowThreads)
dense_finalize_kernel(const float* __restrict__ part, int splits, int64_t split_stride, int rows, int N,
                      const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                      int relu, float* __restrict__ out, float* __restrict__ xhat, float* __restrict__ rstd,
                      int rows_train, __nv_bfloat16* __restrict__ out16 = nullptr) {
  pdl_sync();
  __shared__ float red[kRowThreads / 32];
  const int r = blockIdx.x;
  float z[kRowMaxPerThread];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = threadIdx.x + j * kRowThreads;
    z[j] = 0.f;
    if (n < N) {
      float v = 0.f;
      for (int sp = 0; sp < splits; ++sp) v += part[(int64_t)sp * split_stride + (int64_t)r * N + n];
      z[j] = v + bias[n];
      s += z[j];
    }
  }
  float rs = 0.f;
  if (ln_g) {
    const float mean = block_sum_256(s, red) / (float)N;
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = threadIdx.x + j * kRowThreads;
      if (n < N) {
        z[j] -= mean;
        s2 += z[j] * z[j];
      }
    }
    rs = rsqrtf(block_sum_256(s2, red) / (float)N + kLnEps);
  }
  const bool save = ln_g && xhat && r < rows_train;
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = threadIdx.x + j * kRowThreads;
    if (n < N) {
      float y = z[j];
      if (ln_g) {
        const float xh = z[j] * rs;
        if (save) xhat[(int64_t)r * N + n] = xh;
        y = xh * ln_g[n] + ln_b[n];
      }
      y = relu ? fmaxf(y, 0.f) : y;
      out[(int64_t)r * N + n] = y;
      if (out16) out16[(int64_t)r * N + n] = __float2bfloat16_rn(y);
    }
  }
  if (save && threadIdx.x == 0) rstd[r] = rs;
}

// --------------------------------------------------------------------------------- LN + ReLU backward
// In place: d (= dL/d out, post-ReLU) -> dz (= dL/d pre-LayerNorm conv/dense output).
//   mask = out > 0 ; dy = d*mask ; g = dy*gamma ; dz = rstd * (g - mean(g) - xhat*mean(g*xhat))
// Column partial sums per CTA (fixed order => deterministic): [0]=sum dz (dbias) [1]=sum dy*xhat (dgamma)
// [2]=sum dy (dbeta).  Without LayerNorm: dz = dy, only [0] is meaningful.
// Variant A: C <= 256, one warp per row, R rows in flight per warp (all loads of the R rows are issued before any
// reduction: the kernel is pure streaming and needs the memory-level parallelism).
// STORE_D = false: only the bf16 copy of dz is written (the tensor-core path never reads the fp32 one again).
template <int MAXJ, int R, bool STORE_D>
__global__ void __launch_bounds__(256)
ln_relu_bwd_warp_kernel(float* __restrict__ d, const float* __restrict__ xhat, const float* __restrict__ rstd,
                        const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                        const float* __restrict__ act, int rows, int C, float* __restrict__ colpart,
                        __nv_bfloat16* __restrict__ dz16, const __nv_bfloat16* __restrict__ act16) {
  pdl_sync();
  __shared__ float sm[8][3][32 * MAXJ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float c0[MAXJ], c1[MAXJ], c2[MAXJ], gam[MAXJ], bet[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    c0[j] = c1[j] = c2[j] = 0.f;
    const int n = lane + 32 * j;
    gam[j] = (ln_g && n < C) ? ln_g[n] : 0.f;
    bet[j] = (ln_g && n < C) ? ln_b[n] : 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int r0 = (blockIdx.x * 8 + warp) * R; r0 < rows; r0 += gridDim.x * 8 * R) {
    float dv[R][MAXJ], xh[R][MAXJ], rs[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int r = r0 + q;
      rs[q] = (ln_g && r < rows) ? rstd[r] : 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int n = lane + 32 * j;
        dv[q][j] = xh[q][j] = 0.f;
        if (r < rows && n < C) {
          const int64_t idx = (int64_t)r * C + n;
          dv[q][j] = d[idx];
          if (ln_g) xh[q][j] = xhat[idx];
          else xh[q][j] = act16 ? __bfloat162float(act16[idx]) : act[idx];  // (post-ReLU output: only its sign is used)
        }
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int r = r0 + q;
      float dy[MAXJ];
      float sg = 0.f, sgx = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        if (ln_g) {
          dy[j] = (xh[q][j] * gam[j] + bet[j] > 0.f) ? dv[q][j] : 0.f;
          const float g = dy[j] * gam[j];
          sg += g;
          sgx += g * xh[q][j];
        } else {
          dy[j] = xh[q][j] > 0.f ? dv[q][j] : 0.f;
          xh[q][j] = 0.f;
        }
      }
      float mg = 0.f, mgx = 0.f;
      if (ln_g) {
        mg = warp_sum(sg) * inv_c;
        mgx = warp_sum(sgx) * inv_c;
      }
      if (r < rows) {
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          const int n = lane + 32 * j;
          if (n < C) {
            const float dz = ln_g ? rs[q] * (dy[j] * gam[j] - mg - xh[q][j] * mgx) : dy[j];
            if (STORE_D) d[(int64_t)r * C + n] = dz;
            if (dz16) dz16[(int64_t)r * C + n] = __float2bfloat16_rn(dz);
            c0[j] += dz;
            c1[j] += dy[j] * xh[q][j];
            c2[j] += dy[j];
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    sm[warp][0][lane + 32 * j] = c0[j];
    sm[warp][1][lane + 32 * j] = c1[j];
    sm[warp][2][lane + 32 * j] = c2[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += 256) {
    const int which = i / C, n = i - which * C;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][which][n];
    colpart[((int64_t)blockIdx.x * 3 + which) * C + n] = t;
  }
}

// host-side dispatch over the channel count (C <= 256)
static inline cudaError_t launch_ln_relu_bwd_warp(int ctas, cudaStream_t s, float* d, const float* xhat, const float* rstd,
                                                  const float* ln_g, const float* ln_b, const float* act, int rows, int C,
                                                  float* colpart, __nv_bfloat16* dz16, const __nv_bfloat16* act16,
                                                  bool store_d = true) {
#define ISDQN_LN_BWD(MAXJ, R)                                                                                          \
  if (store_d || !dz16) {                                                                                              \
    co_resident_with_tc(ln_relu_bwd_warp_kernel<MAXJ, R, true>);                                                       \
    return launch_pdl((ln_relu_bwd_warp_kernel<MAXJ, R, true>), dim3(ctas), dim3(256), 0, s, d, xhat, rstd, ln_g, ln_b, act, \
                      rows, C, colpart, dz16, act16);                                                                  \
  }                                                                                                                    \
  co_resident_with_tc(ln_relu_bwd_warp_kernel<MAXJ, R, false>);                                                        \
  return launch_pdl((ln_relu_bwd_warp_kernel<MAXJ, R, false>), dim3(ctas), dim3(256), 0, s, d, xhat, rstd, ln_g, ln_b, act, \
                    rows, C, colpart, dz16, act16)
  if (C <= 32) { ISDQN_LN_BWD(1, 4); }
  if (C <= 64) { ISDQN_LN_BWD(2, 4); }
  if (C <= 128) { ISDQN_LN_BWD(4, 2); }
  ISDQN_LN_BWD(8, 1);
#undef ISDQN_LN_BWD
}

// Variant B: any C <= 2048 (dense layers), one CTA walks its rows, thread t owns columns t, t+256, ...
static __global__ void __launch_bounds__(kRowThreads)
ln_relu_bwd_block_kernel(float* __restrict__ d, const float* __restrict__ xhat, const float* __restrict__ rstd,
                         const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                         const float* __restrict__ act, int rows, int C, float* __restrict__ colpart,
                         __nv_bfloat16* __restrict__ dz16 = nullptr, const __nv_bfloat16* __restrict__ act16 = nullptr) {
  pdl_sync();
  __shared__ float red[kRowThreads / 32];
  float c0[kRowMaxPerThread], c1[kRowMaxPerThread], c2[kRowMaxPerThread], gam[kRowMaxPerThread], bet[kRowMaxPerThread];
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    c0[j] = c1[j] = c2[j] = 0.f;
    const int n = threadIdx.x + j * kRowThreads;
    gam[j] = (ln_g && n < C) ? ln_g[n] : 0.f;
    bet[j] = (ln_g && n < C) ? ln_b[n] : 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    float dy[kRowMaxPerThread], xh[kRowMaxPerThread];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = threadIdx.x + j * kRowThreads;
      dy[j] = xh[j] = 0.f;
      if (n < C) {
        const int64_t idx = (int64_t)r * C + n;
        const float dv = d[idx];
        if (ln_g) {
          xh[j] = xhat[idx];
          dy[j] = (xh[j] * gam[j] + bet[j] > 0.f) ? dv : 0.f;
          const float g = dy[j] * gam[j];
          sg += g;
          sgx += g * xh[j];
        } else {
          const float a = act16 ? __bfloat162float(act16[idx]) : act[idx];
          dy[j] = a > 0.f ? dv : 0.f;
        }
      }
    }
    float rs = 0.f, mg = 0.f, mgx = 0.f;
    if (ln_g) {
      mg = block_sum_256(sg, red) * inv_c;
      mgx = block_sum_256(sgx, red) * inv_c;
      rs = rstd[r];
    }
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = threadIdx.x + j * kRowThreads;
      if (n < C) {
        const float dz = ln_g ? rs * (dy[j] * gam[j] - mg - xh[j] * mgx) : dy[j];
        d[(int64_t)r * C + n] = dz;
        if (dz16) dz16[(int64_t)r * C + n] = __float2bfloat16_rn(dz);
        c0[j] += dz;
        c1[j] += dy[j] * xh[j];
        c2[j] += dy[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = threadIdx.x + j * kRowThreads;
    if (n < C) {
      colpart[((int64_t)blockIdx.x * 3 + 0) * C + n] = c0[j];
      colpart[((int64_t)blockIdx.x * 3 + 1) * C + n] = c1[j];
      colpart[((int64_t)blockIdx.x * 3 + 2) * C + n] = c2[j];
    }
  }
}

// Whole backward pass of the head layer + the ReLU/LayerNorm backward of the last hidden layer in ONE launch (small
// batches: three dependent launches of ~10 us each otherwise).  Two roles by blockIdx.x:
//   [0, row_ctas)        rows b = blockIdx.x, += row_ctas ...:  d[b][n] = sum_k dq[b][k] Wh[n][k]  (input gradient of the
//                        head layer), then exactly ln_relu_bwd_block_kernel's LayerNorm/ReLU backward of that row;
//   [row_ctas, gridDim)  one thread per element of the head kernel gradient: dWh[n][k] = sum_b act[b][n] dq[b][k].
// Both sums run over their index in ascending order with fmaf, as gemm_strided_kernel does, and skip terms whose dq
// factor is exactly zero — which leaves the fp32 result bit-identical (x + 0*w == x) and makes use of the TD loss
// gradient being zero outside the taken action of every head (isdqn.py:97-103).
static __global__ void __launch_bounds__(kRowThreads)
head_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ wh, const float* __restrict__ act_in, int B, int C,
                int NH, int row_ctas, float* __restrict__ d, const float* __restrict__ xhat, const float* __restrict__ rstd,
                const float* __restrict__ ln_g, const float* __restrict__ ln_b, float* __restrict__ colpart,
                __nv_bfloat16* __restrict__ dz16, float* __restrict__ dwh) {
  pdl_sync();
  __shared__ float red[kRowThreads / 32];
  __shared__ float nzv[128];
  __shared__ int nzk[128];
  __shared__ int nnz_sh;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= row_ctas) {  // ---- head kernel gradient
    const int o = ((int)blockIdx.x - row_ctas) * kRowThreads + tid;
    if (o < C * NH) {
      const int n = o / NH, k = o - n * NH;
      float acc = 0.f;
#pragma unroll 8
      for (int b = 0; b < B; ++b) {
        const float g = dq[(int64_t)b * NH + k];
        if (g != 0.f) acc = fmaf(act_in[(int64_t)b * C + n], g, acc);
      }
      dwh[o] = acc;
    }
    return;
  }
  float c0[kRowMaxPerThread], c1[kRowMaxPerThread], c2[kRowMaxPerThread], gam[kRowMaxPerThread], bet[kRowMaxPerThread];
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    c0[j] = c1[j] = c2[j] = 0.f;
    const int n = tid + j * kRowThreads;
    gam[j] = (ln_g && n < C) ? ln_g[n] : 0.f;
    bet[j] = (ln_g && n < C) ? ln_b[n] : 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int r = blockIdx.x; r < B; r += row_ctas) {
    __syncthreads();  // (the previous row's list is no longer read)
    if (tid < 32) {   // compact the non-zero entries of dq[r][:] in ascending k (NH <= 128)
      int base = 0;
      for (int c = 0; c < NH; c += 32) {
        const int k = c + tid;
        const float v = k < NH ? dq[(int64_t)r * NH + k] : 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
        if (v != 0.f) {
          const int pos = base + __popc(m & ((1u << tid) - 1u));
          nzv[pos] = v;
          nzk[pos] = k;
        }
        base += __popc(m);
      }
      if (tid == 0) nnz_sh = base;
    }
    __syncthreads();
    const int nnz = nnz_sh;
    float dy[kRowMaxPerThread], xh[kRowMaxPerThread];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = tid + j * kRowThreads;
      dy[j] = xh[j] = 0.f;
      if (n < C) {
        float dv = 0.f;
        // four head-kernel loads in flight at a time (nnz is ~K: one dependent L2 round trip per entry otherwise); the
        // additions keep their order
        for (int i0 = 0; i0 < nnz; i0 += 4) {
          float w4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) w4[u] = i0 + u < nnz ? __ldg(wh + (int64_t)n * NH + nzk[i0 + u]) : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + u < nnz) dv = fmaf(nzv[i0 + u], w4[u], dv);
        }
        const int64_t idx = (int64_t)r * C + n;
        if (ln_g) {
          xh[j] = xhat[idx];
          dy[j] = (xh[j] * gam[j] + bet[j] > 0.f) ? dv : 0.f;
          const float g = dy[j] * gam[j];
          sg += g;
          sgx += g * xh[j];
        } else {
          dy[j] = act_in[idx] > 0.f ? dv : 0.f;
        }
      }
    }
    float rs = 0.f, mg = 0.f, mgx = 0.f;
    if (ln_g) {
      mg = block_sum_256(sg, red) * inv_c;
      mgx = block_sum_256(sgx, red) * inv_c;
      rs = rstd[r];
    }
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = tid + j * kRowThreads;
      if (n < C) {
        const float dz = ln_g ? rs * (dy[j] * gam[j] - mg - xh[j] * mgx) : dy[j];
        d[(int64_t)r * C + n] = dz;
        if (dz16) dz16[(int64_t)r * C + n] = __float2bfloat16_rn(dz);
        c0[j] += dz;
        c1[j] += dy[j] * xh[j];
        c2[j] += dy[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = tid + j * kRowThreads;
    if (n < C) {
      colpart[((int64_t)blockIdx.x * 3 + 0) * C + n] = c0[j];
      colpart[((int64_t)blockIdx.x * 3 + 1) * C + n] = c1[j];
      colpart[((int64_t)blockIdx.x * 3 + 2) * C + n] = c2[j];
    }
  }
}

// heads_td_loss_kernel + head_bwd_kernel in ONE launch (small batches: the TD loss is a 6 us hop of the dependent chain for
// 288 subtractions).  d(loss)/dq is non-zero only at the taken action of every online head, so it is K numbers per sample,
// cheap enough for every CTA to recompute what it needs from the Q-values:
//   [0, row_ctas)                 input gradient + LayerNorm / ReLU backward of the hidden layer (as head_bwd_kernel), the row's
//                                 K TD errors computed on the spot
//   [row_ctas, row_ctas + wg)     head kernel gradient; the CTA first computes the whole [B][K] TD matrix into shared memory
//   last CTA                      per-head loss means (+ cumulated sums), head bias gradient, |TD| matrix, Adam step counter
// Same arithmetic per element as the two kernels it replaces (isdqn.py:97-109); the sums run in a fixed order.
constexpr int kMaxActions = 32;
constexpr int kTailMaxB = 256;
constexpr int kHbTdMax = 4096;  // B * K values staged per CTA
__device__ __forceinline__ float td_error(const float* __restrict__ q_all, int n_out, int B, int b, int k, int A, int a, float r,
                                          float coef) {
  const float* qn = q_all + (int64_t)(B + b) * n_out + k * A;  // Q_k(s', .)
  // every load is issued before the first comparison (a `mx = fmaxf(mx, qn[j])` loop is a chain of A dependent L2 round trips)
  const float qa = q_all[(int64_t)b * n_out + (k + 1) * A + a];
  float qv[kMaxActions];
#pragma unroll
  for (int j = 0; j < kMaxActions; ++j) qv[j] = j < A ? qn[j] : qn[0];
  float mx = qv[0];
#pragma unroll
  for (int j = 1; j < kMaxActions; ++j) mx = fmaxf(mx, qv[j]);  // (same maximum: max is exact in any order)
  const float target = r + coef * mx;
  return qa - target;
}

static __global__ void __launch_bounds__(kRowThreads)
head_bwd_td_kernel(const float* __restrict__ q_all, const int64_t* __restrict__ action, const double* __restrict__ reward,
                   const uint8_t* __restrict__ terminal, float gamma_n, int B, int B_global, int K, int A,
                   const float* __restrict__ is_weights, float* __restrict__ td_abs, float* __restrict__ losses,
                   double* __restrict__ cumulated, float* __restrict__ dbias, int32_t* count, const float* __restrict__ wh,
                   const float* __restrict__ act_in, int C, int NH, int row_ctas, int wg_ctas, const float* __restrict__ xhat,
                   const float* __restrict__ rstd, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                   float* __restrict__ colpart, __nv_bfloat16* __restrict__ dz16, float* __restrict__ dwh) {
  pdl_sync();
  __shared__ float red[kRowThreads / 32];
  __shared__ float s_val[kHbTdMax];  // row CTAs use the first K entries
  __shared__ float s_sq[kHbTdMax];
  __shared__ int s_act[kTailMaxB];
  const int tid = threadIdx.x;
  const float inv_b = 1.0f / (float)B_global;
  if ((int)blockIdx.x >= row_ctas) {
    // ---- the whole TD matrix: val[b][k] = d(loss)/dq at (b, head k + 1, a_b), sq[b][k] = w_b td^2
    for (int i = tid; i < B; i += kRowThreads) s_act[i] = (int)action[i];
    __syncthreads();
    for (int i = tid; i < B * K; i += kRowThreads) {
      const int b = i / K, k = i - b * K;
      const float rw = (float)reward[b];
      const float coef = (float)(1 - (int)terminal[b]) * gamma_n;
      const float td = td_error(q_all, (1 + K) * A, B, b, k, A, s_act[b], rw, coef);
      const float wb = is_weights ? is_weights[b] : 1.0f;
      s_val[i] = 2.0f * td * inv_b * wb;
      s_sq[i] = (wb * td) * td;
      if ((int)blockIdx.x == row_ctas + wg_ctas && td_abs) td_abs[(int64_t)k * B + b] = fabsf(td);
    }
    __syncthreads();
    if ((int)blockIdx.x == row_ctas + wg_ctas) {  // ---- cross-sample tail
      if (tid < K) {
        float t = 0.f;
        for (int b = 0; b < B; ++b) t += s_sq[b * K + tid];
        const float l = t * inv_b;
        losses[tid] = l;
        if (cumulated) cumulated[tid] += (double)l;
      }
      for (int c = tid; c < NH; c += kRowThreads) {
        const int hd = c / A - 1, a = c - (hd + 1) * A;
        float t = 0.f;
        if (hd >= 0)
          for (int b = 0; b < B; ++b)
            if (s_act[b] == a) t += s_val[b * K + hd];
        dbias[c] = t;
      }
      if (count && tid == 0) *count += 1;
      return;
    }
    // ---- head kernel gradient: dWh[n][c] = sum_b act[b][n] dq[b][c]
    const int o = ((int)blockIdx.x - row_ctas) * kRowThreads + tid;
    if (o < C * NH) {
      const int n = o / NH, c = o - n * NH;
      const int hd = c / A - 1, a = c - (hd + 1) * A;
      float acc = 0.f;
      if (hd >= 0) {
        for (int b = 0; b < B; ++b) {
          if (s_act[b] != a) continue;
          const float g = s_val[b * K + hd];
          if (g != 0.f) acc = fmaf(act_in[(int64_t)b * C + n], g, acc);
        }
      }
      dwh[o] = acc;
    }
    return;
  }
  // ---- row CTAs
  float c0[kRowMaxPerThread], c1[kRowMaxPerThread], c2[kRowMaxPerThread], gam[kRowMaxPerThread], bet[kRowMaxPerThread];
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    c0[j] = c1[j] = c2[j] = 0.f;
    const int n = tid + j * kRowThreads;
    gam[j] = (ln_g && n < C) ? ln_g[n] : 0.f;
    bet[j] = (ln_g && n < C) ? ln_b[n] : 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int r = blockIdx.x; r < B; r += row_ctas) {
    __syncthreads();  // (the previous row's values are no longer read)
    const int a_r = (int)action[r];
    if (tid < K) {
      const float rw = (float)reward[r];
      const float coef = (float)(1 - (int)terminal[r]) * gamma_n;
      const float td = td_error(q_all, (1 + K) * A, B, r, tid, A, a_r, rw, coef);
      const float wb = is_weights ? is_weights[r] : 1.0f;
      s_val[tid] = 2.0f * td * inv_b * wb;
    }
    __syncthreads();
    float dy[kRowMaxPerThread], xh[kRowMaxPerThread];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = tid + j * kRowThreads;
      dy[j] = xh[j] = 0.f;
      if (n < C) {
        float dv = 0.f;
        const float* wrow = wh + (int64_t)n * NH + A + a_r;  // head kernel row n at the taken action of head 1, 2, ...
        for (int i0 = 0; i0 < K; i0 += 8) {  // loads of 8 heads in flight; zero factors are skipped (x + 0 * w == x)
          float wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = (i0 + u < K) ? __ldg(wrow + (i0 + u) * A) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float v = (i0 + u < K) ? s_val[i0 + u] : 0.f;
            if (v != 0.f) dv = fmaf(v, wv[u], dv);
          }
        }
        const int64_t idx = (int64_t)r * C + n;
        if (ln_g) {
          xh[j] = xhat[idx];
          dy[j] = (xh[j] * gam[j] + bet[j] > 0.f) ? dv : 0.f;
          const float g = dy[j] * gam[j];
          sg += g;
          sgx += g * xh[j];
        } else {
          dy[j] = act_in[idx] > 0.f ? dv : 0.f;
        }
      }
    }
    float rs = 0.f, mg = 0.f, mgx = 0.f;
    if (ln_g) {
      mg = block_sum_256(sg, red) * inv_c;
      mgx = block_sum_256(sgx, red) * inv_c;
      rs = rstd[r];
    }
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = tid + j * kRowThreads;
      if (n < C) {
        const float dz = ln_g ? rs * (dy[j] * gam[j] - mg - xh[j] * mgx) : dy[j];
        if (dz16) dz16[(int64_t)r * C + n] = __float2bfloat16_rn(dz);
        c0[j] += dz;
        c1[j] += dy[j] * xh[j];
        c2[j] += dy[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = tid + j * kRowThreads;
    if (n < C) {
      colpart[((int64_t)blockIdx.x * 3 + 0) * C + n] = c0[j];
      colpart[((int64_t)blockIdx.x * 3 + 1) * C + n] = c1[j];
      colpart[((int64_t)blockIdx.x * 3 + 2) * C + n] = c2[j];
    }
  }
}

// --------------------------------------------------------------------------- deterministic partial reduce
constexpr int kMaxSegments = 40;
struct Segment {
  const float* src;  // partial p of element i at src[p*stride + i]
  float* dst;
  int64_t stride;
  int n, parts;
  // != 0: the partials are the weight gradient of the first convolution in 4x4 space-to-depth order
  // [(by, bx, py, px, c)][s2d_cout]; dst is the HWIO kernel [(ky = 4 by + py, kx = 4 bx + px, c)][s2d_cout]
  int s2d_cout;
};
// element i of a space-to-depth ordered conv0 kernel -> its index in the 8x8x4 HWIO kernel (cout % 4 == 0)
__host__ __device__ inline int s2d_to_hwio(int i, int cout) {
  const int kp = i / cout, co = i - kp * cout;
  const int tap = kp >> 6, q = kp & 63;
  const int by = tap >> 1, bx = tap & 1, py = q >> 4, px = (q >> 2) & 3, c = q & 3;
  return (((4 * by + py) * 8 + 4 * bx + px) * 4 + c) * cout + co;
}
// Cross-sample tail of the fused head step (head_mid_kernel): per-head loss means, head bias gradient, head kernel
// gradient, computed by extra CTAs of the partial-reduction launch.  tdq is [B][2K]: [b][k] = d(loss)/dq at the taken
// action of head k+1, [b][K + k] = the (weighted) squared TD error.
struct HeadTail {
  int ctas;  // 0: none
  const float* tdq;
  const int64_t* action;
  const float* act;  // [B][C] hidden activations (fp32)
  int B, K, A, C, NH;
  float inv_b;
  float* losses;
  double* cumulated;
  float* dbias;  // [NH]
  float* dwh;    // [C][NH]
};
static inline int head_tail_ctas(int C, int NH) { return 1 + (C * NH + 255) / 256; }

// (the per-sample values are staged in shared memory first: the sums below then run over shared memory in a fixed order
// instead of over a chain of dependent global loads)
__device__ __forceinline__ void head_tail_block(const HeadTail& h, int j) {
  const int tid = threadIdx.x;
  __shared__ float s_tdq[kTailMaxB * 2 * 16 > 4096 ? 4096 : kTailMaxB * 2 * 16];  // [B][2K] while it fits, else global
  __shared__ int s_act[kTailMaxB];
  const int n_tdq = h.B * 2 * h.K;
  const bool staged = n_tdq <= 4096 && h.B <= kTailMaxB;
  if (staged) {
    for (int i = tid; i < n_tdq; i += 256) s_tdq[i] = h.tdq[i];
    for (int i = tid; i < h.B; i += 256) s_act[i] = (int)h.action[i];
    __syncthreads();
  }
  auto tdq = [&](int b, int k) { return staged ? s_tdq[b * 2 * h.K + k] : h.tdq[(int64_t)b * 2 * h.K + k]; };
  auto act_of = [&](int b) { return staged ? s_act[b] : (int)h.action[b]; };
  if (j == 0) {
    // losses[k] = mean_b w_b td^2 (isdqn.py:102), sequential over b: deterministic
    if (tid < h.K) {
      float t = 0.f;
      for (int b = 0; b < h.B; ++b) t += tdq(b, h.K + tid);
      const float l = t * h.inv_b;
      h.losses[tid] = l;
      if (h.cumulated) h.cumulated[tid] += (double)l;
    }
    // head bias gradient: column sums of dq; head 0 (the frozen target) receives none
    for (int c = tid; c < h.NH; c += 256) {
      const int hd = c / h.A - 1, a = c - (hd + 1) * h.A;
      float t = 0.f;
      if (hd >= 0)
        for (int b = 0; b < h.B; ++b)
          if (act_of(b) == a) t += tdq(b, hd);
      h.dbias[c] = t;
    }
    return;
  }
  const int o = (j - 1) * 256 + tid;
  if (o >= h.C * h.NH) return;
  const int n = o / h.NH, c = o - n * h.NH;
  const int hd = c / h.A - 1, a = c - (hd + 1) * h.A;
  float acc = 0.f;
  if (hd >= 0) {
    for (int b = 0; b < h.B; ++b) {
      if (act_of(b) != a) continue;
      const float g = tdq(b, hd);
      if (g != 0.f) acc = fmaf(h.act[(int64_t)b * h.C + n], g, acc);
    }
  }
  h.dwh[o] = acc;
}

struct SegmentList {
  int count;
  int tile_start[kMaxSegments + 1];  // prefix sum of ceil(n / 32) over the segments (filled by finish_segments)
  Segment s[kMaxSegments];
};
// tile = 32 lanes x VEC elements; returns the grid size.  VEC = 4 needs every segment 16-byte aligned (checked by
// segments_vec4_ok)
static inline int finish_segments(SegmentList* l, int vec) {
  int t = 0;
  for (int i = 0; i < l->count; ++i) {
    l->tile_start[i] = t;
    t += (l->s[i].n + 32 * vec - 1) / (32 * vec);
  }
  l->tile_start[l->count] = t;
  return t;
}
static inline bool segments_vec4_ok(const SegmentList& l) {
  for (int i = 0; i < l.count; ++i) {
    const Segment& g = l.s[i];
    if ((g.n & 3) || (g.stride & 3) || (reinterpret_cast<uintptr_t>(g.src) & 15) || (reinterpret_cast<uintptr_t>(g.dst) & 15))
      return false;
  }
  return true;
}

// One CTA sums ONE tile of 32 x VEC elements of one segment (grid = all tiles of all segments, so every partial of the
// step is in flight at once): its 8 warps take the partials p = w, w+8, ... (independent loads), then the 8 warp sums
// are combined in a fixed order => deterministic, and identical for VEC = 1 and VEC = 4.
template <int VEC>
__global__ void __launch_bounds__(256) reduce_segments_kernel(const SegmentList list, const HeadTail tail, int n_tiles) {
  // Trigger BEFORE the wait: the successor is the streaming optimiser kernel, whose tile pipeline touches nothing this
  // kernel or its predecessor (the last weight gradient) writes and which only waits for THIS grid before its rest-leaf
  // phase (adam_stream.cu).  Launched now, its CTAs take the SMs as the weight-gradient CTAs leave them and stream the
  // Dense kernel's state while the partial sums are still being folded (ISDQN_REDUCE_LATE_TRIGGER=1 at build time: the
  // usual wait-then-trigger).
  trace_kernel_start();
#if defined(ISDQN_REDUCE_LATE_TRIGGER)
  pdl_wait();
  pdl_trigger();
#else
  pdl_trigger();
  pdl_wait();
#endif
  if ((int)blockIdx.x >= n_tiles) {
    head_tail_block(tail, (int)blockIdx.x - n_tiles);
    return;
  }
  __shared__ float sm[8][32 * VEC + 4];
  int seg = 0;
  while (seg + 1 < list.count && (int)blockIdx.x >= list.tile_start[seg + 1]) ++seg;
  const Segment sg = list.s[seg];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = (((int)blockIdx.x - list.tile_start[seg]) * 32 + lane) * VEC;
  float t[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) t[v] = 0.f;
  if (i < sg.n) {
    const float* src = sg.src + i;
    auto load = [&](int p, float (&x)[VEC]) {
      if (VEC == 4) {
        const float4 q = *reinterpret_cast<const float4*>(src + (int64_t)p * sg.stride);
        x[0] = q.x; x[VEC > 1 ? 1 : 0] = q.y; x[VEC > 2 ? 2 : 0] = q.z; x[VEC > 3 ? 3 : 0] = q.w;
      } else {
        x[0] = src[(int64_t)p * sg.stride];
      }
    };
    int p = w;
    for (; p + 24 < sg.parts; p += 32) {  // four independent loads per warp iteration
      float a0[VEC], a1[VEC], a2[VEC], a3[VEC];
      load(p, a0);
      load(p + 8, a1);
      load(p + 16, a2);
      load(p + 24, a3);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        t[v] += a0[v];
        t[v] += a1[v];
        t[v] += a2[v];
        t[v] += a3[v];
      }
    }
    for (; p < sg.parts; p += 8) {
      float a0[VEC];
      load(p, a0);
#pragma unroll
      for (int v = 0; v < VEC; ++v) t[v] += a0[v];
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) sm[w][lane * VEC + v] = t[v];
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * VEC; e += 256) {
    const int idx = (((int)blockIdx.x - list.tile_start[seg]) * 32) * VEC + e;
    if (idx < sg.n) {
      float tot = sm[0][e];
#pragma unroll
      for (int k = 1; k < 8; ++k) tot += sm[k][e];
      sg.dst[sg.s2d_cout ? s2d_to_hwio(idx, sg.s2d_cout) : idx] = tot;
    }
  }
}

static inline cudaError_t launch_reduce_segments(SegmentList& segs, cudaStream_t s, const HeadTail* tail = nullptr) {
  co_resident_with_tc(reduce_segments_kernel<4>);
  co_resident_with_tc(reduce_segments_kernel<1>);
  HeadTail none = {};
  const HeadTail& ht = tail ? *tail : none;
  if (segments_vec4_ok(segs)) {
    const int n_tiles = finish_segments(&segs, 4);
    return launch_pdl((reduce_segments_kernel<4>), dim3(n_tiles + ht.ctas), dim3(256), 0, s, segs, ht, n_tiles);
  }
  const int n_tiles = finish_segments(&segs, 1);
  return launch_pdl((reduce_segments_kernel<1>), dim3(n_tiles + ht.ctas), dim3(256), 0, s, segs, ht, n_tiles);
}

// ------------------------------------------------------------------------------ K-head TD loss fwd + bwd
// isdqn.py:97-109.  One CTA per online head k+1 (regressing onto head k), fixed reduction order => deterministic.
// CTA 0 also bumps the Adam step counter (so the adam_kernel of the same step reads count+1 with no race); every
// CTA produces its head's slice of d(loss)/dq and of the head layer's bias gradient.
constexpr int kLossThreads = 256;
constexpr int kMaxHeads = 64;

static __global__ void __launch_bounds__(kLossThreads)
heads_td_loss_kernel(const float* __restrict__ q_all, const int64_t* __restrict__ action,
                     const double* __restrict__ reward, const uint8_t* __restrict__ terminal, float gamma_n, int B,
                     int B_global, int K, int A, float* __restrict__ losses, float* __restrict__ dq,
                     float* __restrict__ dbias, int32_t* count, double* __restrict__ cumulated = nullptr,
                     const float* __restrict__ is_weights = nullptr, float* __restrict__ td_abs = nullptr) {
  pdl_sync();
  __shared__ float red[kLossThreads / 32];
  __shared__ float dbw[kLossThreads / 32][kMaxActions];
  const int k = blockIdx.x;
  const int n_out = (1 + K) * A;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (dq) {  // d(loss)/dq is zero except at (b, head k+1, a_b); head 0 receives no gradient at all
    for (int i = tid; i < B * A; i += kLossThreads) {
      const int b = i / A, a = i - b * A;
      dq[(int64_t)b * n_out + (k + 1) * A + a] = 0.f;
      if (k == 0) dq[(int64_t)b * n_out + a] = 0.f;
    }
  }
  if (tid < kMaxActions)
    for (int w = 0; w < kLossThreads / 32; ++w) dbw[w][tid] = 0.f;
  __syncthreads();
  const float inv_b = 1.0f / (float)B_global;
  float part = 0.f;
  for (int b0 = 0; b0 < B; b0 += kLossThreads) {
    const int b = b0 + tid;
    float val = 0.f;
    int a = -1;
    if (b < B) {
      a = (int)action[b];
      const float r = (float)reward[b];                             // f64 -> f32 at the jit boundary
      const float coef = (float)(1 - (int)terminal[b]) * gamma_n;   // ((1 - d) * gamma^n) in fp32
      const float td = td_error(q_all, n_out, B, b, k, A, a, r, coef);  // Q_{k+1}(s, a) - (r + coef max_a' Q_k(s', a'))
      // importance weight of a prioritized batch (1 without: both products are then exact, the reference's loss bit for bit)
      const float wb = is_weights ? is_weights[b] : 1.0f;
      part += (wb * td) * td;
      val = 2.0f * td * inv_b * wb;
      if (dq) dq[(int64_t)b * n_out + (k + 1) * A + a] = val;
      if (td_abs) td_abs[(int64_t)k * B + b] = fabsf(td);
    }
    if (dbias) {
      for (int j = 0; j < A; ++j) {
        const float sj = warp_sum(a == j ? val : 0.f);
        if (lane == 0) dbw[warp][j] += sj;
      }
    }
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) t += red[w];
    losses[k] = t * inv_b;
    if (cumulated) cumulated[k] += (double)(t * inv_b);
  }
  if (dbias && tid < A) {
    float t = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) t += dbw[w][tid];
    dbias[(k + 1) * A + tid] = t;
    if (k == 0) dbias[tid] = 0.f;
  }
  if (count && k == 0 && tid == 0) *count += 1;
}

// Head layer forward (tiny: N = (1+K)A <= 128 columns): one CTA per kHeadRows rows (the head kernel is read once per
// CTA, not once per row), 4 k-groups x 128 columns; per row the sum runs in the same order as for a single row.
template <int kHeadRows>
__global__ void __launch_bounds__(512)
head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int K, int N,
                float* __restrict__ out, int rows) {
  pdl_sync();
  extern __shared__ float xs[];  // [kHeadRows][K]
  __shared__ float part[4][kHeadRows][128];
  const int r0 = blockIdx.x * kHeadRows, tid = threadIdx.x;
  const int nr = min(kHeadRows, rows - r0);
  for (int i = tid; i < kHeadRows * K; i += 512) {
    const int r = i / K;
    xs[i] = r < nr ? x[(int64_t)(r0 + r) * K + (i - r * K)] : 0.f;
  }
  __syncthreads();
  const int n = tid & 127, g = tid >> 7;
  float acc[kHeadRows];
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.f;
  if (n < N) {
    const int kq = (K + 3) / 4;
    const int k_end = min(K, (g + 1) * kq);
#pragma unroll(kHeadRows == 1 ? 16 : 4)
    for (int k = g * kq; k < k_end; ++k) {
      const float wv = __ldg(w + (int64_t)k * N + n);
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) acc[r] = fmaf(xs[r * K + k], wv, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) part[g][r][n] = acc[r];
  __syncthreads();
  for (int i = tid; i < kHeadRows * 128; i += 512) {
    const int r = i >> 7, c = i & 127;
    if (r < nr && c < N) out[(int64_t)(r0 + r) * N + c] = ((part[0][r][c] + part[1][r][c]) + (part[2][r][c] + part[3][r][c])) + bias[c];
  }
}

// rows per CTA: 1 while the launch is small (parallelism first), 8 for large batches (K <= kHeadMaxK: 96 KB of rows)
constexpr int kHeadMaxK = 3072;
static inline cudaError_t launch_head_fwd(cudaStream_t s, const float* x, const float* w, const float* bias, int K, int N,
                                          float* out, int rows) {
  if (rows <= 1024) return launch_pdl((head_fwd_kernel<1>), dim3(rows), dim3(512), K * sizeof(float), s, x, w, bias, K, N, out, rows);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(head_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kHeadMaxK * 4);
    if (e != cudaSuccess) return e;
  }
  return launch_pdl((head_fwd_kernel<8>), dim3(ceil_div(rows, 8)), dim3(512), 8 * K * sizeof(float), s, x, w, bias, K, N, out, rows);
}

// dense_finalize_kernel + head_fwd_kernel in one launch (the tensor-core path at small batch is bound by the number of
// dependent launches): one CTA per row finishes the hidden Dense layer (split-K partial sum, bias, LayerNorm, ReLU; the
// first 256 threads, same arithmetic order as dense_finalize_kernel), keeps the row in shared memory and multiplies
// it with the head kernel (all 512 threads, same order as head_fwd_kernel).  N <= 2048 hidden units, NH <= 128 outputs.
static __global__ void __launch_bounds__(512)
dense_finalize_head_kernel(const float* __restrict__ part, int splits, int64_t split_stride, int N,
                           const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                           int relu, float* __restrict__ out, float* __restrict__ xhat, float* __restrict__ rstd, int rows_train,
                           const float* __restrict__ wh, const float* __restrict__ bh, int NH, float* __restrict__ out_head) {
  pdl_sync();
  extern __shared__ float xs[];  // [N]
  __shared__ float red[16];
  __shared__ float hpart[4][128];
  const int r = blockIdx.x, tid = threadIdx.x;
  const bool fin = tid < kRowThreads;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kRowThreads / 32; ++w) t += red[w];  // (warps 8..15 hold zeros)
    __syncthreads();
    return t;
  };
  float z[kRowMaxPerThread];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) z[j] = 0.f;
  if (fin) {  // split partials in ascending order, the loads of 8 splits in flight together (see head_mid_kernel)
    const float* prow = part + (int64_t)r * N + tid;
    for (int sp0 = 0; sp0 < splits; sp0 += 8) {
      float t[8][kRowMaxPerThread];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < kRowMaxPerThread; ++j)
          t[u][j] = (sp0 + u < splits && tid + j * kRowThreads < N) ? prow[(int64_t)(sp0 + u) * split_stride + j * kRowThreads] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < kRowMaxPerThread; ++j) z[j] += t[u][j];
    }
  }
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = tid + j * kRowThreads;
    if (fin && n < N) {
      z[j] += bias[n];
      s += z[j];
    }
  }
  float rs = 0.f;
  if (ln_g) {
    const float mean = block_sum(s) / (float)N;
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = tid + j * kRowThreads;
      if (fin && n < N) {
        z[j] -= mean;
        s2 += z[j] * z[j];
      }
    }
    rs = rsqrtf(block_sum(s2) / (float)N + kLnEps);
  }
  const bool save = ln_g && xhat && r < rows_train;
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = tid + j * kRowThreads;
    if (fin && n < N) {
      float y = z[j];
      if (ln_g) {
        const float xh = z[j] * rs;
        if (save) xhat[(int64_t)r * N + n] = xh;
        y = xh * ln_g[n] + ln_b[n];
      }
      y = relu ? fmaxf(y, 0.f) : y;
      out[(int64_t)r * N + n] = y;
      xs[n] = y;
    }
  }
  if (save && tid == 0) rstd[r] = rs;
  __syncthreads();
  const int n = tid & 127, g = tid >> 7;
  float acc = 0.f;
  if (n < NH) {
    const int kq = (N + 3) / 4;
    const int k_end = min(N, (g + 1) * kq);
#pragma unroll 32
    for (int k = g * kq; k < k_end; ++k) acc = fmaf(xs[k], __ldg(wh + (int64_t)k * NH + n), acc);
  }
  hpart[g][n] = acc;
  __syncthreads();
  if (g == 0 && n < NH) out_head[(int64_t)r * NH + n] = ((hpart[0][n] + hpart[1][n]) + (hpart[2][n] + hpart[3][n])) + bh[n];
}

// The whole middle of a small-batch step in ONE launch, one CTA per sample b (everything here is local to a sample):
//   A  finish the last hidden Dense layer for rows b (s) and B + b (s'): split-K partial sum, bias, LayerNorm, ReLU
//      (threads 0-255 row b, 256-511 row B + b; arithmetic order of dense_finalize_kernel)
//   B  head layer for both rows, every head-kernel element loaded once for the two rows
//   C  iterated TD targets and errors of the K online heads (isdqn.py:97-109), d(loss)/dq, |TD|
//   D  head input gradient + ReLU / LayerNorm backward of the hidden layer for row b (the normalised values never
//      leave the registers), bf16 dz for the tensor-core kernels below, column partials for the bias / LayerNorm gradients
// It replaces dense_finalize_head + heads_td_loss + head_bwd (three dependent launches of the batch-32 chain); what
// needs every sample — the K loss means, the head bias and kernel gradients — is left to head_tail_block.
static __global__ void __launch_bounds__(512)
head_mid_kernel(const float* __restrict__ part, int splits, int64_t split_stride, int N, const float* __restrict__ bias,
                const float* __restrict__ ln_g, const float* __restrict__ ln_b, int relu, float* __restrict__ out,
                const float* __restrict__ wh, const float* __restrict__ bh, int NH, float* __restrict__ out_head,
                const int64_t* __restrict__ action, const double* __restrict__ reward, const uint8_t* __restrict__ terminal,
                float gamma_n, int B, int B_global, int K, int A, const float* __restrict__ is_weights,
                float* __restrict__ td_abs, float* __restrict__ tdq, float* __restrict__ colpart,
                __nv_bfloat16* __restrict__ dz16, int32_t* count) {
  pdl_sync();
  extern __shared__ float xs[];  // [2][N]
  __shared__ float red[16];
  __shared__ float hpart[2][4][128];
  __shared__ float qs[2][128];
  __shared__ float vals[kMaxHeads];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int half = tid >> 8, ht = tid & 255;
  const int r = half ? B + b : b;
  auto half_sum = [&](float v) {  // sum over the 256 threads of this thread's half, fixed order
    v = warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[half * 8 + w];
    __syncthreads();
    return t;
  };
  if (count && b == 0 && tid == 0) *count += 1;  // the Adam step number of this update (read many launches later)
  // ---- A
  float z[kRowMaxPerThread], xh[kRowMaxPerThread], gam[kRowMaxPerThread], bet[kRowMaxPerThread];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) z[j] = xh[j] = gam[j] = bet[j] = 0.f;
  {
    // split-K partial sums in ascending split order; the loads of 8 splits x every column of the thread are issued
    // together (a plain `v += part[...]` loop is a chain of dependent L2 round trips: 37 splits cost 37 latencies)
    const float* prow = part + (int64_t)r * N + ht;
    for (int sp0 = 0; sp0 < splits; sp0 += 8) {
      float t[8][kRowMaxPerThread];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < kRowMaxPerThread; ++j)
          t[u][j] = (sp0 + u < splits && ht + j * kRowThreads < N) ? prow[(int64_t)(sp0 + u) * split_stride + j * kRowThreads] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < kRowMaxPerThread; ++j) z[j] += t[u][j];
    }
  }
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = ht + j * kRowThreads;
    if (n < N) {
      z[j] += bias[n];
      s += z[j];
      if (ln_g) {
        gam[j] = ln_g[n];
        bet[j] = ln_b[n];
      }
    }
  }
  float rs = 0.f;
  if (ln_g) {
    const float mean = half_sum(s) / (float)N;
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = ht + j * kRowThreads;
      if (n < N) {
        z[j] -= mean;
        s2 += z[j] * z[j];
      }
    }
    rs = rsqrtf(half_sum(s2) / (float)N + kLnEps);
  }
  float* xrow = xs + half * N;
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = ht + j * kRowThreads;
    if (n < N) {
      float y = z[j];
      if (ln_g) {
        xh[j] = z[j] * rs;
        y = xh[j] * gam[j] + bet[j];
      }
      y = relu ? fmaxf(y, 0.f) : y;
      out[(int64_t)r * N + n] = y;
      xrow[n] = y;
    }
  }
  __syncthreads();
  // ---- B
  {
    const int n = tid & 127, g = tid >> 7;
    float a0 = 0.f, a1 = 0.f;
    if (n < NH) {
      const int kq = (N + 3) / 4;
      const int k_end = min(N, (g + 1) * kq);
#pragma unroll 16
      for (int k = g * kq; k < k_end; ++k) {
        const float w = __ldg(wh + (int64_t)k * NH + n);
        a0 = fmaf(xs[k], w, a0);
        a1 = fmaf(xs[N + k], w, a1);
      }
    }
    hpart[0][g][n] = a0;
    hpart[1][g][n] = a1;
    __syncthreads();
    if (g < 2 && n < NH) {
      const float q = ((hpart[g][0][n] + hpart[g][1][n]) + (hpart[g][2][n] + hpart[g][3][n])) + bh[n];
      qs[g][n] = q;
      out_head[(int64_t)(g ? B + b : b) * NH + n] = q;
    }
    __syncthreads();
  }
  // ---- C
  const int act_b = (int)action[b];
  if (tid < K) {
    const int k = tid;
    const float rw = (float)reward[b];                             // f64 -> f32 at the jit boundary
    const float coef = (float)(1 - (int)terminal[b]) * gamma_n;    // ((1 - d) * gamma^n) in fp32
    const float* qn = &qs[1][k * A];                               // Q_k(s', .)
    float mx = qn[0];
    for (int j = 1; j < A; ++j) mx = fmaxf(mx, qn[j]);
    const float target = rw + coef * mx;
    const float td = qs[0][(k + 1) * A + act_b] - target;
    const float wb = is_weights ? is_weights[b] : 1.0f;
    const float val = 2.0f * td * (1.0f / (float)B_global) * wb;
    vals[k] = val;
    tdq[(int64_t)b * 2 * K + k] = val;
    tdq[(int64_t)b * 2 * K + K + k] = (wb * td) * td;
    if (td_abs) td_abs[(int64_t)k * B + b] = fabsf(td);
  }
  __syncthreads();
  // ---- D (row b: the first half of the CTA holds its normalised values)
  float dy[kRowMaxPerThread];
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int j = 0; j < kRowMaxPerThread; ++j) {
    const int n = ht + j * kRowThreads;
    dy[j] = 0.f;
    if (half == 0 && n < N) {
      float dv = 0.f;
      const float* wrow = wh + (int64_t)n * NH + A + act_b;  // head kernel row n at the taken action of head 1, 2, ...
      for (int i0 = 0; i0 < K; i0 += 8) {  // (loads of 8 heads in flight; zero factors are skipped: x + 0 * w == x)
        float wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) wv[u] = (i0 + u < K) ? __ldg(wrow + (i0 + u) * A) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float v = (i0 + u < K) ? vals[i0 + u] : 0.f;
          if (v != 0.f) dv = fmaf(v, wv[u], dv);
        }
      }
      if (ln_g) {
        dy[j] = (xh[j] * gam[j] + bet[j] > 0.f) ? dv : 0.f;
        const float g = dy[j] * gam[j];
        sg += g;
        sgx += g * xh[j];
      } else {
        dy[j] = xs[n] > 0.f ? dv : 0.f;
      }
    }
  }
  float mg = 0.f, mgx = 0.f;
  if (ln_g) {
    const float inv_c = 1.0f / (float)N;
    mg = half_sum(sg) * inv_c;
    mgx = half_sum(sgx) * inv_c;
  }
  if (half == 0) {
#pragma unroll
    for (int j = 0; j < kRowMaxPerThread; ++j) {
      const int n = ht + j * kRowThreads;
      if (n < N) {
        const float dz = ln_g ? rs * (dy[j] * gam[j] - mg - xh[j] * mgx) : dy[j];
        dz16[(int64_t)b * N + n] = __float2bfloat16_rn(dz);
        colpart[((int64_t)b * 3 + 0) * N + n] = dz;
        colpart[((int64_t)b * 3 + 1) * N + n] = dy[j] * xh[j];
        colpart[((int64_t)b * 3 + 2) * N + n] = dy[j];
      }
    }
  }
}

// 1 / (1 - b^t) for both Adam decays, in double.  Two threads of different warps (the last lane of the last two warps)
// evaluate one power each by repeated squaring (t is a positive step count: <= 62 dependent multiplications, against a
// few hundred FP64 instructions for pow()) while the rest of the CTA goes on; the result agrees with pow() to ~1e-15
// relative before it is rounded to float.
__device__ __forceinline__ double pow_int(double b, int t) {
  double r = 1.0;
  for (unsigned e = (unsigned)max(t, 0); e; e >>= 1) {
    if (e & 1u) r *= b;
    b *= b;
  }
  return r;
}
__device__ __forceinline__ void adam_bias_corrections(const int32_t* count, float b1, float b2, float* s_c /*[2], shared*/) {
  const int who = (int)blockDim.x - 1 - (int)threadIdx.x;
  if (who == 0 || who == 32 || (who == 1 && blockDim.x <= 32)) {
    const int t = *count;
    const int j = who == 0 ? 0 : 1;
    s_c[j] = (float)(1.0 / (1.0 - pow_int((double)(j == 0 ? b1 : b2), t)));
  }
}
// The element update.  optax: p -= lr * (mu / c1) / (sqrt(nu / c2) + eps).  The two bias-correction divisions are
// multiplications by the reciprocals above and the quotient / square root use the SFU (MUFU.RCP / MUFU.SQRT, <= 2 ulp
// each): 12 instructions per element instead of 68 with three IEEE-rounded operations (ncu: the IEEE version made
// BOTH Adam kernels issue-bound at 43-47 % issue utilisation).  The step differs from the IEEE evaluation by < 1e-6
// relative, i.e. < 1e-9 of a parameter — far inside the 1e-5 parity bar — and stays bit-reproducible run to run.
__device__ __forceinline__ float adam_sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------- Adam
// optax 0.2.4 scale_by_adam + scale(-lr): mu = b1 mu + (1-b1) g ; nu = b2 nu + (1-b2) g^2 ;
// p -= lr * (mu / (1-b1^t)) / (sqrt(nu / (1-b2^t)) + eps) with t = *count (already incremented); c1, c2 below are the
// RECIPROCALS 1 / (1-b^t) (adam_bias_corrections).
static __global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu, float* __restrict__ nu,
            const int32_t* __restrict__ count, float lr, float b1, float b2, float eps, int64_t n4,
            __nv_bfloat16* __restrict__ shadow, int64_t skip_begin4, int64_t skip_len4) {
  // n4 float4 groups are updated; the groups [skip_begin4, skip_begin4 + skip_len4) of the vector are passed over
  // (they were already updated by an earlier launch of the same step)
  pdl_sync();
  __shared__ float s_c[2];
  adam_bias_corrections(count, b1, b2, s_c);
  __syncthreads();
  const float c1 = s_c[0], c2 = s_c[1];
  const float ob1 = 1.0f - b1, ob2 = 1.0f - b2;
  const uint64_t keep = l2_policy_evict_last(), stream_once = l2_policy_evict_first();
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j < skip_begin4 ? j : j + skip_len4;
    const float4 gv = ld_f4_hint(reinterpret_cast<const float4*>(g) + i, stream_once);
    float4 mv = ld_f4_hint(reinterpret_cast<const float4*>(mu) + i, keep);
    float4 vv = ld_f4_hint(reinterpret_cast<const float4*>(nu) + i, keep);
    float4 pv = ld_f4_hint(reinterpret_cast<const float4*>(p) + i, keep);
#define ISDQN_ADAM1(c)                                    \
  mv.c = ob1 * gv.c + b1 * mv.c;                          \
  vv.c = ob2 * (gv.c * gv.c) + b2 * vv.c;                 \
  pv.c -= lr * __fdividef(mv.c * c1, adam_sqrt_approx(vv.c * c2) + eps);
    ISDQN_ADAM1(x) ISDQN_ADAM1(y) ISDQN_ADAM1(z) ISDQN_ADAM1(w)
#undef ISDQN_ADAM1
    st_f4_hint(reinterpret_cast<float4*>(mu) + i, mv, keep);
    st_f4_hint(reinterpret_cast<float4*>(nu) + i, vv, keep);
    st_f4_hint(reinterpret_cast<float4*>(p) + i, pv, keep);
    if (shadow) {  // bf16 copy of the updated parameters for the tensor-core path of the next step
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      st_u2_hint(reinterpret_cast<uint2*>(shadow) + i, make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi)),
                 keep);
    }
  }
}

#define ISDQN_ADAM_ELEM(G, M, V, P, c)                   \
  M.c = ob1 * G.c + b1 * M.c;                            \
  V.c = ob2 * (G.c * G.c) + b2 * V.c;                    \
  P.c -= lr * __fdividef(M.c * c1, adam_sqrt_approx(V.c * c2) + eps);

// Small-batch hidden Dense kernel (97 % of the Atari network's parameters): its gradient is the rank-B product
// g[k][n] = sum_b act[b][k] dz[b][n] of two small bf16 matrices, so it is recomputed on chip here and fed straight into
// the Adam update — the 4 B/param gradient is never written nor read back (26 B/param instead of 4 + 30), and the
// separate weight-gradient launch disappears.  The kernel is HBM-bound; the product is 0.25 GFLOP, so it runs on
// warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate — the same products and accumulator type as the tcgen05 path;
// a first version with FFMA was issue-bound at 47 % issue utilisation, ncu profiles/r01_adam_ncu.md).  Not a tcgen05
// problem: one 16-row tile is 2 k-steps of work per warp between two streaming phases.
// CTA (256 threads) = tiles of 16 kernel rows x 256 columns over a contiguous range of rows: dz[B][256] and its slice
// of act^T are staged once in shared memory (bf16, padded rows: conflict-free ldmatrix), zero-filled up to a multiple
// of 16 batch rows.  Per tile: warp w computes the 16 x 32 block of columns 32 w.. (8 MMAs), stores it to a
// double-buffered fp32 tile in shared memory; then every thread owns 4 rows x 4 columns: 12 independent 16-byte loads
// of p / mu / nu are issued BEFORE the barrier that publishes the gradient tile, so they are in flight while it waits.
// grid.x = rest CTAs + row CTAs, grid.y = ceil(N / 256); the `rest` CTAs (the first ones, blockIdx.y == 0 only) run
// the plain Adam update of every OTHER leaf of the flat vector, so the step ends with one launch.
constexpr int kDwaThreads = 256, kDwaCols = 256, kDwaTileRows = 16, kDwaMaxB = 64;
constexpr int kDwaLdB = kDwaCols + 8;  // bf16 elements per staged dz row
constexpr int kDwaLdC = kDwaCols + 8;  // fp32 elements per row of the gradient tile

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
static inline size_t dwa_smem_bytes(int B, int tiles_per_cta) {
  const int Bp = (B + 15) / 16 * 16;
  return (size_t)Bp * kDwaLdB * 2 + (size_t)Bp * (tiles_per_cta * kDwaTileRows + 8) * 2 + 2 * (size_t)kDwaTileRows * kDwaLdC * 4;
}

static __global__ void __launch_bounds__(kDwaThreads, 3)
dense_wgrad_adam_kernel(float* __restrict__ p_all, const float* __restrict__ g_all, float* __restrict__ mu_all,
                        float* __restrict__ nu_all, const int32_t* __restrict__ count, float lr, float b1, float b2, float eps,
                        __nv_bfloat16* __restrict__ shadow_all, int64_t w_off, int64_t n_total4,
                        const __nv_bfloat16* __restrict__ act, int64_t lda, const __nv_bfloat16* __restrict__ dz, int B, int Kin,
                        int N, int tiles_per_cta, int rest_ctas) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char dwa_smem[];
  __shared__ float s_c[2];
  adam_bias_corrections(count, b1, b2, s_c);
  const float ob1 = 1.0f - b1, ob2 = 1.0f - b2;
  if ((int)blockIdx.x < rest_ctas) {
    // ---- every other leaf: the plain update over [0, n_total4) minus the Dense kernel's range
    if (blockIdx.y != 0) return;
    __syncthreads();
    const float c1 = s_c[0], c2 = s_c[1];
    const int64_t skip_begin4 = w_off / 4, skip_len4 = (int64_t)Kin * N / 4;
    const int64_t n4 = n_total4 - skip_len4;
    const int64_t stride = (int64_t)rest_ctas * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += stride) {
      const int64_t i = j < skip_begin4 ? j : j + skip_len4;
      const float4 gv = reinterpret_cast<const float4*>(g_all)[i];
      float4 mv = reinterpret_cast<const float4*>(mu_all)[i];
      float4 vv = reinterpret_cast<const float4*>(nu_all)[i];
      float4 pv = reinterpret_cast<const float4*>(p_all)[i];
      ISDQN_ADAM_ELEM(gv, mv, vv, pv, x) ISDQN_ADAM_ELEM(gv, mv, vv, pv, y) ISDQN_ADAM_ELEM(gv, mv, vv, pv, z)
      ISDQN_ADAM_ELEM(gv, mv, vv, pv, w)
      reinterpret_cast<float4*>(mu_all)[i] = mv;
      reinterpret_cast<float4*>(nu_all)[i] = vv;
      reinterpret_cast<float4*>(p_all)[i] = pv;
      if (shadow_all) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
        reinterpret_cast<uint2*>(shadow_all)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
    return;
  }
  float* __restrict__ p = p_all + w_off;
  float* __restrict__ mu = mu_all + w_off;
  float* __restrict__ nu = nu_all + w_off;
  __nv_bfloat16* __restrict__ shadow = shadow_all ? shadow_all + w_off : nullptr;
  const int Bp = (B + 15) / 16 * 16;
  const int rows_cta = tiles_per_cta * kDwaTileRows;
  const int lda_s = rows_cta + 8;  // (lda_s / 8 is odd: the 8 rows of an ldmatrix land in 8 different 16-byte banks)
  __nv_bfloat16* s_dz = reinterpret_cast<__nv_bfloat16*>(dwa_smem);                                   // [Bp][kDwaLdB]
  __nv_bfloat16* s_act = s_dz + (size_t)Bp * kDwaLdB;                                                 // [Bp][lda_s]
  float* s_g = reinterpret_cast<float*>(s_act + (size_t)Bp * lda_s);                                  // [2][16][kDwaLdC]
  const int n_tiles = (Kin + kDwaTileRows - 1) / kDwaTileRows;
  const int tile0 = ((int)blockIdx.x - rest_ctas) * tiles_per_cta;
  const int tile1 = min(n_tiles, tile0 + tiles_per_cta);
  if (tile0 >= tile1) return;
  const int col0 = blockIdx.y * kDwaCols;
  const int row0 = tile0 * kDwaTileRows;
  for (int i = threadIdx.x; i < Bp * (kDwaCols / 8); i += kDwaThreads) {  // 16-byte pieces of dz (N % 8 == 0)
    const int b = i / (kDwaCols / 8), c = (i - b * (kDwaCols / 8)) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (b < B && col0 + c < N) v = *reinterpret_cast<const uint4*>(dz + (int64_t)b * N + col0 + c);
    *reinterpret_cast<uint4*>(s_dz + (size_t)b * kDwaLdB + c) = v;
  }
  for (int i = threadIdx.x; i < Bp * (rows_cta / 8); i += kDwaThreads) {  // 16-byte pieces of act (Kin, lda % 8 == 0)
    const int b = i / (rows_cta / 8), r = (i - b * (rows_cta / 8)) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (b < B && row0 + r < Kin) v = *reinterpret_cast<const uint4*>(act + (int64_t)b * lda + row0 + r);
    *reinterpret_cast<uint4*>(s_act + (size_t)b * lda_s + r) = v;
  }
  __syncthreads();
  const float c1 = s_c[0], c2 = s_c[1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;  // this thread's 4 columns / 4 rows of a tile
  const int col = col0 + cg * 4;
  const bool col_ok = col < N;
  // ldmatrix row addresses (see the fragment layouts of mma.m16n8k16): lane -> matrix lane / 8, row lane % 8
  const int lm = lane >> 3, lj = lane & 7;
  const __nv_bfloat16* a_base = s_act + (size_t)(lj + (lm >> 1) * 8) * lda_s + (lm & 1) * 8;           // + k0 * lda_s + m0
  const __nv_bfloat16* b_base = s_dz + (size_t)(lj + (lm & 1) * 8) * kDwaLdB + warp * 32 + (lm >> 1) * 8;  // + k0 * ld + 16 j
  for (int tile = tile0; tile < tile1; ++tile) {
    const int m0 = (tile - tile0) * kDwaTileRows;
    float* gt = s_g + (size_t)((tile - tile0) & 1) * kDwaTileRows * kDwaLdC;
    // ---- gradient tile: 16 rows x (32 columns per warp)
    {
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
      for (int k0 = 0; k0 < Bp; k0 += 16) {
        uint32_t af[4], bf0[4], bf1[4];
        ldmatrix_x4_trans(af, a_base + (size_t)k0 * lda_s + m0);
        ldmatrix_x4_trans(bf0, b_base + (size_t)k0 * kDwaLdB);
        ldmatrix_x4_trans(bf1, b_base + (size_t)k0 * kDwaLdB + 16);
        mma_bf16_m16n8k16(acc[0], af, bf0[0], bf0[1]);
        mma_bf16_m16n8k16(acc[1], af, bf0[2], bf0[3]);
        mma_bf16_m16n8k16(acc[2], af, bf1[0], bf1[1]);
        mma_bf16_m16n8k16(acc[3], af, bf1[2], bf1[3]);
      }
      const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float* dst = gt + (size_t)gq * kDwaLdC + warp * 32 + nt * 8 + tq * 2;
        *reinterpret_cast<float2*>(dst) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(dst + 8 * kDwaLdC) = make_float2(acc[nt][2], acc[nt][3]);
      }
    }
    // ---- streaming loads of this thread's 4 x 4 block (rows beyond Kin / columns beyond N: clamped loads, no stores)
    const int row = row0 + m0 + rg * 4;
    const int64_t i0 = (int64_t)min(row, Kin - 1) * N + (col_ok ? col : 0);
    float4 pv[4], mv[4], vv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t i = row + r < Kin ? i0 + (int64_t)r * N : i0;
      pv[r] = *reinterpret_cast<const float4*>(p + i);
      mv[r] = *reinterpret_cast<const float4*>(mu + i);
      vv[r] = *reinterpret_cast<const float4*>(nu + i);
    }
    __syncthreads();  // the gradient tile is complete (and, double-buffered, nobody still reads the one written next)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (col_ok && row + r < Kin) {
        const float4 g = *reinterpret_cast<const float4*>(gt + (size_t)(rg * 4 + r) * kDwaLdC + cg * 4);
        const int64_t i = i0 + (int64_t)r * N;
        ISDQN_ADAM_ELEM(g, mv[r], vv[r], pv[r], x) ISDQN_ADAM_ELEM(g, mv[r], vv[r], pv[r], y)
        ISDQN_ADAM_ELEM(g, mv[r], vv[r], pv[r], z) ISDQN_ADAM_ELEM(g, mv[r], vv[r], pv[r], w)
        *reinterpret_cast<float4*>(mu + i) = mv[r];
        *reinterpret_cast<float4*>(nu + i) = vv[r];
        *reinterpret_cast<float4*>(p + i) = pv[r];
        if (shadow) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(pv[r].x, pv[r].y), hi = __floats2bfloat162_rn(pv[r].z, pv[r].w);
          *reinterpret_cast<uint2*>(shadow + i) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
      }
    }
  }
}

static __global__ void count_inc_kernel(int32_t* count) { *count += 1; }

// isdqn.py:111-125: one CTA per kernel row (+ one for the bias): row[j] = row[j + A] for j < K*A
static __global__ void __launch_bounds__(256) shift_heads_kernel(float* kernel, float* bias, int n_in, int K, int A) {
  float* row = blockIdx.x < n_in ? kernel + (int64_t)blockIdx.x * (1 + K) * A : bias;
  const int n = K * A;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + j * 256;
    v[j] = i < n ? row[i + A] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + j * 256;
    if (i < n) row[i] = v[j];
  }
}

// isdqn.py:133-135: argmax over the actions of head 1+idx (first maximum, like jnp.argmax)
static __global__ void argmax_head_kernel(const float* __restrict__ q, int A, int head, int32_t* out) {
  if (threadIdx.x == 0) {
    const float* h = q + head * A;
    int best = 0;
    for (int a = 1; a < A; ++a)
      if (h[a] > h[best]) best = a;
    *out = best;
  }
}

// the greedy action of EVERY head of one row of Q-values (acting path: the head is chosen on the host afterwards)
static __global__ void argmax_heads_kernel(const float* __restrict__ q, int n_heads_total, int A, int32_t* out) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < n_heads_total) {
    const float* v = q + h * A;
    int best = 0;
    for (int a = 1; a < A; ++a)
      if (v[a] > v[best]) best = a;
    out[h] = best;
  }
}

}  // namespace isdqn
