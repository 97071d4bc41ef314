// fp32 Dense layers at training-batch sizes (<= 64 rows through the forward, <= 32 / 64 with a backward pass): the hidden
// layer of the Atari network is a [rows x 7744] x [7744 x 512] product whose 15.8 MB kernel is read ONCE per pass, so
// these are memory-shaped kernels (coalesced kernel reads, the small operand broadcast out of shared memory), not the
// tiled GEMM of gemm_strided_kernel, which at 32-64 rows spends its time in half-empty tiles and un-coalesced W^T reads.
// Same arithmetic (fp32 FMA, fixed summation order => deterministic); replaces DQNNet's nn.Dense forward
// (architectures/dqn.py:92-99) and the two halves of its gradient on the 1e-5 path.
#pragma once
#include "common.cuh"

namespace isdqn {

constexpr int kDsK = 112;  // reduction elements staged per pass of the forward kernel

// part[split][r][n] = sum_{k in split} x[r][k] * W[k][n];  rows <= 64, N % 4 == 0.  grid (ceil(N / 128), splits), 256
// threads: thread = (four columns c4 = tid % 32, row group tid / 32 of 8 rows): one 16-byte kernel load feeds 32 FMAs, and a
// warp reads 512 contiguous bytes of a kernel row.  x rows [0, n0) come from in0, the rest from in1.
static __global__ void __launch_bounds__(256, 2)
dense_fwd_small_kernel(const float* __restrict__ in0, const float* __restrict__ in1, int n0, int rows, int K, int N,
                       const float* __restrict__ W, float* __restrict__ part, int64_t split_stride, int k_per_split) {
  __shared__ float4 As[kDsK][17];  // [k][row / 4], one float4 of padding per k: the transposing stores below are 4-way, not 32-way, conflicted
  const int c4 = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * 128 + c4 * 4;
  const int k_begin = blockIdx.y * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = k_begin; k0 < k_end; k0 += kDsK) {
    // stage x[0..64)[k0 .. k0 + kDsK) transposed: k fastest across the lanes, so the global reads are coalesced
    // (all 28 loads of a thread are issued before the first store: one exposed round trip per pass, not four)
    float sv[64 * kDsK / 256];
#pragma unroll
    for (int j = 0; j < 64 * kDsK / 256; ++j) {
      const int idx = threadIdx.x + j * 256;
      const int kk = idx % kDsK, r = idx / kDsK;
      const int k = k0 + kk;
      float v = 0.f;
      if (r < rows && k < k_end) v = r < n0 ? __ldg(in0 + (int64_t)r * K + k) : __ldg(in1 + (int64_t)(r - n0) * K + k);
      sv[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 64 * kDsK / 256; ++j) {
      const int idx = threadIdx.x + j * 256;
      reinterpret_cast<float*>(&As[idx % kDsK][0])[idx / kDsK] = sv[j];
    }
    __syncthreads();
    const int kmax = min(kDsK, k_end - k0);
    for (int kk0 = 0; kk0 < kmax; kk0 += 8) {
      float4 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        w[u] = (n < N && kk0 + u < kmax) ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)(k0 + kk0 + u) * N + n))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 a = As[kk0 + u < kDsK ? kk0 + u : 0][rg * 2 + q];
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4& t = acc[4 * q + j];
            t.x = fmaf(av[j], w[u].x, t.x);
            t.y = fmaf(av[j], w[u].y, t.y);
            t.z = fmaf(av[j], w[u].z, t.z);
            t.w = fmaf(av[j], w[u].w, t.w);
          }
        }
      }
    }
    __syncthreads();
  }
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = rg * 8 + i;
      if (r < rows) *reinterpret_cast<float4*>(part + (int64_t)blockIdx.y * split_stride + (int64_t)r * N + n) = acc[i];
    }
  }
}

// dx[b][n] = sum_k dz[b][k] * W[n][k]   (W is the [N = in][K = out] kernel, rows contiguous in k);  B <= 32.
// grid ceil(N / 32), 256 threads: thread = (n = tid % 32, batch group tid / 32 of 4 rows).
static __global__ void __launch_bounds__(256)
dense_dgrad_small_kernel(const float* __restrict__ dz, const float* __restrict__ W, float* __restrict__ dx, int B, int N, int K) {
  __shared__ float Ws[64][33];    // [k][n]
  __shared__ float4 Ds[64][9];    // [k][b / 4] (+ one float4 of padding: 4-way conflicts on the transposing stores)
  const int nl = threadIdx.x & 31, bg = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float wr[8], dr[8];  // this thread's share of the next chunk (8 kernel elements, 8 dz elements)
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = threadIdx.x + j * 256;
      const int kk = idx & 63, nn = idx >> 6;  // kernel tile: k fastest across the lanes (coalesced)
      wr[j] = (n0 + nn < N && k0 + kk < K) ? __ldg(W + (int64_t)(n0 + nn) * K + k0 + kk) : 0.f;
      dr[j] = (nn < B && k0 + kk < K) ? __ldg(dz + (int64_t)nn * K + k0 + kk) : 0.f;  // dz tile: the same (row, k) mapping
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += 64) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = threadIdx.x + j * 256;
      Ws[idx & 63][idx >> 6] = wr[j];
      reinterpret_cast<float*>(&Ds[idx & 63][0])[idx >> 6] = dr[j];
    }
    __syncthreads();
    if (k0 + 64 < K) fetch(k0 + 64);
#pragma unroll 8
    for (int kk = 0; kk < 64; ++kk) {
      const float w = Ws[kk][nl];
      const float4 d = Ds[kk][bg];
      acc[0] = fmaf(d.x, w, acc[0]);
      acc[1] = fmaf(d.y, w, acc[1]);
      acc[2] = fmaf(d.z, w, acc[2]);
      acc[3] = fmaf(d.w, w, acc[3]);
    }
    __syncthreads();
  }
  if (n0 + nl < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = bg * 4 + i;
      if (b < B) dx[(int64_t)b * N + n0 + nl] = acc[i];
    }
  }
}

// dW[k][n] = sum_b x[b][k] * dz[b][n];  B <= 64, N % 4 == 0.  grid (ceil(Kin / 16), ceil(N / 512)), 128 threads: thread = four
// columns, 16 kernel rows per CTA; dz is read through L2 by every CTA (64 KB), the kernel gradient is written once.
static __global__ void __launch_bounds__(128)
dense_wgrad_small_kernel(const float* __restrict__ x, const float* __restrict__ dz, float* __restrict__ dw, int B, int Kin, int N) {
  __shared__ float4 Xs[64][4];  // [b][16 k / 4]
  const int k0 = blockIdx.x * 16;
  for (int idx = threadIdx.x; idx < 64 * 16; idx += 128) {
    const int kk = idx & 15, b = idx >> 4;
    reinterpret_cast<float*>(&Xs[b][0])[kk] = (b < B && k0 + kk < Kin) ? x[(int64_t)b * Kin + k0 + kk] : 0.f;
  }
  __syncthreads();
  const int n4 = blockIdx.y * 128 + threadIdx.x;
  const int N4 = N >> 2;
  if (n4 >= N4) return;
  float4 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* dz4 = reinterpret_cast<const float4*>(dz);
  for (int b0 = 0; b0 < B; b0 += 4) {
    float4 d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u] = b0 + u < B ? __ldg(dz4 + (int64_t)(b0 + u) * N4 + n4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = Xs[b0 + u][q];
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4& t = acc[4 * q + j];
          t.x = fmaf(av[j], d[u].x, t.x);
          t.y = fmaf(av[j], d[u].y, t.y);
          t.z = fmaf(av[j], d[u].z, t.z);
          t.w = fmaf(av[j], d[u].w, t.w);
        }
      }
    }
  }
  float4* dw4 = reinterpret_cast<float4*>(dw);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (k0 + i < Kin) dw4[(int64_t)(k0 + i) * N4 + n4] = acc[i];
}

}  // namespace isdqn
