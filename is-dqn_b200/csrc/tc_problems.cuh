// The GEMM-shaped pieces of the learner expressed as problems of the tcgen05 tile engine (tc_engine.cuh):
// operand gathers (im2col forward, transposed-data weight gradient, tap-gathered input gradient, plain matrices) and
// epilogues (fused bias + LayerNorm + ReLU, plain fp32 store).  bf16 operands, fp32 accumulation in TMEM.
//
// Gathers never divide: every CTA first builds, in shared memory, one table entry per 16-byte chunk of its
// reduction axis (offset of the chunk relative to the row's anchor pixel + the tap coordinates for the bounds
// test); the per-chunk work of a loader is then one LDS, two compares and one cp.async.
#pragma once
#include "tc_engine.cuh"

namespace isdqn {
namespace tc {

typedef __nv_bfloat16 bf16;

constexpr int kMaxChunks = 512;  // 16-byte chunks of a reduction axis (K <= 4096 elements)

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// x/255 (architectures/dqn.py:51) for a packed uint8 pair -> packed bf16 pair
__device__ __forceinline__ uint32_t norm_pair_bf16(uint32_t b0, uint32_t b1) {
  return pack_bf16(__fmul_rn((float)b0, 1.0f / 255.0f), __fmul_rn((float)b1, 1.0f / 255.0f));
}

struct ChunkEntry {
  int off;  // element offset of the chunk relative to the row's anchor
  int yx;   // (dy << 16) | dx : tap coordinates (or -1: chunk is past the end of the reduction axis)
};

// ------------------------------------------------------------------------------------------- generic loaders
// K-major operand from a row-major matrix [rows][ld] (reduction index contiguous): thread t fills row t.
__device__ __forceinline__ void load_rows_kmajor(const bf16* __restrict__ src, int64_t ld, int row0, int n_rows_valid,
                                                 int k0, int k_end, uint32_t stage, int tid, int rows_in_tile) {
  for (int r = tid; r < rows_in_tile; r += kThreads) {
    const int row = row0 + r;
    const bool rv = row < n_rows_valid;
    const bf16* base = src + (int64_t)(rv ? row : 0) * ld;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int k = k0 + 8 * c;
      const bool v = rv && k < k_end;
      cp_async16(stage + kmajor_off(r, c), base + (v ? k : 0), v);
    }
  }
}
// MN-major operand from a row-major matrix [k rows][ld] (row = reduction index, MN contiguous): 64 k-rows per stage
__device__ __forceinline__ void load_rows_mnmajor(const bf16* __restrict__ src, int64_t ld, int k0, int k_end, int mn0,
                                                  int mn_end, uint32_t stage, int tid, int mn_chunks) {
  for (int idx = tid; idx < 64 * mn_chunks; idx += kThreads) {
    const int kk = idx & 63, c = idx >> 6;
    const int k = k0 + kk, mn = mn0 + 8 * c;
    const bool v = k < k_end && mn < mn_end;
    cp_async16(stage + mnmajor_off(kk, c), src + (v ? (int64_t)k * ld + mn : 0), v);
  }
}

// plain fp32 store of the accumulator tile: thread t owns row t
template <int BN>
__device__ __forceinline__ void store_rows_f32(uint32_t tmem_lane_base, bool has_acc, float* __restrict__ dst, int64_t ld,
                                               bool row_valid, int n0, int n_end) {
#pragma unroll 1
  for (int cb = 0; cb < BN / 32; ++cb) {
    float v[32];
    if (has_acc) {
      tmem_ld32(tmem_lane_base + cb * 32, v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    if (row_valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const int n = n0 + cb * 32 + i;
        if (n + 3 < n_end) {
          *reinterpret_cast<float4*>(dst + n) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < n_end) dst[n + j] = v[i + j];
        }
      }
    }
  }
  (void)ld;
}

// --------------------------------------------------------------------------------------- plain GEMM (+ split-K)
// D[M][N] = sum_k A(m,k) B(n,k); A: K-major [M][lda] or MN-major [K][lda]; B likewise.  fp32 output (partials).
template <int BN_, bool A_MN_, bool B_MN_>
struct GemmTC {
  static constexpr int BN = BN_, STAGES = 3, EXTRA_BYTES = 0;
  static constexpr bool A_MN = A_MN_, B_MN = B_MN_;
  const bf16* A; int64_t lda;
  const bf16* B; int64_t ldb;
  float* C; int64_t ldc; int64_t split_stride;
  int M, N, K, chunks_per_split;
  struct Ctx {};
  __device__ void init(Ctx&, uint8_t*, int, int, int) const {}
  __device__ void k_range(int split, int& b, int& e) const {
    const int total = (K + kBK - 1) / kBK;
    b = split * chunks_per_split;
    e = min(total, b + chunks_per_split);
    if (e < b) e = b;
  }
  __device__ void load_a(const Ctx&, uint32_t stage, int kc, int tid) const {
    const int m0 = blockIdx.x * kBM;
    if (A_MN) load_rows_mnmajor(A, lda, kc * kBK, K, m0, M, stage, tid, kBM / 8);
    else load_rows_kmajor(A, lda, m0, M, kc * kBK, K, stage, tid, kBM);
  }
  __device__ void load_b(const Ctx&, uint32_t stage, int kc, int tid) const {
    const int n0 = blockIdx.y * BN;
    if (B_MN) load_rows_mnmajor(B, ldb, kc * kBK, K, n0, N, stage, tid, BN / 8);
    else load_rows_kmajor(B, ldb, n0, N, kc * kBK, K, stage, tid, BN);
  }
  __device__ void epilogue(const Ctx&, uint32_t tmem_lane_base, bool has_acc, int m0, int n0, int tid, int split) const {
    const int m = m0 + tid;
    float* dst = C + (int64_t)split * split_stride + (int64_t)(m < M ? m : 0) * ldc;
    store_rows_f32<BN>(tmem_lane_base, has_acc, dst, ldc, m < M, n0, N);
  }
};

// im2col chunk table shared by the conv forward and weight-gradient problems.
// bf16 input (Cin % 8 == 0): chunk = 8 channels of one tap.   uint8 input with Cin == 4: chunk = 2 taps (kx, kx+1).
// entry.off = ((ky*W + kx)*Cin + c0) elements, entry.yx = (ky << 16) | kx.
__device__ __forceinline__ void build_im2col_table(ChunkEntry* tab, int n_chunks, int K, int Cin, int ksz, int W, int tid) {
  for (int ch = tid; ch < n_chunks; ch += kThreads) {
    const int k = 8 * ch;
    ChunkEntry e;
    if (k < K) {
      const int c0 = k % Cin;
      const int t = k / Cin;
      const int kx = t % ksz, ky = t / ksz;
      e.off = (ky * W + kx) * Cin + c0;
      e.yx = (ky << 16) | kx;
    } else {
      e.off = 0;
      e.yx = -1;
    }
    tab[ch] = e;
  }
}

// one 16-byte chunk of an im2col row: `anchor` = ELEMENT offset of input element (iy0, ix0, 0) of the row's image
// relative to `base` (may be negative: it is only added to the pointer after the bounds test)
template <bool IN_U8>
__device__ __forceinline__ void gather_im2col_chunk(uint32_t dst, const uint8_t* base, int64_t anchor, bool row_valid, int iy0,
                                                    int ix0, const ChunkEntry e, int H, int W, int Cin, const void* any_valid_ptr) {
  const int ky = e.yx >> 16, kx = e.yx & 0xffff;
  const int iy = iy0 + ky, ix = ix0 + kx;
  const bool rowok = row_valid && e.yx >= 0 && (unsigned)iy < (unsigned)H;
  if (!IN_U8) {
    const bool v = rowok && (unsigned)ix < (unsigned)W;
    cp_async16(dst, v ? base + (anchor + e.off) * 2 : reinterpret_cast<const uint8_t*>(any_valid_ptr), v);
  } else {
    // Cin == 4: two adjacent taps of 4 uint8 channels each (kx even, ksz even: both in the same kernel row)
    const bool v0 = rowok && (unsigned)ix < (unsigned)W, v1 = rowok && (unsigned)(ix + 1) < (unsigned)W;
    const uint32_t p0 = v0 ? *reinterpret_cast<const uint32_t*>(base + (anchor + e.off)) : 0u;
    const uint32_t p1 = v1 ? *reinterpret_cast<const uint32_t*>(base + (anchor + e.off + 4)) : 0u;
    st_shared_v4(dst, norm_pair_bf16(p0 & 0xff, (p0 >> 8) & 0xff), norm_pair_bf16((p0 >> 16) & 0xff, p0 >> 24),
                 norm_pair_bf16(p1 & 0xff, (p1 >> 8) & 0xff), norm_pair_bf16((p1 >> 16) & 0xff, p1 >> 24));
  }
  (void)Cin;
}

// ------------------------------------------------------------------------------------------------ conv forward
// out[m][co] = ReLU(LN(sum_k im2col(x)[m][k] W[k][co] + bias)); A gathered K-major, B = W (HWIO = [K][Cout]) MN-major.
// IN_U8 requires Cin == 4 (the stacked Atari frames); bf16 input requires Cin % 8 == 0.
template <int BN_, bool IN_U8_>
struct ConvFwdTC {
  static constexpr int BN = BN_, STAGES = 3, EXTRA_BYTES = kMaxChunks * (int)sizeof(ChunkEntry);
  static constexpr bool A_MN = false, B_MN = true;
  const void* in0; const void* in1; int n_img0;
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x, M, K;
  const bf16* w;  // [K][Cout]
  const float* bias; const float* ln_g; const float* ln_b; int relu;
  bf16* out; float* xhat; float* rstd; int m_train;
  struct Ctx {
    const uint8_t* base;  // the half of concat(s, s') this row's image lives in
    int64_t anchor;       // element offset of input element (iy0, ix0, 0) of this row's image (can be negative)
    int iy0, ix0;
    bool valid;
    const ChunkEntry* tab;
  };
  __device__ void init(Ctx& c, uint8_t* extra, int m0, int, int tid) const {
    ChunkEntry* tab = reinterpret_cast<ChunkEntry*>(extra);
    build_im2col_table(tab, ((K + kBK - 1) / kBK) * 8, K, Cin, ksz, W, tid);  // padded to whole stages
    c.tab = tab;
    const int m = m0 + tid;
    // (plain locals, assigned to the context once at the end: conditional stores into the by-reference context were
    //  observed to be dropped by nvcc 12.9 at -O3)
    const bool valid = m < M;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(in0);
    int64_t anchor = 0;
    int iy0 = 0, ix0 = 0;
    if (valid) {
      const int img = m / (OH * OW);
      const int rem = m - img * (OH * OW);
      const int oy = rem / OW, ox = rem - oy * OW;
      int64_t li = img;
      if (img >= n_img0) {
        base = reinterpret_cast<const uint8_t*>(in1);
        li = img - n_img0;
      }
      iy0 = oy * stride - pad_y;
      ix0 = ox * stride - pad_x;
      anchor = ((li * H + iy0) * W + ix0) * (int64_t)Cin;
    }
    c.valid = valid;
    c.base = base;
    c.anchor = anchor;
    c.iy0 = iy0;
    c.ix0 = ix0;
  }
  __device__ void k_range(int, int& b, int& e) const { b = 0; e = (K + kBK - 1) / kBK; }
  __device__ void load_a(const Ctx& c, uint32_t stage, int kc, int tid) const {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
      gather_im2col_chunk<IN_U8_>(stage + kmajor_off(tid, ch), c.base, c.anchor, c.valid, c.iy0, c.ix0, c.tab[kc * 8 + ch], H, W,
                                  Cin, in0);
  }
  __device__ void load_b(const Ctx&, uint32_t stage, int kc, int tid) const {
    load_rows_mnmajor(w, Cout, kc * kBK, K, 0, Cout, stage, tid, BN / 8);
  }
  __device__ void epilogue(const Ctx&, uint32_t tmem_lane_base, bool, int m0, int, int tid, int) const {
    const int m = m0 + tid;
    float mean = 0.f, rs = 1.f;
    if (ln_g) {  // flax LayerNorm: var = max(0, E[x^2] - E[x]^2), eps = 1e-6
      float s = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        float v[32];
        tmem_ld32(tmem_lane_base + cb * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float z = v[i] + __ldg(bias + cb * 32 + i);
          s += z;
          s2 += z * z;
        }
      }
      mean = s / (float)BN;
      rs = rsqrtf(fmaxf(s2 / (float)BN - mean * mean, 0.f) + 1e-6f);
    }
    const bool valid = m < M;
    const bool save = valid && ln_g != nullptr && xhat != nullptr && m < m_train;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane_base + cb * 32, v);
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float y[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = cb * 32 + i + j;
          float z = v[i + j] + __ldg(bias + n);
          if (ln_g) {
            z = (z - mean) * rs;
            v[i + j] = z;  // normalised value, saved for the backward pass
            z = z * __ldg(ln_g + n) + __ldg(ln_b + n);
          }
          y[j] = relu ? fmaxf(z, 0.f) : z;
        }
        packed[i / 2] = pack_bf16(y[0], y[1]);
      }
      if (valid) {
        uint4* o = reinterpret_cast<uint4*>(out + (int64_t)m * BN + cb * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      }
      if (save) {
        float4* x = reinterpret_cast<float4*>(xhat + (int64_t)m * BN + cb * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    if (save) rstd[m] = rs;
  }
};

// --------------------------------------------------------------------------------------- conv weight gradient
// dW[kc][co] (partial of split z) = sum_{pixels of the split} im2col(x)[pix][kc] dz[pix][co]
// A' = the im2col rows taken MN-major (row index = reduction), B' = dz MN-major.
template <int BN_, bool IN_U8_>
struct ConvWgradTC {
  static constexpr int BN = BN_, STAGES = 3, EXTRA_BYTES = 16 * (int)sizeof(ChunkEntry);
  static constexpr bool A_MN = true, B_MN = true;
  const void* in;  // layer input of the rows with a backward pass
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x, M, K;
  const bf16* dz;  // [M][Cout]
  float* part;     // [splits][K][Cout]
  int chunks_per_split;
  struct Ctx {
    const ChunkEntry* tab;  // the 16 column chunks of this CTA's 128-column tile
  };
  __device__ void init(Ctx& c, uint8_t* extra, int m0, int, int tid) const {
    ChunkEntry* tab = reinterpret_cast<ChunkEntry*>(extra);
    if (tid < 16) {
      const int k = m0 + 8 * tid;
      ChunkEntry e;
      if (k < K) {
        const int c0 = k % Cin;
        const int t = k / Cin;
        const int kx = t % ksz, ky = t / ksz;
        e.off = (ky * W + kx) * Cin + c0;
        e.yx = (ky << 16) | kx;
      } else {
        e.off = 0;
        e.yx = -1;
      }
      tab[tid] = e;
    }
    c.tab = tab;
  }
  __device__ void k_range(int split, int& b, int& e) const {
    const int total = (M + kBK - 1) / kBK;
    b = split * chunks_per_split;
    e = min(total, b + chunks_per_split);
    if (e < b) e = b;
  }
  __device__ void load_a(const Ctx& c, uint32_t stage, int kc, int tid) const {
    const int kk = tid & 63, half = tid >> 6;
    const int m = kc * kBK + kk;  // pixel row (reduction index)
    const bool mv = m < M;
    int iy0 = 0, ix0 = 0;
    int64_t anchor = 0;
    if (mv) {
      const int img = m / (OH * OW);
      const int rem = m - img * (OH * OW);
      const int oy = rem / OW, ox = rem - oy * OW;
      iy0 = oy * stride - pad_y;
      ix0 = ox * stride - pad_x;
      anchor = (((int64_t)img * H + iy0) * W + ix0) * (int64_t)Cin;
    }
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const int ch = half * 8 + cc;  // 16-byte chunk of the MN (= im2col column) axis
      gather_im2col_chunk<IN_U8_>(stage + mnmajor_off(kk, ch), reinterpret_cast<const uint8_t*>(in), anchor, mv, iy0, ix0,
                                  c.tab[ch], H, W, Cin, in);
    }
  }
  __device__ void load_b(const Ctx&, uint32_t stage, int kc, int tid) const {
    load_rows_mnmajor(dz, Cout, kc * kBK, M, blockIdx.y * BN, Cout, stage, tid, BN / 8);
  }
  __device__ void epilogue(const Ctx&, uint32_t tmem_lane_base, bool has_acc, int m0, int n0, int tid, int split) const {
    const int k = m0 + tid;
    float* dst = part + ((int64_t)split * K + (k < K ? k : 0)) * Cout;
    store_rows_f32<BN>(tmem_lane_base, has_acc, dst, Cout, k < K, n0, Cout);
  }
};

// ---------------------------------------------------------------------------------------- conv input gradient
// dX[pix_in][c] = sum_{tap, co} dz[pix_out(pix_in, tap)][co] W[tap][c][co]; rows grouped by stride parity class
// (blockIdx.z) so that only taps that hit the pixel are multiplied.  A gathered K-major, B = W K-major per tap.
// Table entry per chunk k = (tap, co0): off = -(ty*OW + tx)*Cout + co0 (relative to the row's anchor output pixel),
// yx = (ty << 16) | tx, or -1 when the tap does not exist for this parity class.  A second table holds the weight
// offsets ((ky*ksz + kx)*Cin)*Cout + co0.
template <int BN_>
struct ConvDgradTC {
  static constexpr int BN = BN_, STAGES = 3, EXTRA_BYTES = kMaxChunks * (int)(sizeof(ChunkEntry) + sizeof(int));
  static constexpr bool A_MN = false, B_MN = false;
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x, n_img, taps, Kd;
  const bf16* dz;  // [n_img*OH*OW][Cout]
  const bf16* w;   // [ksz][ksz][Cin][Cout]
  float* dx;       // [n_img*H*W][Cin]
  struct Ctx {
    int img, oy, ox, pix;
    int64_t anchor;  // element offset of dz[(img, oy, ox, 0)]
    const ChunkEntry* tab;
    const int* wtab;
  };
  __device__ void init(Ctx& c, uint8_t* extra, int m0, int, int tid) const {
    const int s = stride;
    const int ry = blockIdx.z / s, rx = blockIdx.z % s;
    ChunkEntry* tab = reinterpret_cast<ChunkEntry*>(extra);
    int* wtab = reinterpret_cast<int*>(extra + kMaxChunks * sizeof(ChunkEntry));
    const int n_chunks = ((Kd + kBK - 1) / kBK) * 8;  // padded to whole stages
    for (int ch = tid; ch < n_chunks; ch += kThreads) {
      const int k = 8 * ch;
      const int co = k % Cout;
      const int t = k / Cout;
      const int tx = t % taps, ty = t / taps;
      const int ky = ry + s * ty, kx = rx + s * tx;
      ChunkEntry e;
      if (k < Kd && ky < ksz && kx < ksz) {
        e.off = -(ty * OW + tx) * Cout + co;
        e.yx = (ty << 16) | tx;
        wtab[ch] = ((ky * ksz + kx) * Cin) * Cout + co;
      } else {
        e.off = 0;
        e.yx = -1;
        wtab[ch] = -1;
      }
      tab[ch] = e;
    }
    c.tab = tab;
    c.wtab = wtab;
    const int iy_first = ((ry - pad_y) % s + s) % s, ix_first = ((rx - pad_x) % s + s) % s;
    const int ny = iy_first < H ? (H - iy_first + s - 1) / s : 0;
    const int nx = ix_first < W ? (W - ix_first + s - 1) / s : 0;
    const int rows = n_img * ny * nx;
    const int m = m0 + tid;
    // (plain locals, assigned to the context once at the end: conditional stores into the by-reference context were
    //  observed to be dropped by nvcc 12.9 at -O3)
    int img = -1, oy = 0, ox = 0, pix = 0;
    int64_t anchor = 0;
    if (m < rows) {
      img = m / (ny * nx);
      const int rem = m - img * (ny * nx);
      const int iyc = rem / nx, ixc = rem - iyc * nx;
      const int iy = iy_first + s * iyc, ix = ix_first + s * ixc;
      oy = (iy + pad_y - ry) / s;
      ox = (ix + pad_x - rx) / s;
      pix = (img * H + iy) * W + ix;
      anchor = (((int64_t)img * OH + oy) * OW + ox) * (int64_t)Cout;
    }
    c.img = img;
    c.oy = oy;
    c.ox = ox;
    c.pix = pix;
    c.anchor = anchor;
  }
  __device__ void k_range(int, int& b, int& e) const { b = 0; e = (Kd + kBK - 1) / kBK; }
  __device__ void load_a(const Ctx& c, uint32_t stage, int kc, int tid) const {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const ChunkEntry e = c.tab[kc * 8 + ch];
      const int oy = c.oy - (e.yx >> 16), ox = c.ox - (e.yx & 0xffff);
      const bool v = c.img >= 0 && e.yx >= 0 && (unsigned)oy < (unsigned)OH && (unsigned)ox < (unsigned)OW;
      cp_async16(stage + kmajor_off(tid, ch), v ? dz + (c.anchor + e.off) : dz, v);
    }
  }
  __device__ void load_b(const Ctx& c, uint32_t stage, int kc, int tid) const {
    const int n0 = blockIdx.y * BN;
    for (int r = tid; r < BN; r += kThreads) {
      const int cin = n0 + r;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const int wo = c.wtab[kc * 8 + ch];
        const bool v = cin < Cin && wo >= 0;
        cp_async16(stage + kmajor_off(r, ch), v ? w + wo + (int64_t)cin * Cout : w, v);
      }
    }
  }
  __device__ void epilogue(const Ctx& c, uint32_t tmem_lane_base, bool has_acc, int, int n0, int, int) const {
    float* dst = dx + (int64_t)c.pix * Cin;
    store_rows_f32<BN>(tmem_lane_base, has_acc, dst, Cin, c.img >= 0, n0, Cin);
  }
};

}  // namespace tc
}  // namespace isdqn
