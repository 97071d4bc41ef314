// The GEMM-shaped pieces of the learner expressed as problems of the tcgen05 tile engine (tc_engine.cuh):
// operand gathers (im2col forward, transposed-data weight gradient, tap-gathered input gradient, plain matrices) and
// epilogues (fused bias + LayerNorm + ReLU, plain fp32 store).  bf16 operands, fp32 accumulation in TMEM.
//
// Gather design (what ncu asked for):
//  * coalesced: 8 consecutive lanes fetch the 8 16-byte chunks of ONE 128-byte row (one cache line per 8 lanes instead
//    of one per lane) and write them into a 128B-swizzled UMMA stage, where those 8 chunks land in 8 different bank
//    groups — the same trick TMA + SWIZZLE_128B plays, done by hand because the rows are gathered (im2col);
//  * no per-element arithmetic: the producers build, once per CTA, a table with one entry per 16-byte chunk of the
//    reduction axis (offset relative to the row's anchor pixel + tap coordinates for the bounds test), and publish
//    the per-row anchors of each tile (or chunk) in shared memory; a gather task is then 2 LDS + 2 compares + 1 cp.async;
//  * uint8 frames (first conv): the 8-byte loads of all tasks of a thread are issued back to back, then converted
//    (x/255 -> bf16) and stored — the asm memory clobbers of the stores would otherwise serialise them.
// Addresses are formed as base + (anchor + offset) with a SIGNED element offset: out-of-range pointers are never
// materialised (nvcc 12.9 -O3 was observed to drop conditional stores into a context that held such pointers).
#pragma once
#include "tc_engine.cuh"

namespace isdqn {
namespace tc {

typedef __nv_bfloat16 bf16;

constexpr int kMaxChunks = 512;  // table entries: 16-byte chunks of a reduction axis (x stride-parity classes)

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// A uint8 pair -> packed bf16 pair holding the INTEGER values: 0..255 need 8 significant bits, bf16 has 8, so the
// conversion is exact (the high half of the fp32 encoding); the x/255 of architectures/dqn.py:51 is applied to the
// fp32 accumulator in the epilogue instead (acc_scale) — fewer instructions and no input rounding at all.
__device__ __forceinline__ uint32_t int_pair_bf16(uint32_t b0, uint32_t b1) {
  return __byte_perm(__float_as_uint((float)b0), __float_as_uint((float)b1), 0x7632);
}

struct ChunkEntry {
  int off;  // element offset of the chunk relative to the row's anchor
  int yx;   // (dy << 16) | dx : tap coordinates (or -1: no such chunk)
};
struct __align__(16) RowInfo {
  int64_t anchor;  // element offset of the row's anchor relative to its base pointer (can be negative)
  int y0;          // anchor coordinates; y0 = kInvalidRow for rows past the end
  short x0;
  short second;    // 1: the row's image lives in the second input pointer (the s' half of concat(s, s'))
};
constexpr int kInvalidRow = -(1 << 28);
// (y0, x0) of a row packed into one register for the per-thread row copies: an invalid row keeps a y0 no tap can lift
// back into the image
__device__ __forceinline__ int pack_row_yx(const RowInfo& ri) {
  const int y = ri.y0 == kInvalidRow ? -0x4000 : ri.y0;
  return (y << 16) | ((int)ri.x0 & 0xffff);
}
__device__ __forceinline__ int row_y(int yx) { return yx >> 16; }
__device__ __forceinline__ int row_x(int yx) { return (int)(short)(yx & 0xffff); }
constexpr int kRowInfoBytes = 2 * kBM * (int)sizeof(RowInfo);  // double buffered (tile / chunk parity)

// ------------------------------------------------------------------------------------------- generic loaders
// K-major operand from a row-major matrix [rows][ld] (reduction index contiguous); 8 lanes = one 128-byte row.
template <int ROWS, int PT, bool SW>
__device__ __forceinline__ void load_rows_kmajor(const bf16* __restrict__ src, int64_t ld, int row0, int n_rows_valid,
                                                 int k0, int k_end, uint32_t stage, int ptid) {
#pragma unroll
  for (int t = ptid; t < ROWS * 8; t += PT) {
    const int c = t & 7, r = t >> 3;
    const int row = row0 + r, k = k0 + 8 * c;
    const bool v = row < n_rows_valid && k < k_end;
    cp_async16(stage + kmajor_off<SW>(r, c), src + (v ? (int64_t)row * ld + k : 0), v);
  }
}
// MN-major operand from a row-major matrix [k rows][ld] (row = reduction index, MN contiguous): 64 k-rows per stage
template <int MN_CHUNKS, int PT, bool SW>
__device__ __forceinline__ void load_rows_mnmajor(const bf16* __restrict__ src, int64_t ld, int k0, int k_end, int mn0,
                                                  int mn_end, uint32_t stage, int ptid) {
#pragma unroll
  for (int t = ptid; t < 64 * MN_CHUNKS; t += PT) {
    const int c = t % MN_CHUNKS, kk = t / MN_CHUNKS;
    const int k = k0 + kk, mn = mn0 + 8 * c;
    const bool v = k < k_end && mn < mn_end;
    cp_async16(stage + mnmajor_off<SW>(kk, c), src + (v ? (int64_t)k * ld + mn : 0), v);
  }
}

// Coalesced store of a 32-row x 16-word block that the warp owns row-per-thread (the layout tcgen05.ld 32x32b delivers):
// staged in shared memory (20-word pitch), then written eight rows (8 x 64 contiguous bytes) per 128-bit store instruction:
// 4 full-sector store instructions instead of 4 that touch 32 half-used sectors each.
// row_ptr: where word 0 of the block goes in THIS thread's row; valid: this thread's row is stored; n_words: words of the
// block inside the matrix (ragged N).  Warp-collective: every lane must call it.
constexpr int kStagePitch = 20;  // words: 16 + 4 (keeps every row 16-byte aligned for the 128-bit accesses)
__device__ __forceinline__ void warp_store_block16(uint32_t* stg, const uint32_t (&w)[16], void* row_ptr, bool valid,
                                                   int n_words = 16) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  uint4* mine4 = reinterpret_cast<uint4*>(stg + lane * kStagePitch);
#pragma unroll
  for (int q = 0; q < 4; ++q) mine4[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  __syncwarp();
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  const unsigned long long mine = (unsigned long long)(uintptr_t)row_ptr;
  const int sub = lane >> 2, c4 = lane & 3;  // 8 rows x 4 quads per instruction: 8 x 64 contiguous bytes
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = 8 * it + sub;
    const unsigned long long rp = __shfl_sync(0xffffffffu, mine, r);
    const uint4 v = *reinterpret_cast<const uint4*>(stg + r * kStagePitch + 4 * c4);
    if (((vmask >> r) & 1u) && 4 * c4 < n_words) reinterpret_cast<uint4*>((uintptr_t)rp)[c4] = v;  // (n_words % 4 == 0)
  }
}

// fp32 store of the accumulator tile through warp_store_block16 (stg != nullptr) — thread t owns row t
template <int BN>
__device__ __forceinline__ void store_rows_f32_coalesced(uint32_t tmem_lane_base, float* __restrict__ dst, bool row_valid, int n0,
                                                         int n_end, uint32_t* stg, float scale = 1.0f) {
#pragma unroll 1
  for (int cb = 0; cb < BN / 32; ++cb) {
    float v[32];
    tmem_ld32(tmem_lane_base + cb * 32, v);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(v[16 * h + i] * scale);
      const int n = n0 + cb * 32 + 16 * h;
      const int left = n_end - n;
      warp_store_block16(stg, w, dst + n, row_valid && left > 0, left < 16 ? left : 16);
    }
  }
}

// plain fp32 store of the accumulator tile: thread t owns row t
template <int BN>
__device__ __forceinline__ void store_rows_f32(uint32_t tmem_lane_base, float* __restrict__ dst, bool row_valid, int n0,
                                               int n_end, float scale = 1.0f) {
#pragma unroll 1
  for (int cb = 0; cb < BN / 32; ++cb) {
    float v[32];
    tmem_ld32(tmem_lane_base + cb * 32, v);
    if (scale != 1.0f) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= scale;
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
}

// ---- LayerNorm / ReLU backward of the layer BELOW, fused into an input-gradient epilogue -----------------------------------
// The accumulator row of thread t is d(loss)/d(output) of one pixel of the layer below (all its C = BN channels), so its
// LayerNorm backward is thread-local; only the three column sums that become the bias / scale / shift gradients cross rows.
// Column sums of 32 values per lane over the 32 lanes of a warp with 31 shuffles (recursive halving: after the step with
// offset o a lane keeps the half of its values whose index has bit o equal to its own lane bit): lane l ends up with the
// sum of value l over the warp, in a fixed order.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

constexpr int ln_bwd_ep_floats(int bn) { return bn <= 64 ? (2 + 12 + 3) * bn : 0; }  // gamma | beta | part[4][3 bn] | acc[3 bn]
struct LnBwdFuse {
  const float* xhat = nullptr;  // [rows][C] normalised values of the layer below (null: no fusion, plain fp32 input gradient)
  const float* rstd = nullptr;  // [rows]
  const float* ln_g = nullptr;  // [C]
  const float* ln_b = nullptr;  // [C]
  bf16* dz16 = nullptr;         // [rows][C] gradient w.r.t. the pre-activation of the layer below
  float* colpart = nullptr;     // [CTAs of the launch][3][C] column sums (dz | dy * xhat | dy) of every CTA
};
struct LnBwdCtx {
  float* prm;   // shared memory: gamma[BN] | beta[BN]
  float* part;  // [4 warps][3 BN]
  float* acc;   // [3 BN] running column sums of this CTA
};
template <int BN>
__device__ __forceinline__ void ln_bwd_init(const LnBwdFuse& f, LnBwdCtx& c, float* ep_sm, int etid) {
  c.prm = ep_sm;
  c.part = ep_sm + 2 * BN;
  c.acc = ep_sm + 14 * BN;
  if (f.xhat) {
    for (int i = etid; i < BN; i += 32 * kEpilogueWarps) {
      ep_sm[i] = f.ln_g[i];
      ep_sm[BN + i] = f.ln_b[i];
    }
    for (int i = etid; i < 3 * BN; i += 32 * kEpilogueWarps) c.acc[i] = 0.f;
  }
}
// one tile: thread etid owns LayerNorm row m (valid or not); warp-collective and epilogue-collective (two named barriers).
// KEEP_XH: the normalised values stay in registers between the two passes (single-wave launches, 255 registers available);
// otherwise they are read a second time (L1 / L2 hits), which keeps two CTAs per SM resident in the throughput shape.
template <int BN, bool KEEP_XH>
__device__ __forceinline__ void ln_bwd_tile(const LnBwdFuse& f, const LnBwdCtx& c, uint32_t tmem_lane_base, bool valid, int64_t m,
                                            uint32_t* stg, int etid) {
  static_assert(BN == 32 || BN == 64, "fused LayerNorm backward: 32 or 64 channels");
  const int lane = etid & 31, warp = etid >> 5;
  const float4* xr = reinterpret_cast<const float4*>(f.xhat + (valid ? m : 0) * BN);
  float dy[BN];
  float xk[KEEP_XH ? BN : 1];
  float sg = 0.f, sgx = 0.f;
  // the normalised values of this thread's row, 16 at a time
  // (measured: staging these loads through shared memory for full-sector instructions — warp_load_block16 — made the
  // batch-4096 input-gradient kernels 20 % SLOWER: they are bound by the L1 LSU data pipe, and the staging adds a
  // st.shared + ld.shared per 16 bytes; direct 16-byte loads of the thread's own row it is)
  auto load_xhat16 = [&](int first, float (&x)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 x4 = valid ? __ldg(xr + first / 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      x[4 * q] = x4.x; x[4 * q + 1] = x4.y; x[4 * q + 2] = x4.z; x[4 * q + 3] = x4.w;
    }
  };
#pragma unroll
  for (int cb = 0; cb < BN / 32; ++cb) {
    float v[32];
    tmem_ld32(tmem_lane_base + cb * 32, v);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float xx[16];
      load_xhat16(cb * 32 + 16 * h, xx);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int ch = cb * 32 + 16 * h + i;
        const float gg = c.prm[ch], bb = c.prm[BN + ch];
        const float d = (valid && xx[i] * gg + bb > 0.f) ? v[16 * h + i] : 0.f;
        dy[ch] = d;
        if (KEEP_XH) xk[KEEP_XH ? ch : 0] = xx[i];
        const float g = d * gg;
        sg += g;
        sgx += g * xx[i];
      }
    }
  }
  const float inv_c = 1.0f / (float)BN;
  const float mg = sg * inv_c, mgx = sgx * inv_c;
  const float rs = valid ? f.rstd[m] : 0.f;
  named_bar_sync(2, 32 * kEpilogueWarps);  // the previous tile's partial sums have been folded into acc
#pragma unroll
  for (int cb = 0; cb < BN / 32; ++cb) {
    float xx[32];
    if (KEEP_XH) {
#pragma unroll
      for (int i = 0; i < 32; ++i) xx[i] = xk[KEEP_XH ? cb * 32 + i : 0];
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float x16[16];
        load_xhat16(cb * 32 + 16 * h, x16);
#pragma unroll
        for (int i = 0; i < 16; ++i) xx[16 * h + i] = x16[i];
      }
    }
    float a[32];
    // gradient w.r.t. the pre-activation: stored as bf16, column-summed for the bias gradient
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = rs * (dy[cb * 32 + i] * c.prm[cb * 32 + i] - mg - xx[i] * mgx);
    {
      uint32_t w[16];  // 32 bf16 = one 64-byte segment of the row
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16(a[2 * i], a[2 * i + 1]);
      warp_store_block16(stg, w, f.dz16 + (valid ? m : 0) * BN + cb * 32, valid);
    }
    const float r0 = warp_transpose_sum32(a);
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = dy[cb * 32 + i] * xx[i];  // LayerNorm scale gradient terms
    const float r1 = warp_transpose_sum32(a);
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = dy[cb * 32 + i];          // LayerNorm shift gradient terms
    const float r2 = warp_transpose_sum32(a);
    float* pw = c.part + warp * 3 * BN + cb * 32 + lane;
    pw[0] = r0;
    pw[BN] = r1;
    pw[2 * BN] = r2;
  }
  named_bar_sync(2, 32 * kEpilogueWarps);
  for (int t = etid; t < 3 * BN; t += 32 * kEpilogueWarps)
    c.acc[t] += (c.part[t] + c.part[3 * BN + t]) + (c.part[6 * BN + t] + c.part[9 * BN + t]);
}
template <int BN>
__device__ __forceinline__ void ln_bwd_finish(const LnBwdFuse& f, const LnBwdCtx& c, int cta, int etid) {
  if (!f.xhat) return;
  named_bar_sync(2, 32 * kEpilogueWarps);
  for (int t = etid; t < 3 * BN; t += 32 * kEpilogueWarps) f.colpart[(int64_t)cta * 3 * BN + t] = c.acc[t];
}

// --------------------------------------------------------------------------------------- plain GEMM (+ split-K)
// D[M][N] = sum_k A(m,k) B(n,k); A: K-major [M][lda] or MN-major [K][lda]; B likewise.  fp32 output (partials).
// WIDE_ (every problem): the latency shape for launches that are a single partial wave (batch 32) — 8 producer warps
// (the gather is bound by instruction issue: twice the warps, half the work each), an 8-stage ring for the narrow
// tiles, one CTA per SM.  Default: the throughput shape — 4 producer warps, 4 stages, two CTAs per SM when they fit.
template <int BN_, bool A_MN_, bool B_MN_, bool WIDE_ = false>
struct GemmTC {
  static constexpr int BN = BN_, STAGES = (WIDE_ && BN_ <= 64) ? 8 : 4, PRODUCER_WARPS = WIDE_ ? 8 : 4, EXTRA_BYTES = 0, EP_FLOATS = 0;
  static constexpr int MIN_CTAS = WIDE_ ? 1 : (BN_ <= 64 ? 2 : 1);
  static constexpr int PT = 32 * PRODUCER_WARPS;
  static constexpr bool A_MN = A_MN_, B_MN = B_MN_, CHUNK_SYNC = false, SYNC_STORES = false, B_SW = BN_ >= 64;
  const bf16* A; int64_t lda;
  const bf16* B; int64_t ldb;
  float* C; int64_t ldc; int64_t split_stride;
  int M, N, K, chunks_per_split;
  struct PCtx {
    int m0, n0;
  };
  struct ECtx {};
  __device__ void init_cta(uint8_t*, int) const {}
  __device__ void tile_producer(PCtx& c, uint8_t*, int m0, int n0, int, int, int) const {
    c.m0 = m0;
    c.n0 = n0;
  }
  __device__ void tile_rows(PCtx&, int) const {}
  __device__ void init_epilogue(ECtx&, float*, int) const {}
  __device__ void chunk_producer(PCtx&, uint8_t*, int, int, int) const {}
  __device__ void tile_epilogue(ECtx&, int, int, int, int) const {}
  __device__ void k_range(int split, int& b, int& e) const {
    const int total = (K + kBK - 1) / kBK;
    b = split * chunks_per_split;
    e = min(total, b + chunks_per_split);
  }
  __device__ void load_a(const PCtx& c, uint32_t stage, int kc, int ptid) const {
    if (A_MN) load_rows_mnmajor<kBM / 8, PT, true>(A, lda, kc * kBK, K, c.m0, M, stage, ptid);
    else load_rows_kmajor<kBM, PT, true>(A, lda, c.m0, M, kc * kBK, K, stage, ptid);
  }
  __device__ void load_b(const PCtx& c, uint32_t stage, int kc, int ptid) const {
    if (B_MN) load_rows_mnmajor<BN / 8, PT, B_SW>(B, ldb, kc * kBK, K, c.n0, N, stage, ptid);
    else load_rows_kmajor<BN, PT, B_SW>(B, ldb, c.n0, N, kc * kBK, K, stage, ptid);
  }
  __device__ void epilogue(const ECtx&, uint32_t tmem_lane_base, int m0, int n0, int split, int etid, uint32_t* stg) const {
    const int m = m0 + etid;
    float* dst = C + (int64_t)split * split_stride + (int64_t)(m < M ? m : 0) * ldc;
    store_rows_f32<BN>(tmem_lane_base, dst, m < M, n0, N);
  }
};

// The same GEMM with both operand stages filled by TMA (2-D tensor maps over the row-major matrices, 128-byte swizzle,
// out-of-range rows/columns zero-filled by the copy engine — ragged M, N and K need no predicates): one producer thread,
// no LSU traffic.  BN >= 64 (a 32-wide B stage is narrower than a swizzle atom: GemmTC handles it).
template <int BN_, bool A_MN_, bool B_MN_, bool WIDE_ = false>
struct GemmTmaTC {
  static constexpr int BN = BN_, STAGES = (WIDE_ && BN_ <= 64) ? 8 : 4, PRODUCER_WARPS = 1, EXTRA_BYTES = 0;
  static constexpr int EP_FLOATS = ln_bwd_ep_floats(BN_);
  static constexpr int MIN_CTAS = WIDE_ ? 1 : (BN_ <= 64 ? 2 : 1);
  static constexpr bool A_MN = A_MN_, B_MN = B_MN_, CHUNK_SYNC = false, SYNC_STORES = false, B_SW = true, TMA = true, EP_STAGE = true;
  static constexpr bool EP_FINISH = true;
  static_assert(BN_ % 64 == 0, "swizzled B stage");
  CUtensorMap tm_a;  // A K-major: dims {K, M}, box {64, 128};  A MN-major: dims {M, K}, box {64, 64}
  CUtensorMap tm_b;  // B K-major: dims {K, N}, box {64, BN};   B MN-major: dims {N, K}, box {64, 64}
  float* C; int64_t ldc; int64_t split_stride;
  int M, N, K, chunks_per_split;
  // Dense input gradient with the LayerNorm / ReLU backward of the layer below fused (ln.xhat != null, BN == its channel
  // count): row m of D holds ln_rows_per_m = N / BN pixels of BN channels each; tile column n0 is pixel n0 / BN
  LnBwdFuse ln;
  int ln_rows_per_m = 1;
  struct PCtx {
    int m0, n0;
  };
  struct ECtx {
    LnBwdCtx lc;
  };
  __device__ void tma_prefetch() const {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  __device__ void tma_tile(PCtx& c, int tx, int ty, int) const {
    c.m0 = tx * kBM;
    c.n0 = ty * BN;
  }
  __device__ uint32_t stage_tx_bytes(const PCtx&) const { return (uint32_t)(kABytes + BN * 128); }
  __device__ void k_range(int split, int& b, int& e) const {
    const int total = (K + kBK - 1) / kBK;
    b = split * chunks_per_split;
    e = min(total, b + chunks_per_split);
  }
  __device__ void tma_load(const PCtx& c, uint32_t stage_a, uint32_t stage_b, uint64_t* bar, int kc) const {
    const int m0 = c.m0, n0 = c.n0, k0 = kc * kBK;
    if (A_MN) {
      tma_load_2d(stage_a, &tm_a, m0, k0, bar);
      tma_load_2d(stage_a + 8192, &tm_a, m0 + 64, k0, bar);
    } else {
      tma_load_2d(stage_a, &tm_a, k0, m0, bar);
    }
    if (B_MN) {
#pragma unroll
      for (int g = 0; g < BN / 64; ++g) tma_load_2d(stage_b + g * 8192, &tm_b, n0 + 64 * g, k0, bar);
    } else {
      tma_load_2d(stage_b, &tm_b, k0, n0, bar);
    }
  }
  __device__ void init_epilogue(ECtx& e, float* ep_sm, int etid) const {
    if constexpr (BN <= 64) ln_bwd_init<BN>(ln, e.lc, ep_sm, etid);
  }
  __device__ void tile_epilogue(ECtx&, int, int, int, int) const {}
  __device__ void epilogue(const ECtx& e, uint32_t tmem_lane_base, int m0, int n0, int split, int etid, uint32_t* stg) const {
    const int m = m0 + etid;
    if constexpr (BN <= 64) {
      if (ln.xhat) {
        ln_bwd_tile<BN, WIDE_>(ln, e.lc, tmem_lane_base, m < M, (int64_t)m * ln_rows_per_m + n0 / BN, stg, etid);
        return;
      }
    }
    float* dst = C + (int64_t)split * split_stride + (int64_t)(m < M ? m : 0) * ldc;
    store_rows_f32_coalesced<BN>(tmem_lane_base, dst, m < M, n0, N, stg);
  }
  __device__ void finish_epilogue(const ECtx& e, int cta, int etid) const {
    if constexpr (BN <= 64) ln_bwd_finish<BN>(ln, e.lc, cta, etid);
  }
};

// im2col chunk table shared by the conv forward and weight-gradient problems.
// bf16 input (Cin % 8 == 0): chunk = 8 channels of one tap.   uint8 input with Cin == 4: chunk = 2 taps (kx, kx+1).
// entry.off = ((ky*W + kx)*Cin + c0) elements, entry.yx = (ky << 16) | kx.
__device__ __forceinline__ void build_im2col_table(ChunkEntry* tab, int n_chunks, int K, int Cin, int ksz, int W, int ptid,
                                                   int pt) {
  for (int ch = ptid; ch < n_chunks; ch += pt) {
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

// anchor of output pixel m of a conv: input element (oy*stride - pad_y, ox*stride - pad_x, 0) of its image
__device__ __forceinline__ RowInfo conv_row_info(int m, int M, int OH, int OW, int H, int W, int Cin, int stride, int pad_y,
                                                 int pad_x, int n_img0) {
  RowInfo ri;
  int64_t anchor = 0;
  int y0 = kInvalidRow, x0 = 0, second = 0;
  if (m < M) {
    const int img = m / (OH * OW);
    const int rem = m - img * (OH * OW);
    const int oy = rem / OW, ox = rem - oy * OW;
    int64_t li = img;
    if (img >= n_img0) {
      second = 1;
      li = img - n_img0;
    }
    y0 = oy * stride - pad_y;
    x0 = ox * stride - pad_x;
    anchor = ((li * H + y0) * W + x0) * (int64_t)Cin;
  }
  ri.anchor = anchor;
  ri.y0 = y0;
  ri.x0 = (short)x0;
  ri.second = (short)second;
  return ri;
}

// bf16 gather task: one 16-byte chunk (8 channels of one tap) of one im2col row
__device__ __forceinline__ void gather_chunk_bf16(uint32_t dst, const void* in0, const void* in1, const RowInfo ri,
                                                  const ChunkEntry e, int H, int W) {
  const int iy = ri.y0 + (e.yx >> 16), ix = ri.x0 + (e.yx & 0xffff);
  const bool v = e.yx >= 0 && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
  const bf16* base = reinterpret_cast<const bf16*>(ri.second ? in1 : in0);
  cp_async16(dst, v ? base + (ri.anchor + e.off) : reinterpret_cast<const bf16*>(in0), v);
}

// uint8 gather task, phase 1: fetch the 8 bytes (2 taps x 4 channels); phase 2 converts and stores.
__device__ __forceinline__ uint2 fetch_chunk_u8(const void* in0, const void* in1, const RowInfo ri, const ChunkEntry e, int H,
                                                int W) {
  const int iy = ri.y0 + (e.yx >> 16), ix = ri.x0 + (e.yx & 0xffff);
  const bool rowok = e.yx >= 0 && (unsigned)iy < (unsigned)H;
  const bool v0 = rowok && (unsigned)ix < (unsigned)W, v1 = rowok && (unsigned)(ix + 1) < (unsigned)W;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(ri.second ? in1 : in0) + (ri.anchor + e.off);
  uint2 r = make_uint2(0u, 0u);
  if (v0 && v1 && ((reinterpret_cast<uintptr_t>(p) & 7) == 0)) {
    r = __ldg(reinterpret_cast<const uint2*>(p));
  } else {
    if (v0) r.x = __ldg(reinterpret_cast<const uint32_t*>(p));
    if (v1) r.y = __ldg(reinterpret_cast<const uint32_t*>(p + 4));
  }
  return r;
}
__device__ __forceinline__ void store_chunk_u8(uint32_t dst, const uint2 p) {
  st_shared_v4(dst, int_pair_bf16(p.x & 0xff, (p.x >> 8) & 0xff), int_pair_bf16((p.x >> 16) & 0xff, p.x >> 24),
               int_pair_bf16(p.y & 0xff, (p.y >> 8) & 0xff), int_pair_bf16((p.y >> 16) & 0xff, p.y >> 24));
}

// Epilogue of a forward convolution for ONE accumulator row (thread = row): bias + LayerNorm + ReLU, bf16 activation
// out, fp32 x-hat and rstd for the rows with a backward pass.  Every lane of the warp must call it (tcgen05.ld).
struct ConvEpilogueArgs {
  const float* prm;  // shared memory: bias[BN] | ln_g[BN] | ln_b[BN]
  bool ln_g;         // LayerNorm present
  int relu;
  bf16* out; float* xhat; float* rstd; int m_train;
  float acc_scale;
};
// m_out: row index of the bf16 activation (differs from m when the activation is stored with a padded row pitch);
// zero_left: also write a zero pixel at m_out - 1 (the left padding column of that layout)
template <int BN>
__device__ __forceinline__ void conv_fwd_epilogue_row(const ConvEpilogueArgs& e, uint32_t tmem_lane_base, int64_t m, bool valid,
                                                      int64_t m_out, bool zero_left, uint32_t* stg = nullptr) {
    const float4* pb = reinterpret_cast<const float4*>(e.prm);
    const float4* pg = reinterpret_cast<const float4*>(e.prm + BN);
    const float4* pbeta = reinterpret_cast<const float4*>(e.prm + 2 * BN);
    float mean = 0.f, rs = 1.f;
    if (e.ln_g) {  // flax LayerNorm: var = max(0, E[x^2] - E[x]^2), eps = 1e-6
      float s = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        float v[32];
        tmem_ld32(tmem_lane_base + cb * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = pb[cb * 8 + q];
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float z = fmaf(v[4 * q + j], e.acc_scale, bb[j]);
            s += z;
            s2 += z * z;
          }
        }
      }
      mean = s / (float)BN;
      rs = rsqrtf(fmaxf(s2 / (float)BN - mean * mean, 0.f) + 1e-6f);
    }
    const bool save = valid && e.ln_g && e.xhat != nullptr && m < e.m_train;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane_base + cb * 32, v);
      uint32_t packed[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = pb[cb * 8 + q], g4 = pg[cb * 8 + q], e4 = pbeta[cb * 8 + q];
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float z = fmaf(v[4 * q + j], e.acc_scale, bb[j]);
          if (e.ln_g) {
            z = (z - mean) * rs;
            v[4 * q + j] = z;  // normalised value, saved for the backward pass
            z = z * gg[j] + ee[j];
          }
          y[j] = e.relu ? fmaxf(z, 0.f) : z;
        }
        packed[2 * q] = pack_bf16(y[0], y[1]);
        packed[2 * q + 1] = pack_bf16(y[2], y[3]);
      }
      if (stg) {  // coalesced through shared memory (warp-collective: no early outs above)
        warp_store_block16(stg, packed, e.out + m_out * BN + cb * 32, valid);
        if (e.xhat != nullptr && e.ln_g) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(v[16 * h + i]);
            warp_store_block16(stg, w, e.xhat + (int64_t)m * BN + cb * 32 + 16 * h, save);
          }
        }
        if (valid && zero_left) {
          uint4* o = reinterpret_cast<uint4*>(e.out + (m_out - 1) * BN + cb * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
        }
      } else {
        if (valid) {
          uint4* o = reinterpret_cast<uint4*>(e.out + m_out * BN + cb * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
          if (zero_left) {
#pragma unroll
            for (int q = 0; q < 4; ++q) (o - BN / 8)[q] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        if (save) {
          float4* x = reinterpret_cast<float4*>(e.xhat + (int64_t)m * BN + cb * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
    if (save) e.rstd[m] = rs;
  }

// The same row epilogue for the TMA-store shape: the bf16 activation row and the fp32 normalised row go into shared-memory
// tiles (128-byte-swizzled rows, or dense 64-byte rows for 32 bf16 channels) that ONE thread then hands to the copy engine
// (ConvFwdTmaTC::epilogue) — one st.shared per 16 bytes instead of st.shared + ld.shared + st.global through the LSU.
//   tile_act : [128 rows][BN bf16]                      tile_xhat: BN / 32 sub-tiles of [128 rows][32 fp32]
template <int BN>
__device__ __forceinline__ void conv_fwd_epilogue_row_ts(const ConvEpilogueArgs& e, uint32_t tmem_lane_base, int row, int64_t m,
                                                         bool valid, bool save, uint32_t tile_act, uint32_t tile_xhat) {
  static_assert(BN == 32 || BN == 64, "TMA-store epilogue: 32 or 64 output channels");
  const float4* pb = reinterpret_cast<const float4*>(e.prm);
  const float4* pg = reinterpret_cast<const float4*>(e.prm + BN);
  const float4* pbeta = reinterpret_cast<const float4*>(e.prm + 2 * BN);
  float mean = 0.f, rs = 1.f;
  if (e.ln_g) {  // flax LayerNorm: var = max(0, E[x^2] - E[x]^2), eps = 1e-6
    float s = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem_lane_base + cb * 32, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = pb[cb * 8 + q];
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float z = fmaf(v[4 * q + j], e.acc_scale, bb[j]);
          s += z;
          s2 += z * z;
        }
      }
    }
    mean = s / (float)BN;
    rs = rsqrtf(fmaxf(s2 / (float)BN - mean * mean, 0.f) + 1e-6f);
  }
  const int sw = row & 7;
#pragma unroll 1
  for (int cb = 0; cb < BN / 32; ++cb) {
    float v[32];
    tmem_ld32(tmem_lane_base + cb * 32, v);
    uint32_t packed[16];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b4 = pb[cb * 8 + q], g4 = pg[cb * 8 + q], e4 = pbeta[cb * 8 + q];
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w}, ee[4] = {e4.x, e4.y, e4.z, e4.w};
      float y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float z = fmaf(v[4 * q + j], e.acc_scale, bb[j]);
        if (e.ln_g) {
          z = (z - mean) * rs;
          v[4 * q + j] = z;  // normalised value, saved for the backward pass
          z = z * gg[j] + ee[j];
        }
        y[j] = e.relu ? fmaxf(z, 0.f) : z;
      }
      packed[2 * q] = pack_bf16(y[0], y[1]);
      packed[2 * q + 1] = pack_bf16(y[2], y[3]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // 32 bf16 = four 16-byte chunks of the activation row
      const uint32_t a = BN == 64 ? tile_act + row * 128 + (((cb * 4 + q) ^ sw) << 4) : tile_act + row * 64 + (q << 4);
      st_shared_v4(a, packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    }
    if (save) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {  // 32 fp32 = eight 16-byte chunks of one 128-byte-swizzled sub-tile row
        const uint32_t a = tile_xhat + cb * (kBM * 128) + row * 128 + ((q ^ sw) << 4);
        st_shared_v4(a, __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                     __float_as_uint(v[4 * q + 3]));
      }
    }
  }
  if (save && valid) e.rstd[m] = rs;
}

// ------------------------------------------------------------------------------------------------ conv forward
// out[m][co] = ReLU(LN(sum_k im2col(x)[m][k] W[k][co] + bias)); A gathered K-major, B = W (HWIO = [K][Cout]) MN-major.
// IN_U8 requires Cin == 4 (the stacked Atari frames); bf16 input requires Cin % 8 == 0.
// SEG4: one image row of a window is only 64 contiguous bytes (4 chunks; the 8x8x4 first convolution): the lanes of a
// warp then take 4 chunks x 8 CONSECUTIVE output pixels, whose windows overlap in memory, instead of 8 chunks (two
// image rows) x 4 pixels — a third of the distinct 128-byte lines per LDGSTS, which is what bounds that kernel (the
// L1 data pipe: profiles/r01_summary.md).
template <int BN_, bool IN_U8_, bool SEG4_ = false, bool WIDE_ = false>
struct ConvFwdTC {
  __device__ static int task_ch(int ptid) { return SEG4_ ? ((ptid >> 5) & 1) * 4 + (ptid & 3) : (ptid & 7); }
  __device__ static int task_r0(int ptid) { return SEG4_ ? (ptid >> 6) * 8 + ((ptid & 31) >> 2) : (ptid >> 3); }
  static constexpr int BN = BN_, STAGES = (WIDE_ && BN_ <= 64) ? 8 : 4, PRODUCER_WARPS = (IN_U8_ || WIDE_) ? 8 : 4;
  static constexpr int MIN_CTAS = WIDE_ ? 1 : (BN_ <= 64 ? 2 : 1);
  static constexpr int PT = 32 * PRODUCER_WARPS, TASKS = kBM * 8 / PT, EP_FLOATS = 3 * BN_;
  static constexpr int EXTRA_BYTES = kMaxChunks * (int)sizeof(ChunkEntry) + kRowInfoBytes;
  static constexpr bool A_MN = false, B_MN = true, CHUNK_SYNC = false, SYNC_STORES = IN_U8_, B_SW = BN_ >= 64;
  const void* in0; const void* in1; int n_img0;
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x, M, K;
  const bf16* w;  // [K][Cout]
  const float* bias; const float* ln_g; const float* ln_b; int relu;
  bf16* out; float* xhat; float* rstd; int m_train;
  float acc_scale;  // 1/255 when the A operand holds raw uint8 pixel values, else 1
  struct PCtx {
    const ChunkEntry* tab;
    const RowInfo* rows;
    // bf16 gather: thread ptid always fills 16-byte column (ptid & 7) of the rows (ptid >> 3) + i * PT/8
    uint64_t row_addr[TASKS];  // byte address of the row's anchor element
    int row_yx[TASKS];         // pack_row_yx
  };
  struct ECtx {
    const float* prm;  // shared memory: bias[BN] | ln_g[BN] | ln_b[BN]
  };
  __device__ void init_epilogue(ECtx& e, float* ep_sm, int etid) const {
    for (int i = etid; i < BN; i += kThreads) {
      ep_sm[i] = bias[i];
      ep_sm[BN + i] = ln_g ? ln_g[i] : 1.f;
      ep_sm[2 * BN + i] = ln_g ? ln_b[i] : 0.f;
    }
    e.prm = ep_sm;
  }
  __device__ void init_cta(uint8_t* extra, int ptid) const {
    build_im2col_table(reinterpret_cast<ChunkEntry*>(extra), ((K + kBK - 1) / kBK) * 8, K, Cin, ksz, W, ptid, PT);
  }
  __device__ void tile_rows(PCtx& c, int ptid) const {
    if (IN_U8_) return;
#pragma unroll
    for (int i = 0; i < TASKS; ++i) {
      const RowInfo ri = c.rows[task_r0(ptid) + i * (PT / 8)];
      c.row_addr[i] = (uint64_t)(uintptr_t)(ri.second ? in1 : in0) + (uint64_t)(ri.anchor * 2);
      c.row_yx[i] = pack_row_yx(ri);
    }
  }
  __device__ void tile_producer(PCtx& c, uint8_t* extra, int m0, int, int, int ptid, int ti) const {
    RowInfo* rows = reinterpret_cast<RowInfo*>(extra + kMaxChunks * sizeof(ChunkEntry)) + (ti & 1) * kBM;
    if (ptid < kBM) rows[ptid] = conv_row_info(m0 + ptid, M, OH, OW, H, W, Cin, stride, pad_y, pad_x, n_img0);
    c.tab = reinterpret_cast<const ChunkEntry*>(extra);
    c.rows = rows;
  }
  __device__ void chunk_producer(PCtx&, uint8_t*, int, int, int) const {}
  __device__ void tile_epilogue(ECtx&, int, int, int, int) const {}
  __device__ void k_range(int, int& b, int& e) const { b = 0; e = (K + kBK - 1) / kBK; }
  __device__ void load_a(const PCtx& c, uint32_t stage, int kc, int ptid) const {
    if (!IN_U8_) {
      const int ch = task_ch(ptid);
      const ChunkEntry e = c.tab[kc * 8 + ch];
      const int dy = e.yx >> 16, dx = e.yx & 0xffff;
      const int64_t eoff = (int64_t)e.off * 2;
      const uint32_t dst = stage + kmajor_off<true>(task_r0(ptid), ch);  // rows 8 apart share the swizzle phase
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int iy = row_y(c.row_yx[i]) + dy, ix = row_x(c.row_yx[i]) + dx;
        const bool v = e.yx >= 0 && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
        cp_async16(dst + i * (PT / 8) * 128, v ? reinterpret_cast<const void*>((uintptr_t)(c.row_addr[i] + eoff)) : in0, v);
      }
    } else {
      uint2 px[TASKS];
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int t = ptid + i * PT;
        px[i] = fetch_chunk_u8(in0, in1, c.rows[t >> 3], c.tab[kc * 8 + (t & 7)], H, W);
      }
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int t = ptid + i * PT;
        store_chunk_u8(stage + kmajor_off<true>(t >> 3, t & 7), px[i]);
      }
    }
  }
  __device__ void load_b(const PCtx&, uint32_t stage, int kc, int ptid) const {
    load_rows_mnmajor<BN / 8, PT, B_SW>(w, Cout, kc * kBK, K, 0, Cout, stage, ptid);
  }
  __device__ void epilogue(const ECtx& ec, uint32_t tmem_lane_base, int m0, int, int, int etid, uint32_t* stg) const {
    const int m = m0 + etid;
    ConvEpilogueArgs e;
    e.prm = ec.prm; e.ln_g = ln_g != nullptr; e.relu = relu; e.out = out; e.xhat = xhat; e.rstd = rstd; e.m_train = m_train;
    e.acc_scale = acc_scale;
    conv_fwd_epilogue_row<BN>(e, tmem_lane_base, m, m < M, m, false);
  }
};

// ------------------------------------------------------------------------------- conv forward, TMA-fed (stride 1)
// Same contraction and epilogue as ConvFwdTC, but no gather code: one M tile = `th` whole output rows of ONE image
// (th * OW <= 128 output pixels; accumulator rows beyond that are computed on stale shared memory and never stored); per
// 64-deep K chunk — one tap (ky, kx) x 64 input channels — a single 4-D tensor-map copy of the box {64 channels, OW, th,
// 1 image} starting at (c0, kx - pad, y0 + ky - pad, image) brings the whole A stage: rows land at 128-byte pitch with the
// hardware 128-byte swizzle (= kmajor_off<true>), pixels outside the image are zero-filled by the TMA unit.  The weight
// chunk is BN/64 2-D copies in the MN-major swizzled layout, or (B_KMAJOR_, for BN = 32: narrower than a swizzle atom
// the other way round) one copy of a TRANSPOSED weight matrix [Cout][K] in the K-major layout.  Nothing goes through the
// LSU.
// TS_ (BN <= 64): the epilogue stores the activation and the normalised values with TMA out of shared-memory tiles
// (`UTMASTG`): the LSU only sees one st.shared per 16 bytes.  The tiles cost 48 KB (24 KB at 32 channels) of shared memory,
// so the ring has 6 stages in the single-wave shape and the throughput shape runs one CTA per SM (its two TMEM
// accumulators still overlap the epilogue of tile i with the main loop of tile i + 1).
template <int BN_, bool WIDE_ = false, bool B_KMAJOR_ = false, bool TS_ = false>
struct ConvFwdTmaTC {
  static constexpr bool TS = TS_ && BN_ <= 64;
  static constexpr int BN = BN_, STAGES = TS ? (WIDE_ ? 6 : 4) : ((WIDE_ && BN_ <= 64) ? 8 : 4), PRODUCER_WARPS = 1;
  static constexpr int MIN_CTAS = (WIDE_ || TS) ? 1 : (BN_ <= 64 ? 2 : 1);
  static constexpr int EP_TILE_BYTES = TS ? kBM * BN_ * 2 + kBM * BN_ * 4 : 0;  // bf16 activation tile + fp32 normalised tile
  static constexpr bool EP_FINISH = true;
  static constexpr int EXTRA_BYTES = 0, EP_FLOATS = 3 * BN_;
  static constexpr bool A_MN = false, B_MN = !B_KMAJOR_, CHUNK_SYNC = false, SYNC_STORES = false, B_SW = true, TMA = true, EP_STAGE = true;
  static_assert(B_KMAJOR_ || BN_ % 64 == 0, "the MN-major weight stage is filled in 64-column swizzle groups");
  CUtensorMap tm_x;  // activations [N][H][W][Cin] bf16, box {64, OW, th, 1}, SWIZZLE_128B
  CUtensorMap tm_w;  // weights [K][Cout] bf16, box {64, 64} — or transposed [Cout][K], box {64, BN}
  int n_img, pix, OW, OH, th, tpi, cchunks;  // tpi = tiles per image, cchunks = Cin / 64
  int ksz_y, ksz_x, sy, pad_y, pad_x;        // taps and row stride OF THE VIEW the tensor map describes (column stride 1)
  int out_pitch;                             // 0, or pixels per stored output row (OW + 1: a zero column on the left)
  const float* bias; const float* ln_g; const float* ln_b; int relu;
  bf16* out; float* xhat; float* rstd; int m_train;
  float acc_scale;
  // TS: output tensor maps — activation [N][OH][OW or out_pitch][C] bf16, box {C, OW, th, 1} (128-byte swizzle at 64
  // channels, none at 32); normalised values [n_train images][OH][OW][C] fp32, box {32, OW, th, 1}, 128-byte swizzle
  CUtensorMap tm_out, tm_xhat;
  int n_train_img;
  struct PCtx {
    int img, y0;         // tile
    int ky, kx, cc;      // running tap / channel-chunk counters (chunks are visited in order)
  };
  struct ECtx {
    const float* prm;
    uint32_t tile;       // TS: shared-memory address of the epilogue tiles
  };
  __device__ void set_ep_tile(ECtx& e, uint32_t addr) const { e.tile = addr; }
  __device__ void tma_prefetch() const {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    if constexpr (TS) {
      tma_prefetch_desc(&tm_out);
      tma_prefetch_desc(&tm_xhat);
    }
  }
  __device__ void tma_tile(PCtx& c, int tile, int, int) const {
    c.img = tile / tpi;
    c.y0 = (tile - c.img * tpi) * th * sy - pad_y;
    c.ky = c.kx = c.cc = 0;
  }
  __device__ uint32_t stage_tx_bytes(const PCtx&) const { return (uint32_t)(th * OW * 128 + BN * 128); }
  __device__ void k_range(int, int& b, int& e) const { b = 0; e = ksz_y * ksz_x * cchunks; }
  __device__ void tma_load(PCtx& c, uint32_t stage_a, uint32_t stage_b, uint64_t* bar, int kc) const {
    tma_load_4d(stage_a, &tm_x, c.cc * 64, c.kx - pad_x, c.y0 + c.ky, c.img, bar);
    if (++c.cc == cchunks) {  // next chunk: next 64 channels, then next tap (kx fastest) — no divisions in the loop
      c.cc = 0;
      if (++c.kx == ksz_x) {
        c.kx = 0;
        ++c.ky;
      }
    }
    if (B_KMAJOR_) {
      tma_load_2d(stage_b, &tm_w, kc * 64, 0, bar);
    } else {
#pragma unroll
      for (int g = 0; g < BN / 64; ++g) tma_load_2d(stage_b + g * 8192, &tm_w, g * 64, kc * 64, bar);
    }
  }
  __device__ void init_epilogue(ECtx& e, float* ep_sm, int etid) const {
    for (int i = etid; i < BN; i += kThreads) {
      ep_sm[i] = bias[i];
      ep_sm[BN + i] = ln_g ? ln_g[i] : 1.f;
      ep_sm[2 * BN + i] = ln_g ? ln_b[i] : 0.f;
    }
    e.prm = ep_sm;
  }
  __device__ void tile_epilogue(ECtx&, int, int, int, int) const {}
  __device__ void epilogue(const ECtx& ec, uint32_t tmem_lane_base, int m0, int, int, int etid, uint32_t* stg) const {
    const int tile = m0 / kBM;  // (the engine numbers tiles in units of 128 rows)
    const int img = tile / tpi, y0 = (tile - img * tpi) * th;
    const int r = etid / OW, ox = etid - r * OW, oy = y0 + r;
    ConvEpilogueArgs e;
    e.prm = ec.prm; e.ln_g = ln_g != nullptr; e.relu = relu; e.out = out; e.xhat = xhat; e.rstd = rstd; e.m_train = m_train;
    e.acc_scale = acc_scale;
    const int64_t m = (int64_t)img * pix + oy * OW + ox;
    const int64_t m_out = out_pitch ? ((int64_t)img * OH + oy) * out_pitch + ox + 1 : m;
    const bool valid = r < th && oy < OH && img < n_img;
    if constexpr (TS) {
      const bool train_tile = xhat != nullptr && ln_g != nullptr && img < n_train_img;  // (uniform: a tile is one image)
      const uint32_t tile_act = ec.tile, tile_xhat = ec.tile + kBM * BN * 2;
      if (etid == 0) bulk_wait_read_all();      // the previous tile's stores have read the tiles
      named_bar_sync(2, 32 * kEpilogueWarps);
      conv_fwd_epilogue_row_ts<BN>(e, tmem_lane_base, etid, m, valid, train_tile, tile_act, tile_xhat);
      fence_proxy_async();                      // st.shared (generic proxy) -> the copy engine's reads (async proxy)
      named_bar_sync(2, 32 * kEpilogueWarps);
      if (etid == 0 && img < n_img) {
        tma_store_4d(&tm_out, tile_act, 0, out_pitch ? 1 : 0, y0, img);
        if (train_tile) {
#pragma unroll
          for (int cb = 0; cb < BN / 32; ++cb) tma_store_4d(&tm_xhat, tile_xhat + cb * (kBM * 128), cb * 32, 0, y0, img);
        }
        bulk_commit_group();
      }
      if (valid && out_pitch != 0 && ox == 0) {  // the zero column on the left of the padded layout
        uint4* o = reinterpret_cast<uint4*>(out + (m_out - 1) * BN);
#pragma unroll
        for (int q = 0; q < BN / 8; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      return;
    }
    conv_fwd_epilogue_row<BN>(e, tmem_lane_base, m, valid, m_out, out_pitch != 0 && ox == 0, stg);
  }
  __device__ void finish_epilogue(const ECtx&, int, int etid) const {
    if constexpr (TS) {
      if (etid == 0) bulk_wait_all();  // the stores have been performed before the CTA (and with it the grid) completes
    }
  }
};

// --------------------------------------------------------------------------------------- conv weight gradient
// dW[kc][co] (partial of split z) = sum_{pixels of the split} im2col(x)[pix][kc] dz[pix][co]
// A' = the im2col rows taken MN-major (row index = reduction), B' = dz MN-major.  The 64 reduction rows (pixels) of
// a chunk change every chunk, so their anchors are published per chunk (CHUNK_SYNC).
template <int BN_, bool IN_U8_, bool SEG4_ = false, bool WIDE_ = false>
struct ConvWgradTC {
  static constexpr int BN = BN_, STAGES = (WIDE_ && BN_ <= 64) ? 8 : 4, PRODUCER_WARPS = (IN_U8_ || WIDE_) ? 8 : 4;
  static constexpr int MIN_CTAS = WIDE_ ? 1 : (BN_ <= 64 ? 2 : 1);
  // (SEG4: as in ConvFwdTC — 4 chunks of one image row x 8 consecutive pixels per warp; warp w takes image row w & 3 of
  // the window, pixel group w >> 2)
  __device__ static int task_ch(int ptid) { return SEG4_ ? ((ptid >> 5) & 3) * 4 + (ptid & 3) : (ptid & 15); }
  __device__ static int task_k0(int ptid) { return SEG4_ ? (ptid >> 7) * 8 + ((ptid & 31) >> 2) : (ptid >> 4); }
  static constexpr int kTaskRowStep = SEG4_ ? (PRODUCER_WARPS / 4) * 8 : (32 * PRODUCER_WARPS) / 16;
  static constexpr int PT = 32 * PRODUCER_WARPS, TASKS = kBK * 16 / PT, EP_FLOATS = 0;
  static constexpr int EXTRA_BYTES = kMaxChunks * (int)sizeof(ChunkEntry) + kRowInfoBytes;
  static constexpr bool A_MN = true, B_MN = true, CHUNK_SYNC = true, SYNC_STORES = IN_U8_, B_SW = BN_ >= 64;
  const void* in;  // layer input of the rows with a backward pass
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x, M, K;
  const bf16* dz;  // [M][Cout]
  float* part;     // [splits][K][Cout]
  int chunks_per_split;
  float acc_scale;  // 1/255 when the im2col operand holds raw uint8 pixel values, else 1
  struct PCtx {
    const ChunkEntry* tab;  // the 16 column chunks of this tile
    const RowInfo* rows;    // the 64 pixel rows of the current chunk
    int n0;
  };
  struct ECtx {};
  __device__ void init_cta(uint8_t* extra, int ptid) const {  // whole K axis, padded to whole 128-column tiles
    build_im2col_table(reinterpret_cast<ChunkEntry*>(extra), ((K + kBM - 1) / kBM) * 16, K, Cin, ksz, W, ptid, PT);
  }
  __device__ void tile_producer(PCtx& c, uint8_t* extra, int m0, int n0, int, int, int) const {
    c.tab = reinterpret_cast<const ChunkEntry*>(extra) + m0 / 8;
    c.n0 = n0;
  }
  __device__ void tile_rows(PCtx&, int) const {}
  __device__ void init_epilogue(ECtx&, float*, int) const {}
  __device__ void chunk_producer(PCtx& c, uint8_t* extra, int kc, int ptid, int j) const {
    RowInfo* rows = reinterpret_cast<RowInfo*>(extra + kMaxChunks * sizeof(ChunkEntry)) + (j & 1) * kBM;
    if (ptid < kBK) rows[ptid] = conv_row_info(kc * kBK + ptid, M, OH, OW, H, W, Cin, stride, pad_y, pad_x, 1 << 30);
    c.rows = rows;
  }
  __device__ void tile_epilogue(ECtx&, int, int, int, int) const {}
  __device__ void k_range(int split, int& b, int& e) const {
    const int total = (M + kBK - 1) / kBK;
    b = split * chunks_per_split;
    e = min(total, b + chunks_per_split);
  }
  __device__ void load_a(const PCtx& c, uint32_t stage, int, int ptid) const {
    if (!IN_U8_) {
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int ch = task_ch(ptid), kk = task_k0(ptid) + i * kTaskRowStep;
        gather_chunk_bf16(stage + mnmajor_off<true>(kk, ch), in, in, c.rows[kk], c.tab[ch], H, W);
      }
    } else {
      uint2 px[TASKS];
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int t = ptid + i * PT;
        px[i] = fetch_chunk_u8(in, in, c.rows[t >> 4], c.tab[t & 15], H, W);
      }
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int t = ptid + i * PT;
        store_chunk_u8(stage + mnmajor_off<true>(t >> 4, t & 15), px[i]);
      }
    }
  }
  __device__ void load_b(const PCtx& c, uint32_t stage, int kc, int ptid) const {
    load_rows_mnmajor<BN / 8, PT, B_SW>(dz, Cout, kc * kBK, M, c.n0, Cout, stage, ptid);
  }
  __device__ void epilogue(const ECtx&, uint32_t tmem_lane_base, int m0, int n0, int split, int etid, uint32_t* stg) const {
    const int k = m0 + etid;
    float* dst = part + ((int64_t)split * K + (k < K ? k : 0)) * Cout;
    store_rows_f32<BN>(tmem_lane_base, dst, k < K, n0, Cout, acc_scale);
  }
};

// ---------------------------------------------------------------------------------------- conv input gradient
// dX[pix_in][c] = sum_{tap, co} dz[pix_out(pix_in, tap)][co] W[tap][c][co]; rows grouped by stride parity class
// (tile z) so that only taps that hit the pixel are multiplied.  A gathered K-major, B = W K-major per tap.
// Table entry per (class, chunk k = (tap, co0)): off = -(ty*OW + tx)*Cout + co0 (relative to the row's anchor output
// pixel), yx = (ty << 16) | tx, or -1 when the tap does not exist for this class; wtab = weight offset of the chunk.
template <int BN_, bool WIDE_ = false>
struct ConvDgradTC {
  static constexpr int BN = BN_, STAGES = (WIDE_ && BN_ <= 64) ? 8 : 4, PRODUCER_WARPS = WIDE_ ? 8 : 4;
  static constexpr int MIN_CTAS = WIDE_ ? 1 : (BN_ <= 64 ? 2 : 1);
  static constexpr int PT = 32 * PRODUCER_WARPS, TASKS = kBM * 8 / PT, EP_FLOATS = 0;
  static constexpr int EXTRA_BYTES = kMaxChunks * (int)(sizeof(ChunkEntry) + sizeof(int)) + kRowInfoBytes;
  static constexpr bool A_MN = false, B_MN = false, CHUNK_SYNC = false, SYNC_STORES = false, B_SW = BN_ >= 64;
  int H, W, Cin, OH, OW, Cout, ksz, stride, pad_y, pad_x, n_img, taps, Kd;
  const bf16* dz;  // [n_img*OH*OW][Cout]
  const bf16* w;   // [ksz][ksz][Cin][Cout]
  float* dx;       // [n_img*H*W][Cin]
  struct PCtx {
    int n0;
    const ChunkEntry* tab;
    const int* wtab;
    const RowInfo* rows;  // anchor = element offset of dz[(img, oy, ox, 0)], (y0, x0) = (oy, ox)
    uint64_t row_addr[TASKS];  // per-thread copies, as in ConvFwdTC
    int row_yx[TASKS];
  };
  __device__ void tile_rows(PCtx& c, int ptid) const {
#pragma unroll
    for (int i = 0; i < TASKS; ++i) {
      const RowInfo ri = c.rows[(ptid >> 3) + i * (PT / 8)];
      c.row_addr[i] = (uint64_t)(uintptr_t)dz + (uint64_t)(ri.anchor * 2);
      c.row_yx[i] = pack_row_yx(ri);
    }
  }
  struct ECtx {
    int pix;
    bool valid;
  };
  __device__ int chunks_padded() const { return ((Kd + kBK - 1) / kBK) * 8; }
  // input pixel m of parity class cls -> (valid, output anchor (oy, ox), input pixel index)
  __device__ void decode_row(int m, int cls, int& img, int& oy, int& ox, int& pix) const {
    const int s = stride;
    const int ry = cls / s, rx = cls % s;
    const int iy_first = ((ry - pad_y) % s + s) % s, ix_first = ((rx - pad_x) % s + s) % s;
    const int ny = iy_first < H ? (H - iy_first + s - 1) / s : 0;
    const int nx = ix_first < W ? (W - ix_first + s - 1) / s : 0;
    img = -1;
    oy = ox = pix = 0;
    if (m < n_img * ny * nx) {
      img = m / (ny * nx);
      const int rem = m - img * (ny * nx);
      const int iyc = rem / nx, ixc = rem - iyc * nx;
      const int iy = iy_first + s * iyc, ix = ix_first + s * ixc;
      oy = (iy + pad_y - ry) / s;
      ox = (ix + pad_x - rx) / s;
      pix = (img * H + iy) * W + ix;
    }
  }
  __device__ void init_cta(uint8_t* extra, int ptid) const {
    ChunkEntry* tab = reinterpret_cast<ChunkEntry*>(extra);
    int* wtab = reinterpret_cast<int*>(extra + kMaxChunks * sizeof(ChunkEntry));
    const int s = stride, n_chunks = chunks_padded();
    for (int i = ptid; i < s * s * n_chunks; i += PT) {
      const int cls = i / n_chunks, ch = i - cls * n_chunks;
      const int ry = cls / s, rx = cls % s;
      const int k = 8 * ch;
      const int co = k % Cout;
      const int t = k / Cout;
      const int tx = t % taps, ty = t / taps;
      const int ky = ry + s * ty, kx = rx + s * tx;
      ChunkEntry e;
      int wo = -1;
      if (k < Kd && ky < ksz && kx < ksz) {
        e.off = -(ty * OW + tx) * Cout + co;
        e.yx = (ty << 16) | tx;
        wo = ((ky * ksz + kx) * Cin) * Cout + co;
      } else {
        e.off = 0;
        e.yx = -1;
      }
      tab[i] = e;
      wtab[i] = wo;
    }
  }
  __device__ void tile_producer(PCtx& c, uint8_t* extra, int m0, int n0, int cls, int ptid, int ti) const {
    RowInfo* rows = reinterpret_cast<RowInfo*>(extra + kMaxChunks * (sizeof(ChunkEntry) + sizeof(int))) + (ti & 1) * kBM;
    if (ptid < kBM) {
      int img, oy, ox, pix;
      decode_row(m0 + ptid, cls, img, oy, ox, pix);
      RowInfo ri;
      ri.anchor = img >= 0 ? (((int64_t)img * OH + oy) * OW + ox) * (int64_t)Cout : 0;
      ri.y0 = img >= 0 ? oy : kInvalidRow;
      ri.x0 = (short)ox;
      ri.second = 0;
      rows[ptid] = ri;
    }
    c.n0 = n0;
    c.tab = reinterpret_cast<const ChunkEntry*>(extra) + cls * chunks_padded();
    c.wtab = reinterpret_cast<const int*>(extra + kMaxChunks * sizeof(ChunkEntry)) + cls * chunks_padded();
    c.rows = rows;
  }
  __device__ void init_epilogue(ECtx&, float*, int) const {}
  __device__ void chunk_producer(PCtx&, uint8_t*, int, int, int) const {}
  __device__ void tile_epilogue(ECtx& e, int m0, int, int cls, int etid) const {
    int img, oy, ox, pix;
    decode_row(m0 + etid, cls, img, oy, ox, pix);
    e.pix = pix;
    e.valid = img >= 0;
  }
  __device__ void k_range(int, int& b, int& e) const { b = 0; e = (Kd + kBK - 1) / kBK; }
  __device__ void load_a(const PCtx& c, uint32_t stage, int kc, int ptid) const {
    const int ch = ptid & 7;
    const ChunkEntry e = c.tab[kc * 8 + ch];
    const int dy = e.yx >> 16, dx = e.yx & 0xffff;
    const int64_t eoff = (int64_t)e.off * 2;
    const uint32_t dst = stage + kmajor_off<true>(ptid >> 3, ch);
#pragma unroll
    for (int i = 0; i < TASKS; ++i) {
      const int oy = row_y(c.row_yx[i]) - dy, ox = row_x(c.row_yx[i]) - dx;
      const bool v = e.yx >= 0 && (unsigned)oy < (unsigned)OH && (unsigned)ox < (unsigned)OW;
      cp_async16(dst + i * (PT / 8) * 128, v ? reinterpret_cast<const void*>((uintptr_t)(c.row_addr[i] + eoff)) : dz, v);
    }
  }
  __device__ void load_b(const PCtx& c, uint32_t stage, int kc, int ptid) const {
#pragma unroll
    for (int t = ptid; t < BN * 8; t += PT) {
      const int ch = t & 7, r = t >> 3;
      const int cin = c.n0 + r;
      const int wo = c.wtab[kc * 8 + ch];
      const bool v = cin < Cin && wo >= 0;
      cp_async16(stage + kmajor_off<B_SW>(r, ch), v ? w + wo + (int64_t)cin * Cout : w, v);
    }
  }
  __device__ void epilogue(const ECtx& e, uint32_t tmem_lane_base, int, int n0, int, int, uint32_t* stg) const {
    store_rows_f32<BN>(tmem_lane_base, dx + (int64_t)e.pix * Cin, e.valid, n0, Cin);
  }
};

// ------------------------------------------------------------------------------- conv input gradient, TMA-fed
// One M tile = (image, stride-parity class): the class's input pixels form an ny x nx grid, and for tap (ty, tx) of the
// class the output pixels they touch are the same grid shifted by (-ty, -tx) — a plain box of dz, whatever the stride.
// Per chunk (valid tap x 64 output channels): one 4-D copy {64, nx, ny, 1} of dz at (co0, ox0 - tx, oy0 - ty, image) with
// out-of-range output pixels zero-filled, and one 2-D copy {64 co, BN input channels} of the tap's weights (K-major B).
// Up to 4 classes (stride <= 2): one tensor map per class because the box extents differ.
template <int BN_, bool WIDE_ = false>
struct ConvDgradTmaTC {
  static constexpr int BN = BN_, STAGES = (WIDE_ && BN_ <= 64) ? 8 : 4, PRODUCER_WARPS = 1;
  static constexpr int MIN_CTAS = WIDE_ ? 1 : (BN_ <= 64 ? 2 : 1);
  static constexpr int EXTRA_BYTES = 0, EP_FLOATS = ln_bwd_ep_floats(BN_);
  static constexpr bool A_MN = false, B_MN = false, CHUNK_SYNC = false, SYNC_STORES = false, B_SW = true, TMA = true, EP_STAGE = true;
  static constexpr bool EP_FINISH = true;
  CUtensorMap tm_dz[4];  // dz [N][OH][OW][Cout] bf16, box {64, nx(cls), ny(cls), 1}
  CUtensorMap tm_w;      // weights as [ksz*ksz*Cin rows][Cout], box {64, BN}
  int H, W, Cin, Cout, ksz, stride, pad_y, pad_x, n_img, cchunks;  // cchunks = Cout / 64
  float* dx;             // [n_img*H*W][Cin] (unused when the LayerNorm backward is fused)
  LnBwdFuse ln;          // xhat != null (BN == Cin <= 64): the epilogue goes on to the pre-activation gradient of the layer below
  struct Cls {
    int ry, rx, iy_first, ix_first, ny, nx, oy0, ox0, n_ty, n_tx;
  };
  struct PCtx {
    Cls c;
    int img, n0, cls;
    int ty, tx, cc;  // running counters over (tap row, tap column, 64-channel chunk)
  };
  struct ECtx {
    LnBwdCtx lc;
  };
  __device__ Cls cls_of(int cls) const {
    Cls c;
    const int s = stride;
    c.ry = cls / s;
    c.rx = cls % s;
    c.iy_first = ((c.ry - pad_y) % s + s) % s;
    c.ix_first = ((c.rx - pad_x) % s + s) % s;
    c.ny = c.iy_first < H ? (H - c.iy_first + s - 1) / s : 0;
    c.nx = c.ix_first < W ? (W - c.ix_first + s - 1) / s : 0;
    c.oy0 = (c.iy_first + pad_y - c.ry) / s;
    c.ox0 = (c.ix_first + pad_x - c.rx) / s;
    c.n_ty = c.ry < ksz ? (ksz - c.ry + s - 1) / s : 0;
    c.n_tx = c.rx < ksz ? (ksz - c.rx + s - 1) / s : 0;
    return c;
  }
  __device__ void tma_prefetch() const {
    for (int i = 0; i < stride * stride; ++i) tma_prefetch_desc(&tm_dz[i]);
    tma_prefetch_desc(&tm_w);
  }
  __device__ void tma_tile(PCtx& p, int img, int ty_n, int cls) const {
    p.c = cls_of(cls);
    p.img = img;
    p.n0 = ty_n * BN;
    p.cls = cls;
    p.ty = p.tx = p.cc = 0;
  }
  __device__ uint32_t stage_tx_bytes(const PCtx& p) const { return (uint32_t)(p.c.ny * p.c.nx * 128 + BN * 128); }
  __device__ void k_range(int cls, int& b, int& e) const {
    const Cls c = cls_of(cls);
    b = 0;
    e = c.n_ty * c.n_tx * cchunks;
  }
  __device__ void tma_load(PCtx& p, uint32_t stage_a, uint32_t stage_b, uint64_t* bar, int) const {
    const int ky = p.c.ry + stride * p.ty, kx = p.c.rx + stride * p.tx, co0 = p.cc * 64;
    tma_load_4d(stage_a, &tm_dz[p.cls], co0, p.c.ox0 - p.tx, p.c.oy0 - p.ty, p.img, bar);
    tma_load_2d(stage_b, &tm_w, co0, (ky * ksz + kx) * Cin + p.n0, bar);
    if (++p.cc == cchunks) {
      p.cc = 0;
      if (++p.tx == p.c.n_tx) {
        p.tx = 0;
        ++p.ty;
      }
    }
  }
  __device__ void init_epilogue(ECtx& e, float* ep_sm, int etid) const {
    if constexpr (BN <= 64) ln_bwd_init<BN>(ln, e.lc, ep_sm, etid);
  }
  __device__ void tile_epilogue(ECtx&, int, int, int, int) const {}
  __device__ void epilogue(const ECtx& e, uint32_t tmem_lane_base, int m0, int n0, int cls, int etid, uint32_t* stg) const {
    const Cls c = cls_of(cls);
    const int img = m0 / kBM;
    const bool valid = etid < c.ny * c.nx && img < n_img;
    const int iyc = etid / (c.nx > 0 ? c.nx : 1), ixc = etid - iyc * c.nx;
    const int64_t pix = valid ? ((int64_t)img * H + (c.iy_first + stride * iyc)) * W + (c.ix_first + stride * ixc) : 0;
    if constexpr (BN <= 64) {
      if (ln.xhat) {
        ln_bwd_tile<BN, WIDE_>(ln, e.lc, tmem_lane_base, valid, pix, stg, etid);
        return;
      }
    }
    store_rows_f32_coalesced<BN>(tmem_lane_base, dx + pix * Cin, valid, n0, Cin, stg);
  }
  __device__ void finish_epilogue(const ECtx& e, int cta, int etid) const {
    if constexpr (BN <= 64) ln_bwd_finish<BN>(ln, e.lc, cta, etid);
  }
};

// ------------------------------------------------------------------------------- conv weight gradient, TMA-fed
// dW[k][co] = sum over output pixels of im2col(x)[pix][k] dz[pix][co]; both operands MN-major, the reduction axis is the
// pixel axis.  A reduction chunk is a box of wb x rpc = 64 pixel positions of ONE image (wb = 16/32/64 columns >= OW,
// rpc = 64 / wb output rows): the box columns beyond OW are out of range for dz, so the copy engine zero-fills those rows
// of the dz stage and they contribute nothing, whatever the input stage holds there.  A 128-wide tile of k is two 64-wide
// groups = two (tap, 64-channel chunk) pairs, each ONE 4-D copy of the input view at the tap's offset (row traversal stride
// sy); the dz stage is Cout/64 copies.  63-69 % of the rows are real pixels (OW = 11: 44 of 64) — the price for having no
// gather at all.
template <int BN_, bool WIDE_ = false>
struct ConvWgradTmaTC {
  static constexpr int BN = BN_, STAGES = (WIDE_ && BN_ <= 64) ? 8 : 4, PRODUCER_WARPS = 1;
  static constexpr int MIN_CTAS = WIDE_ ? 1 : (BN_ <= 64 ? 2 : 1);
  static constexpr int EXTRA_BYTES = 0, EP_FLOATS = 0;
  static constexpr bool A_MN = true, B_MN = true, CHUNK_SYNC = false, SYNC_STORES = false, B_SW = true, TMA = true, EP_STAGE = true;
  static_assert(BN_ % 64 == 0, "swizzled dz stage");
  CUtensorMap tm_x;   // input view [N][H][W][Cin] bf16, box {64, wb, sy*(rpc-1)+1, 1}, traversal stride sy along H
  CUtensorMap tm_dz;  // dz [N][OH][OW][Cout] bf16, box {64, wb, rpc, 1}
  int K, Cout, cchunks, ksz_x, sy, pad_y, pad_x, rpc, cpi, total_chunks, chunks_per_split;  // cpi = chunks per image
  float* part;        // [splits][K][Cout]
  float acc_scale;
  struct PCtx {
    int c0[2], kx[2], ky[2], n_groups;  // the (up to) two 64-wide k groups of the tile
    int img, yc;                        // running reduction chunk: image, chunk inside the image
  };
  struct ECtx {};
  __device__ void tma_prefetch() const {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dz);
  }
  __device__ void tma_tile(PCtx& c, int tx, int, int split) const {
    c.n_groups = 0;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int gi = 2 * tx + g;
      const int tap = gi / cchunks;
      c.c0[g] = (gi - tap * cchunks) * 64;
      c.ky[g] = tap / ksz_x;
      c.kx[g] = tap - c.ky[g] * ksz_x;
      if (gi * 64 < K) c.n_groups = g + 1;
    }
    const int kb = split * chunks_per_split;
    c.img = kb / cpi;
    c.yc = kb - c.img * cpi;
  }
  __device__ uint32_t stage_tx_bytes(const PCtx& c) const { return (uint32_t)(c.n_groups * 8192 + (BN / 64) * 8192); }
  __device__ void k_range(int split, int& b, int& e) const {
    b = split * chunks_per_split;
    e = min(total_chunks, b + chunks_per_split);
  }
  __device__ void tma_load(PCtx& c, uint32_t stage_a, uint32_t stage_b, uint64_t* bar, int) const {
    const int y0 = c.yc * rpc;
    for (int g = 0; g < c.n_groups; ++g)
      tma_load_4d(stage_a + g * 8192, &tm_x, c.c0[g], c.kx[g] - pad_x, sy * y0 + c.ky[g] - pad_y, c.img, bar);
#pragma unroll
    for (int g = 0; g < BN / 64; ++g) tma_load_4d(stage_b + g * 8192, &tm_dz, 64 * g, 0, y0, c.img, bar);
    if (++c.yc == cpi) {
      c.yc = 0;
      ++c.img;
    }
  }
  __device__ void init_epilogue(ECtx&, float*, int) const {}
  __device__ void tile_epilogue(ECtx&, int, int, int, int) const {}
  __device__ void epilogue(const ECtx&, uint32_t tmem_lane_base, int m0, int n0, int split, int etid, uint32_t* stg) const {
    const int k = m0 + etid;
    float* dst = part + ((int64_t)split * K + (k < K ? k : 0)) * Cout;
    store_rows_f32_coalesced<BN>(tmem_lane_base, dst, k < K, n0, Cout, stg, acc_scale);
  }
};

}  // namespace tc
}  // namespace isdqn
