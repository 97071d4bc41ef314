// bf16 tensor-core learner path (tcgen05 / TMEM): the conv torso and the hidden Dense layers of
// iSDQN.learn_on_batch (slimdqn/networks/isdqn.py:82-103, architectures/dqn.py:55-103) as implicit GEMMs on the
// 5th-generation tensor cores; the tiny head layer, the K-head TD loss, LayerNorm backward, the deterministic
// reductions and Adam stay on the fp32 kernels of learner_kernels.cuh.  Tolerance of this path: 2e-2 (north_star).
#include <cstdio>
#include <cstdlib>
#include "learner_kernels.cuh"
#include "plan.cuh"
#include "tc_problems.cuh"

using namespace isdqn;
using isdqn::tc::bf16;

bool isdqn_dense_wgrad_adam_ok(int B, int Kin, int N, int64_t w_off);
int isdqn_dense_wgrad_adam_launch(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count,
                                  float lr, float b1, float b2, float eps, void* d_shadow_bf16, int64_t n_total, int64_t w_off,
                                  const void* d_act_bf16, int64_t lda, const void* d_dz_bf16, int B, int Kin, int N,
                                  void* stream);
int isdqn_dense_wgrad_adam_stream_launch(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count,
                                         float lr, float b1, float b2, float eps, void* d_shadow_bf16, int64_t n_total,
                                         int64_t w_off, const void* d_act_bf16, int64_t lda, const void* d_dz_bf16, int B,
                                         int Kin, int N, void* stream);
int isdqn_adam_launch(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count, float lr,
                      float b1, float b2, float eps, int64_t n, void* d_shadow_bf16, void* stream, int64_t skip_begin,
                      int64_t skip_len, int max_ctas = 0);
int isdqn_dp_allreduce_rest(void* comm, float* d_buf, int64_t n, int64_t skip_off, int64_t skip_n, void* stream);

namespace isdqn {
constexpr int kSideCtas = 64;  // grid cap of the tensor-core weight-gradient kernels that run beside the critical path
}

namespace isdqn {

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(tc::pack_bf16(v.x, v.y), tc::pack_bf16(v.z, v.w));
  }
}

// uint8 frames -> bf16 holding the exact integer pixel values 0..255 (the 1/255 of `x / 255` is applied to the fp32
// accumulator in the epilogues): each pixel is converted ONCE here instead of once per convolution window it is part of,
// and the first convolution then takes the same asynchronous-copy gather as the others.  16 pixels per thread.
__global__ void __launch_bounds__(256)
u8_frames_to_bf16_kernel(const uint8_t* __restrict__ s0, const uint8_t* __restrict__ s1, bf16* __restrict__ dst, int64_t n16_each) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n16_each; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = i < n16_each ? __ldg(reinterpret_cast<const uint4*>(s0) + i) : __ldg(reinterpret_cast<const uint4*>(s1) + (i - n16_each));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      o[2 * q] = tc::int_pair_bf16(w[q] & 0xff, (w[q] >> 8) & 0xff);
      o[2 * q + 1] = tc::int_pair_bf16((w[q] >> 16) & 0xff, w[q] >> 24);
    }
    uint4* d = reinterpret_cast<uint4*>(dst) + 2 * i;
    d[0] = make_uint4(o[0], o[1], o[2], o[3]);
    d[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

// The first convolution (8x8, stride 4, SAME => 2 pixels of padding, 4 stacked frames) as an ordinary 64-channel
// convolution: the frames are stored 4x4 space-to-depth with the padding folded in,
//     dst[n][Y][X][(py*4 + px)*4 + c] = frame[n][4Y - 2 + py][4X - 2 + px][c]   (0 outside the frame),  Y, X in [0, H/4 + 1),
// which turns conv0 into a 2x2, stride-1, unpadded convolution over (H/4+1) x (W/4+1) x 64: one tap is one 128-byte line
// (TMA-able; 4 shared-memory wavefronts per LDGSTS instead of 12.7 for the 64-byte window rows of the raw layout).
// One thread = one (Y, X, py): 16 input bytes -> 16 bf16.  Blocks >= frame_blocks re-order the conv0 kernel of the bf16
// shadow to the matching K order, as [K'][Cout] (wp) and transposed [Cout][K'] (wt).
__global__ void __launch_bounds__(256)
u8_frames_to_s2d_kernel(const uint8_t* __restrict__ s0, const uint8_t* __restrict__ s1, bf16* __restrict__ dst, int n_each, int H,
                        int W, int frame_blocks, const bf16* __restrict__ w_hwio, bf16* __restrict__ wp, bf16* __restrict__ wt,
                        int cout) {
  pdl_sync();
  if ((int)blockIdx.x >= frame_blocks) {
    const int n = 256 * cout;
    for (int i = ((int)blockIdx.x - frame_blocks) * 256 + threadIdx.x; i < n; i += ((int)gridDim.x - frame_blocks) * 256) {
      const bf16 v = w_hwio[s2d_to_hwio(i, cout)];
      const int kp = i / cout, co = i - kp * cout;
      wp[i] = v;
      wt[co * 256 + kp] = v;
    }
    return;
  }
  const int GY = H / 4 + 1, GX = W / 4 + 1;
  const int64_t total = (int64_t)2 * n_each * GY * GX * 4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)frame_blocks * 256) {
    const int py = (int)(i & 3);
    int64_t r = i >> 2;
    const int X = (int)(r % GX);
    r /= GX;
    const int Y = (int)(r % GY);
    const int n = (int)(r / GY);
    const int iy = 4 * Y - 2 + py, x0 = 4 * X - 2;
    const uint8_t* img = n < n_each ? s0 + (int64_t)n * H * W * 4 : s1 + (int64_t)(n - n_each) * H * W * 4;
    uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
    if ((unsigned)iy < (unsigned)H) {
      const uint8_t* p = img + ((int64_t)iy * W + x0) * 4;
      if (x0 >= 0) lo = __ldg(reinterpret_cast<const uint2*>(p));
      if (x0 + 3 < W) hi = __ldg(reinterpret_cast<const uint2*>(p + 8));
    }
    const uint32_t w[4] = {lo.x, lo.y, hi.x, hi.y};
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      o[2 * q] = tc::int_pair_bf16(w[q] & 0xff, (w[q] >> 8) & 0xff);
      o[2 * q + 1] = tc::int_pair_bf16((w[q] >> 16) & 0xff, w[q] >> 24);
    }
    uint4* d = reinterpret_cast<uint4*>(dst + (((int64_t)n * GY + Y) * GX + X) * 64 + py * 16);
    d[0] = make_uint4(o[0], o[1], o[2], o[3]);
    d[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

}  // namespace isdqn

namespace {

inline float* wsp(void* ws, int64_t off) { return off < 0 ? nullptr : reinterpret_cast<float*>(ws) + off; }

// Persistent launch: the tiles (x fastest) are dealt round-robin to min(#tiles, SMs x resident CTAs) CTAs.
template <class P>
int launch_tc(const P& p, int tiles_x, int tiles_y, int tiles_z, cudaStream_t s, const char* tag, int max_ctas = 0,
              int* grid_out = nullptr) {
  constexpr size_t smem = tc::smem_bytes_of<P>();
  constexpr int threads = 32 * (tc::kFirstProducerWarp + P::PRODUCER_WARPS);
  static PerDevice<int> ctas_per_sm_dev;  // (the opt-in attributes below are per device context)
  int& ctas_per_sm = ctas_per_sm_dev.get();
  if (!ctas_per_sm) {
    ISDQN_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_gemm_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ISDQN_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_gemm_kernel<P>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
    // Resident CTAs per SM from the kernel's own resources (228 KB shared memory with 1 KB reserved per CTA, 64 K
    // registers allocated per warp in units of 256, 2048 threads, 512 TMEM columns: every CTA owns two accumulators).
    // (cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for these kernels whatever the carve-out.)
    cudaFuncAttributes fa;
    ISDQN_CUDA_CHECK(cudaFuncGetAttributes(&fa, tc::tc_gemm_kernel<P>));
    const int by_smem = (int)((228 * 1024) / (smem + fa.sharedSizeBytes + 1024));
    const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
    const int by_regs = 65536 / (regs_per_warp * (threads / 32));
    const int by_threads = 2048 / threads;
    const int by_tmem = 512 / tc::tmem_cols_for(2 * P::BN);
    int occ = by_smem < by_regs ? by_smem : by_regs;
    if (occ > by_threads) occ = by_threads;
    if (occ > by_tmem) occ = by_tmem;
    if (getenv("ISDQN_DEBUG_OCC"))
      fprintf(stderr, "[isdqn] %s: regs %d static smem %zu dyn %zu threads %d -> smem %d regs %d thr %d tmem %d\n", tag,
              fa.numRegs, fa.sharedSizeBytes, smem, threads, by_smem, by_regs, by_threads, by_tmem);
    ctas_per_sm = occ < 1 ? 1 : occ;
  }
  const int64_t n_tiles = (int64_t)tiles_x * tiles_y * tiles_z;
  if (n_tiles < 1 || n_tiles > 0x7fffffff) return ISDQN_E_INVALID;
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  if (max_ctas > 0 && cap > max_ctas) cap = max_ctas;  // side-stream work: leave most SMs to the critical path
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
  if (grid_out) *grid_out = grid;
  ISDQN_PROF(s, tag);
  ISDQN_CUDA_CHECK(launch_pdl((tc::tc_gemm_kernel<P>), dim3(grid), dim3(threads), smem, s, p, tiles_x, tiles_y, tiles_z));
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

// A launch that is a single partial wave is latency bound (batch 32): it takes the WIDE shape of its problem
// (tc_problems.cuh).  ISDQN_WIDE=0 keeps the throughput shape everywhere.
bool wide_launch(int64_t n_tiles) {
  static const bool on = [] {
    const char* e = getenv("ISDQN_WIDE");
    return !(e && e[0] == '0');
  }();
  return on && n_tiles <= kNumSMs;
}

int pick_bn(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }
// widest N tile that still gives the grid about one CTA per SM (batch-32 shapes are latency bound: parallelism first)
int pick_bn_parallel(int n, int other_ctas) {
  int bn = pick_bn(n);
  while (bn > 32 && other_ctas * ceil_div(n, bn) < 120) bn >>= 1;
  return bn;
}

// ---- tensor maps (the driver entry point is fetched at run time: no link dependency on libcuda) -----------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    const char* e = getenv("ISDQN_TMA");
    if (e && e[0] == '0') return (EncodeTiledFn) nullptr;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}


// 2-D bf16 row-major matrix [rows][ld] as a tensor map with a {box_inner, box_rows} box and the 128-byte swizzle
int encode_matrix_map(CUtensorMap* tm, const bf16* base, int64_t inner, int64_t rows, int64_t ld, int box_inner, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  return tensor_map_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? ISDQN_OK
             : ISDQN_E_CUDA;
}

// D[M][N] (fp32, + split partials) = A B^T with the four operand-major combinations
bool gemm_tma_operands_ok(const bf16* A, int64_t lda, const bf16* B, int64_t ldb) {
  return tensor_map_encoder() != nullptr && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
         lda % 8 == 0 && ldb % 8 == 0;
}

// ln (optional): fuse the LayerNorm / ReLU backward of the layer whose output is D (64 channels per pixel, N / 64 pixels per
// row) into the epilogue — requires gemm_tma_operands_ok and splits == 1; *ln_parts receives the number of column partials
template <bool A_MN, bool B_MN>
int launch_gemm_tc(const bf16* A, int64_t lda, const bf16* B, int64_t ldb, float* C, int64_t ldc, int64_t split_stride, int M,
                   int N, int K, int splits, cudaStream_t s, const char* tag, int max_ctas = 0, const tc::LnBwdFuse* ln = nullptr,
                   int* ln_parts = nullptr) {
  const int total_chunks = ceil_div(K, tc::kBK);
  const int cps = ceil_div(total_chunks, splits);
  const int real_splits = ceil_div(total_chunks, cps);
  // (a capped side-stream launch takes the narrow tile: half the shared memory, so it can share an SM)
  const int bn = ln ? 64 : max_ctas > 0 ? pick_bn(N < 64 ? N : 64) : pick_bn_parallel(N, ceil_div(M, tc::kBM) * real_splits);
  if (ln && (real_splits != 1 || N % 64 != 0 || !gemm_tma_operands_ok(A, lda, B, ldb))) return ISDQN_E_UNSUPPORTED;
#define ISDQN_GEMM_TC_W(BN, WIDE)                                              \
  {                                                                            \
    tc::GemmTC<BN, A_MN, B_MN, WIDE> p;                                        \
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;          \
    p.split_stride = split_stride; p.M = M; p.N = N; p.K = K;                  \
    p.chunks_per_split = cps;                                                  \
    return launch_tc(p, ceil_div(M, tc::kBM), ceil_div(N, BN), real_splits, s, tag, max_ctas); \
  }
#define ISDQN_GEMM_TC(BN)                                                      \
  if (wide_launch((int64_t)ceil_div(M, tc::kBM) * ceil_div(N, BN) * real_splits)) ISDQN_GEMM_TC_W(BN, true) \
  else ISDQN_GEMM_TC_W(BN, false)
#define ISDQN_GEMM_TMA_W(BN, WIDE)                                             \
  {                                                                            \
    tc::GemmTmaTC<BN, A_MN, B_MN, WIDE> p;                                     \
    p.tm_a = tm_a; p.tm_b = tm_b; p.C = C; p.ldc = ldc;                        \
    p.split_stride = split_stride; p.M = M; p.N = N; p.K = K;                  \
    p.chunks_per_split = cps;                                                  \
    if (ln) { p.ln = *ln; p.ln_rows_per_m = N / BN; }                          \
    return launch_tc(p, ceil_div(M, tc::kBM), ceil_div(N, BN), real_splits, s, tag, max_ctas, ln_parts); \
  }
#define ISDQN_GEMM_TMA(BN)                                                     \
  if (wide_launch((int64_t)ceil_div(M, tc::kBM) * ceil_div(N, BN) * real_splits)) ISDQN_GEMM_TMA_W(BN, true) \
  else ISDQN_GEMM_TMA_W(BN, false)
  if (bn >= 64 && gemm_tma_operands_ok(A, lda, B, ldb)) {
    CUtensorMap tm_a, tm_b;
    int rc = A_MN ? encode_matrix_map(&tm_a, A, M, K, lda, 64, 64) : encode_matrix_map(&tm_a, A, K, M, lda, 64, tc::kBM);
    if (rc) return rc;
    rc = B_MN ? encode_matrix_map(&tm_b, B, N, K, ldb, 64, 64) : encode_matrix_map(&tm_b, B, K, N, ldb, 64, bn);
    if (rc) return rc;
    switch (bn) {
      case 64: ISDQN_GEMM_TMA(64)
      case 128: ISDQN_GEMM_TMA(128)
      default: ISDQN_GEMM_TMA(256)
    }
  }
#undef ISDQN_GEMM_TMA_W
#undef ISDQN_GEMM_TMA
  switch (bn) {
    case 32: ISDQN_GEMM_TC(32)
    case 64: ISDQN_GEMM_TC(64)
    case 128: ISDQN_GEMM_TC(128)
    default: ISDQN_GEMM_TC(256)
  }
#undef ISDQN_GEMM_TC_W
#undef ISDQN_GEMM_TC
}

// ---------------------------------------------------------------------------------------------- workspace
struct TcWorkspace {
  int64_t act16[ISDQN_MAX_FEATURES + 1];  // byte offsets; bf16 [rows*pix][out_dim] for every non-final layer
  int64_t dz16[ISDQN_MAX_FEATURES + 1];   // bf16 [B*pix][out_dim] gradient w.r.t. the pre-activation of every non-final layer
  int64_t x16;                            // bf16 [2B][H][W][4] integer-valued copy of the uint8 frames (or -1); with
                                          // frames_s2d: [2B][H/4+1][W/4+1][64] (4x4 space-to-depth, padding folded in)
  int64_t w0p, w0t;                       // frames_s2d: conv0 kernel in space-to-depth K order [256][Cout] / [Cout][256]
  bool pair_view;                         // act16[0] is stored [rows][OH][OW + 1][C] with a zero column on the left, so that the
                                          // second convolution reads it as [rows][OH][(OW + 1) / 2][2C] through TMA
  int64_t total;                          // bytes
};

// The first convolution reads a bf16 copy of the frames when a 16-byte gather chunk (two horizontally adjacent taps of
// the 4 stacked frames) can never straddle the left/right image border; otherwise it converts uint8 in its producers.
bool frames_as_bf16(const Layer& L) {
  return L.type == 0 && L.Cin == 4 && L.pad_x % 2 == 0 && L.W % 2 == 0 && L.ksz % 2 == 0 && L.stride % 2 == 0 &&
         ((int64_t)L.H * L.W * L.Cin) % 16 == 0;
}

// conv0 = 8x8 / stride 4 / padding 2 over 4 stacked frames: stored 4x4 space-to-depth it is a 2x2 stride-1 convolution over
// 64 channels (u8_frames_to_s2d_kernel).  Taken when the TMA path is available.
EncodeTiledFn tensor_map_encoder();
bool frames_s2d(const Layer& L) {
  return tensor_map_encoder() != nullptr && L.type == 0 && L.Cin == 4 && L.ksz == 8 && L.stride == 4 && L.pad_y == 2 &&
         L.pad_x == 2 && L.H % 4 == 0 && L.W % 4 == 0 && L.OH == L.H / 4 && L.OW == L.W / 4 && L.OW <= tc::kBM &&
         (L.out_dim == 32 || L.out_dim == 64 || L.out_dim == 128 || L.out_dim == 256);
}
// the same layer as the engine sees it on the space-to-depth image
Layer s2d_layer(const Layer& L) {
  Layer S = L;
  S.H = L.H / 4 + 1;
  S.W = L.W / 4 + 1;
  S.Cin = 64;
  S.ksz = 2;
  S.stride = 1;
  S.pad_y = S.pad_x = 0;
  return S;
}

// Second convolution (4x4, stride 2, SAME => 1 pixel of padding left/top) read through TMA: with its input stored with one
// zero column on the left, two horizontally adjacent pixels are one 2C-channel pixel of a view with half the width; the
// window columns 2ox-1 .. 2ox+2 are the view pixels ox, ox+1 (column stride 1, 2 taps), rows keep stride 2 (traversal
// stride of the tensor map).  The K order (ky, kx, c) is unchanged, so the weights are used as they are.  Needs conv0 on the
// space-to-depth TMA path (its epilogue writes the padded layout) with LayerNorm (nothing else reads act16[0] flat).
bool conv1_pair_view(const Plan& p) {
  if (p.n_layers < 2) return false;
  const Layer& A = p.L[0];
  const Layer& L = p.L[1];
  return frames_s2d(A) && A.has_ln && L.type == 0 && L.ksz == 4 && L.stride == 2 && L.pad_y == 1 && L.pad_x == 1 &&
         (L.Cin == 32 || L.Cin == 64 || L.Cin == 128) && (L.W + 1) % 2 == 0 && L.OW == (L.W + 1) / 2 && L.OW <= tc::kBM &&
         (L.out_dim == 64 || L.out_dim == 128 || L.out_dim == 256);
}

void carve_tc(const Plan& p, int rows, int B, TcWorkspace* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    const int64_t o = off;
    off = (off + bytes + 255) & ~(int64_t)255;
    return o;
  };
  w->pair_view = conv1_pair_view(p);
  for (int l = 0; l < p.n_layers; ++l) {
    const Layer& L = p.L[l];
    const int64_t pix = (l == 0 && w->pair_view) ? (int64_t)L.OH * (L.OW + 1) : L.pix;
    w->act16[l] = l + 1 < p.n_layers ? take((int64_t)rows * pix * L.out_dim * 2) : -1;
  }
  w->x16 = (!frames_s2d(p.L[0]) && frames_as_bf16(p.L[0])) ? take((int64_t)rows * p.L[0].H * p.L[0].W * p.L[0].Cin * 2) : -1;
  w->w0p = w->w0t = -1;
  if (frames_s2d(p.L[0])) {
    w->x16 = take((int64_t)rows * (p.L[0].H / 4 + 1) * (p.L[0].W / 4 + 1) * 64 * 2);
    w->w0p = take((int64_t)256 * p.L[0].out_dim * 2);
    w->w0t = take((int64_t)256 * p.L[0].out_dim * 2);
  }
  // one buffer per layer (not a ping-pong): the weight gradient of layer l runs on the side stream while the main
  // stream is already producing the gradients of the layers below it
  for (int l = 0; l < p.n_layers; ++l) {
    const Layer& L = p.L[l];
    w->dz16[l] = (B > 0 && l + 1 < p.n_layers) ? take((int64_t)B * L.pix * L.out_dim * 2) : -1;
  }
  w->total = off;
}

inline bf16* w16(void* ws, int64_t off) { return off < 0 ? nullptr : reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(ws) + off); }

bool tc_eligible(const Plan& p, const isdqn_net* net) {
  if (net->arch != ISDQN_ARCH_CNN || p.n_layers < 5) return false;
  if (net->obs_c != 4) return false;  // the uint8 gather of the first conv packs two 4-channel taps per 16-byte chunk
  for (int l = 0; l + 1 < p.n_layers; ++l) {
    const Layer& L = p.L[l];
    if (L.type == 0) {
      if (!(L.out_dim == 32 || L.out_dim == 64 || L.out_dim == 128 || L.out_dim == 256)) return false;
      const int taps = ceil_div(L.ksz, L.stride);
      // chunk tables: whole K axis for the weight gradient, (parity classes x tap chunks) for the input gradient
      if (ceil_div(L.in_dim, tc::kBM) * 16 > tc::kMaxChunks) return false;
      if (l > 0 && L.stride * L.stride * ceil_div(taps * taps * L.out_dim, tc::kBK) * 8 > tc::kMaxChunks) return false;  // (no input gradient for the first layer)
    } else {
      if (L.in_dim % 8 || L.out_dim % 64 || L.out_dim > kRowThreads * kRowMaxPerThread) return false;
    }
  }
  return true;
}

template <bool U8, bool SEG4 = false>
int launch_conv_fwd_tc(const Layer& L, const void* in0, const void* in1, int n0, int rows, const bf16* w, const float* params,
                       bf16* out, float* xhat, float* rstd, int m_train, cudaStream_t s, float in_scale = 1.0f) {
#define ISDQN_CONV_FWD_TC_W(BN, WIDE)                                                                  \
  {                                                                                                    \
    tc::ConvFwdTC<BN, U8, SEG4, WIDE> p;                                                               \
    p.in0 = in0; p.in1 = in1; p.n_img0 = n0;                                                           \
    p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.OH = L.OH; p.OW = L.OW; p.Cout = L.out_dim;                \
    p.ksz = L.ksz; p.stride = L.stride; p.pad_y = L.pad_y; p.pad_x = L.pad_x;                          \
    p.M = rows * L.pix; p.K = L.in_dim; p.w = w;                                                       \
    p.bias = params + L.b_off;                                                                         \
    p.ln_g = L.has_ln ? params + L.g_off : nullptr;                                                    \
    p.ln_b = L.has_ln ? params + L.beta_off : nullptr;                                                 \
    p.relu = L.relu; p.out = out; p.xhat = xhat; p.rstd = rstd; p.m_train = m_train;                   \
    p.acc_scale = U8 ? 1.0f / 255.0f : in_scale;                                                       \
    return launch_tc(p, ceil_div(p.M, tc::kBM), 1, 1, s, "tc_conv_fwd");                               \
  }
#define ISDQN_CONV_FWD_TC(BN)                                                                          \
  if (!U8 && wide_launch(ceil_div(rows * L.pix, tc::kBM))) ISDQN_CONV_FWD_TC_W(BN, true)               \
  else ISDQN_CONV_FWD_TC_W(BN, false)
  switch (L.out_dim) {
    case 32: ISDQN_CONV_FWD_TC(32)
    case 64: ISDQN_CONV_FWD_TC(64)
    case 128: ISDQN_CONV_FWD_TC(128)
    case 256: ISDQN_CONV_FWD_TC(256)
    default: return ISDQN_E_UNSUPPORTED;
  }
#undef ISDQN_CONV_FWD_TC_W
#undef ISDQN_CONV_FWD_TC
}

// ---- TMA-fed forward convolution (stride 1, Cin % 64 == 0, one image = one M tile) -----------------------------------
bool conv_fwd_tma_ok(const Layer& L) {
  return tensor_map_encoder() != nullptr && L.type == 0 && L.stride == 1 && L.Cin % 64 == 0 && L.OW <= tc::kBM && L.W <= 256 &&
         L.H <= 256 && (L.out_dim == 64 || L.out_dim == 128 || L.out_dim == 256);
}

// w: [K][Cout] (MN-major B), or — BN = 32 — the transposed kernel wt [Cout][K] (K-major B)
// `L` describes the convolution on the VIEW the tensor map is built over (H, W, Cin, taps ksz x ksz_x, row stride sy,
// column stride 1); ksz_x = 0 means a square stride-1 kernel.  out_pitch: see ConvFwdTmaTC.
int launch_conv_fwd_tma(const Layer& L, const bf16* x, int n_img, const bf16* w, const bf16* wt, const float* params, bf16* out,
                        float* xhat, float* rstd, int m_train, float in_scale, cudaStream_t s, int ksz_x = 0, int sy = 1,
                        int out_pitch = 0) {
  if (ksz_x == 0) ksz_x = L.ksz;
  EncodeTiledFn enc = tensor_map_encoder();
  CUtensorMap tm_x, tm_w;
  int th = tc::kBM / L.OW;  // whole output rows per M tile
  if (th > L.OH) th = L.OH;
  int tpi = ceil_div(L.OH, th);
  // few images (batch 32): more, shorter tiles so that the launch covers the SMs — the K loop of one CTA is bound by the
  // shared-memory fill rate of ONE SM, so halving the rows of a tile halves its latency
  while (n_img * tpi < 96 && th > 1) {
    ++tpi;
    th = ceil_div(L.OH, tpi);
  }
  tpi = ceil_div(L.OH, th);
  {
    const cuuint64_t dims[4] = {(cuuint64_t)L.Cin, (cuuint64_t)L.W, (cuuint64_t)L.H, (cuuint64_t)n_img};
    const cuuint64_t strides[3] = {(cuuint64_t)L.Cin * 2, (cuuint64_t)L.W * L.Cin * 2, (cuuint64_t)L.H * L.W * L.Cin * 2};
    // (traversal stride sy along H: the box spans sy*(th-1)+1 rows of which every sy-th is copied)
    const cuuint32_t box[4] = {64, (cuuint32_t)L.OW, (cuuint32_t)(sy * (th - 1) + 1), 1};
    const cuuint32_t es[4] = {1, 1, (cuuint32_t)sy, 1};
    if (enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISDQN_E_CUDA;
  }
  const bool kmajor_b = L.out_dim == 32;
  if (kmajor_b ? encode_matrix_map(&tm_w, wt, L.in_dim, L.out_dim, L.in_dim, 64, 32)
               : encode_matrix_map(&tm_w, w, L.out_dim, L.in_dim, L.out_dim, 64, 64))
    return ISDQN_E_CUDA;
  const int n_tiles = n_img * tpi;
  // TMA-store epilogue (<= 64 output channels): tensor maps of the two outputs
  static const bool ts_on = [] {
    const char* e = getenv("ISDQN_TMA_STORE");
    return !(e && e[0] == '0');
  }();
  // single-wave launches only: measured at batch 4096 the one-CTA-per-SM TMA-store shape loses to two staged-store CTAs per
  // SM (torso forward 0.66 vs 0.58 ms: with one CTA the four epilogue warps are latency bound), at batch 32 it is neutral
  // to slightly faster (129.7 vs 130.7 us per update)
  const bool ts = ts_on && wide_launch(n_tiles) && L.out_dim <= 64 && L.OW <= 256 && th <= 256;
  CUtensorMap tm_out = tm_x, tm_xhat = tm_x;
  const int n_train_img = (xhat != nullptr && L.has_ln) ? m_train / L.pix : 0;
  if (ts) {
    const int C = L.out_dim, OWp = out_pitch ? out_pitch : L.OW;
    {
      const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)OWp, (cuuint64_t)L.OH, (cuuint64_t)n_img};
      const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)OWp * C * 2, (cuuint64_t)L.OH * OWp * C * 2};
      const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)L.OW, (cuuint32_t)th, 1};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      if (enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return ISDQN_E_CUDA;
    }
    if (n_train_img > 0) {
      const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)L.OW, (cuuint64_t)L.OH, (cuuint64_t)n_train_img};
      const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)L.OW * C * 4, (cuuint64_t)L.OH * L.OW * C * 4};
      const cuuint32_t box[4] = {32, (cuuint32_t)L.OW, (cuuint32_t)th, 1};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      if (enc(&tm_xhat, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, xhat, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return ISDQN_E_CUDA;
    }
  }
#define ISDQN_CONV_FWD_TMA_W(BN, WIDE, KB, TS)                                                         \
  {                                                                                                    \
    tc::ConvFwdTmaTC<BN, WIDE, KB, TS> p;                                                              \
    p.tm_out = tm_out; p.tm_xhat = tm_xhat; p.n_train_img = n_train_img;                               \
    p.tm_x = tm_x; p.tm_w = tm_w; p.n_img = n_img; p.pix = L.pix; p.OW = L.OW; p.OH = L.OH;            \
    p.th = th; p.tpi = tpi; p.ksz_y = L.ksz; p.ksz_x = ksz_x; p.sy = sy; p.out_pitch = out_pitch;      \
    p.pad_y = L.pad_y; p.pad_x = L.pad_x; p.cchunks = L.Cin / 64;                                      \
    p.bias = params + L.b_off;                                                                         \
    p.ln_g = L.has_ln ? params + L.g_off : nullptr;                                                    \
    p.ln_b = L.has_ln ? params + L.beta_off : nullptr;                                                 \
    p.relu = L.relu; p.out = out; p.xhat = xhat; p.rstd = rstd; p.m_train = m_train;                   \
    p.acc_scale = in_scale;                                                                            \
    return launch_tc(p, n_tiles, 1, 1, s, "tc_conv_fwd_tma");                                          \
  }
#define ISDQN_CONV_FWD_TMA(BN, KB)                                                                     \
  if (ts) ISDQN_CONV_FWD_TMA_W(BN, true, KB, true)                                                     \
  else if (wide_launch(n_tiles)) ISDQN_CONV_FWD_TMA_W(BN, true, KB, false)                             \
  else ISDQN_CONV_FWD_TMA_W(BN, false, KB, false)
  switch (L.out_dim) {
    case 32: ISDQN_CONV_FWD_TMA(32, true)
    case 64: ISDQN_CONV_FWD_TMA(64, false)
    case 128: ISDQN_CONV_FWD_TMA(128, false)
    case 256: ISDQN_CONV_FWD_TMA(256, false)
    default: return ISDQN_E_UNSUPPORTED;
  }
#undef ISDQN_CONV_FWD_TMA_W
#undef ISDQN_CONV_FWD_TMA
}

// TMA-fed weight gradient.  V: the convolution on the view the input tensor map is built over (as in launch_conv_fwd_tma);
// dz is [n_img][OH][OW][Cout].  Partials come out as [real_splits][K][Cout].
bool conv_wgrad_tma_ok(const Layer& V) {
  return tensor_map_encoder() != nullptr && V.type == 0 && V.Cin % 64 == 0 && V.OW <= 64 && V.W <= 256 && V.H <= 256 &&
         (V.out_dim == 64 || V.out_dim == 128 || V.out_dim == 256);
}

int launch_conv_wgrad_tma(const Layer& V, const bf16* x, const bf16* dz, float* part, int n_img, int splits, int* real_splits,
                          cudaStream_t s, float in_scale, int ksz_x, int sy) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (ksz_x == 0) ksz_x = V.ksz;
  const int wb = V.OW <= 16 ? 16 : V.OW <= 32 ? 32 : 64, rpc = 64 / wb;
  const int cpi = ceil_div(V.OH, rpc);
  const int total_chunks = n_img * cpi;
  const int cps = ceil_div(total_chunks, splits);
  *real_splits = ceil_div(total_chunks, cps);
  CUtensorMap tm_x, tm_dz;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)V.Cin, (cuuint64_t)V.W, (cuuint64_t)V.H, (cuuint64_t)n_img};
    const cuuint64_t strides[3] = {(cuuint64_t)V.Cin * 2, (cuuint64_t)V.W * V.Cin * 2, (cuuint64_t)V.H * V.W * V.Cin * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)wb, (cuuint32_t)(sy * (rpc - 1) + 1), 1};
    const cuuint32_t es[4] = {1, 1, (cuuint32_t)sy, 1};
    if (enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISDQN_E_CUDA;
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)V.out_dim, (cuuint64_t)V.OW, (cuuint64_t)V.OH, (cuuint64_t)n_img};
    const cuuint64_t strides[3] = {(cuuint64_t)V.out_dim * 2, (cuuint64_t)V.OW * V.out_dim * 2, (cuuint64_t)V.OH * V.OW * V.out_dim * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)wb, (cuuint32_t)rpc, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tm_dz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(dz), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISDQN_E_CUDA;
  }
  const int tiles_k = ceil_div(V.in_dim, tc::kBM);
#define ISDQN_CONV_WGRAD_TMA_W(BN, WIDE)                                                               \
  {                                                                                                    \
    tc::ConvWgradTmaTC<BN, WIDE> p;                                                                    \
    p.tm_x = tm_x; p.tm_dz = tm_dz; p.K = V.in_dim; p.Cout = V.out_dim; p.cchunks = V.Cin / 64;        \
    p.ksz_x = ksz_x; p.sy = sy; p.pad_y = V.pad_y; p.pad_x = V.pad_x; p.rpc = rpc; p.cpi = cpi;        \
    p.total_chunks = total_chunks; p.chunks_per_split = cps; p.part = part; p.acc_scale = in_scale;    \
    return launch_tc(p, tiles_k, 1, *real_splits, s, "tc_conv_wgrad_tma");                             \
  }
#define ISDQN_CONV_WGRAD_TMA(BN)                                                                       \
  if (wide_launch((int64_t)tiles_k * *real_splits)) ISDQN_CONV_WGRAD_TMA_W(BN, true)                   \
  else ISDQN_CONV_WGRAD_TMA_W(BN, false)
  switch (V.out_dim) {
    case 64: ISDQN_CONV_WGRAD_TMA(64)
    case 128: ISDQN_CONV_WGRAD_TMA(128)
    case 256: ISDQN_CONV_WGRAD_TMA(256)
    default: return ISDQN_E_UNSUPPORTED;
  }
#undef ISDQN_CONV_WGRAD_TMA_W
#undef ISDQN_CONV_WGRAD_TMA
}

template <bool U8, bool SEG4 = false>
int launch_conv_wgrad_tc(const Layer& L, const void* in, const bf16* dz, float* part, int rows, int splits, int* real_splits,
                         cudaStream_t s, float in_scale = 1.0f, int max_ctas = 0) {
  const int total_chunks = ceil_div(rows, tc::kBK);
  const int cps = ceil_div(total_chunks, splits);
  *real_splits = ceil_div(total_chunks, cps);
#define ISDQN_CONV_WGRAD_TC_W(BN, WIDE)                                                                \
  {                                                                                                    \
    tc::ConvWgradTC<BN, U8, SEG4, WIDE> p;                                                             \
    p.in = in; p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.OH = L.OH; p.OW = L.OW; p.Cout = L.out_dim;     \
    p.ksz = L.ksz; p.stride = L.stride; p.pad_y = L.pad_y; p.pad_x = L.pad_x;                          \
    p.M = rows; p.K = L.in_dim; p.dz = dz; p.part = part; p.chunks_per_split = cps;                    \
    p.acc_scale = U8 ? 1.0f / 255.0f : in_scale;                                                       \
    return launch_tc(p, ceil_div(L.in_dim, tc::kBM), 1, *real_splits, s, "tc_conv_wgrad", max_ctas);   \
  }
#define ISDQN_CONV_WGRAD_TC(BN)                                                                        \
  if (!U8 && wide_launch((int64_t)ceil_div(L.in_dim, tc::kBM) * *real_splits)) ISDQN_CONV_WGRAD_TC_W(BN, true) \
  else ISDQN_CONV_WGRAD_TC_W(BN, false)
  switch (L.out_dim) {
    case 32: ISDQN_CONV_WGRAD_TC(32)
    case 64: ISDQN_CONV_WGRAD_TC(64)
    case 128: ISDQN_CONV_WGRAD_TC(128)
    case 256: ISDQN_CONV_WGRAD_TC(256)
    default: return ISDQN_E_UNSUPPORTED;
  }
#undef ISDQN_CONV_WGRAD_TC_W
#undef ISDQN_CONV_WGRAD_TC
}

// TMA-fed input gradient: stride <= 2, Cout % 64 == 0, every parity class of one image fits one M tile
bool conv_dgrad_tma_ok(const Layer& L) {
  if (tensor_map_encoder() == nullptr || L.type != 0 || L.stride > 2 || L.out_dim % 64 != 0) return false;
  if (!(L.Cin == 32 || L.Cin == 64 || L.Cin == 128 || L.Cin == 256)) return false;
  const int ny = ceil_div(L.H, L.stride), nx = ceil_div(L.W, L.stride);
  return ny * nx <= tc::kBM && nx <= 256 && ny <= 256;
}

int launch_conv_dgrad_tma(const Layer& L, const bf16* dz, const bf16* w, float* dx, int B, cudaStream_t s,
                          const tc::LnBwdFuse* ln = nullptr, int* ln_parts = nullptr) {
  EncodeTiledFn enc = tensor_map_encoder();
  CUtensorMap tm_dz[4], tm_w;
  const int st = L.stride;
  for (int cls = 0; cls < st * st; ++cls) {
    const int ry = cls / st, rx = cls % st;
    const int iy_first = ((ry - L.pad_y) % st + st) % st, ix_first = ((rx - L.pad_x) % st + st) % st;
    int ny = iy_first < L.H ? (L.H - iy_first + st - 1) / st : 0;
    int nx = ix_first < L.W ? (L.W - ix_first + st - 1) / st : 0;
    if (ny < 1) ny = 1;  // (an empty class issues no copies; the map only has to be valid)
    if (nx < 1) nx = 1;
    const cuuint64_t dims[4] = {(cuuint64_t)L.out_dim, (cuuint64_t)L.OW, (cuuint64_t)L.OH, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)L.out_dim * 2, (cuuint64_t)L.OW * L.out_dim * 2, (cuuint64_t)L.OH * L.OW * L.out_dim * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)nx, (cuuint32_t)ny, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tm_dz[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(dz), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISDQN_E_CUDA;
  }
  for (int cls = st * st; cls < 4; ++cls) tm_dz[cls] = tm_dz[0];
  const int bn = pick_bn(L.Cin);
  if (encode_matrix_map(&tm_w, w, L.out_dim, (int64_t)L.ksz * L.ksz * L.Cin, L.out_dim, 64, bn)) return ISDQN_E_CUDA;
#define ISDQN_CONV_DGRAD_TMA_W(BN, WIDE)                                                               \
  {                                                                                                    \
    tc::ConvDgradTmaTC<BN, WIDE> p;                                                                    \
    for (int i = 0; i < 4; ++i) p.tm_dz[i] = tm_dz[i];                                                 \
    p.tm_w = tm_w; p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.Cout = L.out_dim; p.ksz = L.ksz;             \
    p.stride = L.stride; p.pad_y = L.pad_y; p.pad_x = L.pad_x; p.n_img = B; p.cchunks = L.out_dim / 64; \
    p.dx = dx;                                                                                         \
    if (ln) p.ln = *ln;                                                                                \
    return launch_tc(p, B, ceil_div(L.Cin, BN), st * st, s, "tc_conv_dgrad_tma", 0, ln_parts);         \
  }
#define ISDQN_CONV_DGRAD_TMA(BN)                                                                       \
  if (wide_launch((int64_t)B * ceil_div(L.Cin, BN) * st * st)) ISDQN_CONV_DGRAD_TMA_W(BN, true)        \
  else ISDQN_CONV_DGRAD_TMA_W(BN, false)
  switch (bn) {
    case 32: ISDQN_CONV_DGRAD_TMA(32)
    case 64: ISDQN_CONV_DGRAD_TMA(64)
    case 128: ISDQN_CONV_DGRAD_TMA(128)
    default: ISDQN_CONV_DGRAD_TMA(256)
  }
#undef ISDQN_CONV_DGRAD_TMA_W
#undef ISDQN_CONV_DGRAD_TMA
}

// ---- weight gradient + input gradient of one layer in ONE launch (small batches) -------------------------------------------
// Both consume the same dz and neither reads what the other writes.  At batch 32 each of them is a fraction of a wave and
// costs a full launch latency on the critical path of the backward chain; together they fill the GPU once.
template <class P1, class P2>
int launch_tc2(const P1& p1, int t1x, int t1y, int t1z, int ctas1, const P2& p2, int t2x, int t2y, int t2z, int ctas2,
               cudaStream_t s, const char* tag) {
  constexpr size_t sm1 = tc::smem_bytes_of<P1>(), sm2 = tc::smem_bytes_of<P2>();
  constexpr size_t smem = sm1 > sm2 ? sm1 : sm2;
  constexpr int pw = P1::PRODUCER_WARPS > P2::PRODUCER_WARPS ? P1::PRODUCER_WARPS : P2::PRODUCER_WARPS;
  constexpr int threads = 32 * (tc::kFirstProducerWarp + pw);
  ISDQN_CUDA_CHECK(cudaFuncSetAttribute(tc::tc_gemm2_kernel<P1, P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ctas1 < 1 || ctas2 < 1) return ISDQN_E_INVALID;
  ISDQN_PROF(s, tag);
  ISDQN_CUDA_CHECK(launch_pdl((tc::tc_gemm2_kernel<P1, P2>), dim3(ctas1 + ctas2), dim3(threads), smem, s, p1, t1x, t1y, t1z, p2,
                              t2x, t2y, t2z, ctas1));
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

bool dp_overlap_on() {
  static const bool on = [] {
    const char* e = getenv("ISDQN_DP_OVERLAP");
    return !(e && e[0] == '0');
  }();
  return on;
}

bool pair_launch_enabled() {
  static const bool on = [] {
    const char* e = getenv("ISDQN_PAIR");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <int BNW>
int build_conv_wgrad_tma(const Layer& V, const bf16* x, const bf16* dz, float* part, int n_img, int splits, float in_scale,
                         int ksz_x, int sy, tc::ConvWgradTmaTC<BNW, true>* p, int* tiles_k, int* real_splits) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (ksz_x == 0) ksz_x = V.ksz;
  const int wb = V.OW <= 16 ? 16 : V.OW <= 32 ? 32 : 64, rpc = 64 / wb;
  const int cpi = ceil_div(V.OH, rpc);
  const int total_chunks = n_img * cpi;
  if (splits > total_chunks) splits = total_chunks;
  if (splits < 1) splits = 1;
  const int cps = ceil_div(total_chunks, splits);
  *real_splits = ceil_div(total_chunks, cps);
  {
    const cuuint64_t dims[4] = {(cuuint64_t)V.Cin, (cuuint64_t)V.W, (cuuint64_t)V.H, (cuuint64_t)n_img};
    const cuuint64_t strides[3] = {(cuuint64_t)V.Cin * 2, (cuuint64_t)V.W * V.Cin * 2, (cuuint64_t)V.H * V.W * V.Cin * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)wb, (cuuint32_t)(sy * (rpc - 1) + 1), 1};
    const cuuint32_t es[4] = {1, 1, (cuuint32_t)sy, 1};
    if (enc(&p->tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISDQN_E_CUDA;
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)V.out_dim, (cuuint64_t)V.OW, (cuuint64_t)V.OH, (cuuint64_t)n_img};
    const cuuint64_t strides[3] = {(cuuint64_t)V.out_dim * 2, (cuuint64_t)V.OW * V.out_dim * 2, (cuuint64_t)V.OH * V.OW * V.out_dim * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)wb, (cuuint32_t)rpc, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&p->tm_dz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(dz), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISDQN_E_CUDA;
  }
  *tiles_k = ceil_div(V.in_dim, tc::kBM);
  p->K = V.in_dim; p->Cout = V.out_dim; p->cchunks = V.Cin / 64;
  p->ksz_x = ksz_x; p->sy = sy; p->pad_y = V.pad_y; p->pad_x = V.pad_x; p->rpc = rpc; p->cpi = cpi;
  p->total_chunks = total_chunks; p->chunks_per_split = cps; p->part = part; p->acc_scale = in_scale;
  return ISDQN_OK;
}

template <int BND>
int build_conv_dgrad_tma(const Layer& L, const bf16* dz, const bf16* w, float* dx, int B, tc::ConvDgradTmaTC<BND, true>* p,
                         const tc::LnBwdFuse* ln = nullptr) {
  if (ln) p->ln = *ln;
  EncodeTiledFn enc = tensor_map_encoder();
  const int st = L.stride;
  for (int cls = 0; cls < st * st; ++cls) {
    const int ry = cls / st, rx = cls % st;
    const int iy_first = ((ry - L.pad_y) % st + st) % st, ix_first = ((rx - L.pad_x) % st + st) % st;
    int ny = iy_first < L.H ? (L.H - iy_first + st - 1) / st : 0;
    int nx = ix_first < L.W ? (L.W - ix_first + st - 1) / st : 0;
    if (ny < 1) ny = 1;
    if (nx < 1) nx = 1;
    const cuuint64_t dims[4] = {(cuuint64_t)L.out_dim, (cuuint64_t)L.OW, (cuuint64_t)L.OH, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)L.out_dim * 2, (cuuint64_t)L.OW * L.out_dim * 2, (cuuint64_t)L.OH * L.OW * L.out_dim * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)nx, (cuuint32_t)ny, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&p->tm_dz[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(dz), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return ISDQN_E_CUDA;
  }
  for (int cls = st * st; cls < 4; ++cls) p->tm_dz[cls] = p->tm_dz[0];
  if (encode_matrix_map(&p->tm_w, w, L.out_dim, (int64_t)L.ksz * L.ksz * L.Cin, L.out_dim, 64, BND)) return ISDQN_E_CUDA;
  p->H = L.H; p->W = L.W; p->Cin = L.Cin; p->Cout = L.out_dim; p->ksz = L.ksz;
  p->stride = L.stride; p->pad_y = L.pad_y; p->pad_x = L.pad_x; p->n_img = B; p->cchunks = L.out_dim / 64;
  p->dx = dx;
  return ISDQN_OK;
}

// V: the weight gradient's view of the layer (as for launch_conv_wgrad_tma); L: the layer itself (input gradient).
bool conv_bwd_pair_ok(const Layer& V, const Layer& L, int B) {
  if (!pair_launch_enabled() || B > 64 || !conv_wgrad_tma_ok(V) || !conv_dgrad_tma_ok(L)) return false;
  if (!(V.out_dim == 64 || V.out_dim == 128)) return false;
  const int bnd = pick_bn(L.Cin);
  if (!(bnd == 32 || bnd == 64 || bnd == 128)) return false;
  const int64_t dgrad_tiles = (int64_t)B * ceil_div(L.Cin, bnd) * L.stride * L.stride;
  return dgrad_tiles <= 2 * kNumSMs;
}

template <int BNW, int BND>
int launch_conv_bwd_pair_t(const Layer& V, const bf16* x, const bf16* dz, float* part, int B, int max_splits, int* real_splits,
                           float in_scale, int ksz_x, int sy, const Layer& L, const bf16* w, float* dx, cudaStream_t s,
                           const tc::LnBwdFuse* ln, int* ln_parts) {
  tc::ConvWgradTmaTC<BNW, true> pw;
  tc::ConvDgradTmaTC<BND, true> pd;
  const int st = L.stride;
  const int d_tiles = B * ceil_div(L.Cin, BND) * st * st;
  const int taps = ceil_div(L.ksz, st);
  const int wb = V.OW <= 16 ? 16 : V.OW <= 32 ? 32 : 64, cpi = ceil_div(V.OH, 64 / wb);
  const int tiles_k0 = ceil_div(V.in_dim, tc::kBM);
  // split the SMs in proportion to the chunk loads of the two problems (+ ~4 chunk times per tile for its epilogue)
  const double work_d = (double)d_tiles * (taps * taps * (L.out_dim / 64) + 4);
  const double work_w = (double)tiles_k0 * B * cpi + 4.0 * tiles_k0 * 8;
  int ctas_d = (int)(kNumSMs * work_d / (work_d + work_w) + 0.5);
  if (ctas_d < 1) ctas_d = 1;
  if (ctas_d > d_tiles) ctas_d = d_tiles;
  ctas_d = ceil_div(d_tiles, ceil_div(d_tiles, ctas_d));  // every CTA the same number of tiles
  // with the LayerNorm backward fused the input-gradient tiles are bound by their epilogue: one tile per CTA when the
  // weight gradient can still have a CTA per 128-row tile of K and 4 splits (ISDQN_PAIR_D1=0: the proportional split)
  static const bool d1 = [] {
    const char* e = getenv("ISDQN_PAIR_D1");
    return !(e && e[0] == '0');
  }();
  if (d1 && ln && d_tiles + 4 * tiles_k0 <= kNumSMs) ctas_d = d_tiles;
  if (ctas_d > kNumSMs - tiles_k0) ctas_d = kNumSMs - tiles_k0;
  int splits = (kNumSMs - ctas_d) / tiles_k0;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int tiles_k = 0;
  int rc = build_conv_wgrad_tma<BNW>(V, x, dz, part, B, splits, in_scale, ksz_x, sy, &pw, &tiles_k, real_splits);
  if (rc) return rc;
  rc = build_conv_dgrad_tma<BND>(L, dz, w, dx, B, &pd, ln);
  if (rc) return rc;
  if (ln_parts) *ln_parts = ctas_d;
  return launch_tc2(pw, tiles_k, 1, *real_splits, tiles_k * *real_splits, pd, B, ceil_div(L.Cin, BND), st * st, ctas_d, s,
                    "tc_conv_wgrad_dgrad");
}

int launch_conv_bwd_pair(const Layer& V, const bf16* x, const bf16* dz, float* part, int B, int max_splits, int* real_splits,
                         float in_scale, int ksz_x, int sy, const Layer& L, const bf16* w, float* dx, cudaStream_t s,
                         const tc::LnBwdFuse* ln = nullptr, int* ln_parts = nullptr) {
  const int bnd = pick_bn(L.Cin);
#define ISDQN_PAIR(BNW, BND) \
  return launch_conv_bwd_pair_t<BNW, BND>(V, x, dz, part, B, max_splits, real_splits, in_scale, ksz_x, sy, L, w, dx, s, ln, ln_parts);
  if (V.out_dim == 64) {
    if (bnd == 32) ISDQN_PAIR(64, 32)
    if (bnd == 64) ISDQN_PAIR(64, 64)
    if (bnd == 128) ISDQN_PAIR(64, 128)
  } else if (V.out_dim == 128) {
    if (bnd == 32) ISDQN_PAIR(128, 32)
    if (bnd == 64) ISDQN_PAIR(128, 64)
    if (bnd == 128) ISDQN_PAIR(128, 128)
  }
#undef ISDQN_PAIR
  return ISDQN_E_UNSUPPORTED;
}

int launch_conv_dgrad_tc(const Layer& L, const bf16* dz, const bf16* w, float* dx, int B, cudaStream_t s,
                         const tc::LnBwdFuse* ln = nullptr, int* ln_parts = nullptr) {
  if (conv_dgrad_tma_ok(L)) return launch_conv_dgrad_tma(L, dz, w, dx, B, s, ln, ln_parts);
  if (ln) return ISDQN_E_UNSUPPORTED;
  const int taps = ceil_div(L.ksz, L.stride);
  const int rows_max = B * ceil_div(L.H, L.stride) * ceil_div(L.W, L.stride);
  const int bn = pick_bn(L.Cin);
#define ISDQN_CONV_DGRAD_TC_W(BN, WIDE)                                                                \
  {                                                                                                    \
    tc::ConvDgradTC<BN, WIDE> p;                                                                       \
    p.H = L.H; p.W = L.W; p.Cin = L.Cin; p.OH = L.OH; p.OW = L.OW; p.Cout = L.out_dim;                \
    p.ksz = L.ksz; p.stride = L.stride; p.pad_y = L.pad_y; p.pad_x = L.pad_x; p.n_img = B;             \
    p.taps = taps; p.Kd = taps * taps * L.out_dim; p.dz = dz; p.w = w; p.dx = dx;                      \
    return launch_tc(p, ceil_div(rows_max, tc::kBM), ceil_div(L.Cin, BN), L.stride * L.stride, s, "tc_conv_dgrad"); \
  }
#define ISDQN_CONV_DGRAD_TC(BN)                                                                        \
  if (wide_launch((int64_t)ceil_div(rows_max, tc::kBM) * ceil_div(L.Cin, BN) * L.stride * L.stride))   \
    ISDQN_CONV_DGRAD_TC_W(BN, true)                                                                    \
  else ISDQN_CONV_DGRAD_TC_W(BN, false)
  switch (bn) {
    case 32: ISDQN_CONV_DGRAD_TC(32)
    case 64: ISDQN_CONV_DGRAD_TC(64)
    case 128: ISDQN_CONV_DGRAD_TC(128)
    default: ISDQN_CONV_DGRAD_TC(256)
  }
#undef ISDQN_CONV_DGRAD_TC_W
#undef ISDQN_CONV_DGRAD_TC
}

int launch_simt_gemm(const GemmArgs& g, cudaStream_t s, const char* tag) {
  dim3 grid(ceil_div(g.M, 64), ceil_div(g.N, 64), g.split_stride ? ceil_div(g.K, g.k_per_split) : 1);
  ISDQN_PROF(s, tag);
  const bool a_kfast = g.sak == 1, b_nfast = g.sbn == 1;
  if (a_kfast && b_nfast) ISDQN_CUDA_CHECK(launch_pdl((gemm_strided_kernel<true, true>), dim3(grid), dim3(kGemmThreads), 0, s, g));
  else if (a_kfast) ISDQN_CUDA_CHECK(launch_pdl((gemm_strided_kernel<true, false>), dim3(grid), dim3(kGemmThreads), 0, s, g));
  else if (b_nfast) ISDQN_CUDA_CHECK(launch_pdl((gemm_strided_kernel<false, true>), dim3(grid), dim3(kGemmThreads), 0, s, g));
  else ISDQN_CUDA_CHECK(launch_pdl((gemm_strided_kernel<false, false>), dim3(grid), dim3(kGemmThreads), 0, s, g));
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

int tc_train(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, bool backward, bool update, float* q_out,
             void* stream) {
  Plan p;
  if (!net) return ISDQN_E_INVALID;
  if (net->n_heads > kMaxHeads || net->n_actions > kMaxActions) return ISDQN_E_TOO_LARGE;
  int rc = build_plan(net, &p);
  if (rc) return rc;
  if (!tc_eligible(p, net)) return ISDQN_E_UNSUPPORTED;
  if (!tr || !b || !tr->d_params || !tr->d_losses || !tr->d_workspace || !tr->d_workspace_tc || !tr->d_params_bf16 || tr->batch < 1 ||
      tr->batch_global < tr->batch)
    return ISDQN_E_INVALID;
  if (!b->d_state || !b->d_next_state || !b->d_action || !b->d_reward || !b->d_terminal) return ISDQN_E_INVALID;
  if (backward && !tr->d_grads) return ISDQN_E_INVALID;
  if (update && (!tr->d_mu || !tr->d_nu || !tr->d_count)) return ISDQN_E_INVALID;
  const int B = tr->batch, rows = 2 * B;
  Workspace w;
  carve_workspace(p, rows, B, &w);
  TcWorkspace t;
  carve_tc(p, rows, B, &t);
  if (w.total * (int64_t)sizeof(float) > tr->workspace_bytes || t.total > tr->workspace_tc_bytes) return ISDQN_E_INVALID;
  cudaStream_t s = as_stream(stream);
  void* ws = tr->d_workspace;
  void* wt = tr->d_workspace_tc;
  const float* params = tr->d_params;
  float* grads = tr->d_grads;
  const int nl = p.n_layers;

  // bf16 shadow of the parameters (the fp32 master copy stays the truth; Adam updates it)
  bf16* shadow = reinterpret_cast<bf16*>(tr->d_params_bf16);
  if (tr->refresh_shadow) {
    const int64_t n4 = p.layout.total / 4;
    int64_t grid = ceil_div<int64_t>(n4, 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    ISDQN_PROF(s, "cast_params_bf16");
    ISDQN_CUDA_CHECK(launch_pdl(cast_f32_bf16_kernel, dim3((unsigned)grid), dim3(256), 0, s, params, shadow, n4));
    ISDQN_LAUNCH_CHECK();
  }
  // ------------------------------------------------------------------------------------------ forward
  const int rows_train = backward ? B : 0;
  bool mid_done = false;  // head_mid_kernel ran: TD loss + head backward are done, their cross-sample tail is pending
  for (int l = 0; l < nl; ++l) {
    const Layer& L = p.L[l];
    const float* ln_g = L.has_ln ? params + L.g_off : nullptr;
    const float* ln_b = L.has_ln ? params + L.beta_off : nullptr;
    float* xhat = rows_train > 0 ? wsp(ws, w.xhat[l]) : nullptr;
    float* rstd = rows_train > 0 ? wsp(ws, w.rstd[l]) : nullptr;
    if (L.type == 0) {
      if (l == 0 && t.w0p >= 0) {  // space-to-depth frames: conv0 is a TMA-fed 2x2 convolution over 64 channels
        const Layer S = s2d_layer(L);
        const int64_t threads = (int64_t)rows * S.H * S.W * 4;
        int64_t fb = ceil_div<int64_t>(threads, 256);
        if (fb > kNumSMs * 16) fb = kNumSMs * 16;
        ISDQN_PROF(s, "frames_to_bf16");
        ISDQN_CUDA_CHECK(launch_pdl(u8_frames_to_s2d_kernel, dim3((unsigned)fb + 8), dim3(256), 0, s,
                                    reinterpret_cast<const uint8_t*>(b->d_state), reinterpret_cast<const uint8_t*>(b->d_next_state),
                                    w16(wt, t.x16), B, L.H, L.W, (int)fb, shadow + L.w_off, w16(wt, t.w0p), w16(wt, t.w0t), L.out_dim));
        rc = launch_conv_fwd_tma(S, w16(wt, t.x16), rows, w16(wt, t.w0p), w16(wt, t.w0t), params, w16(wt, t.act16[l]), xhat, rstd,
                                 rows_train * L.pix, 1.0f / 255.0f, s, 0, 1, t.pair_view ? L.OW + 1 : 0);
      } else if (l == 0 && t.x16 >= 0) {
        const int64_t n16 = (int64_t)B * L.H * L.W * L.Cin / 16;
        int64_t grid = ceil_div<int64_t>(2 * n16, 256);
        if (grid > kNumSMs * 8) grid = kNumSMs * 8;
        ISDQN_PROF(s, "frames_to_bf16");
        ISDQN_CUDA_CHECK(launch_pdl(u8_frames_to_bf16_kernel, dim3((unsigned)grid), dim3(256), 0, s,
                                    reinterpret_cast<const uint8_t*>(b->d_state), reinterpret_cast<const uint8_t*>(b->d_next_state),
                                    w16(wt, t.x16), n16));
        if (L.ksz * L.Cin == 32)  // 64-byte image rows per window: the lane mapping that keeps a warp on few cache lines
          rc = launch_conv_fwd_tc<false, true>(L, w16(wt, t.x16), nullptr, rows, rows, shadow + L.w_off, params,
                                               w16(wt, t.act16[l]), xhat, rstd, rows_train * L.pix, s, 1.0f / 255.0f);
        else
          rc = launch_conv_fwd_tc<false>(L, w16(wt, t.x16), nullptr, rows, rows, shadow + L.w_off, params, w16(wt, t.act16[l]),
                                         xhat, rstd, rows_train * L.pix, s, 1.0f / 255.0f);
      } else if (l == 0)
        rc = launch_conv_fwd_tc<true>(L, b->d_state, b->d_next_state, B, rows, shadow + L.w_off, params, w16(wt, t.act16[l]),
                                      xhat, rstd, rows_train * L.pix, s);
      else if (l == 1 && t.pair_view) {
        Layer V = L;  // the view: half the width, twice the channels, 4 x 2 taps, row stride 2, no left padding
        V.W = (L.W + 1) / 2;
        V.Cin = 2 * L.Cin;
        V.pad_x = 0;
        rc = launch_conv_fwd_tma(V, w16(wt, t.act16[l - 1]), rows, shadow + L.w_off, nullptr, params, w16(wt, t.act16[l]), xhat,
                                 rstd, rows_train * L.pix, 1.0f, s, /*ksz_x=*/2, /*sy=*/2);
      } else if (conv_fwd_tma_ok(L))
        rc = launch_conv_fwd_tma(L, w16(wt, t.act16[l - 1]), rows, shadow + L.w_off, nullptr, params, w16(wt, t.act16[l]), xhat,
                                 rstd, rows_train * L.pix, 1.0f, s);
      else
        rc = launch_conv_fwd_tc<false>(L, w16(wt, t.act16[l - 1]), nullptr, rows, rows, shadow + L.w_off, params,
                                       w16(wt, t.act16[l]), xhat, rstd, rows_train * L.pix, s);
      if (rc) return rc;
    } else if (l + 1 < nl) {  // hidden Dense: split-K tensor-core GEMM -> partials -> bias + LN + ReLU
      const int splits = dense_fwd_splits(rows, L.out_dim, L.in_dim);
      const int64_t split_stride = (int64_t)rows * L.out_dim;
      const int total_chunks = ceil_div(L.in_dim, tc::kBK);
      const int cps = ceil_div(total_chunks, splits);
      const int real_splits = ceil_div(total_chunks, cps);
      rc = launch_gemm_tc<false, true>(w16(wt, t.act16[l - 1]), L.in_dim, shadow + L.w_off, L.out_dim, wsp(ws, w.fwd_part),
                                       L.out_dim, split_stride, rows, L.out_dim, L.in_dim, splits, s, "tc_dense_fwd");
      if (rc) return rc;
      const Layer& Hd = p.L[l + 1];
      static const bool mid_on = [] {
        // opt-in: measured 124.7 us of in-graph intervals against 119.2 us for the three kernels it replaces — its CTAs are
        // bound by the same dependent L2 round trips, and the cross-sample tail it leaves to the partial reduction delays
        // the optimiser (profiles/r02_summary.md)
        const char* e = getenv("ISDQN_MID");
        return e && e[0] == '1';
      }();
      if (mid_on && backward && l + 2 == nl && B <= 256 && Hd.out_dim <= 128 && L.out_dim <= kRowThreads * kRowMaxPerThread &&
          w.wsplits[nl - 1] == 1 && w.col_ctas[l] == B && p.n_out >= 2 * net->n_heads && L.relu) {
        // small batch: finish the hidden layer, head layer, TD loss and the head / hidden-layer backward in ONE launch
        static PerDeviceOnce attr_once;
        if (attr_once.first()) {
          ISDQN_CUDA_CHECK(cudaFuncSetAttribute(head_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                2 * kRowThreads * kRowMaxPerThread * (int)sizeof(float)));
        }
        ISDQN_PROF(s, "head_mid");
        ISDQN_CUDA_CHECK(launch_pdl(head_mid_kernel, dim3(B), dim3(512), 2 * L.out_dim * sizeof(float), s, wsp(ws, w.fwd_part),
                                    real_splits, split_stride, L.out_dim, params + L.b_off, ln_g, ln_b, L.relu, wsp(ws, w.act[l]),
                                    params + Hd.w_off, params + Hd.b_off, Hd.out_dim, wsp(ws, w.act[l + 1]), b->d_action,
                                    b->d_reward, b->d_terminal, tr->gamma_n, B, tr->batch_global, net->n_heads, net->n_actions,
                                    tr->d_is_weights, tr->d_td_abs, wsp(ws, w.dq), wsp(ws, w.colpart[l]), w16(wt, t.dz16[l]),
                                    update ? tr->d_count : nullptr));
        ISDQN_LAUNCH_CHECK();
        mid_done = true;
        ++l;  // the head layer is done
      } else if (l + 2 == nl && rows <= 1024 && Hd.out_dim <= 128 && L.out_dim <= kRowThreads * kRowMaxPerThread) {
        // last hidden layer: finish it and apply the (fp32) head layer in the same launch
        ISDQN_PROF(s, "dense_finalize_head");
        ISDQN_CUDA_CHECK(launch_pdl(dense_finalize_head_kernel, dim3(rows), dim3(512), L.out_dim * sizeof(float), s,
                                    wsp(ws, w.fwd_part), real_splits, split_stride, L.out_dim, params + L.b_off, ln_g, ln_b, L.relu,
                                    wsp(ws, w.act[l]), xhat, rstd, rows_train, params + Hd.w_off, params + Hd.b_off, Hd.out_dim,
                                    wsp(ws, w.act[l + 1])));
        ++l;  // the head layer is done
      } else {  // (large batches: two launches, each with all its threads busy)
        ISDQN_PROF(s, "dense_finalize");
        ISDQN_CUDA_CHECK(launch_pdl(dense_finalize_kernel, dim3(rows), dim3(kRowThreads), 0, s, wsp(ws, w.fwd_part), real_splits,
                                    split_stride, rows, L.out_dim, params + L.b_off, ln_g, ln_b, L.relu, wsp(ws, w.act[l]), xhat,
                                    rstd, rows_train, w16(wt, t.act16[l])));
      }
    } else if (L.out_dim <= 128 && L.in_dim <= kHeadMaxK) {  // head layer: fp32 (N = (1+K)A is tiny, not 16-byte aligned)
      ISDQN_PROF(s, "head_fwd");
      ISDQN_CUDA_CHECK(launch_head_fwd(s, wsp(ws, w.act[l - 1]), params + L.w_off, params + L.b_off, L.in_dim, L.out_dim,
                                       wsp(ws, w.act[l]), rows));
    } else {
      GemmArgs g;
      g.A = wsp(ws, w.act[l - 1]); g.sam = L.in_dim; g.sak = 1;
      g.B = params + L.w_off; g.sbk = L.out_dim; g.sbn = 1;
      g.C = wsp(ws, w.act[l]); g.ldc = L.out_dim; g.split_stride = 0;
      g.M = rows; g.N = L.out_dim; g.K = L.in_dim; g.k_per_split = ceil_div(L.in_dim, kBK) * kBK;
      g.bias = params + L.b_off;
      rc = launch_simt_gemm(g, s, "head_fwd_gemm");
      if (rc) return rc;
    }
  }
  const float* q_all = wsp(ws, w.act[nl - 1]);
  const Layer& last = p.L[nl - 1];
  // small batch: the TD loss runs inside the head-backward launch (head_bwd_td_kernel) instead of as a hop of its own
  // (opt-in: measured 131.3 us per update against 126.7 us with the separate loss kernel — the folded kernel's row CTAs
  // gain two round trips on the critical path and every weight-gradient CTA recomputes the TD matrix; profiles/r02_summary.md)
  static const bool td_fold_on = [] {
    const char* e = getenv("ISDQN_TD_FOLD");
    return e && e[0] == '1';
  }();
  const bool td_fold = td_fold_on && backward && !mid_done && !q_out && nl >= 2 && p.L[nl - 2].type == 1 && B <= 256 &&
                       B <= kTailMaxB && B * net->n_heads <= kHbTdMax && w.wsplits[nl - 1] == 1 && last.out_dim <= 128 &&
                       p.L[nl - 2].out_dim <= kRowThreads * kRowMaxPerThread;
  if (!mid_done && !td_fold) {
  ISDQN_PROF(s, "heads_td_loss");
  ISDQN_CUDA_CHECK(launch_pdl(heads_td_loss_kernel, dim3(net->n_heads), dim3(kLossThreads), 0, s, q_all, b->d_action, b->d_reward, b->d_terminal, tr->gamma_n, B,
                                                  tr->batch_global, net->n_heads, net->n_actions, tr->d_losses,
                                                  backward ? wsp(ws, w.dq) : nullptr, backward ? grads + last.b_off : nullptr,
                                                  update ? tr->d_count : nullptr, update ? tr->d_cumulated : nullptr,
                                                  tr->d_is_weights, tr->d_td_abs));
  ISDQN_LAUNCH_CHECK();
  }
  if (q_out)
    ISDQN_CUDA_CHECK(cudaMemcpyAsync(q_out, q_all, sizeof(float) * (size_t)rows * p.n_out, cudaMemcpyDeviceToDevice, s));
  if (!backward) return ISDQN_OK;

  // ----------------------------------------------------------------------------------------- backward
  SegmentList segs;
  segs.count = 0;
  auto add_seg = [&](const float* src, float* dst, int64_t stride, int n, int parts) {
    Segment& sg = segs.s[segs.count++];
    sg.src = src; sg.dst = dst; sg.stride = stride; sg.n = n; sg.parts = parts; sg.s2d_cout = 0;
  };
  float* dz32 = wsp(ws, w.dq);
  // Two streams: the chain  input gradient -> LayerNorm/ReLU backward -> input gradient ...  is the critical path; the
  // weight gradient of every layer but the first only feeds the optimiser, so it runs on a side stream (forked off an
  // event, joined before the partial reduction).  The Adam update of a hidden Dense kernel (97 % of the parameters of
  // the Atari network) follows its weight gradient on the side stream as soon as the input-gradient GEMM that reads
  // the same weights has finished, and the final Adam launch passes over that range.
  cudaStream_t s2 = s, s3 = s;
  // fork_mode(): 0 = one stream; 1 = weight gradients + early Adam on side streams; 2 = only the early Adam
  const int fmode = (!g_profile_on && !tr->nccl_comm && side_stream(0) != nullptr && side_stream(1) != nullptr) ? fork_mode() : 0;
  const bool fork = fmode == 1;
  if (fmode) {
    if (fork) s2 = side_stream(0);  // weight gradients
    s3 = side_stream(1);            // early Adam
  }
  // small batches are latency bound (every kernel is a fraction of a wave): keep the side work narrow so that it shares
  // the SMs with the critical path instead of queueing in front of it; large batches are throughput bound: no cap
  const bool narrow_side = fmode && B <= 256;
  const int side_cap = narrow_side ? kSideCtas : 0;
  int n_ev = 0;
  auto fork_to_side = [&]() -> int {
    cudaEvent_t e = side_event(n_ev++);
    if (!e) return ISDQN_E_CUDA;
    ISDQN_CUDA_CHECK(cudaEventRecord(e, s));
    ISDQN_CUDA_CHECK(cudaStreamWaitEvent(s2, e, 0));
    return ISDQN_OK;
  };
  int64_t pend_off = -1, pend_n = 0;    // Dense kernel whose early Adam waits for the next fork point
  int64_t early_off = 0, early_n = 0;   // range already updated on the side stream
  int fused_l = -1;                     // small batch: Dense layer whose weight gradient is recomputed inside its Adam update
  int64_t dp_early_off = 0, dp_early_n = 0;  // data parallel: gradient range already being all-reduced on the communication stream
  int col_parts[ISDQN_MAX_FEATURES + 1];  // column partials every layer's LayerNorm / ReLU backward produced
  for (int l = 0; l < nl; ++l) col_parts[l] = w.col_ctas[l];
  static const bool ln_fuse_on = [] {
    const char* e = getenv("ISDQN_LNFUSE");
    return !(e && e[0] == '0');
  }();
  // The LayerNorm / ReLU backward of conv layer l - 1 (<= 64 channels) runs in the epilogue of layer l's input gradient:
  // the accumulator row of an epilogue thread IS one pixel of layer l - 1 with all its channels (tc_problems.cuh).
  auto ln_fuse_for = [&](int l, tc::LnBwdFuse* f) -> bool {
    if (!ln_fuse_on || l < 1 || fmode) return false;
    const Layer& P = p.L[l - 1];
    if (P.type != 0 || !P.has_ln || !P.relu || !(P.out_dim == 32 || P.out_dim == 64)) return false;
    f->xhat = wsp(ws, w.xhat[l - 1]);
    f->rstd = wsp(ws, w.rstd[l - 1]);
    f->ln_g = params + P.g_off;
    f->ln_b = params + P.beta_off;
    f->dz16 = w16(wt, t.dz16[l - 1]);
    f->colpart = wsp(ws, w.colpart[l - 1]);
    return true;
  };
  for (int l = nl - 1; l >= 0; --l) {
    const Layer& L = p.L[l];
    const bf16* dz16 = w16(wt, t.dz16[l]);
    const int rows_l = B * L.pix;
    bool ln_fused = false;  // layer l - 1's LayerNorm / ReLU backward was done by this layer's input-gradient launch
    if (mid_done && l == nl - 1) continue;  // head backward + the hidden layer's LayerNorm/ReLU backward: head_mid_kernel
    if (l == nl - 1 && l >= 1 && p.L[l - 1].type == 1 && B <= 256 && w.wsplits[l] == 1 && L.out_dim <= 128 &&
        p.L[l - 1].out_dim <= kRowThreads * kRowMaxPerThread) {
      // small batch: head weight gradient + head input gradient + LayerNorm/ReLU backward of the hidden layer, one launch
      const Layer& P = p.L[l - 1];
      float* dprev = wsp(ws, w.dbuf[l & 1]);
      const int row_ctas = w.col_ctas[l - 1];
      const int wg_ctas = ceil_div(P.out_dim * L.out_dim, kRowThreads);
      if (td_fold) {
        ISDQN_PROF(s, "head_bwd_td");
        ISDQN_CUDA_CHECK(launch_pdl(head_bwd_td_kernel, dim3(row_ctas + wg_ctas + 1), dim3(kRowThreads), 0, s, q_all, b->d_action,
                                    b->d_reward, b->d_terminal, tr->gamma_n, B, tr->batch_global, net->n_heads, net->n_actions,
                                    tr->d_is_weights, tr->d_td_abs, tr->d_losses, update ? tr->d_cumulated : nullptr,
                                    grads + L.b_off, update ? tr->d_count : nullptr, params + L.w_off, wsp(ws, w.act[l - 1]),
                                    P.out_dim, L.out_dim, row_ctas, wg_ctas, wsp(ws, w.xhat[l - 1]), wsp(ws, w.rstd[l - 1]),
                                    P.has_ln ? params + P.g_off : nullptr, P.has_ln ? params + P.beta_off : nullptr,
                                    wsp(ws, w.colpart[l - 1]), w16(wt, t.dz16[l - 1]), grads + L.w_off));
        ISDQN_LAUNCH_CHECK();
        dz32 = dprev;
        continue;
      }
      ISDQN_PROF(s, "head_bwd");
      ISDQN_CUDA_CHECK(launch_pdl(head_bwd_kernel, dim3(row_ctas + wg_ctas), dim3(kRowThreads), 0, s, dz32, params + L.w_off,
                                  wsp(ws, w.act[l - 1]), B, P.out_dim, L.out_dim, row_ctas, dprev, wsp(ws, w.xhat[l - 1]),
                                  wsp(ws, w.rstd[l - 1]), P.has_ln ? params + P.g_off : nullptr,
                                  P.has_ln ? params + P.beta_off : nullptr, wsp(ws, w.colpart[l - 1]), w16(wt, t.dz16[l - 1]),
                                  grads + L.w_off));
      dz32 = dprev;
      continue;
    }
    const bool side = fork && l > 0;
    cudaStream_t sw = side ? s2 : s;
    bool paired_dgrad = false;  // the input gradient of this layer was computed by the weight-gradient launch
    if (side || (fmode == 2 && pend_n > 0)) {
      rc = fork_to_side();  // everything the main stream has produced so far (dz of this layer, the input gradient above)
      if (rc) return rc;
      if (pend_n > 0) {
        // after the weight gradient it consumes (side stream) and the input-gradient GEMM reading the same weights (main)
        cudaEvent_t e = side_event(n_ev++);
        if (!e) return ISDQN_E_CUDA;
        ISDQN_CUDA_CHECK(cudaEventRecord(e, s2));
        if (s2 != s) ISDQN_CUDA_CHECK(cudaStreamWaitEvent(s3, e, 0));
        ISDQN_CUDA_CHECK(cudaStreamWaitEvent(s3, side_event(n_ev - 2), 0));
        rc = isdqn_adam_launch(tr->d_params + pend_off, tr->d_grads + pend_off, tr->d_mu + pend_off, tr->d_nu + pend_off,
                               tr->d_count, tr->lr, tr->b1, tr->b2, tr->eps, pend_n, shadow + pend_off, s3, 0, 0, narrow_side ? 2 * kNumSMs : 0);
        if (rc) return rc;
        early_off = pend_off;
        early_n = pend_n;
        pend_n = 0;
      }
    }
    // ---- weight gradient
    if (l == nl - 1) {
      GemmArgs g;
      g.A = wsp(ws, w.act[l - 1]); g.sam = 1; g.sak = L.in_dim;
      g.B = dz32; g.sbk = L.out_dim; g.sbn = 1;
      g.C = grads + L.w_off; g.ldc = L.out_dim; g.split_stride = 0;
      g.M = L.in_dim; g.N = L.out_dim; g.K = B; g.k_per_split = ceil_div(B, kBK) * kBK; g.bias = nullptr;
      if (w.wsplits[l] > 1) {  // long batch axis: split it over CTAs, partial sums folded by reduce_segments
        g.k_per_split = ceil_div(ceil_div(B, w.wsplits[l]), kBK) * kBK;
        const int real_splits = ceil_div(B, (int)g.k_per_split);
        g.C = wsp(ws, w.wpart[l]);
        g.split_stride = (int64_t)L.in_dim * L.out_dim;
        add_seg(g.C, grads + L.w_off, g.split_stride, L.in_dim * L.out_dim, real_splits);
      }
      rc = launch_simt_gemm(g, sw, "head_wgrad_gemm");
    } else if (L.type == 1 && update && fmode == 0 && !tr->nccl_comm && fused_l < 0 && l > 0 &&
               isdqn_dense_wgrad_adam_ok(B, L.in_dim, L.out_dim, L.w_off)) {
      fused_l = l;  // no gradient launch: dense_wgrad_adam_kernel (below, after the last reader of this kernel's shadow)
    } else if (L.type == 1) {
      rc = launch_gemm_tc<true, true>(w16(wt, t.act16[l - 1]), L.in_dim, dz16, L.out_dim, grads + L.w_off, L.out_dim, 0,
                                      L.in_dim, L.out_dim, B, 1, sw, "tc_dense_wgrad", side ? side_cap : 0);
      if (!rc && tr->nccl_comm && update && dp_overlap_on() && dp_early_n == 0 && l > 0 && side_stream(0) && !g_profile_on) {
        // data parallel: this kernel's gradient is 97 % of the bytes and is complete NOW, before the convolution backward
        // has even started — all-reduce it on the communication stream under the rest of the backward pass
        cudaStream_t sc = side_stream(0);
        cudaEvent_t e0 = side_event(kSideEvents - 1), e1 = side_event(kSideEvents - 2);
        if (!e0 || !e1) return ISDQN_E_CUDA;
        ISDQN_CUDA_CHECK(cudaEventRecord(e0, s));
        ISDQN_CUDA_CHECK(cudaStreamWaitEvent(sc, e0, 0));
        rc = isdqn_dp_allreduce_f32(tr->nccl_comm, tr->d_grads + L.w_off, (int64_t)L.in_dim * L.out_dim, sc);
        if (rc) return rc;
        ISDQN_CUDA_CHECK(cudaEventRecord(e1, sc));
        dp_early_off = L.w_off;
        dp_early_n = (int64_t)L.in_dim * L.out_dim;
      }
      // (the early range must be the only one: the final Adam launch passes over a single range)
      if ((side || fmode == 2) && l > 0 && update && early_n == 0 && pend_n == 0 && ((int64_t)L.in_dim * L.out_dim) % 4 == 0 &&
          L.w_off % 4 == 0) {
        pend_off = L.w_off;
        pend_n = (int64_t)L.in_dim * L.out_dim;
      }
    } else {
      int real_splits = 1;
      float* part = wsp(ws, w.wpart[l]);
      static const bool wgrad_tma = [] {
        const char* e = getenv("ISDQN_TMA_WGRAD");
        return !(e && e[0] == '0');
      }();
      if (l > 0 && !side && wgrad_tma) {  // small batch: this layer's weight gradient and input gradient in one launch
        Layer V = L;
        int kx = 0, sy = 1;
        bool view_ok = L.stride == 1;
        if (l == 1 && t.pair_view) {  // the pair-of-pixels view of the padded activation (as in the forward pass)
          V.W = (L.W + 1) / 2;
          V.Cin = 2 * L.Cin;
          V.pad_x = 0;
          kx = 2;
          sy = 2;
          view_ok = true;
        }
        if (view_ok && conv_bwd_pair_ok(V, L, B)) {
          tc::LnBwdFuse lf;
          int parts = 0;
          const bool fuse = ln_fuse_for(l, &lf) && pick_bn(L.Cin) == p.L[l - 1].out_dim;
          rc = launch_conv_bwd_pair(V, w16(wt, t.act16[l - 1]), dz16, part, B, w.wsplits_tc[l], &real_splits, 1.0f, kx, sy, L,
                                    shadow + L.w_off, wsp(ws, w.dbuf[l & 1]), s, fuse ? &lf : nullptr, &parts);
          paired_dgrad = true;
          if (fuse) {
            ln_fused = true;
            col_parts[l - 1] = parts;
          }
        }
      }
      if (paired_dgrad) {
      } else if (l == 0 && t.w0p >= 0) {  // space-to-depth frames (128-byte taps); partials come out in s2d K order
        const Layer S = s2d_layer(L);
        if (wgrad_tma && conv_wgrad_tma_ok(S))
          rc = launch_conv_wgrad_tma(S, w16(wt, t.x16), dz16, part, B, w.wsplits_tc[l], &real_splits, sw, 1.0f / 255.0f, 0, 1);
        else
          rc = launch_conv_wgrad_tc<false>(S, w16(wt, t.x16), dz16, part, rows_l, w.wsplits_tc[l], &real_splits, sw,
                                           1.0f / 255.0f, side ? side_cap : 0);
      } else if (l == 1 && t.pair_view && wgrad_tma) {
        Layer V = L;  // the pair-of-pixels view of the padded activation (as in the forward pass)
        V.W = (L.W + 1) / 2;
        V.Cin = 2 * L.Cin;
        V.pad_x = 0;
        rc = conv_wgrad_tma_ok(V)
                 ? launch_conv_wgrad_tma(V, w16(wt, t.act16[l - 1]), dz16, part, B, w.wsplits_tc[l], &real_splits, sw, 1.0f, 2, 2)
                 : ISDQN_E_UNSUPPORTED;
      } else if (l > 0 && wgrad_tma && L.stride == 1 && conv_wgrad_tma_ok(L)) {
        rc = launch_conv_wgrad_tma(L, w16(wt, t.act16[l - 1]), dz16, part, B, w.wsplits_tc[l], &real_splits, sw, 1.0f, 0, 1);
      } else if (l == 0 && t.x16 >= 0 && L.ksz * L.Cin == 32)
        rc = launch_conv_wgrad_tc<false, true>(L, w16(wt, t.x16), dz16, part, rows_l, w.wsplits_tc[l], &real_splits, sw,
                                               1.0f / 255.0f, side ? side_cap : 0);
      else if (l == 0 && t.x16 >= 0)
        rc = launch_conv_wgrad_tc<false>(L, w16(wt, t.x16), dz16, part, rows_l, w.wsplits_tc[l], &real_splits, sw, 1.0f / 255.0f,
                                         side ? side_cap : 0);
      else if (l == 0) rc = launch_conv_wgrad_tc<true>(L, b->d_state, dz16, part, rows_l, w.wsplits_tc[l], &real_splits, sw);
      else if (l == 1 && t.pair_view) {
        Layer G = L;  // the same pixels at stored column x + 1: one more column, no left padding
        G.W = L.W + 1;
        G.pad_x = L.pad_x - 1;
        rc = launch_conv_wgrad_tc<false>(G, w16(wt, t.act16[l - 1]), dz16, part, rows_l, w.wsplits_tc[l], &real_splits, sw, 1.0f,
                                         side ? side_cap : 0);
      } else
        rc = launch_conv_wgrad_tc<false>(L, w16(wt, t.act16[l - 1]), dz16, part, rows_l, w.wsplits_tc[l], &real_splits, sw, 1.0f,
                                         side ? side_cap : 0);
      add_seg(part, grads + L.w_off, (int64_t)L.in_dim * L.out_dim, L.in_dim * L.out_dim, real_splits);
      if (l == 0 && t.w0p >= 0) segs.s[segs.count - 1].s2d_cout = L.out_dim;
    }
    if (rc) return rc;
    if (L.relu) {
      const float* cp = wsp(ws, w.colpart[l]);
      const int64_t st = 3 * (int64_t)L.out_dim;
      add_seg(cp, grads + L.b_off, st, L.out_dim, col_parts[l]);
      if (L.has_ln) {
        add_seg(cp + L.out_dim, grads + L.g_off, st, L.out_dim, col_parts[l]);
        add_seg(cp + 2 * L.out_dim, grads + L.beta_off, st, L.out_dim, col_parts[l]);
      }
    }
    if (l == 0) break;
    // ---- input gradient, then the previous layer's ReLU + LayerNorm backward (fp32), which also emits bf16 dz
    const Layer& P = p.L[l - 1];
    float* dprev = wsp(ws, w.dbuf[l & 1]);
    if (l == nl - 1) {
      GemmArgs g;
      g.A = dz32; g.sam = L.out_dim; g.sak = 1;
      g.B = params + L.w_off; g.sbk = 1; g.sbn = L.out_dim;
      g.C = dprev; g.ldc = L.in_dim; g.split_stride = 0;
      g.M = B; g.N = L.in_dim; g.K = L.out_dim; g.k_per_split = ceil_div(L.out_dim, kBK) * kBK; g.bias = nullptr;
      rc = launch_simt_gemm(g, s, "head_dgrad_gemm");
    } else if (L.type == 1) {
      tc::LnBwdFuse lf;
      int parts = 0;
      if (ln_fuse_for(l, &lf) && p.L[l - 1].out_dim == 64 && L.in_dim % 64 == 0 &&
          gemm_tma_operands_ok(dz16, L.out_dim, shadow + L.w_off, L.out_dim)) {
        rc = launch_gemm_tc<false, false>(dz16, L.out_dim, shadow + L.w_off, L.out_dim, dprev, L.in_dim, 0, B, L.in_dim,
                                          L.out_dim, 1, s, "tc_dense_dgrad_ln", 0, &lf, &parts);
        ln_fused = true;
        col_parts[l - 1] = parts;
      } else {
        rc = launch_gemm_tc<false, false>(dz16, L.out_dim, shadow + L.w_off, L.out_dim, dprev, L.in_dim, 0, B, L.in_dim,
                                          L.out_dim, 1, s, "tc_dense_dgrad");
      }
    } else if (!paired_dgrad) {
      tc::LnBwdFuse lf;
      int parts = 0;
      if (ln_fuse_for(l, &lf) && conv_dgrad_tma_ok(L) && pick_bn(L.Cin) == p.L[l - 1].out_dim) {
        rc = launch_conv_dgrad_tc(L, dz16, shadow + L.w_off, dprev, B, s, &lf, &parts);
        ln_fused = true;
        col_parts[l - 1] = parts;
      } else {
        rc = launch_conv_dgrad_tc(L, dz16, shadow + L.w_off, dprev, B, s);
      }
    }
    if (rc) return rc;
    if (ln_fused && col_parts[l - 1] > kColpartFusedCap) return ISDQN_E_INVALID;  // (grids are <= 4 CTAs per SM)
    if (!ln_fused) {
      const int rows_p = B * P.pix;
      const float* g_ = P.has_ln ? params + P.g_off : nullptr;
      const float* b_ = P.has_ln ? params + P.beta_off : nullptr;
      bf16* dz16_prev = w16(wt, t.dz16[l - 1]);
      // the ReLU mask of a layer without LayerNorm needs its post-activation output: fp32 for Dense, bf16 for conv
      ISDQN_PROF(s, "ln_relu_bwd");
      if (ln_bwd_use_warp(P.out_dim)) {
        ISDQN_CUDA_CHECK(launch_ln_relu_bwd_warp(w.col_ctas[l - 1], s, dprev, wsp(ws, w.xhat[l - 1]), wsp(ws, w.rstd[l - 1]), g_, b_,
                                P.type == 1 ? wsp(ws, w.act[l - 1]) : nullptr, rows_p, P.out_dim, wsp(ws, w.colpart[l - 1]),
                                dz16_prev, P.type == 0 ? w16(wt, t.act16[l - 1]) : nullptr, /*store_d=*/false));
      } else {
        ISDQN_CUDA_CHECK(launch_pdl(ln_relu_bwd_block_kernel, dim3(w.col_ctas[l - 1]), dim3(kRowThreads), 0, s, 
            dprev, wsp(ws, w.xhat[l - 1]), wsp(ws, w.rstd[l - 1]), g_, b_, P.type == 1 ? wsp(ws, w.act[l - 1]) : nullptr, rows_p,
            P.out_dim, wsp(ws, w.colpart[l - 1]), dz16_prev, P.type == 0 ? w16(wt, t.act16[l - 1]) : nullptr));
      }
      ISDQN_LAUNCH_CHECK();
    }
    dz32 = dprev;
  }
  if (fmode) {  // join
    for (cudaStream_t x : {s2, s3}) {
      if (x == s) continue;
      if (x == s3 && early_n == 0) continue;  // (never forked: not part of a capture)
      cudaEvent_t e = side_event(n_ev++);
      if (!e) return ISDQN_E_CUDA;
      ISDQN_CUDA_CHECK(cudaEventRecord(e, x));
      ISDQN_CUDA_CHECK(cudaStreamWaitEvent(s, e, 0));
    }
  }
  if (segs.count > 0) {
    HeadTail tail = {};
    if (mid_done) {  // loss means, head bias / kernel gradients: the part of the head step that needs every sample
      const Layer& Hp = p.L[nl - 2];
      tail.ctas = head_tail_ctas(Hp.out_dim, last.out_dim);
      tail.tdq = wsp(ws, w.dq);
      tail.action = b->d_action;
      tail.act = wsp(ws, w.act[nl - 2]);
      tail.B = B; tail.K = net->n_heads; tail.A = net->n_actions; tail.C = Hp.out_dim; tail.NH = last.out_dim;
      tail.inv_b = 1.0f / (float)tr->batch_global;
      tail.losses = tr->d_losses;
      tail.cumulated = update ? tr->d_cumulated : nullptr;
      tail.dbias = grads + last.b_off;
      tail.dwh = grads + last.w_off;
    }
    ISDQN_PROF(s, "reduce_segments");
    ISDQN_CUDA_CHECK(launch_reduce_segments(segs, s, mid_done ? &tail : nullptr));
  }
  if (!update) return ISDQN_OK;
  if (tr->nccl_comm) {
    ISDQN_PROF(s, "nccl_allreduce");
    if (dp_early_n > 0) {  // the big range is in flight (or done) on the communication stream: the small rest, then join
      rc = isdqn_dp_allreduce_rest(tr->nccl_comm, tr->d_grads, p.layout.total, dp_early_off, dp_early_n, stream);
      if (rc) return rc;
      ISDQN_CUDA_CHECK(cudaStreamWaitEvent(s, side_event(kSideEvents - 2), 0));
    } else {
      rc = isdqn_dp_allreduce_f32(tr->nccl_comm, tr->d_grads, p.layout.total, stream);
      if (rc) return rc;
    }
  }
  if (fused_l >= 0) {  // the Dense kernel from its recomputed gradient + every other leaf from `grads`: one launch
    const Layer& L = p.L[fused_l];
    // the bulk-copy pipelined version first (adam_stream.cu); shapes it does not take go to the per-thread-load kernel
    rc = isdqn_dense_wgrad_adam_stream_launch(tr->d_params, tr->d_grads, tr->d_mu, tr->d_nu, tr->d_count, tr->lr, tr->b1, tr->b2,
                                              tr->eps, shadow, p.layout.total, L.w_off, w16(wt, t.act16[fused_l - 1]), L.in_dim,
                                              w16(wt, t.dz16[fused_l]), B, L.in_dim, L.out_dim, stream);
    if (rc != ISDQN_E_UNSUPPORTED) return rc;
    return isdqn_dense_wgrad_adam_launch(tr->d_params, tr->d_grads, tr->d_mu, tr->d_nu, tr->d_count, tr->lr, tr->b1, tr->b2,
                                         tr->eps, shadow, p.layout.total, L.w_off, w16(wt, t.act16[fused_l - 1]), L.in_dim,
                                         w16(wt, t.dz16[fused_l]), B, L.in_dim, L.out_dim, stream);
  }
  return isdqn_adam_launch(tr->d_params, tr->d_grads, tr->d_mu, tr->d_nu, tr->d_count, tr->lr, tr->b1, tr->b2, tr->eps,
                           p.layout.total, shadow, stream, early_off, early_n);
}

}  // namespace

// Entry used by the learner dispatch in learner.cu
// ---- building blocks for callers outside this file (the impala torso in learner.cu): one convolution forward / weight
// gradient / input gradient on the tile engine, bf16 NHWC activations, bf16 HWIO kernels out of the parameter shadow
bool isdqn_tc_conv_ok(const Layer& L) {
  if (L.type != 0 || L.Cin % 8 != 0) return false;
  if (!(L.out_dim == 32 || L.out_dim == 64 || L.out_dim == 128 || L.out_dim == 256)) return false;
  const int taps = ceil_div(L.ksz, L.stride);
  if (ceil_div(L.in_dim, tc::kBM) * 16 > tc::kMaxChunks) return false;
  if (L.stride * L.stride * ceil_div(taps * taps * L.out_dim, tc::kBK) * 8 > tc::kMaxChunks) return false;
  return true;
}
// out16 = act(conv(x16) + bias) with L.relu / L.b_off (no LayerNorm in the epilogue: L.has_ln must be 0)
static bool blocks_tma_on() {  // TMA-fed problems where the shape allows (A/B: ISDQN_IMPALA_TMA=0 keeps the gathers)
  static const bool on = [] {
    const char* e = getenv("ISDQN_IMPALA_TMA");
    return !(e && e[0] == '0');
  }();
  return on;
}
int isdqn_tc_conv_fwd(const Layer& L, const void* x16, int rows, const void* w16_, const float* params, void* out16, cudaStream_t s,
                      float in_scale) {
  if (blocks_tma_on() && conv_fwd_tma_ok(L))
    return launch_conv_fwd_tma(L, reinterpret_cast<const bf16*>(x16), rows, reinterpret_cast<const bf16*>(w16_), nullptr, params,
                               reinterpret_cast<bf16*>(out16), nullptr, nullptr, 0, in_scale, s);
  return launch_conv_fwd_tc<false>(L, x16, nullptr, rows, rows, reinterpret_cast<const bf16*>(w16_), params,
                                   reinterpret_cast<bf16*>(out16), nullptr, nullptr, 0, s, in_scale);
}
int isdqn_tc_conv_wgrad(const Layer& L, const void* x16, const void* dz16, float* part, int rows_l, int splits, int* real_splits,
                        cudaStream_t s, float in_scale) {
  if (blocks_tma_on() && conv_wgrad_tma_ok(L))
    return launch_conv_wgrad_tma(L, reinterpret_cast<const bf16*>(x16), reinterpret_cast<const bf16*>(dz16), part, rows_l / L.pix,
                                 splits, real_splits, s, in_scale, L.ksz, 1);
  return launch_conv_wgrad_tc<false>(L, x16, reinterpret_cast<const bf16*>(dz16), part, rows_l, splits, real_splits, s, in_scale);
}
int isdqn_tc_conv_dgrad(const Layer& L, const void* dz16, const void* w16_, float* dx, int B, cudaStream_t s) {
  return launch_conv_dgrad_tc(L, reinterpret_cast<const bf16*>(dz16), reinterpret_cast<const bf16*>(w16_), dx, B, s);
}
int isdqn_cast_bf16_launch(const float* src, void* dst16, int64_t n, cudaStream_t s) {
  if (n & 3) return ISDQN_E_INVALID;
  const int64_t n4 = n / 4;
  if (n4 == 0) return ISDQN_OK;
  int64_t grid = ceil_div<int64_t>(n4, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  ISDQN_PROF(s, "cast_bf16");
  cast_f32_bf16_kernel<<<(unsigned)grid, 256, 0, s>>>(src, reinterpret_cast<bf16*>(dst16), n4);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

int64_t isdqn_impala_tc_bytes(const isdqn_net* net, int batch);  // learner.cu

int isdqn_tc_train_dispatch(const isdqn_net* net, const isdqn_train* tr, const isdqn_batch* b, bool backward, bool update,
                            float* q_out, void* stream) {
  return tc_train(net, tr, b, backward, update, q_out, stream);
}

extern "C" int64_t isdqn_learn_workspace_tc_bytes(const isdqn_net* net, int32_t batch) {
  if (net && net->arch == ISDQN_ARCH_IMPALA) return isdqn_impala_tc_bytes(net, batch);
  Plan p;
  if (build_plan(net, &p) || batch < 1) return -1;
  if (!tc_eligible(p, net)) return 0;
  TcWorkspace t;
  carve_tc(p, 2 * batch, batch, &t);
  return t.total;
}

extern "C" int isdqn_cast_f32_to_bf16(const float* d_src, void* d_dst_bf16, int64_t n, void* stream) {
  if (!d_src || !d_dst_bf16 || n < 0 || (n & 3)) return ISDQN_E_INVALID;
  if (n == 0) return ISDQN_OK;
  const int64_t n4 = n / 4;
  int64_t grid = ceil_div<int64_t>(n4, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  ISDQN_PROF(as_stream(stream), "cast_params_bf16");
  cast_f32_bf16_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(d_src, reinterpret_cast<bf16*>(d_dst_bf16), n4);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

// Test / building-block entry: D[M][N] (fp32) = A B^T on the tensor cores, bf16 operands.
//   a_mn_major = 0: A is [M][lda] (K contiguous);  1: A is [K][lda] (M contiguous)
//   b_mn_major = 0: B is [N][ldb] (K contiguous);  1: B is [K][ldb] (N contiguous)
// splits > 1 writes `splits` partial products at d_c + z*M*N (the caller reduces them).
extern "C" int isdqn_tc_gemm_bf16(const void* d_a, int64_t lda, int32_t a_mn_major, const void* d_b, int64_t ldb,
                                  int32_t b_mn_major, float* d_c, int32_t M, int32_t N, int32_t K, int32_t splits,
                                  void* stream) {
  if (!d_a || !d_b || !d_c || M < 1 || N < 1 || K < 1 || splits < 1) return ISDQN_E_INVALID;
  if ((lda % 8) || (ldb % 8) || (N % 8)) return ISDQN_E_INVALID;  // 16-byte cp.async granularity
  const bf16* A = reinterpret_cast<const bf16*>(d_a);
  const bf16* B = reinterpret_cast<const bf16*>(d_b);
  cudaStream_t s = as_stream(stream);
  const int64_t ss = (int64_t)M * N;
  if (a_mn_major && b_mn_major) return launch_gemm_tc<true, true>(A, lda, B, ldb, d_c, N, ss, M, N, K, splits, s, "tc_gemm");
  if (a_mn_major) return launch_gemm_tc<true, false>(A, lda, B, ldb, d_c, N, ss, M, N, K, splits, s, "tc_gemm");
  if (b_mn_major) return launch_gemm_tc<false, true>(A, lda, B, ldb, d_c, N, ss, M, N, K, splits, s, "tc_gemm");
  return launch_gemm_tc<false, false>(A, lda, B, ldb, d_c, N, ss, M, N, K, splits, s, "tc_gemm");
}

int isdqn_trace_set_tc(unsigned long long* buf) { return isdqn::trace_set_local(buf) == cudaSuccess ? ISDQN_OK : ISDQN_E_CUDA; }
