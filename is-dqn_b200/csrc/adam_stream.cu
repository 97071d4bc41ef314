// Fused rank-B Dense weight gradient + Adam, streaming version (see dense_wgrad_adam_kernel in learner_kernels.cuh for
// what is computed; this is the same arithmetic, element for element, behind a bulk-copy pipeline).
//
// The first fused kernel loads p / mu / nu with per-thread 16-byte loads: every CTA alternates between "all loads in
// flight" and "all threads computing", and ncu showed it latency-bound at half of the HBM peak (profiles/r01_adam_ncu.md).
// Here the memory side is decoupled from the warps, the Blackwell way:
//   * one persistent CTA per SM, 8 consumer warps + 1 producer warp, a ring of STAGES shared-memory stages;
//   * a tile is R whole kernel rows = 4096 CONTIGUOUS elements (8 rows at N = 512); a stage holds that tile of p, mu and
//     nu (3 x 16 KB) and the bf16 shadow tile that goes out;
//   * one producer lane moves a tile with three 16 KB 1-D bulk async copies (cp.async.bulk -> UBLKCP, the TMA unit) onto
//     the stage's "full" mbarrier (byte-counted) and — once the consumers have arrived on the stage's "computed"
//     mbarrier — stores p, mu, nu and the shadow back with four bulk copies from shared memory (a first version with
//     one 1 KB copy per row and array spent 6.6 us per tile issuing 112 small copies: the copy RATE of the unit bound
//     it); a stage is refilled one tile after its stores were committed (cp.async.bulk.wait_group.read 1), so
//     STAGES - 1 tiles (96 KB per SM) are in flight all the time; the act^T slice of a tile rides on the same barrier
//     through cp.async + cp.async.mbarrier.arrive.noinc;
//   * the consumers compute the R x N gradient tile with mma.sync.m16n8k16 (bf16 operands staged once per CTA / per
//     tile, fp32 accumulation), then update the stage in place: conflict-free 16-byte shared-memory accesses, the Adam
//     element update of learner_kernels.cuh, fence.proxy.async, arrive.
// Tiles are dealt round-robin over the CTAs, so at any moment the whole GPU streams one contiguous window of each array.  Every other leaf of the flat vector is updated by the consumers after their last tile.
#include <cstdlib>

#include "learner_kernels.cuh"
#include "tc_engine.cuh"

using namespace isdqn;
using isdqn::tc::fence_proxy_async;
using isdqn::tc::mbar_arrive;
using isdqn::tc::mbar_arrive_expect_tx;
using isdqn::tc::mbar_fence_init;
using isdqn::tc::mbar_init;
using isdqn::tc::mbar_wait;
using isdqn::tc::named_bar_sync;
using isdqn::tc::smem_u32;

namespace {

constexpr int kTpConsumers = 256, kTpThreads = kTpConsumers + 32;
constexpr int kTpTileElems = 4096;                                  // R kernel rows x N columns, contiguous in memory
constexpr int kTpTileBytes = kTpTileElems * 4;                      // one fp32 tile
constexpr int kTpStageBytes = 3 * kTpTileBytes + kTpTileElems * 2;  // p, mu, nu + bf16 shadow out
constexpr int kTpLdA = 24;  // bf16 per staged act row (16 + 8: the 8 rows of an ldmatrix fall in 8 different banks)
constexpr int kTpMaxStages = 3;

// pol != 0: an L2 eviction-priority policy (createpolicy) rides on the copy.  The optimiser state of the Dense kernel
// (p, mu, nu: 48 MB) and its bf16 shadow (8 MB) are written here and read again by the next step — 56 MB of the 126 MB
// L2: marked evict_last they are still resident when the next update (and the next forward / input-gradient GEMM, for
// the shadow) comes for them, so the pass streams from L2 instead of HBM (ISDQN_ADAM_L2=0: no hints).
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  if (pol)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
  if (pol)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem),
                 "r"(bytes), "l"(pol)
                 : "memory");
  else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all_but_one() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

size_t tp_smem_bytes(int B, int N, int stages) {
  const int Bp = (B + 15) / 16 * 16, R = kTpTileElems / N;
  return (size_t)stages * kTpStageBytes + (size_t)Bp * (N + 8) * 2 + (size_t)stages * Bp * kTpLdA * 2 + (size_t)R * (N + 8) * 4 +
         2 * kTpMaxStages * sizeof(uint64_t) + 128;
}

__global__ void __launch_bounds__(kTpThreads, 1)
dense_wgrad_adam_stream_kernel(float* __restrict__ p_all, const float* __restrict__ g_all, float* __restrict__ mu_all,
                               float* __restrict__ nu_all, const int32_t* __restrict__ count, float lr, float b1, float b2,
                               float eps, __nv_bfloat16* __restrict__ shadow_all, int64_t w_off, int64_t n_total4,
                               const __nv_bfloat16* __restrict__ act, int64_t lda, const __nv_bfloat16* __restrict__ dz, int B,
                               int Kin, int N, int stages, int l2_hint) {
  // Programmatic dependent launch: this kernel's predecessor is the partial reduction, which only produces `g_all` —
  // the gradients of the OTHER leaves.  p / mu / nu / shadow of the Dense kernel, dz, act and the step counter were all
  // written by grids that completed before the predecessor even started (every kernel of the chain waits for its own
  // predecessor), and the predecessor touches none of them.  So the tile pipeline starts at once, under the
  // predecessor's execution; only the consumers' rest-leaf phase at the end waits (griddepcontrol.wait) for `g_all`.
#ifndef ISDQN_ADAM_EARLY
#define ISDQN_ADAM_EARLY 1  // 0: wait for the predecessor at the top like every other kernel (A/B builds)
#endif
#if ISDQN_ADAM_EARLY
  trace_kernel_start();
#else
  pdl_sync();
#endif
  extern __shared__ __align__(128) unsigned char tp_smem[];
  __shared__ float s_c[2];
  const int Bp = (B + 15) / 16 * 16;
  const int R = kTpTileElems / N;  // 8 rows at N = 512, 16 at N = 256
  const int ldb = N + 8;           // bf16 per staged dz row / fp32 per gradient row
  unsigned char* stage0 = tp_smem;
  __nv_bfloat16* s_dz = reinterpret_cast<__nv_bfloat16*>(tp_smem + (size_t)stages * kTpStageBytes);   // [Bp][ldb]
  __nv_bfloat16* s_act = s_dz + (size_t)Bp * ldb;                                                     // [stages][Bp][kTpLdA]
  float* s_g = reinterpret_cast<float*>(s_act + (size_t)stages * Bp * kTpLdA);                        // [R][ldb]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_g + (size_t)R * ldb);
  uint64_t* computed = full + kTpMaxStages;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full + s, 1 + 32);            // the expect_tx arrival + every producer lane once its act pieces landed
      mbar_init(computed + s, kTpConsumers);
    }
    mbar_fence_init();
  }
  __syncthreads();  // barriers initialised: the producer starts the first tiles while the consumers stage dz
  const int n_tiles = Kin / R;  // (Kin % R == 0, checked by the launcher)
  const int n_my = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == kTpConsumers / 32) {
    // ------------------------------------------------------------------------------------------ producer warp
    const uint64_t pol = l2_hint ? l2_policy_evict_last() : 0ull;
    auto issue_loads = [&](int it) {
      const int s = it % stages;
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      if (lane == 0) {
        mbar_arrive_expect_tx(full + s, 3u * kTpTileBytes);
        const int64_t off = w_off + tile * kTpTileElems;
        const uint32_t dst = smem_u32(stage0 + (size_t)s * kTpStageBytes);
        bulk_g2s(dst, p_all + off, kTpTileBytes, full + s, pol);
        bulk_g2s(dst + kTpTileBytes, mu_all + off, kTpTileBytes, full + s, pol);
        bulk_g2s(dst + 2 * kTpTileBytes, nu_all + off, kTpTileBytes, full + s, pol);
      }
      // act^T of these R rows: R / 8 16-byte pieces per batch row, asynchronous (the arrival is performed by the copy)
      const int64_t row0 = tile * R;
      const int pieces = R / 8;
      for (int i = lane; i < Bp * pieces; i += 32) {
        const int b = i / pieces, h = i - b * pieces;
        isdqn::tc::cp_async16(smem_u32(s_act + ((size_t)s * Bp + b) * kTpLdA + h * 8), act + (int64_t)min(b, B - 1) * lda + row0 + h * 8,
                              b < B);
      }
      isdqn::tc::cp_async_mbar_arrive_noinc(full + s);
    };
    for (int it = 0; it < min(stages, n_my); ++it) issue_loads(it);
    for (int j = 0; j < n_my; ++j) {
      const int s = j % stages;
      mbar_wait(computed + s, (uint32_t)(j / stages) & 1u);
      if (lane == 0) {
        const int64_t off = w_off + ((int64_t)blockIdx.x + (int64_t)j * gridDim.x) * kTpTileElems;
        const uint32_t src = smem_u32(stage0 + (size_t)s * kTpStageBytes);
        bulk_s2g(p_all + off, src, kTpTileBytes, pol);
        bulk_s2g(mu_all + off, src + kTpTileBytes, kTpTileBytes, pol);
        bulk_s2g(nu_all + off, src + 2 * kTpTileBytes, kTpTileBytes, pol);
        if (shadow_all) bulk_s2g(shadow_all + off, src + 3 * kTpTileBytes, kTpTileElems * 2, pol);
        bulk_commit();
        // refill the stage whose stores were committed one tile ago: they have had a whole tile to leave shared memory
        if (j >= 1 && j - 1 + stages < n_my) bulk_wait_read_all_but_one();
      }
      __syncwarp();
      if (j >= 1 && j - 1 + stages < n_my) issue_loads(j - 1 + stages);
    }
    if (lane == 0) bulk_wait_all();
    return;
  }
  // ------------------------------------------------------------------------------------------------ consumers
  if (tid == kTpConsumers - 1 || tid == kTpConsumers - 33) {  // 1 / (1 - b^t), one power per thread (two warps)
    const int which = tid == kTpConsumers - 1 ? 0 : 1;
    s_c[which] = (float)(1.0 / (1.0 - pow_int((double)(which == 0 ? b1 : b2), *count)));
  }
  for (int i = tid; i < Bp * (N / 8); i += kTpConsumers) {  // dz[B][N] (N % 8 == 0), zero rows up to Bp
    const int b = i / (N / 8), c = (i - b * (N / 8)) * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (b < B) v = *reinterpret_cast<const uint4*>(dz + (int64_t)b * N + c);
    *reinterpret_cast<uint4*>(s_dz + (size_t)b * ldb + c) = v;
  }
  named_bar_sync(1, kTpConsumers);
  const float c1 = s_c[0], c2 = s_c[1];
  const float ob1 = 1.0f - b1, ob2 = 1.0f - b2;
  const int lm = lane >> 3, lj = lane & 7;
  const int wcols = N / 8;  // columns of the gradient tile per warp (a multiple of 16)
  const __nv_bfloat16* b_base = s_dz + (size_t)(lj + (lm & 1) * 8) * ldb + warp * wcols + (lm >> 1) * 8;
  const int gq = lane >> 2, tq = lane & 3;
  for (int j = 0; j < n_my; ++j) {
    const int s = j % stages;
    mbar_wait(full + s, (uint32_t)(j / stages) & 1u);
    {
      const __nv_bfloat16* a_base = s_act + (size_t)s * Bp * kTpLdA + (size_t)(lj + (lm >> 1) * 8) * kTpLdA + (lm & 1) * 8;
      uint32_t af[kDwaMaxB / 16][4];
#pragma unroll
      for (int ks = 0; ks < kDwaMaxB / 16; ++ks)
        if (ks * 16 < Bp) ldmatrix_x4_trans(af[ks], a_base + (size_t)ks * 16 * kTpLdA);
      for (int n0 = 0; n0 < wcols; n0 += 16) {
        float acc[2][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[0][e] = acc[1][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < kDwaMaxB / 16; ++ks) {
          if (ks * 16 < Bp) {
            uint32_t bf[4];
            ldmatrix_x4_trans(bf, b_base + (size_t)ks * 16 * ldb + n0);
            mma_bf16_m16n8k16(acc[0], af[ks], bf[0], bf[1]);
            mma_bf16_m16n8k16(acc[1], af[ks], bf[2], bf[3]);
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          float* dst = s_g + (size_t)gq * ldb + warp * wcols + n0 + nt * 8 + tq * 2;
          *reinterpret_cast<float2*>(dst) = make_float2(acc[nt][0], acc[nt][1]);
          if (R > 8) *reinterpret_cast<float2*>(dst + 8 * ldb) = make_float2(acc[nt][2], acc[nt][3]);
        }
      }
    }
    named_bar_sync(1, kTpConsumers);  // the gradient tile is complete
    unsigned char* st = stage0 + (size_t)s * kTpStageBytes;
#pragma unroll
    for (int i = 0; i < kTpTileElems / 4 / kTpConsumers; ++i) {
      const int e = (i * kTpConsumers + tid) * 4;  // consecutive threads, consecutive 16-byte pieces: conflict-free
      const int row = e / N, col = e - row * N;
      float4 pv = *reinterpret_cast<const float4*>(st + (size_t)e * 4);
      float4 mv = *reinterpret_cast<const float4*>(st + kTpTileBytes + (size_t)e * 4);
      float4 vv = *reinterpret_cast<const float4*>(st + 2 * kTpTileBytes + (size_t)e * 4);
      const float4 g = *reinterpret_cast<const float4*>(s_g + (size_t)row * ldb + col);
      ISDQN_ADAM_ELEM(g, mv, vv, pv, x) ISDQN_ADAM_ELEM(g, mv, vv, pv, y) ISDQN_ADAM_ELEM(g, mv, vv, pv, z)
      ISDQN_ADAM_ELEM(g, mv, vv, pv, w)
      *reinterpret_cast<float4*>(st + (size_t)e * 4) = pv;
      *reinterpret_cast<float4*>(st + kTpTileBytes + (size_t)e * 4) = mv;
      *reinterpret_cast<float4*>(st + 2 * kTpTileBytes + (size_t)e * 4) = vv;
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      *reinterpret_cast<uint2*>(st + 3 * kTpTileBytes + (size_t)e * 2) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
    fence_proxy_async();  // these shared-memory writes are read by the bulk stores (async proxy)
    mbar_arrive(computed + s);
    named_bar_sync(1, kTpConsumers);  // everybody has read the gradient tile: the next one may be written
  }
  // ---- every other leaf: the plain update over [0, n_total4) minus the Dense kernel's range
#if ISDQN_ADAM_EARLY
  pdl_wait();     // the predecessor's gradients are complete and visible
  pdl_trigger();  // (the successor may start its prologue; it waits for this grid's completion before it reads anything)
#endif
  {
    const int64_t skip_begin4 = w_off / 4, skip_len4 = (int64_t)Kin * N / 4;
    const int64_t n4 = n_total4 - skip_len4;
    const int64_t stride = (int64_t)gridDim.x * kTpConsumers;
    for (int64_t q = (int64_t)blockIdx.x * kTpConsumers + tid; q < n4; q += stride) {
      const int64_t i = q < skip_begin4 ? q : q + skip_len4;
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
  }
}

}  // namespace

int isdqn_trace_set_adam_stream(unsigned long long* buf) { return isdqn::trace_set_local(buf) == cudaSuccess ? ISDQN_OK : ISDQN_E_CUDA; }

// Returns ISDQN_E_UNSUPPORTED when the shape does not fit (the caller falls back to dense_wgrad_adam_kernel).
int isdqn_dense_wgrad_adam_stream_launch(float* d_params, const float* d_grads, float* d_mu, float* d_nu, const int32_t* d_count,
                                         float lr, float b1, float b2, float eps, void* d_shadow_bf16, int64_t n_total,
                                         int64_t w_off, const void* d_act_bf16, int64_t lda, const void* d_dz_bf16, int B,
                                         int Kin, int N, void* stream) {
  static const int mode = [] {
    const char* e = getenv("ISDQN_ADAM_STREAM");  // 0 = off, otherwise on (default)
    return (e && e[0] == '0') ? 0 : 1;
  }();
  if (!mode) return ISDQN_E_UNSUPPORTED;
  // whole kernel rows per tile: N in {256, 512} (16 / 8 rows of 4096 contiguous elements)
  if (B < 1 || B > kDwaMaxB || (N != 256 && N != 512) || Kin % (kTpTileElems / N) || w_off % 8 || lda % 8 || (n_total & 3))
    return ISDQN_E_UNSUPPORTED;
  int stages = kTpMaxStages;
  while (stages > 1 && tp_smem_bytes(B, N, stages) > 225 * 1024) --stages;
  if (stages < 2) return ISDQN_E_UNSUPPORTED;
  const size_t smem = tp_smem_bytes(B, N, stages);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    ISDQN_CUDA_CHECK(cudaFuncSetAttribute(dense_wgrad_adam_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  }
  static const bool l2_hint = [] {
    const char* e = getenv("ISDQN_ADAM_L2");
    return !(e && e[0] == '0');
  }();
  ISDQN_PROF(as_stream(stream), "dense_wgrad_adam");
  ISDQN_CUDA_CHECK(launch_pdl(dense_wgrad_adam_stream_kernel, dim3((unsigned)kNumSMs), dim3(kTpThreads), smem, as_stream(stream), d_params,
                              d_grads, d_mu, d_nu, d_count, lr, b1, b2, eps, reinterpret_cast<__nv_bfloat16*>(d_shadow_bf16), w_off,
                              n_total / 4, reinterpret_cast<const __nv_bfloat16*>(d_act_bf16), lda,
                              reinterpret_cast<const __nv_bfloat16*>(d_dz_bf16), B, Kin, N, stages, l2_hint ? 1 : 0));
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}
