// tcgen05 tile engine: D[128 x BN] (fp32, in TMEM) += A[128 x K] * B[BN x K]^T with bf16 operands staged in shared
// memory by the CTA's own threads (implicit-GEMM gathers), one elected thread issuing tcgen05.mma, the epilogue
// reading the accumulator back with tcgen05.ld — thread t owns output row t, so row-wise LayerNorm is thread-local.
//
// Shared-memory operand layout: the canonical NO-SWIZZLE ("interleave") UMMA layout of 8x16-byte core matrices,
// which — unlike the 128B-swizzle atoms TMA produces — accepts any extent that is a multiple of 8 and is cheap to
// fill from a gather.  One pipeline stage holds 64 elements of the reduction (K) dimension:
//     K-major  operand (reduction contiguous in the source):  byte(row r, 16B-chunk c of K) = (r/8)*1024 + c*128 + (r%8)*16
//     MN-major operand (row index contiguous in the source):  byte(K-row kk, 16B-chunk c of MN) = c*1024 + (kk/8)*128 + (kk%8)*16
// In both cases the descriptor has LBO = 128 B (next core matrix along K), SBO = 1024 B (next core matrix along
// M/N), and one MMA (K = 16) advances the start address by 256 B.  (cute/arch/mma_sm100_desc.hpp and
// cute/atom/mma_traits_sm100.hpp::make_umma_desc document the field semantics.)
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace isdqn {
namespace tc {

constexpr int kThreads = 128;   // 4 warps: all of them load and run the epilogue, thread 0 also issues the MMAs
constexpr int kBM = 128;        // UMMA_M (cta_group::1): accumulator row i lives in TMEM lane i
constexpr int kBK = 64;         // reduction elements per pipeline stage (4 MMAs of K=16)
constexpr int kABytes = kBM * kBK * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) __trap();  // a lost arrive must fault, not hang the GPU
}

// ---- proxies / fences ----------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- cp.async (LDGSTS) ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t dst_smem, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_smem), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- TMEM ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA ----------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor for kind::f16: bf16 x bf16 -> fp32, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__host__ __device__ constexpr int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// stage addressing helpers (bytes)
__device__ __forceinline__ uint32_t kmajor_off(int row, int chunk) { return (uint32_t)((row >> 3) * 1024 + chunk * 128 + (row & 7) * 16); }
__device__ __forceinline__ uint32_t mnmajor_off(int krow, int chunk) { return (uint32_t)(chunk * 1024 + (krow >> 3) * 128 + (krow & 7) * 16); }

template <int BN, int STAGES>
constexpr size_t smem_bytes() {
  return (size_t)STAGES * (kABytes + (size_t)BN * kBK * 2) + 1024;  // + alignment slack
}

// P (the problem) provides:
//   static constexpr int BN, STAGES; static constexpr bool A_MN, B_MN;
//   static constexpr int EXTRA_BYTES (shared scratch for lookup tables);
//   struct Ctx;  __device__ void init(Ctx&, uint8_t* extra, int m0, int n0, int tid) const;   (a __syncthreads follows)
//   __device__ void k_range(int split, int& kc_begin, int& kc_end) const;      (in units of 64-element chunks)
//   __device__ void load_a(const Ctx&, uint32_t stage_smem, int kc, int tid) const;   load_b(...)
//   __device__ void epilogue(const Ctx&, uint32_t tmem_lane_base, bool has_acc, int m0, int n0, int tid, int split) const;
template <class P>
__global__ void __launch_bounds__(kThreads) tc_gemm_kernel(const P p) {
  constexpr int BN = P::BN, STAGES = P::STAGES;
  constexpr int B_BYTES = BN * kBK * 2;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(BN);
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_sh;

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = smem_base + STAGES * kABytes;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN, split = blockIdx.z;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_sh, TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = tmem_base_sh;

  __shared__ __align__(16) uint8_t extra_sm[P::EXTRA_BYTES > 0 ? P::EXTRA_BYTES : 16];
  typename P::Ctx ctx;
  p.init(ctx, extra_sm, m0, n0, tid);
  __syncthreads();
  int kc_begin, kc_end;
  p.k_range(split, kc_begin, kc_end);
  const int nk = kc_end - kc_begin;
  constexpr uint32_t idesc = make_idesc(BN, P::A_MN, P::B_MN);

#pragma unroll
  for (int i = 0; i < STAGES - 1; ++i) {
    if (i < nk) {
      p.load_a(ctx, sA + i * kABytes, kc_begin + i, tid);
      p.load_b(ctx, sB + i * B_BYTES, kc_begin + i, tid);
    }
    cp_async_commit();
  }
  for (int i = 0; i < nk; ++i) {
    const int pf = i + STAGES - 1;  // chunk to prefetch now
    if (pf < nk) {
      const int ps = pf % STAGES;
      if (pf >= STAGES) mbar_wait(&empty_bar[ps], (uint32_t)((pf / STAGES - 1) & 1));  // MMAs of chunk pf-STAGES done
      p.load_a(ctx, sA + ps * kABytes, kc_begin + pf, tid);
      p.load_b(ctx, sB + ps * B_BYTES, kc_begin + pf, tid);
    }
    cp_async_commit();
    cp_async_wait<STAGES - 1>();  // this thread's part of chunk i has landed
    fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const int s = i % STAGES;
#pragma unroll
      for (int j = 0; j < kBK / 16; ++j) {
        const uint64_t adesc = make_smem_desc(sA + s * kABytes + j * 256, 128, 1024);
        const uint64_t bdesc = make_smem_desc(sB + s * B_BYTES + j * 256, 128, 1024);
        umma_bf16(tmem_d, adesc, bdesc, idesc, (i | j) != 0 ? 1u : 0u);
      }
      umma_commit(&empty_bar[s]);
      if (i == nk - 1) umma_commit(&done_bar);
    }
  }
  if (nk > 0) {
    mbar_wait(&done_bar, 0);
    tcgen05_fence_after();
  }
  p.epilogue(ctx, tmem_d + ((uint32_t)(warp * 32) << 16), nk > 0, m0, n0, tid, split);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

}  // namespace tc
}  // namespace isdqn
