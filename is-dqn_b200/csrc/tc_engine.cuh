// tcgen05 tile engine: D[128 x BN] (fp32, in TMEM) += A[128 x K] * B[BN x K]^T with bf16 operands staged in shared
// memory by producer warps (implicit-GEMM gathers), one elected thread issuing tcgen05.mma, epilogue warps reading
// the accumulator back with tcgen05.ld — thread t owns output row t, so row-wise LayerNorm is thread-local.
// Persistent and warp-specialised: see tc_gemm_kernel below.
//
// Shared-memory operand layouts (one pipeline stage holds 64 elements = 128 bytes of the reduction (K) dimension):
//   128-byte swizzle (the A stage always, the B stage when BN >= 64): 8-row atoms of 1024 B, row r at
//       (r/8)*1024 + (r%8)*128, its eight 16-byte chunks XOR-permuted by r%8 — conflict-free for a gather that writes
//       one row per 8 lanes, and what the tensor core reads fastest.  MN-major operands use the transposed atom
//       (64 MN elements x 8 K rows).
//   core-matrix ("interleave", no swizzle) layout for the BN == 32 B stage, which is narrower than a swizzle atom:
//       K-major  byte(row r, 16B-chunk c of K)     = (r/8)*1024 + c*128 + (r%8)*16
//       MN-major byte(K-row kk, 16B-chunk c of MN) = c*1024 + (kk/8)*128 + (kk%8)*16
// kmajor_off / mnmajor_off give the byte offsets, stage_desc the matching shared-memory descriptors; one MMA (K = 16)
// advances the start address inside the stage.  (cute/arch/mma_sm100_desc.hpp and
// cute/atom/mma_traits_sm100.hpp::make_umma_desc document the descriptor fields.)
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"

namespace isdqn {
namespace tc {

constexpr int kThreads = 128;   // rows of a tile = threads of the epilogue = lanes of the accumulator
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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// arrive on `bar` (counted in its expected arrivals) once all prior cp.async of this thread have completed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- TMA (tensor-map bulk copies; completion counted in bytes on an mbarrier) ---------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          dst_smem),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   dst_smem),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// TMA store: shared-memory tile -> global through a tensor map (rows outside the tensor are clipped); bulk-group completion
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory SOURCE of every committed bulk store may be overwritten
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// every committed bulk store has completed (its writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// problems that feed their stages with TMA declare `static constexpr bool TMA = true`
// problems whose epilogue stores through the per-warp shared-memory staging block declare `EP_STAGE = true`
template <class P, class = void>
struct uses_ep_stage : std::false_type {};
template <class P>
struct uses_ep_stage<P, std::void_t<decltype(P::EP_STAGE)>> : std::bool_constant<P::EP_STAGE> {};
constexpr int kEpStageWords = 32 * 20;  // per epilogue warp: 32 rows x 16 words, 20-word pitch

// problems whose epilogue keeps per-CTA state that is written out after the last tile declare `EP_FINISH = true`
template <class P, class = void>
struct uses_ep_finish : std::false_type {};
template <class P>
struct uses_ep_finish<P, std::void_t<decltype(P::EP_FINISH)>> : std::bool_constant<P::EP_FINISH> {};

// problems whose epilogue stores through TMA out of a shared-memory tile declare `EP_TILE_BYTES` (> 0): that many bytes of
// 1024-byte aligned dynamic shared memory behind the stage ring, handed to the problem with set_ep_tile()
template <class P, class = void>
struct ep_tile_bytes : std::integral_constant<int, 0> {};
template <class P>
struct ep_tile_bytes<P, std::void_t<decltype(P::EP_TILE_BYTES)>> : std::integral_constant<int, P::EP_TILE_BYTES> {};

template <class P, class = void>
struct uses_tma : std::false_type {};
template <class P>
struct uses_tma<P, std::void_t<decltype(P::TMA)>> : std::bool_constant<P::TMA> {};

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
// shared-memory matrix descriptor, version 1 (Blackwell); layout_type 0 = no swizzle, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}
// Descriptor of the q-th K=16 slice of one 64-deep stage.
//   swizzled (128B):  K-major : rows of 128 B (64 k), 8-row atoms of 1 KB, SBO = 1 KB; a K slice is 32 B inside the row
//                     MN-major: K-rows of 128 B (64 mn), 8-row atoms, SBO = 1 KB (next 8 k-rows), LBO = 8 KB (next 64 mn);
//                               a K slice is two atoms = 2 KB
//   no swizzle     :  8x16-byte core matrices, LBO = 128 B along K, SBO = 1 KB along M/N; a K slice is 256 B
template <bool MN_MAJOR, bool SWIZZLED>
__device__ __forceinline__ uint64_t stage_desc(uint32_t stage_addr, int q) {
  if (!SWIZZLED) return make_smem_desc(stage_addr + q * 256, 128, 1024, 0);
  if (MN_MAJOR) return make_smem_desc(stage_addr + q * 2048, 8192, 1024, 2);
  return make_smem_desc(stage_addr + q * 32, 16, 1024, 2);
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

// stage addressing helpers (bytes): (row, 16-byte chunk of K) for K-major, (K-row, 16-byte chunk of MN) for MN-major
template <bool SWIZZLED>
__device__ __forceinline__ uint32_t kmajor_off(int row, int chunk) {
  if (SWIZZLED) return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
  return (uint32_t)((row >> 3) * 1024 + chunk * 128 + (row & 7) * 16);
}
template <bool SWIZZLED>
__device__ __forceinline__ uint32_t mnmajor_off(int krow, int chunk) {
  if (SWIZZLED) return (uint32_t)((chunk >> 3) * 8192 + (krow >> 3) * 1024 + (krow & 7) * 128 + (((chunk & 7) ^ (krow & 7)) << 4));
  return (uint32_t)(chunk * 1024 + (krow >> 3) * 128 + (krow & 7) * 16);
}

template <int BN, int STAGES>
constexpr size_t smem_bytes() {
  return (size_t)STAGES * (kABytes + (size_t)BN * kBK * 2) + 1024;  // + alignment slack
}
template <class P>
constexpr size_t smem_bytes_of() {
  return smem_bytes<P::BN, P::STAGES>() + (size_t)ep_tile_bytes<P>::value;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr int kEpilogueWarps = 4;  // warps 0..3: warp w may only touch TMEM lanes 32w..32w+31
constexpr int kMmaWarp = 4;        // warp 4: lane 0 issues every tcgen05.mma
constexpr int kFirstProducerWarp = 5;

// Persistent, warp-specialised kernel.  Tiles (x fastest, then y, then z) are dealt round-robin to the CTAs; inside
// a CTA three roles run concurrently and meet only through mbarriers:
//   producers  gather the operands of k-chunk j into stage j % STAGES with 16-byte cp.async (LDG + convert for the uint8
//              fallback) as soon as the stage is free, and never wait for data: the arrival of each producer thread on
//              full[s] is triggered by the completion of its copies (cp.async.mbarrier.arrive.noinc);
//   MMA warp   waits full[s], issues 4 x tcgen05.mma (K = 16 each) into accumulator a = tile & 1, tcgen05.commit ->
//              empty[s]; after the last chunk of a tile commit -> tmem_full[a];
//   epilogue   waits tmem_full[a], reads the accumulator with tcgen05.ld (thread t owns row t), runs the problem's
//              epilogue, arrives on tmem_empty[a] — so the gather of tile i+1 overlaps the epilogue of tile i.
//
// P (the problem) provides:
//   static constexpr int BN, STAGES, PRODUCER_WARPS, EXTRA_BYTES, EP_FLOATS; static constexpr bool A_MN, B_MN, CHUNK_SYNC,
//   SYNC_STORES (the producers also fill the stage with plain st.shared);
//   __device__ void init_epilogue(ECtx&, float* ep_sm, int etid) const;        (EP_FLOATS > 0 only; an epilogue barrier follows)
//   (the A stage is always 128B-swizzled; the B stage is swizzled when BN >= 64, core-matrix layout for BN == 32)
//   struct PCtx, ECtx;
//   __device__ void init_cta(uint8_t* extra, int ptid) const;                  (producers only; a producer barrier follows)
//   __device__ void tile_producer(PCtx&, uint8_t* extra, int m0, int n0, int z, int ptid, int tile_iter) const;
//   __device__ void chunk_producer(PCtx&, uint8_t* extra, int kc, int ptid, int chunk_iter) const;   (if CHUNK_SYNC)
//   __device__ void tile_epilogue(ECtx&, int m0, int n0, int z, int etid) const;
//   __device__ void k_range(int z, int& kc_begin, int& kc_end) const;          (64-element chunks, never empty)
//   __device__ void load_a(const PCtx&, uint32_t stage_smem, int kc, int ptid) const;   load_b(...)
//   __device__ void epilogue(const ECtx&, uint32_t tmem_lane_base, int m0, int n0, int z, int etid) const;
// shared-memory objects of one CTA besides the dynamic stage ring (declared by the kernel, sized for its problem(s))
struct TcShared {
  uint64_t* full_bar;
  uint64_t* empty_bar;
  uint64_t* tmem_full_bar;
  uint64_t* tmem_empty_bar;
  uint32_t* tmem_base_sh;
  uint8_t* extra_sm;
  float* ep_sm;
  uint32_t* ep_stage;
};

// The whole CTA program for problem P: this CTA is number `cta` of the `n_ctas` that share P's tiles.
template <class P>
__device__ __forceinline__ void tc_gemm_body(const P& p, int tiles_x, int tiles_y, int tiles_z, int cta, int n_ctas,
                                             uint8_t* smem_dyn, const TcShared& sh) {
  constexpr int BN = P::BN, STAGES = P::STAGES;
  constexpr int B_BYTES = BN * kBK * 2;
  constexpr int PT = 32 * P::PRODUCER_WARPS;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(2 * BN);
  static_assert(2 * BN <= 512, "two accumulators must fit the 512 TMEM columns");
  uint64_t* full_bar = sh.full_bar;
  uint64_t* empty_bar = sh.empty_bar;
  uint64_t* tmem_full_bar = sh.tmem_full_bar;
  uint64_t* tmem_empty_bar = sh.tmem_empty_bar;
  uint8_t* extra_sm = sh.extra_sm;
  float* ep_sm = sh.ep_sm;
  uint32_t* ep_stage = sh.ep_stage;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = smem_base + STAGES * kABytes;
  const int n_tiles = tiles_x * tiles_y * tiles_z;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], uses_tma<P>::value ? 1 : PT);  // TMA: one arrive.expect_tx, the bytes complete the phase
      mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 32 * kEpilogueWarps);
    }
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(sh.tmem_base_sh, TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = *sh.tmem_base_sh;
  if (tid == 0) trace_mark(1);  // prologue done (barriers, TMEM)
#if !ISDQN_PDL_LATE
  pdl_trigger();  // (after the TMEM allocation: see common.cuh)
#endif

  if (warp >= kFirstProducerWarp) {
   if constexpr (uses_tma<P>::value) {
    // --------------------------------------------------------------------------------- TMA producer (one thread)
    // No gather code at all: per 64-deep K chunk the thread announces the stage's byte count on full[s] and issues the
    // problem's tensor-map copies, which land swizzled exactly as the MMA descriptors expect and bypass the LSU.
    if (warp == kFirstProducerWarp && lane == 0) {
      p.tma_prefetch();
      pdl_wait_then_trigger();
      int j = 0;
      for (int t = cta; t < n_tiles; t += n_ctas) {
        const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, tz = t / (tiles_x * tiles_y);
        int kb, ke;
        p.k_range(tz, kb, ke);
        typename P::PCtx ctx;
        p.tma_tile(ctx, tx, ty, tz);  // per-tile coordinates, computed once (this thread is the whole producer)
        const uint32_t tx_bytes = p.stage_tx_bytes(ctx);
        for (int kc = kb; kc < ke; ++kc, ++j) {
          const int s = j % STAGES;
          mbar_wait(&empty_bar[s], (uint32_t)(((j / STAGES) & 1) ^ 1));
          if (j == 0) trace_mark(2);
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          p.tma_load(ctx, sA + s * kABytes, sB + s * B_BYTES, &full_bar[s], kc);
        }
      }
    }
   } else {
    // ------------------------------------------------------------------------------------------ producers
    const int ptid = tid - 32 * kFirstProducerWarp;
    if (ptid < PT) {  // (a two-problem kernel is launched with the larger producer count of the two)
    p.init_cta(extra_sm, ptid);  // index tables from the kernel arguments only: overlaps the previous kernel's tail
    pdl_wait_then_trigger();
    named_bar_sync(1, PT);
    typename P::PCtx ctx;
    int j = 0;  // chunk counter of this CTA (stage = j % STAGES)
    int ti = 0;
    for (int t = cta; t < n_tiles; t += n_ctas, ++ti) {
      const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, tz = t / (tiles_x * tiles_y);
      const int m0 = tx * kBM, n0 = ty * BN;
      p.tile_producer(ctx, extra_sm, m0, n0, tz, ptid, ti);  // may publish per-row info in shared memory (parity ti & 1)
      named_bar_sync(1, PT);
      p.tile_rows(ctx, ptid);  // per-thread copy (registers) of the rows this thread gathers for the whole tile
      if (ptid == 0 && ti == 0) trace_mark(2);  // tables + row info ready: first gather is issued next
      int kb, ke;
      p.k_range(tz, kb, ke);
      for (int kc = kb; kc < ke; ++kc, ++j) {
        const int s = j % STAGES;
        if (P::CHUNK_SYNC) {  // per-chunk row info (weight gradient: the reduction rows change every chunk)
          p.chunk_producer(ctx, extra_sm, kc, ptid, j);
          named_bar_sync(1, PT);
        }
        mbar_wait(&empty_bar[s], (uint32_t)(((j / STAGES) & 1) ^ 1));  // stage free (first lap: passes)
        p.load_a(ctx, sA + s * kABytes, kc, ptid);
        p.load_b(ctx, sB + s * B_BYTES, kc, ptid);
        // This thread's arrival on the stage's "full" barrier is performed by the hardware when its asynchronous copies
        // of the chunk have landed: the producer never waits for data, so every free stage of the ring is in flight and
        // a chunk is published the moment it is complete.
        if (P::SYNC_STORES) fence_proxy_async();  // (plain st.shared of this thread, uint8 path)
        cp_async_mbar_arrive_noinc(&full_bar[s]);
      }
    }
    cp_async_wait_all();
    }
   }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, P::A_MN, P::B_MN);
      int j = 0, ti = 0;
      for (int t = cta; t < n_tiles; t += n_ctas, ++ti) {
        const int tz = t / (tiles_x * tiles_y);
        int kb, ke;
        p.k_range(tz, kb, ke);
        const int a = ti & 1;
        mbar_wait(&tmem_empty_bar[a], (uint32_t)((((ti >> 1) & 1)) ^ 1));  // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_d + a * BN;
        for (int kc = kb; kc < ke; ++kc, ++j) {
          const int s = j % STAGES;
          mbar_wait(&full_bar[s], (uint32_t)((j / STAGES) & 1));
          if (j == 0) trace_mark(3);  // first chunk landed in shared memory
          fence_proxy_async();  // the LDGSTS / st.shared writes of the producers -> the tensor core's (async proxy) reads
          tcgen05_fence_after();
#pragma unroll
          for (int q = 0; q < kBK / 16; ++q) {
            const uint64_t adesc = stage_desc<P::A_MN, true>(sA + s * kABytes, q);
            const uint64_t bdesc = stage_desc<P::B_MN, P::B_SW>(sB + s * B_BYTES, q);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kc > kb || q > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[a]);
        if (ti == 0) trace_mark(4);  // all MMAs of the first tile issued
      }
    }
  } else {
    // -------------------------------------------------------------------------------------------- epilogue
    typename P::ECtx ectx;
    int ti = 0;
    if constexpr (ep_tile_bytes<P>::value > 0) p.set_ep_tile(ectx, sB + STAGES * B_BYTES);
    pdl_wait_then_trigger();
    if (P::EP_FLOATS > 0) {  // per-channel epilogue parameters -> shared memory, once per CTA, while the first tile is gathered
      p.init_epilogue(ectx, ep_sm, tid);
      named_bar_sync(2, 32 * kEpilogueWarps);
    }
    for (int t = cta; t < n_tiles; t += n_ctas, ++ti) {
      const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, tz = t / (tiles_x * tiles_y);
      const int m0 = tx * kBM, n0 = ty * BN;
      const int a = ti & 1;
      p.tile_epilogue(ectx, m0, n0, tz, tid);
      mbar_wait(&tmem_full_bar[a], (uint32_t)((ti >> 1) & 1));
      if (tid == 0 && ti == 0) trace_mark(5);  // first accumulator complete
      tcgen05_fence_after();
      p.epilogue(ectx, tmem_d + a * BN + ((uint32_t)(warp * 32) << 16), m0, n0, tz, tid,
                 uses_ep_stage<P>::value ? ep_stage + warp * kEpStageWords : nullptr);
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[a]);
    }
    if constexpr (uses_ep_finish<P>::value) p.finish_epilogue(ectx, cta, tid);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) trace_mark(6);  // all roles of CTA 0 done
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

template <class P>
__global__ void __launch_bounds__(32 * (kFirstProducerWarp + P::PRODUCER_WARPS), P::MIN_CTAS)
tc_gemm_kernel(const __grid_constant__ P p, int tiles_x, int tiles_y, int tiles_z) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[P::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[P::STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_sh;
  __shared__ __align__(16) uint8_t extra_sm[P::EXTRA_BYTES > 0 ? P::EXTRA_BYTES : 16];
  __shared__ __align__(16) float ep_sm[P::EP_FLOATS > 0 ? P::EP_FLOATS : 4];
  __shared__ __align__(16) uint32_t ep_stage[uses_ep_stage<P>::value ? kEpilogueWarps * kEpStageWords : 4];
  trace_kernel_start();
  const TcShared sh = {full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, &tmem_base_sh, extra_sm, ep_sm, ep_stage};
  tc_gemm_body<P>(p, tiles_x, tiles_y, tiles_z, (int)blockIdx.x, (int)gridDim.x, smem_dyn, sh);
}

// Two independent problems in ONE launch: CTAs [0, ctas1) run P1's tiles, the others P2's.  For short dependent chains
// (the batch-32 backward pass) where the weight gradient and the input gradient of a layer consume the same dz: one launch
// latency instead of two, and the two half-empty waves share the GPU.  Both problems must use the same thread count.
__host__ __device__ constexpr int tc_cmax(int a, int b) { return a > b ? a : b; }
template <class P1, class P2>
__global__ void __launch_bounds__(32 * (kFirstProducerWarp + tc_cmax(P1::PRODUCER_WARPS, P2::PRODUCER_WARPS)), 1)
tc_gemm2_kernel(const __grid_constant__ P1 p1, int t1x, int t1y, int t1z, const __grid_constant__ P2 p2, int t2x, int t2y, int t2z,
                int ctas1) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[tc_cmax(P1::STAGES, P2::STAGES)];
  __shared__ __align__(8) uint64_t empty_bar[tc_cmax(P1::STAGES, P2::STAGES)];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_sh;
  __shared__ __align__(16) uint8_t extra_sm[tc_cmax(tc_cmax(P1::EXTRA_BYTES, P2::EXTRA_BYTES), 16)];
  __shared__ __align__(16) float ep_sm[tc_cmax(tc_cmax(P1::EP_FLOATS, P2::EP_FLOATS), 4)];
  __shared__ __align__(16) uint32_t ep_stage[(uses_ep_stage<P1>::value || uses_ep_stage<P2>::value) ? kEpilogueWarps * kEpStageWords : 4];
  trace_kernel_start();
  const TcShared sh = {full_bar, empty_bar, tmem_full_bar, tmem_empty_bar, &tmem_base_sh, extra_sm, ep_sm, ep_stage};
  if ((int)blockIdx.x < ctas1)
    tc_gemm_body<P1>(p1, t1x, t1y, t1z, (int)blockIdx.x, ctas1, smem_dyn, sh);
  else
    tc_gemm_body<P2>(p2, t2x, t2y, t2z, (int)blockIdx.x - ctas1, (int)gridDim.x - ctas1, smem_dyn, sh);
}

}  // namespace tc
}  // namespace isdqn
