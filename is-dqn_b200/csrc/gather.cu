// Frame-stack gather out of the device frame ring: replaces ReplayBuffer.sample's itemgetter + unpack + np.stack
// (slimdqn/sample_collection/replay_buffer.py:198-213).
//
// Fast path (uint8 frames, stack 4 — the Atari layout): one CTA per drawn element.  One elected thread issues a
// 1-D bulk async copy (cp.async.bulk -> UBLKCP, the TMA unit) per UNIQUE frame of the element into shared
// memory (state and next_state share stack-n frames, so 5 copies of 7056 B at n=1 instead of 8) and arms an
// mbarrier with the byte count.  After the wait every thread reads one pixel quad of each of the four frames
// of a stack (4 conflict-free LDS.32), transposes the 4x4 bytes with PRMT into HWC-interleaved order and
// writes one 16-byte, fully coalesced store.  The /255 normalisation to f32/bf16 (architectures/dqn.py:51) is
// fused into the same pass on request.  Zero padding is a ring slot that holds zeros: no branches.
//
// Generic path: any element size / stack (int64 test frames, float32 LunarLander vectors).
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"

namespace isdqn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
// Bounded wait: a byte-count mismatch must fault (trap) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kGatherThreads = 256;

template <int OUT>
struct OutTraits;
template <>
struct OutTraits<ISDQN_OUT_RAW> {
  typedef uint8_t T;
};
template <>
struct OutTraits<ISDQN_OUT_F32> {
  typedef float T;
};
template <>
struct OutTraits<ISDQN_OUT_BF16> {
  typedef __nv_bfloat16 T;
};

__device__ __forceinline__ float norm255(uint32_t b) { return __fdiv_rn((float)b, 255.0f); }

// 4 words (one per frame, 4 pixels each) -> 4 output words (one per pixel, 4 frames each)
__device__ __forceinline__ void transpose4x4_bytes(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3, uint32_t o[4]) {
  const uint32_t a_lo = __byte_perm(f0, f1, 0x5140);  // f0.b0 f1.b0 f0.b1 f1.b1
  const uint32_t a_hi = __byte_perm(f0, f1, 0x7362);  // f0.b2 f1.b2 f0.b3 f1.b3
  const uint32_t b_lo = __byte_perm(f2, f3, 0x5140);
  const uint32_t b_hi = __byte_perm(f2, f3, 0x7362);
  o[0] = __byte_perm(a_lo, b_lo, 0x5410);  // f0.b0 f1.b0 f2.b0 f3.b0
  o[1] = __byte_perm(a_lo, b_lo, 0x7632);
  o[2] = __byte_perm(a_hi, b_hi, 0x5410);
  o[3] = __byte_perm(a_hi, b_hi, 0x7632);
}

template <int OUT>
__global__ void __launch_bounds__(kGatherThreads)
gather_stack4_u8_kernel(const uint8_t* __restrict__ frames, int64_t frame_stride, int frame_bytes,
                        const int32_t* __restrict__ elem_frames, const int64_t* __restrict__ elem_action,
                        const double* __restrict__ elem_reward, const uint8_t* __restrict__ elem_terminal,
                        const int32_t* __restrict__ slots, int n, void* __restrict__ out_state_v,
                        void* __restrict__ out_next_v, int64_t* __restrict__ out_action,
                        double* __restrict__ out_reward, uint8_t* __restrict__ out_terminal) {
  typedef typename OutTraits<OUT>::T T;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int frame_off[8];  // shared-memory byte offset of the copy serving each of the 8 frame refs
  const int tid = threadIdx.x;
  const int fstride_s = (frame_bytes + 15) & ~15;  // 16-byte aligned staging stride
  T* out_state = reinterpret_cast<T*>(out_state_v);
  T* out_next = reinterpret_cast<T*>(out_next_v);

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  uint32_t phase = 0;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int slot = slots[e];
    if (tid == 0) {
      int ref[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ref[j] = elem_frames[(int64_t)slot * 8 + j];
      int n_unique = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int found = -1;
        for (int k = 0; k < j; ++k)
          if (ref[k] == ref[j]) {
            found = frame_off[k];
            break;
          }
        frame_off[j] = found >= 0 ? found : (n_unique++) * fstride_s;
      }
      mbar_arrive_expect_tx(&bar, (uint32_t)(n_unique * fstride_s));
      int issued = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (frame_off[j] == issued * fstride_s) {  // first reference of this unique frame
          bulk_g2s(smem + frame_off[j], frames + (int64_t)ref[j] * frame_stride, (uint32_t)fstride_s, &bar);
          ++issued;
        }
      }
      if (out_action) out_action[e] = elem_action[slot];
      if (out_reward) out_reward[e] = elem_reward[slot];
      if (out_terminal) out_terminal[e] = elem_terminal[slot];
    }
    __syncthreads();  // frame_off visible
    mbar_wait(&bar, phase);
    phase ^= 1;

    const int n_quads = frame_bytes >> 2;  // frame_bytes % 4 == 0 on this path
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      T* dst = (which == 0 ? out_state : out_next) + (int64_t)e * frame_bytes * 4;
      const uint32_t* s0 = reinterpret_cast<const uint32_t*>(smem + frame_off[which * 4 + 0]);
      const uint32_t* s1 = reinterpret_cast<const uint32_t*>(smem + frame_off[which * 4 + 1]);
      const uint32_t* s2 = reinterpret_cast<const uint32_t*>(smem + frame_off[which * 4 + 2]);
      const uint32_t* s3 = reinterpret_cast<const uint32_t*>(smem + frame_off[which * 4 + 3]);
      for (int q = tid; q < n_quads; q += kGatherThreads) {
        uint32_t o[4];
        transpose4x4_bytes(s0[q], s1[q], s2[q], s3[q], o);
        if (OUT == ISDQN_OUT_RAW) {
          reinterpret_cast<uint4*>(dst)[q] = make_uint4(o[0], o[1], o[2], o[3]);
        } else if (OUT == ISDQN_OUT_F32) {
          float4* d4 = reinterpret_cast<float4*>(dst) + (int64_t)q * 4;
#pragma unroll
          for (int p = 0; p < 4; ++p)
            d4[p] = make_float4(norm255(o[p] & 0xff), norm255((o[p] >> 8) & 0xff), norm255((o[p] >> 16) & 0xff),
                                norm255(o[p] >> 24));
        } else {
          uint4* d4 = reinterpret_cast<uint4*>(dst) + (int64_t)q * 2;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t w[4];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              const uint32_t v = o[h * 2 + p];
              __nv_bfloat162 lo = __floats2bfloat162_rn(norm255(v & 0xff), norm255((v >> 8) & 0xff));
              __nv_bfloat162 hi = __floats2bfloat162_rn(norm255((v >> 16) & 0xff), norm255(v >> 24));
              w[p * 2 + 0] = *reinterpret_cast<uint32_t*>(&lo);
              w[p * 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
            }
            d4[h] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    __syncthreads();  // staging buffers are reused by the next element
  }
}

// Generic: out[e][p][j] = frame(ref[j])[p], elements of ES bytes.  One thread per output element, j fastest.
template <int ES>
__global__ void __launch_bounds__(256)
gather_generic_kernel(const uint8_t* __restrict__ frames, int64_t frame_stride, int frame_elems, int stack,
                      const int32_t* __restrict__ elem_frames, const int64_t* __restrict__ elem_action,
                      const double* __restrict__ elem_reward, const uint8_t* __restrict__ elem_terminal,
                      const int32_t* __restrict__ slots, int n, uint8_t* __restrict__ out_state,
                      uint8_t* __restrict__ out_next, int64_t* __restrict__ out_action,
                      double* __restrict__ out_reward, uint8_t* __restrict__ out_terminal) {
  typedef typename std::conditional<ES == 1, uint8_t,
                                    typename std::conditional<ES == 2, uint16_t,
                                                              typename std::conditional<ES == 4, uint32_t, uint64_t>::type>::type>::type E;
  const int e = blockIdx.y;
  const int slot = slots[e];
  const int per = frame_elems * stack;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (out_action) out_action[e] = elem_action[slot];
    if (out_reward) out_reward[e] = elem_reward[slot];
    if (out_terminal) out_terminal[e] = elem_terminal[slot];
  }
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 2 * per; t += gridDim.x * blockDim.x) {
    const int which = t >= per;
    const int r = t - which * per;
    const int p = r / stack, j = r - p * stack;
    uint8_t* outp = which ? out_next : out_state;
    if (!outp) continue;
    const int ref = elem_frames[(int64_t)slot * 2 * stack + which * stack + j];
    const E v = reinterpret_cast<const E*>(frames + (int64_t)ref * frame_stride)[p];
    reinterpret_cast<E*>(outp)[(int64_t)e * per + r] = v;
  }
}

}  // namespace isdqn

using namespace isdqn;

template <int OUT>
static int launch_fast(const uint8_t* d_frames, int64_t frame_stride, int32_t frame_bytes, const int32_t* d_elem_frames,
                       const int64_t* d_elem_action, const double* d_elem_reward, const uint8_t* d_elem_terminal,
                       const int32_t* d_slots, int32_t n, void* d_out_state, void* d_out_next, int64_t* d_out_action,
                       double* d_out_reward, uint8_t* d_out_terminal, cudaStream_t stream) {
  const int fstride_s = (frame_bytes + 15) & ~15;
  const size_t smem = (size_t)8 * fstride_s;
  if (smem > 200 * 1024) return ISDQN_E_TOO_LARGE;
  static size_t configured[3] = {0, 0, 0};
  if (smem > configured[OUT]) {
    ISDQN_CUDA_CHECK(cudaFuncSetAttribute(gather_stack4_u8_kernel<OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    configured[OUT] = smem;
  }
  int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  int grid = kNumSMs * ctas_per_sm;
  if (grid > n) grid = n;
  ISDQN_PROF(stream, "gather_stack4_u8");
  gather_stack4_u8_kernel<OUT><<<grid, kGatherThreads, smem, stream>>>(
      d_frames, frame_stride, frame_bytes, d_elem_frames, d_elem_action, d_elem_reward, d_elem_terminal, d_slots, n,
      d_out_state, d_out_next, d_out_action, d_out_reward, d_out_terminal);
  ISDQN_LAUNCH_CHECK();
  return ISDQN_OK;
}

extern "C" int isdqn_gather_stacks(const uint8_t* d_frames, int64_t frame_stride, int32_t frame_elems,
                                   int32_t elem_size, int32_t stack, const int32_t* d_elem_frames,
                                   const int64_t* d_elem_action, const double* d_elem_reward,
                                   const uint8_t* d_elem_terminal, const int32_t* d_slots, int32_t n, int32_t out_dtype,
                                   void* d_out_state, void* d_out_next, int64_t* d_out_action, double* d_out_reward,
                                   uint8_t* d_out_terminal, void* stream) {
  if (!d_frames || !d_elem_frames || !d_slots || frame_elems < 1 || stack < 1 || n < 0 || frame_stride < 1)
    return ISDQN_E_INVALID;
  if ((d_out_action && !d_elem_action) || (d_out_reward && !d_elem_reward) || (d_out_terminal && !d_elem_terminal))
    return ISDQN_E_INVALID;
  if (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8) return ISDQN_E_INVALID;
  if (out_dtype != ISDQN_OUT_RAW && elem_size != 1) return ISDQN_E_INVALID;
  if (n == 0) return ISDQN_OK;
  cudaStream_t s = as_stream(stream);
  const bool fast = elem_size == 1 && stack == 4 && (frame_elems % 4) == 0 && (frame_stride % 16) == 0 &&
                    ((int64_t)((frame_elems + 15) & ~15) <= frame_stride) && d_out_state && d_out_next;
  if (fast) {
    switch (out_dtype) {
      case ISDQN_OUT_RAW:
        return launch_fast<ISDQN_OUT_RAW>(d_frames, frame_stride, frame_elems, d_elem_frames, d_elem_action,
                                          d_elem_reward, d_elem_terminal, d_slots, n, d_out_state, d_out_next,
                                          d_out_action, d_out_reward, d_out_terminal, s);
      case ISDQN_OUT_F32:
        return launch_fast<ISDQN_OUT_F32>(d_frames, frame_stride, frame_elems, d_elem_frames, d_elem_action,
                                          d_elem_reward, d_elem_terminal, d_slots, n, d_out_state, d_out_next,
                                          d_out_action, d_out_reward, d_out_terminal, s);
      case ISDQN_OUT_BF16:
        return launch_fast<ISDQN_OUT_BF16>(d_frames, frame_stride, frame_elems, d_elem_frames, d_elem_action,
                                           d_elem_reward, d_elem_terminal, d_slots, n, d_out_state, d_out_next,
                                           d_out_action, d_out_reward, d_out_terminal, s);
      default:
        return ISDQN_E_INVALID;
    }
  }
  if (out_dtype != ISDQN_OUT_RAW) return ISDQN_E_UNSUPPORTED;  // fused normalise exists for the uint8 stack-4 layout
  const int per = 2 * frame_elems * stack;
  const int64_t out_stride = (int64_t)frame_elems * stack * elem_size;  // bytes per element per output
  for (int32_t e0 = 0; e0 < n; e0 += 32768) {  // gridDim.y limit
    const int32_t ne = n - e0 < 32768 ? n - e0 : 32768;
    dim3 grid((unsigned)min(ceil_div(per, 256), 64), (unsigned)ne);
    uint8_t* os = d_out_state ? reinterpret_cast<uint8_t*>(d_out_state) + e0 * out_stride : nullptr;
    uint8_t* on = d_out_next ? reinterpret_cast<uint8_t*>(d_out_next) + e0 * out_stride : nullptr;
#define ISDQN_GENERIC(ES)                                                                                       \
  gather_generic_kernel<ES><<<grid, 256, 0, s>>>(                                                               \
      d_frames, frame_stride, frame_elems, stack, d_elem_frames, d_elem_action, d_elem_reward, d_elem_terminal, \
      d_slots + e0, ne, os, on, d_out_action ? d_out_action + e0 : nullptr,                                     \
      d_out_reward ? d_out_reward + e0 : nullptr, d_out_terminal ? d_out_terminal + e0 : nullptr)
    switch (elem_size) {
      case 1: ISDQN_GENERIC(1); break;
      case 2: ISDQN_GENERIC(2); break;
      case 4: ISDQN_GENERIC(4); break;
      default: ISDQN_GENERIC(8); break;
    }
#undef ISDQN_GENERIC
    ISDQN_LAUNCH_CHECK();
  }
  return ISDQN_OK;
}
