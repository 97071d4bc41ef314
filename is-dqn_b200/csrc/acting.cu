// Acting path: iSDQN.best_action (slimdqn/networks/isdqn.py:127-135) — the forward pass of DQNNet
// (slimdqn/networks/architectures/dqn.py:47-103, `cnn` + optional LayerNorm) on ONE uint8 observation stack and the
// greedy action of every head, as ONE kernel.
//
// A batch-of-one forward is 32 MFLOP over 16 MB of fp32 weights: nothing in it is throughput bound, what a chain of
// per-layer launches pays for is launch gaps and five kernel prologues (~50 us for ~5 us of work).  Here one CTA per SM
// (grid = 148, all co-resident) walks the layers and meets at a grid-wide barrier between them:
//   conv layer   : the output pixels are dealt to the CTAs (3 / 1 / 1 per CTA for 21x21, 11x11, 11x11); a CTA gathers
//                  its windows into shared memory, threads = (pixel, channel, K slice) take partial dot products
//                  against the HWIO kernel (coalesced over channels), one warp per pixel does bias + LayerNorm + ReLU.
//   hidden Dense : the [K][N] kernel (15.8 MB for 7744 -> 512: the only real traffic) is cut into (K group, 128-column
//                  group) blocks, one per CTA, every thread streaming one column over half of the block's rows;
//                  partial sums go to global memory.
//   after it     : every consumer CTA re-reduces the partials + bias + LayerNorm + ReLU into shared memory by itself
//                  (75 KB from L2) instead of paying one more barrier.
//   head layer   : CTA h computes the A Q-values of head h and their argmax (first maximum, like jnp.argmax).
// fp32 master weights and fp32 arithmetic whatever the learner's compute dtype (the reference acts in fp32).
#include <chrono>

#include "learner_kernels.cuh"
#include "plan.cuh"

using namespace isdqn;

namespace {

constexpr int kActThreads = 256;
constexpr int kActMaxX = 4096;      // floats of gathered input per CTA (windows of its pixels / its slice of a Dense input)
constexpr int kActMaxOut = 1024;    // (pixels per CTA) x channels of a conv layer
constexpr int kActMaxHidden = 2048; // widest hidden Dense layer
constexpr int kActColsPerCta = 128;

struct ActLayer {
  int type, H, W, Cin, OH, OW, ksz, stride, pad_y, pad_x, in_dim, out_dim, pix, has_ln, relu;
  int pg;          // conv: output pixels per CTA
  int slices;      // conv: K slices per output
  int kg, cg;      // hidden Dense: K groups x column groups (kg * cg <= grid)
  int64_t w_off, b_off, g_off, beta_off;
  int64_t out_off;  // scratch: conv -> activation [pix][out_dim]; hidden Dense -> partials [kg][out_dim]
};
struct ActPlan {
  int n_layers, n_heads_total, n_actions;
  ActLayer L[ISDQN_MAX_FEATURES + 1];
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// All CTAs of the grid are resident (grid <= SM count, checked by the launcher): arrive + spin on a monotonically
// increasing counter; the kernel's last CTA to leave resets it.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    // (a CTA that never arrives — the grid was not co-resident after all — must fault, not hang the GPU; the launch is
    // cooperative, so the runtime refuses a grid that cannot be resident, and this bound is the second line of defence)
    for (unsigned spins = 0; ld_acquire_u32(bar) < target; ++spins)
      if (spins > (1u << 24)) __trap();
  }
  __syncthreads();
}

__device__ __forceinline__ float block_sum_act(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();  // (protects `red` from the previous use)
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kActThreads / 32; ++w) t += red[w];
  return t;
}

// partials [kg][N] (+ bias, LayerNorm, ReLU) of hidden Dense layer P -> hs[N] in shared memory (every thread of the CTA).
// Latency is what counts: a thread owns 4 columns and every other K group, so all its loads are independent 16-byte
// loads issued back to back.
__device__ void finalize_hidden(const ActLayer& P, const float* __restrict__ params, const float* scratch, float* hs, float* tmp,
                                float* red) {
  const int N = P.out_dim, N4 = N / 4;
  const float4* part = reinterpret_cast<const float4*>(scratch + P.out_off);
  for (int base = 0; base < N4; base += kActThreads / 2) {
    const int q = base + (threadIdx.x & (kActThreads / 2 - 1)), ph = threadIdx.x / (kActThreads / 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < N4) {
#pragma unroll 20
      for (int g = ph; g < P.kg; g += 2) {
        const float4 t = __ldcg(part + (int64_t)g * N4 + q);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
    }
    if (ph == 1 && q < N4) *reinterpret_cast<float4*>(tmp + 4 * (q - base)) = v;
    __syncthreads();
    if (ph == 0 && q < N4) {
      const float4 t = *reinterpret_cast<const float4*>(tmp + 4 * (q - base));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(params + P.b_off) + q);
      *reinterpret_cast<float4*>(hs + 4 * q) = make_float4(v.x + t.x + bb.x, v.y + t.y + bb.y, v.z + t.z + bb.z, v.w + t.w + bb.w);
    }
    __syncthreads();
  }
  if (P.has_ln) {
    float s = 0.f;
    for (int n = threadIdx.x; n < N; n += kActThreads) s += hs[n];
    const float mean = block_sum_act(s, red) / (float)N;
    float s2 = 0.f;
    for (int n = threadIdx.x; n < N; n += kActThreads) {
      const float d = hs[n] - mean;
      s2 += d * d;
    }
    const float rs = rsqrtf(block_sum_act(s2, red) / (float)N + kLnEps);
    for (int n = threadIdx.x; n < N; n += kActThreads) {
      const float y = (hs[n] - mean) * rs * __ldg(params + P.g_off + n) + __ldg(params + P.beta_off + n);
      hs[n] = P.relu ? fmaxf(y, 0.f) : y;
    }
  } else if (P.relu) {
    for (int n = threadIdx.x; n < N; n += kActThreads) hs[n] = fmaxf(hs[n], 0.f);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kActThreads, 1)
act_forward_kernel(const __grid_constant__ ActPlan plan, const float* __restrict__ params, const uint8_t* __restrict__ obs,
                   float* scratch, unsigned* bar, float* __restrict__ q_out, int32_t* __restrict__ actions,
                   int* host_flag, int seq) {
  __shared__ __align__(16) float xs[kActMaxX];
  __shared__ __align__(16) float ps[kActMaxOut];
  __shared__ __align__(16) float zs[kActMaxOut];
  __shared__ __align__(16) float hs[kActMaxHidden];
  __shared__ float red[kActThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  unsigned n_bar = 0;
  if (tid == 0) trace_mark(0);
  for (int l = 0; l < plan.n_layers; ++l) {
    const ActLayer& L = plan.L[l];
    if (l > 0 && tid == 0) trace_mark(11 + 2 * (l - 1));  // past the barrier that closed layer l - 1
    if (L.type == 0) {
      // ------------------------------------------------------------------------------------------- conv
      const int p0 = blockIdx.x * L.pg;
      const int np = max(0, min(L.pg, L.pix - p0));
      const int K = L.in_dim, C = L.out_dim;
      const float* in_f = l > 0 ? scratch + plan.L[l - 1].out_off : nullptr;
      for (int e = tid; e < np * K; e += kActThreads) {
        const int i = e / K, k = e - i * K;
        const int tap = k / L.Cin, c = k - tap * L.Cin;
        const int ky = tap / L.ksz, kx = tap - ky * L.ksz;
        const int p = p0 + i;
        const int oy = p / L.OW, ox = p - oy * L.OW;
        const int iy = oy * L.stride - L.pad_y + ky, ix = ox * L.stride - L.pad_x + kx;
        float v = 0.f;
        if (iy >= 0 && iy < L.H && ix >= 0 && ix < L.W) {
          const int64_t idx = ((int64_t)iy * L.W + ix) * L.Cin + c;
          v = l == 0 ? __fdiv_rn((float)obs[idx], 255.0f) : __ldcg(in_f + idx);
        }
        xs[e] = v;
      }
      __syncthreads();
      // threads = (K slice, pixel, channel quad): 16-byte kernel loads, coalesced over the channels; a slice is ~32 taps so
      // that a thread's loads are all in flight after two unrolled batches
      const int C4 = C / 4, OS4 = np * C4, S = L.slices, per = kActThreads / S;
      const int s = tid / per;
      const int kchunk = (K + S - 1) / S;
      const int k_begin = s * kchunk, k_end = min(K, k_begin + kchunk);
      const float4* __restrict__ w4 = reinterpret_cast<const float4*>(params + L.w_off);
      for (int o = tid - s * per; o < OS4; o += per) {
        const int i = o / C4, cq = o - i * C4;
        const float* x = xs + i * K;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
        for (int k = k_begin; k < k_end; ++k) {
          const float4 wv = __ldg(w4 + (int64_t)k * C4 + cq);
          const float xv = x[k];
          acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
        }
        *reinterpret_cast<float4*>(ps + (s * np * C4 + o) * 4) = acc;
      }
      __syncthreads();
      for (int o = tid; o < np * C; o += kActThreads) {
        float v = ps[o];
        for (int q = 1; q < S; ++q) v += ps[q * np * C + o];
        zs[o] = v + __ldg(params + L.b_off + (o % C));
      }
      __syncthreads();
      float* out = scratch + L.out_off;
      for (int i = warp; i < np; i += kActThreads / 32) {  // one warp per pixel: LayerNorm over the channels + ReLU
        const float* z = zs + i * C;
        float mean = 0.f, rs = 1.f;
        if (L.has_ln) {
          float sum = 0.f;
          for (int c = lane; c < C; c += 32) sum += z[c];
          mean = warp_sum(sum) / (float)C;
          float s2 = 0.f;
          for (int c = lane; c < C; c += 32) {
            const float d = z[c] - mean;
            s2 += d * d;
          }
          rs = rsqrtf(warp_sum(s2) / (float)C + kLnEps);
        }
        for (int c = lane; c < C; c += 32) {
          float y = z[c];
          if (L.has_ln) y = (y - mean) * rs * __ldg(params + L.g_off + c) + __ldg(params + L.beta_off + c);
          out[(int64_t)(p0 + i) * C + c] = L.relu ? fmaxf(y, 0.f) : y;
        }
      }
      if (tid == 0) trace_mark(10 + 2 * l);  // this CTA's share of the layer is done
      grid_barrier(bar, ++n_bar * G);
    } else if (l + 1 < plan.n_layers) {
      // ----------------------------------------------------------------------------------- hidden Dense
      const ActLayer& P = plan.L[l - 1];
      const int Kd = L.in_dim, N = L.out_dim;
      if ((int)blockIdx.x < L.kg * L.cg) {
        const int kgi = blockIdx.x / L.cg, cgi = blockIdx.x - kgi * L.cg;
        const int chunk = (Kd + L.kg - 1) / L.kg;
        const int row0 = kgi * chunk, rows = max(0, min(chunk, Kd - row0));
        const float* x;
        if (P.type == 0) {  // flattened NHWC activation of the last conv layer
          for (int r = tid; r < rows; r += kActThreads) xs[r] = __ldcg(scratch + P.out_off + row0 + r);
          __syncthreads();
          x = xs;
        } else {
          finalize_hidden(P, params, scratch, hs, ps, red);
          x = hs + row0;
        }
        // warp = row phase (8 of them), lane = 4 columns: a warp reads 512 contiguous bytes of one kernel row
        const int cq = lane, rp = warp;
        const int col = cgi * kActColsPerCta + cq * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < N) {
          const float* __restrict__ w = params + L.w_off + (int64_t)row0 * N + col;
#pragma unroll 14
          for (int r = rp; r < rows; r += kActThreads / 32) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (int64_t)r * N));
            const float xv = x[r];
            acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
          }
        }
        __syncthreads();  // (ps may still hold the finalize scratch)
        *reinterpret_cast<float4*>(ps + (rp * 32 + cq) * 4) = acc;
        __syncthreads();
        if (tid < kActColsPerCta && cgi * kActColsPerCta + tid < N) {
          float v = ps[tid];
#pragma unroll
          for (int h = 1; h < kActThreads / 32; ++h) v += ps[h * kActColsPerCta + tid];
          scratch[L.out_off + (int64_t)kgi * N + cgi * kActColsPerCta + tid] = v;
        }
      }
      if (tid == 0) trace_mark(10 + 2 * l);
      grid_barrier(bar, ++n_bar * G);
    } else {
      // --------------------------------------------------------------------------------------- head layer
      const ActLayer& P = plan.L[l - 1];
      const int A = plan.n_actions, NH = L.out_dim, Kd = L.in_dim;
      for (int h = blockIdx.x; h < plan.n_heads_total; h += G) {
        const float* x;
        if (P.type == 0) {
          for (int r = tid; r < Kd; r += kActThreads) hs[r] = __ldcg(scratch + P.out_off + r);
          __syncthreads();
          x = hs;
        } else {
          finalize_hidden(P, params, scratch, hs, ps, red);
          x = hs;
        }
        float acc[kMaxActions];
#pragma unroll
        for (int a = 0; a < kMaxActions; ++a) acc[a] = 0.f;
        for (int k = tid; k < Kd; k += kActThreads) {
          const float xv = x[k];
          const float* __restrict__ w = params + L.w_off + (int64_t)k * NH + h * A;
#pragma unroll
          for (int a = 0; a < kMaxActions; ++a)
            if (a < A) acc[a] = fmaf(xv, __ldg(w + a), acc[a]);
        }
        __syncthreads();  // (zs / ps free)
#pragma unroll
        for (int a = 0; a < kMaxActions; ++a) {
          if (a < A) {
            const float v = warp_sum(acc[a]);
            if (lane == 0) ps[warp * kMaxActions + a] = v;
          }
        }
        __syncthreads();
        if (tid < A) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kActThreads / 32; ++w) v += ps[w * kMaxActions + tid];
          v += __ldg(params + L.b_off + h * A + tid);
          zs[tid] = v;
          if (q_out) q_out[h * A + tid] = v;
        }
        __syncthreads();
        if (tid == 0) {
          int best = 0;
          for (int a = 1; a < A; ++a)
            if (zs[a] > zs[best]) best = a;
          actions[h] = best;
        }
        __syncthreads();
      }
    }
  }
  if (tid == 0) trace_mark(10 + 2 * (plan.n_layers - 1));
  // last CTA out re-arms the barrier for the next launch (everybody has passed the final barrier by then)
  // (host_flag: `actions` is mapped host memory and the host spins on *host_flag == seq instead of synchronising an event:
  // every CTA makes its action visible system-wide before it checks out, the last one out raises the flag)
  if (tid == 0) {
    if (host_flag) __threadfence_system();
    else __threadfence();
    if (atomicAdd(bar + 1, 1u) == (unsigned)G - 1u) {
      bar[0] = 0u;
      bar[1] = 0u;
      __threadfence();
      if (host_flag) {
        *reinterpret_cast<volatile int*>(host_flag) = seq;
        __threadfence_system();
      }
    }
  }
}

int floor_pow2(int x) {
  int p = 1;
  while (2 * p <= x) p *= 2;
  return p;
}

// ISDQN_OK and a filled plan, or ISDQN_E_UNSUPPORTED when the network is outside what the single-kernel path covers
// (the caller then keeps the per-layer forward).
int build_act_plan(const isdqn_net* net, const Plan& p, int grid, ActPlan* a, int64_t* scratch_floats) {
  if (net->arch != ISDQN_ARCH_CNN || p.n_layers < 2) return ISDQN_E_UNSUPPORTED;
  if (net->n_actions > kMaxActions || 1 + net->n_heads > kMaxHeads + 1) return ISDQN_E_UNSUPPORTED;
  a->n_layers = p.n_layers;
  a->n_heads_total = 1 + net->n_heads;
  a->n_actions = net->n_actions;
  int64_t off = 4;  // [0..1]: barrier words (as unsigned), 16-byte header
  for (int l = 0; l < p.n_layers; ++l) {
    const Layer& L = p.L[l];
    ActLayer& D = a->L[l];
    D.type = L.type; D.H = L.H; D.W = L.W; D.Cin = L.Cin; D.OH = L.OH; D.OW = L.OW; D.ksz = L.ksz; D.stride = L.stride;
    D.pad_y = L.pad_y; D.pad_x = L.pad_x; D.in_dim = L.in_dim; D.out_dim = L.out_dim; D.pix = L.pix; D.has_ln = L.has_ln;
    D.relu = L.relu; D.w_off = L.w_off; D.b_off = L.b_off; D.g_off = L.g_off; D.beta_off = L.beta_off;
    D.pg = D.slices = D.kg = D.cg = 0;
    D.out_off = -1;
    const bool last = l == p.n_layers - 1;
    if (L.type == 0) {
      if (l > 0 && p.L[l - 1].type != 0) return ISDQN_E_UNSUPPORTED;
      D.pg = ceil_div(L.pix, grid);
      if ((int64_t)D.pg * L.in_dim > kActMaxX || D.pg * L.out_dim > kActMaxOut) return ISDQN_E_UNSUPPORTED;
      if (L.out_dim % 4) return ISDQN_E_UNSUPPORTED;
      int s = kActThreads / (D.pg * (L.out_dim / 4));
      s = s < 1 ? 1 : floor_pow2(s);
      if (s > 32) s = 32;
      while (s > 1 && (int64_t)s * D.pg * L.out_dim > kActMaxOut) s /= 2;
      D.slices = s;
      D.out_off = off;
      off = align4(off + (int64_t)L.pix * L.out_dim);
    } else if (!last) {
      if (L.out_dim > kActMaxHidden || L.out_dim % 4) return ISDQN_E_UNSUPPORTED;
      D.cg = ceil_div(L.out_dim, kActColsPerCta);
      if (D.cg > grid) return ISDQN_E_UNSUPPORTED;
      D.kg = grid / D.cg;
      if (D.kg > L.in_dim) D.kg = L.in_dim;
      if (ceil_div(L.in_dim, D.kg) > kActMaxX) return ISDQN_E_UNSUPPORTED;
      D.out_off = off;
      off = align4(off + (int64_t)D.kg * L.out_dim);
    } else {
      if (L.in_dim > kActMaxHidden) return ISDQN_E_UNSUPPORTED;
    }
  }
  *scratch_floats = off;
  return ISDQN_OK;
}

}  // namespace

int isdqn_trace_set_acting(unsigned long long* buf) { return isdqn::trace_set_local(buf) == cudaSuccess ? ISDQN_OK : ISDQN_E_CUDA; }

// one CTA per SM of THIS device must be co-resident for the grid barrier (checked once per device)
static int act_grid_ctas() {
  static PerDevice<int> grid;
  int& g = grid.get();
  if (g == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, act_forward_kernel, kActThreads, 0) != cudaSuccess) per_sm = 0;
    int dev = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    g = (per_sm >= 1 && coop) ? num_sms() : -1;
  }
  return g;
}

extern "C" int64_t isdqn_act_workspace_bytes(const isdqn_net* net) {
  Plan p;
  if (build_plan(net, &p)) return -1;
  const int G = act_grid_ctas();
  if (G < 1) return 0;  // no co-resident grid on this device: the caller keeps the layer chain
  ActPlan a;
  int64_t floats = 0;
  const int rc = build_act_plan(net, p, G, &a, &floats);
  if (rc == ISDQN_E_UNSUPPORTED) return 0;
  if (rc) return -1;
  return floats * (int64_t)sizeof(float);
}

extern "C" int isdqn_act(const isdqn_net* net, const float* d_params, const uint8_t* d_obs, float* d_q, int32_t* d_actions,
                         void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (!net || !d_params || !d_obs || !d_actions || !d_workspace) return ISDQN_E_INVALID;
  Plan p;
  int rc = build_plan(net, &p);
  if (rc) return rc;
  const int G = act_grid_ctas();
  if (G < 1) return ISDQN_E_UNSUPPORTED;
  ActPlan a;
  int64_t floats = 0;
  rc = build_act_plan(net, p, G, &a, &floats);
  if (rc) return rc;
  if (floats * (int64_t)sizeof(float) > workspace_bytes) return ISDQN_E_INVALID;
  cudaStream_t s = as_stream(stream);
  ISDQN_PROF(s, "act_forward");
  float* scratch = reinterpret_cast<float*>(d_workspace);
  unsigned* bar = reinterpret_cast<unsigned*>(d_workspace);
  // cooperative launch: the runtime guarantees (or refuses) the co-residency the grid barrier relies on
  int* no_flag = nullptr;
  int seq = 0;
  void* args[] = {(void*)&a, (void*)&d_params, (void*)&d_obs, (void*)&scratch, (void*)&bar, (void*)&d_q, (void*)&d_actions,
                  (void*)&no_flag, (void*)&seq};
  ISDQN_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(act_forward_kernel), dim3(G), dim3(kActThreads), args, 0, s));
  return ISDQN_OK;
}

// The env-step call without copy or event traffic: the kernel reads the observation straight out of pinned (mapped) host
// memory and writes the 1 + K greedy actions and a completion flag straight back; the host spins on the flag.
extern "C" int isdqn_act_mapped(const isdqn_net* net, const float* d_params, const uint8_t* h_obs_pinned, float* d_q,
                                int32_t* h_actions_pinned, int32_t* h_flag_pinned, int32_t seq, void* d_workspace,
                                int64_t workspace_bytes, void* stream, int64_t timeout_us) {
  if (!net || !d_params || !h_obs_pinned || !h_actions_pinned || !h_flag_pinned || !d_workspace) return ISDQN_E_INVALID;
  Plan p;
  int rc = build_plan(net, &p);
  if (rc) return rc;
  const int G = act_grid_ctas();
  if (G < 1) return ISDQN_E_UNSUPPORTED;
  ActPlan a;
  int64_t floats = 0;
  rc = build_act_plan(net, p, G, &a, &floats);
  if (rc) return rc;
  if (floats * (int64_t)sizeof(float) > workspace_bytes) return ISDQN_E_INVALID;
  // device-side aliases of the pinned blocks (identical to the host pointers under unified addressing)
  uint8_t* d_obs = nullptr;
  int32_t* d_act = nullptr;
  int* d_flag = nullptr;
  ISDQN_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_obs), const_cast<uint8_t*>(h_obs_pinned), 0));
  ISDQN_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_act), h_actions_pinned, 0));
  ISDQN_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_flag), h_flag_pinned, 0));
  cudaStream_t s = as_stream(stream);
  ISDQN_PROF(s, "act_forward");
  float* scratch = reinterpret_cast<float*>(d_workspace);
  unsigned* bar = reinterpret_cast<unsigned*>(d_workspace);
  const uint8_t* obs_c = d_obs;
  int seq_i = seq;
  void* args[] = {(void*)&a, (void*)&d_params, (void*)&obs_c, (void*)&scratch, (void*)&bar, (void*)&d_q, (void*)&d_act,
                  (void*)&d_flag, (void*)&seq_i};
  ISDQN_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(act_forward_kernel), dim3(G), dim3(kActThreads), args, 0, s));
  if (timeout_us <= 0) return ISDQN_OK;  // the caller waits itself (isdqn_act_wait)
  return isdqn_act_wait(h_flag_pinned, seq, timeout_us);
}

extern "C" int isdqn_act_wait(const int32_t* h_flag_pinned, int32_t seq, int64_t timeout_us) {
  if (!h_flag_pinned) return ISDQN_E_INVALID;
  const volatile int32_t* f = h_flag_pinned;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0; *f != seq; ++spins) {
    if ((spins & 1023u) == 1023u) {
      if (std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > timeout_us) {
        // a faulted kernel never raises the flag: report the fault if the runtime knows of one, a timeout otherwise
        const cudaError_t e = cudaPeekAtLastError();
        set_last_cuda_error(e != cudaSuccess ? e : cudaErrorTimeout, "isdqn_act_wait");
        return ISDQN_E_CUDA;
      }
    }
  }
  return ISDQN_OK;
}

// One call per env step: pinned observation -> device, the kernel, greedy actions -> pinned host, wait for them.
extern "C" int isdqn_act_host(const isdqn_net* net, const float* d_params, const uint8_t* h_obs_pinned, uint8_t* d_obs,
                              int64_t obs_bytes, float* d_q, int32_t* d_actions, int32_t* h_actions_pinned, void* d_workspace,
                              int64_t workspace_bytes, void* stream, void* event) {
  if (!h_obs_pinned || !d_obs || !h_actions_pinned || !event || obs_bytes < 1) return ISDQN_E_INVALID;
  cudaStream_t s = as_stream(stream);
  ISDQN_CUDA_CHECK(cudaMemcpyAsync(d_obs, h_obs_pinned, (size_t)obs_bytes, cudaMemcpyHostToDevice, s));
  const int rc = isdqn_act(net, d_params, d_obs, d_q, d_actions, d_workspace, workspace_bytes, stream);
  if (rc) return rc;
  ISDQN_CUDA_CHECK(cudaMemcpyAsync(h_actions_pinned, d_actions, sizeof(int32_t) * (size_t)(1 + net->n_heads), cudaMemcpyDeviceToHost, s));
  ISDQN_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), s));
  ISDQN_CUDA_CHECK(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
  return ISDQN_OK;
}
