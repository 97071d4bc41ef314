// PCG64 (XSL-RR 128/64) exactly as numpy's Generator uses it, host + device.
// Restated from numpy/random/src/pcg64/pcg64.h; pinned against numpy in tests/test_pcg64.py (oracle) and
// tests/test_samplers_gpu.py (these kernels).
#pragma once
#include <stdint.h>

namespace isdqn {

typedef unsigned __int128 u128;

__host__ __device__ __forceinline__ u128 pcg_mult() {
  return (((u128)0x2360ED051FC65DA4ull) << 64) | (u128)0x4385DF649FCCF645ull;
}

__host__ __device__ __forceinline__ uint64_t pcg_output(u128 s) {
  const uint64_t hi = (uint64_t)(s >> 64), lo = (uint64_t)s;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(hi >> 58);  // == s >> 122
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

// state after `delta` LCG steps (O(log delta) 128-bit multiplies)
__host__ __device__ __forceinline__ u128 pcg_advance(u128 state, u128 inc, uint64_t delta) {
  u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mult(), cur_plus = inc;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  return acc_mult * state + acc_plus;
}

// j-th (0-based) 64-bit output counted from `state`: numpy steps first, then outputs the new state
__host__ __device__ __forceinline__ uint64_t pcg_out64_at(u128 state, u128 inc, uint64_t j) {
  return pcg_output(pcg_advance(state, inc, j + 1));
}

struct PcgMirror {  // layout of the 6 x uint64 device/host mirror
  u128 state, inc;
  uint32_t has_uint32, uinteger;
};

__host__ __device__ __forceinline__ PcgMirror pcg_load(const uint64_t* r) {
  PcgMirror m;
  m.state = (((u128)r[1]) << 64) | (u128)r[0];
  m.inc = (((u128)r[3]) << 64) | (u128)r[2];
  m.has_uint32 = (uint32_t)r[4];
  m.uinteger = (uint32_t)r[5];
  return m;
}

__host__ __device__ __forceinline__ void pcg_store(uint64_t* r, const PcgMirror& m) {
  r[0] = (uint64_t)m.state;
  r[1] = (uint64_t)(m.state >> 64);
  r[2] = (uint64_t)m.inc;
  r[3] = (uint64_t)(m.inc >> 64);
  r[4] = m.has_uint32;
  r[5] = m.uinteger;
}

// c-th (0-based) value of the buffered 32-bit stream (numpy's next_uint32: low half first, high half cached)
__host__ __device__ __forceinline__ uint32_t pcg_next32_at(const PcgMirror& m, uint64_t c) {
  if (m.has_uint32) {
    if (c == 0) return m.uinteger;
    c -= 1;
  }
  const uint64_t o = pcg_out64_at(m.state, m.inc, c >> 1);
  return (c & 1) ? (uint32_t)(o >> 32) : (uint32_t)o;
}

// mirror after consuming `c` values of the 32-bit stream
__host__ __device__ __forceinline__ PcgMirror pcg_after_next32(PcgMirror m, uint64_t c) {
  if (c == 0) return m;
  if (m.has_uint32) {
    m.has_uint32 = 0;
    c -= 1;
  }
  const uint64_t g = (c + 1) >> 1;  // 64-bit outputs generated
  if (g > 0) {
    m.state = pcg_advance(m.state, m.inc, g);
    m.uinteger = (uint32_t)(pcg_output(m.state) >> 32);
    m.has_uint32 = (uint32_t)(c & 1);
  }
  return m;
}

}  // namespace isdqn
