// Counter-based (0,2)-sequence sampler for the wavefront renderer (product code).
//
// rustracer's ZeroTwoSequence (sampler/zerotwosequence.rs) draws from ONE PCG32 stream per 16x16 tile in
// pixel -> sample order with data-dependent draw counts (renderer.rs:83-84), which a wavefront cannot replay
// (SURVEY 7 "Sampler sequentiality").  This sampler keeps its structure — per pixel, the first `dimensions`
// 1-D and 2-D draws are van der Corput / Sobol' (0,2) points over the pixel's spp samples, randomly scrambled
// and visited in a random permutation; later draws are plain random numbers — but keys every draw by
// (pixel, sample index, draw counter) so any sample can be generated independently.  All integer arithmetic;
// the oracle carries a bit-identical twin (oracle/orc_sampler.hpp CounterSampler) for sample-exact parity.
#pragma once
#include "dmath.cuh"

namespace rt {

RT_DEV uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// Kensler, "Correlated multi-jittered sampling" (2013): pseudo-random permutation of [0, l)
RT_DEV uint32_t cmj_permute(uint32_t i, uint32_t l, uint32_t p) {
  uint32_t w = l - 1;
  w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
  do {
    i ^= p; i *= 0xe170893du; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3fu; i ^= p >> 23;
    i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69u; i ^= (i & w) >> 11; i *= 0x74dcb303u; i ^= (i & w) >> 2;
    i *= 0x9e501cc3u; i ^= (i & w) >> 2; i *= 0xc860a3dfu; i &= w; i ^= i >> 5;
  } while (i >= l);
  return (l & w) == 0 ? ((i + p) & w) : (i + p) % l;     // l a power of two (02sequence rounds spp up): the modulo is a mask
}
// Sobol' dimension 2 generator (lowdiscrepancy.rs:141-174): column i = c[i-1] ^ (c[i-1] >> 1), c[0] = 1 << 31
// In bit-reversed order the columns are (1 + x)^i over GF(2), so the product is the substitution x -> x + 1 in the polynomial whose
// coefficients are the bits of idx: five butterfly steps ((x + 1)^(2^k) = x^(2^k) + 1) instead of one iteration per bit of idx.
RT_DEV uint32_t sobol1_eval(uint32_t idx) {
  idx ^= (idx >> 1) & 0x55555555u; idx ^= (idx >> 2) & 0x33333333u; idx ^= (idx >> 4) & 0x0f0f0f0fu; idx ^= (idx >> 8) & 0x00ff00ffu; idx ^= idx >> 16;
  return __brev(idx);
}
RT_DEV float u32_to_unit(uint32_t v) { return fminf((float)v * kU32ToUnit, kOneMinusEpsilon); }     // lowdiscrepancy.rs:16, rng.rs:42-44

RT_DEV uint32_t pixel_hash(int x, int y, uint64_t seed) {
  return mix32((uint32_t)x ^ mix32((uint32_t)y + 0x632be5abu) ^ mix32((uint32_t)seed + 0x9e3779b9u));
}
RT_DEV uint32_t stream_key(uint32_t ph, uint32_t counter, uint32_t tag) { return mix32(ph ^ mix32(counter * 0x9e3779b1u + tag * 0x85ebca6bu + 0x27d4eb2fu)); }

struct SamplerCfg { uint32_t spp, dims, n_arrays; };

RT_DEV float draw_1d(uint32_t ph, uint32_t s, const SamplerCfg& c, uint32_t counter) {
  uint32_t k = stream_key(ph, counter, 1);
  if (counter < c.dims) { uint32_t idx = cmj_permute(s, c.spp, k); return u32_to_unit(__brev(idx) ^ mix32(k + 1)); }
  return u32_to_unit(mix32(k ^ mix32(s * 0x9e3779b1u + 0x68bc21ebu)));
}
RT_DEV P2 draw_2d(uint32_t ph, uint32_t s, const SamplerCfg& c, uint32_t counter) {
  uint32_t k = stream_key(ph, counter, 2);
  if (counter < c.dims) {
    uint32_t idx = cmj_permute(s, c.spp, k);
    return mk2(u32_to_unit(__brev(idx) ^ mix32(k + 1)), u32_to_unit(sobol1_eval(idx) ^ mix32(k + 2)));
  }
  uint32_t a = mix32(k ^ mix32(s * 0x9e3779b1u + 0x68bc21ebu));
  uint32_t b = mix32(a + 0x3c6ef372u + k);
  return mk2(u32_to_unit(a), u32_to_unit(b));
}
// Element j of the n-element 2-D array request number `counter` of sample s (DirectLighting "all").
RT_DEV P2 draw_2d_array(uint32_t ph, uint32_t s, const SamplerCfg& c, uint32_t n, uint32_t j, uint32_t counter) {
  uint32_t k = stream_key(ph, counter, 3);
  const bool stratified = counter < c.n_arrays && counter < 64;
  if (stratified) {
    uint32_t idx = cmj_permute(s * n + j, c.spp * n, k);
    return mk2(u32_to_unit(__brev(idx) ^ mix32(k + 1)), u32_to_unit(sobol1_eval(idx) ^ mix32(k + 2)));
  }
  uint32_t a = mix32(k ^ mix32((s * n + j) * 0x9e3779b1u + 0x68bc21ebu));
  uint32_t b = mix32(a + 0x3c6ef372u + k);
  return mk2(u32_to_unit(a), u32_to_unit(b));
}

// Per-path sampler cursor.
struct SamplerState {
  uint32_t ph, s, d1, d2, da;
  RT_DEV float get_1d(const SamplerCfg& c) { return draw_1d(ph, s, c, d1++); }
  RT_DEV P2 get_2d(const SamplerCfg& c) { return draw_2d(ph, s, c, d2++); }
};

// sampler/lowdiscrepancy.rs:50-93 radical inverse (for the spatial light distribution's Halton points)
RT_DEV float radical_inverse_specialized(uint32_t base, uint64_t a) {
  float inv_base = 1.0f / (float)base;
  uint64_t reversed = 0; float inv_base_n = 1.0f;
  while (a != 0) {
    uint64_t next = a / base, digit = a - next * base;
    reversed = reversed * base + digit;
    inv_base_n *= inv_base;
    a = next;
  }
  return fminf((float)reversed * inv_base_n, kOneMinusEpsilon);
}
RT_DEV float radical_inverse(uint32_t base_index, uint64_t a) {
  switch (base_index) {
    case 0: return (float)__brevll(a) * 5.4210108624275222e-20f;
    case 1: return radical_inverse_specialized(3, a);
    case 2: return radical_inverse_specialized(5, a);
    case 3: return radical_inverse_specialized(7, a);
    case 4: return radical_inverse_specialized(11, a);
    default: return radical_inverse_specialized(13, a);
  }
}

}  // namespace rt
