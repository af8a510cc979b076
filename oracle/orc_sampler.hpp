// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Samplers.
//  * ZeroTwoSequence + PCG32: restatement of sampler/zerotwosequence.rs, sampler/lowdiscrepancy.rs, rng.rs.
//  * CounterSampler: NOT in the reference.  It is the oracle-side twin of the device's counter-based
//    (0,2) sampler (rustracer_b200/csrc/device/sampler.cuh) so that oracle and GPU can be compared
//    sample-for-sample; image parity against ZeroTwoSequence is statistical (SURVEY §7 "Sampler
//    sequentiality").  Integer-only, so both sides are bit-identical.
#pragma once
#include "orc_math.hpp"
#include <vector>
#include <memory>

namespace orc {

struct RNG {                                                    // rng.rs:5-52
  uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
  uint32_t uniform_u32() {
    uint64_t old = state;
    state = old * 0x5851f42d4c957f2dULL + inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
  }
  uint32_t uniform_u32_bounded(uint32_t b) {                    // :32-40 (note `& b`, not `% b`, in the threshold)
    uint32_t threshold = (~b + 1u) & b;
    while (true) { uint32_t r = uniform_u32(); if (r >= threshold) return r % b; }
  }
  float uniform_f32() { return fmin_((float)uniform_u32() * 2.3283064365386963e-10f, ONE_MINUS_EPSILON); }
  void set_sequence(uint64_t seed) {                            // :46-52
    state = 0; inc = (seed << 1) | 1;
    (void)uniform_u32();
    state += 0x853c49e6748fea9bULL;
    (void)uniform_u32();
  }
};

inline uint32_t reverse_bits_32(uint32_t n) {                   // lowdiscrepancy.rs:63-71
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
  n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
  n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
  n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
  return n;
}
inline uint64_t reverse_bits_64(uint64_t n) {                   // :73-77
  uint64_t n0 = reverse_bits_32((uint32_t)n), n1 = reverse_bits_32((uint32_t)(n >> 32));
  return (n0 << 32) | n1;
}
inline float radical_inverse_specialized(uint32_t base, uint64_t a) {   // :79-93
  float inv_base = 1.0f / (float)base;
  uint64_t reversed = 0; float inv_base_n = 1.0f;
  while (a != 0) {
    uint64_t next = a / base, digit = a - next * base;
    reversed = reversed * base + digit;
    inv_base_n *= inv_base;
    a = next;
  }
  return fmin_((float)reversed * inv_base_n, ONE_MINUS_EPSILON);
}
inline float radical_inverse(uint32_t base_index, uint64_t a) {          // :50-61
  switch (base_index) {
    case 0: return (float)reverse_bits_64(a) * 5.4210108624275222e-20f;
    case 1: return radical_inverse_specialized(3, a);
    case 2: return radical_inverse_specialized(5, a);
    case 3: return radical_inverse_specialized(7, a);
    case 4: return radical_inverse_specialized(11, a);
    default: return radical_inverse_specialized(13, a);
  }
}
// Generator matrices (:125-174): van der Corput = identity columns MSB-first; Sobol' dim 2 column i = c[i-1]^(c[i-1]>>1).
inline uint32_t c_vdc(int i) { return 0x80000000u >> i; }
inline uint32_t c_sobol1(int i) { uint32_t c = 0x80000000u; for (int k = 0; k < i; k++) c ^= c >> 1; return c; }
inline int trailing_zeros(uint32_t v) { return v == 0 ? 32 : __builtin_ctz(v); }

template <class T> inline void shuffle(T* samp, uint32_t count, uint32_t n_dim, RNG& rng) {    // :113-123
  for (uint32_t i = 0; i < count; i++) {
    uint32_t other = i + rng.uniform_u32_bounded(count - i);
    for (uint32_t j = 0; j < n_dim; j++) std::swap(samp[n_dim * i + j], samp[n_dim * other + j]);
  }
}
inline void van_der_corput(uint32_t n_per, uint32_t n_pix, float* samples, RNG& rng) {          // :4-24, :95-101
  uint32_t scramble = rng.uniform_u32();
  uint32_t total = n_per * n_pix, v = scramble;
  for (uint32_t i = 0; i < total; i++) {
    samples[i] = fmin_((float)v * 2.3283064365386963e-10f, ONE_MINUS_EPSILON);
    v ^= c_vdc(trailing_zeros(i + 1));
  }
  for (uint32_t i = 0; i < n_pix; i++) shuffle(samples + (size_t)i * n_per, n_per, 1, rng);
  shuffle(samples, n_pix, n_per, rng);
}
inline void sobol_2d(uint32_t n_per, uint32_t n_pix, P2* samples, RNG& rng) {                   // :26-48, :103-111
  uint32_t s0 = rng.uniform_u32(), s1 = rng.uniform_u32();
  uint32_t total = n_per * n_pix, v0 = s0, v1 = s1;
  static uint32_t c1[32]; static bool init = false;
  if (!init) { for (int i = 0; i < 32; i++) c1[i] = c_sobol1(i); init = true; }
  for (uint32_t i = 0; i < total; i++) {
    samples[i].x = fmin_((float)v0 * 2.3283064365386963e-10f, ONE_MINUS_EPSILON);
    samples[i].y = fmin_((float)v1 * 2.3283064365386963e-10f, ONE_MINUS_EPSILON);
    int tz = trailing_zeros(i + 1);
    v0 ^= c_vdc(tz); v1 ^= c1[tz];
  }
  for (uint32_t i = 0; i < n_pix; i++) shuffle(samples + (size_t)i * n_per, n_per, 1, rng);
  shuffle(samples, n_pix, n_per, rng);
}

struct CameraSample { P2 p_film, p_lens; float time; };

struct Sampler {                                                 // sampler/mod.rs:7-22
  size_t spp = 1;
  virtual ~Sampler() {}
  virtual void start_pixel(int x, int y) = 0;
  virtual float get_1d() = 0;
  virtual P2 get_2d() = 0;
  virtual void request_2d_array(size_t n) = 0;
  virtual const P2* get_2d_array(size_t n) = 0;                  // nullptr == None
  virtual bool start_next_sample() = 0;
  virtual void reseed(uint64_t seed) = 0;
  virtual std::unique_ptr<Sampler> clone() const = 0;
  virtual size_t current_sample_number() const = 0;
  size_t round_count(size_t c) const { return next_power_of_two(c); }   // zerotwosequence.rs:194-196
  CameraSample get_camera_sample(int px, int py) {               // zerotwosequence.rs:182-192
    CameraSample cs;
    P2 u = get_2d();
    cs.p_film = P2((float)px + u.x, (float)py + u.y);
    cs.time = get_1d();
    cs.p_lens = get_2d();
    return cs;
  }
  // Oracle-only hooks (no-ops for the reference sampler): recursion-tree node keying for CounterSampler.
  struct Saved { uint32_t a = 0, b = 0, c = 0; };
  virtual Saved enter_node(uint32_t) { return Saved(); }
  virtual void leave_node(Saved) {}
};

struct ZeroTwoSequence : Sampler {                               // zerotwosequence.rs:11-213
  size_t cur_idx = 0;
  std::vector<size_t> a1_sizes, a2_sizes;
  std::vector<std::vector<float>> a1;
  std::vector<std::vector<P2>> a2;
  size_t a1_off = 0, a2_off = 0;
  std::vector<std::vector<float>> s1;
  std::vector<std::vector<P2>> s2;
  size_t d1 = 0, d2 = 0;
  RNG rng;
  ZeroTwoSequence(size_t spp_, size_t dims) {
    spp = next_power_of_two(spp_);
    s1.assign(dims, std::vector<float>(spp, 0.0f));
    s2.assign(dims, std::vector<P2>(spp));
  }
  void start_pixel(int, int) override {                          // :67-108 (does not reset d1/d2)
    for (auto& v : s1) van_der_corput(1, (uint32_t)spp, v.data(), rng);
    for (auto& v : s2) sobol_2d(1, (uint32_t)spp, v.data(), rng);
    for (size_t i = 0; i < a1_sizes.size(); i++) van_der_corput((uint32_t)a1_sizes[i], (uint32_t)spp, a1[i].data(), rng);
    for (size_t i = 0; i < a2_sizes.size(); i++) sobol_2d((uint32_t)a2_sizes[i], (uint32_t)spp, a2[i].data(), rng);
    cur_idx = 0; a1_off = 0; a2_off = 0;
  }
  bool start_next_sample() override { a1_off = 0; a2_off = 0; d1 = 0; d2 = 0; cur_idx += 1; return cur_idx < spp; }   // :110-117
  void request_2d_array(size_t n) override { a2_sizes.push_back(n); a2.emplace_back(n * spp); }                       // :126-132
  const P2* get_2d_array(size_t n) override {                    // :146-156
    if (a2_off == a2.size()) return nullptr;
    const P2* r = a2[a2_off].data() + cur_idx * n;
    a2_off += 1;
    return r;
  }
  float get_1d() override {                                      // :158-166
    if (d1 < s1.size()) { float r = s1[d1][cur_idx]; d1 += 1; return r; }
    return rng.uniform_f32();
  }
  P2 get_2d() override {                                         // :168-180 (fallback returns (second, first) — Q26)
    if (d2 < s2.size()) { P2 r = s2[d2][cur_idx]; d2 += 1; return r; }
    float x = rng.uniform_f32();
    float y = rng.uniform_f32();
    return P2(y, x);
  }
  void reseed(uint64_t seed) override { rng.set_sequence(seed); }
  std::unique_ptr<Sampler> clone() const override { return std::unique_ptr<Sampler>(new ZeroTwoSequence(*this)); }
  size_t current_sample_number() const override { return cur_idx; }
};

// ---------------------------------------------------------------------------------------
// Counter-based twin of the device sampler.  All integer; see sampler.cuh for the device copy.
inline uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
inline uint32_t cmj_permute(uint32_t i, uint32_t l, uint32_t p) {     // Kensler 2013, "Correlated multi-jittered sampling"
  uint32_t w = l - 1;
  w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
  do {
    i ^= p; i *= 0xe170893du; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3fu; i ^= p >> 23;
    i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69u; i ^= (i & w) >> 11; i *= 0x74dcb303u; i ^= (i & w) >> 2;
    i *= 0x9e501cc3u; i ^= (i & w) >> 2; i *= 0xc860a3dfu; i &= w; i ^= i >> 5;
  } while (i >= l);
  return (i + p) % l;
}
inline uint32_t sobol1_eval(uint32_t idx) { uint32_t v = 0, c = 0x80000000u; while (idx) { if (idx & 1) v ^= c; c ^= c >> 1; idx >>= 1; } return v; }
inline float u32_to_unit(uint32_t v) { return fmin_((float)v * 2.3283064365386963e-10f, ONE_MINUS_EPSILON); }

struct CounterSampler : Sampler {
  uint32_t dims; uint64_t seed;
  uint32_t pix_hash = 0, s = 0;
  uint32_t d1 = 0, d2 = 0, da = 0;     // 1-D, 2-D and 2-D-array draw counters
  size_t n_arrays = 0;
  std::vector<P2> scratch;
  CounterSampler(size_t spp_, size_t dims_, uint64_t seed_) : dims((uint32_t)dims_), seed(seed_) { spp = next_power_of_two(spp_); }
  static uint32_t pixel_hash(int x, int y, uint64_t seed) {
    return mix32((uint32_t)x ^ mix32((uint32_t)y + 0x632be5abu) ^ mix32((uint32_t)seed + 0x9e3779b9u));
  }
  static uint32_t stream_key(uint32_t ph, uint32_t counter, uint32_t tag) { return mix32(ph ^ mix32(counter * 0x9e3779b1u + tag * 0x85ebca6bu + 0x27d4eb2fu)); }
  static float draw_1d(uint32_t ph, uint32_t s, uint32_t spp, uint32_t dims, uint32_t counter) {
    uint32_t k = stream_key(ph, counter, 1);
    if (counter < dims) { uint32_t idx = cmj_permute(s, spp, k); return u32_to_unit(reverse_bits_32(idx) ^ mix32(k + 1)); }
    return u32_to_unit(mix32(k ^ mix32(s * 0x9e3779b1u + 0x68bc21ebu)));
  }
  static P2 draw_2d(uint32_t ph, uint32_t s, uint32_t spp, uint32_t dims, uint32_t counter) {
    uint32_t k = stream_key(ph, counter, 2);
    if (counter < dims) {
      uint32_t idx = cmj_permute(s, spp, k);
      return P2(u32_to_unit(reverse_bits_32(idx) ^ mix32(k + 1)), u32_to_unit(sobol1_eval(idx) ^ mix32(k + 2)));
    }
    uint32_t a = mix32(k ^ mix32(s * 0x9e3779b1u + 0x68bc21ebu));
    uint32_t b = mix32(a + 0x3c6ef372u + k);
    return P2(u32_to_unit(a), u32_to_unit(b));
  }
  // n values of 2-D array `counter` for sample s: a (0,2) net over spp*n points when stratified.
  static P2 draw_2d_array(uint32_t ph, uint32_t s, uint32_t spp, uint32_t n, uint32_t j, uint32_t counter, bool stratified) {
    uint32_t k = stream_key(ph, counter, 3);
    if (stratified) {
      uint32_t idx = cmj_permute(s * n + j, spp * n, k);
      return P2(u32_to_unit(reverse_bits_32(idx) ^ mix32(k + 1)), u32_to_unit(sobol1_eval(idx) ^ mix32(k + 2)));
    }
    uint32_t a = mix32(k ^ mix32((s * n + j) * 0x9e3779b1u + 0x68bc21ebu));
    uint32_t b = mix32(a + 0x3c6ef372u + k);
    return P2(u32_to_unit(a), u32_to_unit(b));
  }
  void start_pixel(int x, int y) override { pix_hash = pixel_hash(x, y, seed); s = 0; d1 = d2 = da = 0; }
  bool start_next_sample() override { s += 1; d1 = d2 = da = 0; return s < spp; }
  float get_1d() override { return draw_1d(pix_hash, s, (uint32_t)spp, dims, d1++); }
  P2 get_2d() override { return draw_2d(pix_hash, s, (uint32_t)spp, dims, d2++); }
  void request_2d_array(size_t) override { n_arrays++; }
  const P2* get_2d_array(size_t n) override {
    scratch.resize(n);
    bool strat = da < n_arrays && da < 64;
    for (size_t j = 0; j < n; j++) scratch[j] = draw_2d_array(pix_hash, s, (uint32_t)spp, (uint32_t)n, (uint32_t)j, da, strat);
    da++;
    return scratch.data();
  }
  void reseed(uint64_t) override {}
  std::unique_ptr<Sampler> clone() const override { return std::unique_ptr<Sampler>(new CounterSampler(*this)); }
  size_t current_sample_number() const override { return s; }
  // Recursion-tree keying (Whitted / DirectLighting): node 1 continues the camera counters, node k>1
  // restarts every counter at 64*k so its draws do not depend on traversal order.
  Saved enter_node(uint32_t node) override {
    Saved saved; saved.a = d1; saved.b = d2; saved.c = da;
    if (node > 1) { d1 = d2 = da = 64u * node; }
    return saved;
  }
  void leave_node(Saved saved) override { d1 = saved.a; d2 = saved.b; da = saved.c; }
};

}  // namespace orc
