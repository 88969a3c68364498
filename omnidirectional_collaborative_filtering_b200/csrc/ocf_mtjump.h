// Jump-ahead for NumPy's MT19937 (host side): the polynomial g_J(x) = x^J mod phi(x), phi = characteristic
// polynomial of the generator's word recurrence (degree 19937, primitive over GF(2)).
//
// Why: the reference draws every batch's random split from ONE sequential stream (data_reader.py:120,130). A
// single CTA regenerates it at ~0.84 G draws/s (one dependency chain), which bounded the end-to-end rate of the
// multi-GPU runs. The stream is linear over GF(2): every bit position of the untempered word sequence x_n
// satisfies  sum_j phi_j x_{n+j} = 0,  hence  x_{n+J} = sum_{i : g_J has bit i} x_{n+i}  for all n. A worker
// that holds the 624-word array at word n gets the array at word n + J as a correlation of g_J with the next
// 19 937 + 624 words of its own sequence (k_mt_jump_apply) - no matrix, 2.5 KB per distinct jump length. With
// fixed block sizes only two jump lengths ever occur (ocf_rng in ocf_api.cu), so the stream is produced by M
// CTAs side by side, each skipping the blocks of the others, bit-identical to the sequential stream.
//
// phi is not hard-coded: it is recovered from the generator itself as the minimal polynomial of one output bit
// (Berlekamp-Massey over 2 * 19937 words of the sequence), which also checks that its degree is 19937.
#pragma once

#include <stdint.h>

#include <array>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace ocf {
namespace mtj {

constexpr int DEG = 19937;
constexpr int NW = 312;                      // 64-bit words of a polynomial of degree < 19968
using Poly = std::array<uint64_t, NW>;       // bit i = coefficient of x^i

// the untempered MT19937 word sequence continuing a 624-word array: word(n) for n = 0.. (words 0..623 = the array)
struct Seq {
  std::vector<uint32_t> w;
  explicit Seq(const uint32_t* key) : w(key, key + 624) {}
  void extend(size_t n_words) {
    while (w.size() < n_words) {
      const size_t n = w.size();
      const uint32_t y = (w[n - 624] & 0x80000000u) | (w[n - 623] & 0x7fffffffu);
      w.push_back(w[n - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u));
    }
  }
};

inline uint32_t temper(uint32_t y) {
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
inline uint32_t untemper(uint32_t y) {
  y ^= y >> 18;
  y ^= (y << 15) & 0xefc60000u;
  uint32_t t = y;                                    // y = t ^ ((t << 7) & B): recover 7 bits at a time
  for (int k = 0; k < 5; ++k) t = y ^ ((t << 7) & 0x9d2c5680u);
  y = t;
  t = y;                                             // y = t ^ (t >> 11)
  for (int k = 0; k < 3; ++k) t = y ^ (t >> 11);
  return t;
}

// phi(x): Berlekamp-Massey on bit 0 of the untempered sequence (started a few arrays in, past the 31 unused low
// bits of the very first word). Returns phi with bit DEG set, as NW words (DEG < 64 * NW).
inline const Poly& phi() {
  static Poly P = [] {
    uint32_t key[624];
    uint32_t s = 19650218u;                           // any non-zero array
    for (int i = 0; i < 624; ++i) { key[i] = s; s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u; }
    Seq q(key);
    const int N = 2 * DEG + 64;
    q.extend(1248 + (size_t)N);
    // connection polynomial C (bit i = c_i), previous B; window Wd bit i = s[n-1-i] for the discrepancy
    std::vector<uint64_t> C(NW + 1, 0), B(NW + 1, 0), T(NW + 1, 0), Wd(NW + 1, 0);
    C[0] = B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < N; ++n) {
      const uint64_t bit = q.w[1248 + n] & 1u;
      // d = s[n] + sum_{i=1..L} c_i s[n-i];  Wd holds s[n-1-i] at bit i, so c_i pairs with Wd bit i-1
      uint64_t acc = 0;
      const int words = L / 64 + 1;
      for (int k = 0; k < words; ++k) {
        // (C >> 1) word k
        const uint64_t c = (C[k] >> 1) | (C[k + 1] << 63);
        acc ^= c & Wd[k];
      }
      const uint64_t d = bit ^ (uint64_t)(__builtin_popcountll(acc) & 1);
      if (d) {
        const bool grow = 2 * L <= n;
        if (grow) T = C;
        // C ^= B << m
        const int ws = m / 64, bs = m % 64;
        for (int k = NW; k >= ws; --k) {
          uint64_t v = B[k - ws] << bs;
          if (bs && k - ws - 1 >= 0) v |= B[k - ws - 1] >> (64 - bs);
          C[k] ^= v;
        }
        if (grow) { L = n + 1 - L; B = T; m = 1; } else ++m;
      } else ++m;
      // shift the window: new bit 0 = s[n]
      for (int k = NW; k > 0; --k) Wd[k] = (Wd[k] << 1) | (Wd[k - 1] >> 63);
      Wd[0] = (Wd[0] << 1) | bit;
    }
    Poly out{};
    if (L != DEG) return out;                          // all-zero: callers treat it as a failure
    // phi_j = c_{L-j}
    for (int j = 0; j <= DEG; ++j) {
      const int i = DEG - j;
      if ((C[i / 64] >> (i % 64)) & 1u) out[j / 64] |= uint64_t(1) << (j % 64);
    }
    return out;
  }();
  return P;
}

inline bool phi_ok() { const Poly& p = phi(); return (p[DEG / 64] >> (DEG % 64)) & 1u; }

// t (2 * NW words, degree < 2 * DEG) reduced mod phi into r
inline void reduce(std::vector<uint64_t>& t, Poly& r) {
  static std::vector<std::array<uint64_t, NW + 1>> shifted = [] {
    std::vector<std::array<uint64_t, NW + 1>> s(64);
    const Poly& p = phi();
    for (int sh = 0; sh < 64; ++sh) {
      s[sh].fill(0);
      for (int k = 0; k < NW; ++k) {
        s[sh][k] |= p[k] << sh;
        if (sh) s[sh][k + 1] |= p[k] >> (64 - sh);
      }
    }
    return s;
  }();
  for (int i = 2 * DEG; i >= DEG; --i) {
    if (!((t[i / 64] >> (i % 64)) & 1u)) continue;
    const int off = i - DEG, ws = off / 64, sh = off % 64;
    const auto& ps = shifted[sh];
    for (int k = 0; k <= NW; ++k) t[ws + k] ^= ps[k];
  }
  for (int k = 0; k < NW; ++k) r[k] = t[k];
}

inline uint64_t spread32(uint32_t v) {               // bit i -> bit 2i
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000ffff0000ffffull;
  x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}

// x^J mod phi (left-to-right square and multiply-by-x), cached per J
inline const Poly& x_pow(uint64_t J) {
  static std::mutex mu;
  static std::map<uint64_t, Poly> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(J);
  if (it != cache.end()) return it->second;
  Poly r{};
  r[0] = 1;
  std::vector<uint64_t> t(2 * NW + 2);
  int top = 63;
  while (top > 0 && !((J >> top) & 1u)) --top;
  for (int b = top; b >= 0; --b) {
    std::fill(t.begin(), t.end(), 0);
    for (int k = 0; k < NW; ++k) {                     // squaring over GF(2) = spreading the bits
      t[2 * k] = spread32((uint32_t)r[k]);
      t[2 * k + 1] = spread32((uint32_t)(r[k] >> 32));
    }
    reduce(t, r);
    if ((J >> b) & 1u) {                               // times x
      uint64_t carry = 0;
      for (int k = 0; k < NW; ++k) { const uint64_t nc = r[k] >> 63; r[k] = (r[k] << 1) | carry; carry = nc; }
      if ((r[DEG / 64] >> (DEG % 64)) & 1u) { const Poly& p = phi(); for (int k = 0; k < NW; ++k) r[k] ^= p[k]; }
    }
  }
  return cache.emplace(J, r).first->second;
}

// polynomial as 624 32-bit words (the layout the device kernel reads)
inline void to_words(const Poly& p, uint32_t out[624]) {
  for (int k = 0; k < NW; ++k) { out[2 * k] = (uint32_t)p[k]; out[2 * k + 1] = (uint32_t)(p[k] >> 32); }
}

// host reference of the device jump: array at word n -> array at word n + J, by the window correlation.
// (Word 0 of the result is exact in its top bit only - the only bit of it the generator ever reads.)
inline void apply_host(const uint32_t key[624], const uint32_t poly[624], uint32_t out[624]) {
  Seq q(key);
  q.extend(DEG + 624 + 1);
  std::memset(out, 0, 624 * sizeof(uint32_t));
  for (int i = 0; i < DEG; ++i) {
    if (!((poly[i >> 5] >> (i & 31)) & 1u)) continue;
    const uint32_t* src = q.w.data() + i;
    for (int k = 0; k < 624; ++k) out[k] ^= src[k];
  }
}

}  // namespace mtj
}  // namespace ocf
