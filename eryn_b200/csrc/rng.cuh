// Counter-based random streams of the production ("philox") mode.
//
// The reference draws from two NumPy MT19937 streams on the host (ensemble.py:651-652,
// red_blue.py:124, tempering.py:526-535).  In replay mode those host draws are uploaded and
// consumed verbatim; in philox mode every draw is a pure function of
// (seed, iteration, purpose tag, index), so no RNG state lives in HBM and a captured CUDA graph
// can be replayed (the iteration counter is read from device memory).
// oracle/philox_np.py restates every function of this header in NumPy, bit for bit.
#pragma once
#include <cstdint>
#include <curand_philox4x32_x.h>  // curand_Philox4x32_10(uint4 ctr, uint2 key)

namespace eb {

enum : uint32_t {
  TAG_SPLIT_KEY = 1, TAG_STRETCH = 2, TAG_GAUSS = 3, TAG_ACCEPT = 4, TAG_SWAP_KEY = 5, TAG_SWAP_U = 6,
  TAG_RJ = 7
};

struct RngKey {
  uint32_t seed_lo, seed_hi;
  uint32_t it_lo, it_hi24;  // iteration counter (low 32 bits, next 24 bits)
};

__device__ __forceinline__ RngKey make_rng_key(uint32_t seed_lo, uint32_t seed_hi, unsigned long long it) {
  RngKey k;
  k.seed_lo = seed_lo; k.seed_hi = seed_hi;
  k.it_lo = (uint32_t)(it & 0xFFFFFFFFull);
  k.it_hi24 = (uint32_t)((it >> 32) & 0xFFFFFFull);
  return k;
}

__device__ __forceinline__ uint4 stream(const RngKey& k, uint32_t tag, uint32_t c0, uint32_t c1) {
  uint4 ctr = make_uint4(c0, c1, k.it_lo, (tag << 24) | k.it_hi24);
  return curand_Philox4x32_10(ctr, make_uint2(k.seed_lo, k.seed_hi));
}

// two 32-bit words -> double in the open interval (0,1): (((hi:lo) >> 12) + 0.5) * 2^-52
__device__ __forceinline__ double u01_52(uint32_t lo, uint32_t hi) {
  unsigned long long x = ((unsigned long long)hi << 32) | (unsigned long long)lo;
  return __dmul_rn(__dadd_rn((double)(x >> 12), 0.5), 2.220446049250313e-16);
}

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

// (integer, fraction) of x * n / 2^64 for a 64-bit uniform x: the integer part is uniform on [0, n), the
// fraction is uniform on [0, 1) and independent of it (one draw -> randint + rand)
__device__ __forceinline__ void split_draw(uint32_t lo, uint32_t hi, uint32_t n, uint32_t& ipart, double& frac) {
  const unsigned long long x = ((unsigned long long)hi << 32) | (unsigned long long)lo;
  ipart = (uint32_t)__umul64hi(x, (unsigned long long)n);
  const unsigned long long f = x * (unsigned long long)n;
  frac = __dmul_rn(__dadd_rn((double)(f >> 12), 0.5), 2.220446049250313e-16);
}

// Keyed bijection of [0, n): balanced Feistel network on 2*hb bits + cycle walking.  The four round keys are one Philox
// block.  Domains of up to 6 bits (n <= 64) take eight rounds — rounds 4..7 reuse the keys offset by the golden-ratio
// constant — because four rounds of a 1..3-bit round function reach too few permutations: the frequencies
// P(sigma(p) = w) were measurably non-uniform there (chi-square test in tests/test_host_logic_cpu.py).  MCMC validity
// needs only that the pairing is independent of the walkers' state; fairness is what the extra rounds buy.
constexpr int FEISTEL_ROUNDS = 4;
struct Feistel {
  uint32_t k[FEISTEL_ROUNDS];
  uint32_t n, hb, mask, nr;

  __device__ __forceinline__ void set_size(uint32_t n_) {
    n = n_;
    uint32_t bits = (n_ <= 2u) ? 1u : (32u - (uint32_t)__clz((int)(n_ - 1u)));
    hb = (bits + 1u) >> 1;
    mask = (1u << hb) - 1u;
    nr = hb <= 3u ? 8u : 4u;
  }
  // round keys of bijection number `idx` of stream `tag`
  static __device__ __forceinline__ void make_keys(const RngKey& key, uint32_t tag, uint32_t idx, uint32_t* k4) {
    uint4 w = stream(key, tag, idx, 0u);
    k4[0] = w.x; k4[1] = w.y; k4[2] = w.z; k4[3] = w.w;
  }
  __device__ __forceinline__ void init(const RngKey& key, uint32_t tag, uint32_t idx, uint32_t n_) {
    make_keys(key, tag, idx, k);
    set_size(n_);
  }
  // keys computed once per block and staged in shared memory
  __device__ __forceinline__ void init_from(const uint32_t* k4, uint32_t n_) {
#pragma unroll
    for (int r = 0; r < FEISTEL_ROUNDS; ++r) k[r] = k4[r];
    set_size(n_);
  }
  __device__ __forceinline__ uint32_t round_key(int r) const { return k[r & 3] + 0x9E3779B9u * (uint32_t)(r >> 2); }
  // inverse bijection: rounds undone last to first, cycle walking on the inverse permutation
  __device__ __forceinline__ uint32_t inv(uint32_t x) const {
    if (n <= 1u) return 0u;
    do {
      uint32_t L = x >> hb, R = x & mask;
#pragma unroll
      for (int r = 2 * FEISTEL_ROUNDS - 1; r >= 0; --r) {
        if ((uint32_t)r < nr) {
          uint32_t pl = R ^ (fmix32(L ^ round_key(r)) & mask);
          R = L; L = pl;
        }
      }
      x = (L << hb) | R;
    } while (x >= n);
    return x;
  }
  __device__ __forceinline__ uint32_t operator()(uint32_t x) const {
    if (n <= 1u) return 0u;
    do {
      uint32_t L = x >> hb, R = x & mask;
#pragma unroll
      for (int r = 0; r < FEISTEL_ROUNDS; ++r) {
        uint32_t nr_ = L ^ (fmix32(R ^ k[r]) & mask);
        L = R; R = nr_;
      }
      if (nr > (uint32_t)FEISTEL_ROUNDS) {   // small domains only: four more rounds
#pragma unroll
        for (int r = FEISTEL_ROUNDS; r < 2 * FEISTEL_ROUNDS; ++r) {
          uint32_t nr_ = L ^ (fmix32(R ^ round_key(r)) & mask);
          L = R; R = nr_;
        }
      }
      x = (L << hb) | R;
    } while (x >= n);
    return x;
  }
};

}  // namespace eb
