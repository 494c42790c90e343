// Counter-based Philox4x32-10 (Salmon et al., SC'11) for the device sampler.
//
// Replaces the reference's stateful `Random<RNG>` (include/walnutpie/util.hpp:
// 78-162: std::mt19937_64 behind libstdc++ uniform / bernoulli / normal
// distributions), which cannot be evaluated in parallel.  Every draw is a pure
// function of (seed, global chain id, transition index, kind, index), so the
// result does not depend on how chains are scheduled or sharded over GPUs.
//
//   kind 0  momentum normals; index = pair j -> elements 2j, 2j+1 (Box-Muller)
//           (walnuts.hpp:528)
//   kind 1  scalar decisions in the reference's consumption order: the
//           direction bit (walnuts.hpp:552) and one uniform per Barker /
//           Metropolis merge (walnuts.hpp:378); index = running count
//   kind 2  initial positions (config.hpp:259-268)
//   kind 3  momentum for the initial step-size search (util.hpp:290-293)
#pragma once
#if !defined(__CUDACC_RTC__)
#include <cstdint>
#endif

namespace wb200 {

constexpr uint32_t kPhiloxKey1 = 0x57414C4Eu;
constexpr uint32_t kKindNormal = 0, kKindScalar = 1, kKindInit = 2,
                   kKindStepInit = 3;

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1,
                                                          uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// two words -> double in the open interval (0,1), 53 random bits
__host__ __device__ __forceinline__ double u01_from_words(uint32_t hi, uint32_t lo) {
  uint64_t mant = (static_cast<uint64_t>(hi >> 5) << 26) | (lo >> 6);
  return (static_cast<double>(mant) + 0.5) * (1.0 / 9007199254740992.0);
}

__host__ __device__ __forceinline__ Philox4 philox_draw(uint32_t seed, uint32_t chain,
                                                        uint32_t iter, uint32_t kind,
                                                        uint32_t index) {
  return philox4x32_10(chain, iter, kind, index, seed, kPhiloxKey1);
}

// Box-Muller pair j of stream (seed, chain, iter, kind)
__device__ __forceinline__ void philox_normal_pair(uint32_t seed, uint32_t chain,
                                                   uint32_t iter, uint32_t kind,
                                                   uint32_t j, double& z0, double& z1) {
  Philox4 p = philox_draw(seed, chain, iter, kind, j);
  double u1 = u01_from_words(p.x, p.y);
  double u2 = u01_from_words(p.z, p.w);
  double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincos(6.283185307179586476925286766559 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

__device__ __forceinline__ double philox_uniform(uint32_t seed, uint32_t chain,
                                                 uint32_t iter, uint32_t index) {
  Philox4 p = philox_draw(seed, chain, iter, kKindScalar, index);
  return u01_from_words(p.x, p.y);
}

__device__ __forceinline__ bool philox_bit(uint32_t seed, uint32_t chain,
                                           uint32_t iter, uint32_t index) {
  Philox4 p = philox_draw(seed, chain, iter, kKindScalar, index);
  return (p.x & 1u) != 0;
}

}  // namespace wb200
