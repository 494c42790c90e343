// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// Counter-based Philox4x32-10 generator (Salmon et al., "Parallel random
// numbers: as easy as 1, 2, 3", SC'11) and the draw conventions the B200
// sampler uses.  The reference (util.hpp:78-162) draws from a *stateful*
// std::mt19937_64 through libstdc++ distributions; that stream cannot be
// reproduced on a GPU, so parity of identically-seeded trajectories is defined
// on this stateless stream instead.  The product keeps its own
// independent implementation in walnuts_b200/csrc/philox.cuh; the two are
// cross-checked against the Random123 known-answer vectors in tests/.
//
// Addressing: key = (seed, 0x57414C4E), counter = (chain, iteration, kind, index)
//   kind 0: momentum normals, index = pair j -> elements 2j, 2j+1 (Box-Muller)
//   kind 1: scalar decisions (direction bit, Barker / Metropolis uniforms),
//           index = running count of scalar draws inside the transition, in the
//           reference's consumption order (walnuts.hpp:552, :378).
//   kind 2: initial positions,  kind 3: step-size-initialisation momentum
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace oracle {

struct Philox4x32 {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

  static void block(const uint32_t ctr[4], const uint32_t key[2],
                    uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      uint64_t p0 = static_cast<uint64_t>(M0) * c0;
      uint64_t p1 = static_cast<uint64_t>(M1) * c2;
      uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
      uint32_t n1 = static_cast<uint32_t>(p1);
      uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
      uint32_t n3 = static_cast<uint32_t>(p0);
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

constexpr uint32_t kPhiloxKey1 = 0x57414C4Eu;  // "WALN"
enum PhiloxKind : uint32_t { kKindNormal = 0, kKindScalar = 1, kKindInit = 2,
                             kKindStepInit = 3 };

// two 32-bit words -> double in the open interval (0, 1), 53 random bits
inline double u01_from_words(uint32_t hi, uint32_t lo) {
  uint64_t mant = (static_cast<uint64_t>(hi >> 5) << 26) | (lo >> 6);
  return (static_cast<double>(mant) + 0.5) * (1.0 / 9007199254740992.0);
}

inline void philox_draw(uint32_t seed, uint32_t chain, uint32_t iter,
                        uint32_t kind, uint32_t index, uint32_t out[4]) {
  uint32_t ctr[4] = {chain, iter, kind, index};
  uint32_t key[2] = {seed, kPhiloxKey1};
  Philox4x32::block(ctr, key, out);
}

// Box-Muller pair for (chain, iter, kind, pair j)
inline void philox_normal_pair(uint32_t seed, uint32_t chain, uint32_t iter,
                               uint32_t kind, uint32_t j, double& z0,
                               double& z1) {
  uint32_t w[4];
  philox_draw(seed, chain, iter, kind, j, w);
  double u1 = u01_from_words(w[0], w[1]);
  double u2 = u01_from_words(w[2], w[3]);
  double r = std::sqrt(-2.0 * std::log(u1));
  double a = 6.283185307179586476925286766559 * u2;
  z0 = r * std::cos(a);
  z1 = r * std::sin(a);
}

/// Source of randomness with the reference's `Random<RNG>` interface
/// (util.hpp:78-162) over the stateless Philox stream.
class PhiloxRand {
 public:
  PhiloxRand(uint32_t seed, uint32_t chain)
      : seed_(seed), chain_(chain), iter_(0), scalar_(0) {}
  void begin_transition(uint32_t iter) { iter_ = iter; scalar_ = 0; }
  uint32_t iteration() const { return iter_; }

  double uniform_real_01() {
    uint32_t w[4];
    philox_draw(seed_, chain_, iter_, kKindScalar, scalar_++, w);
    return u01_from_words(w[0], w[1]);
  }
  bool uniform_binary() {
    uint32_t w[4];
    philox_draw(seed_, chain_, iter_, kKindScalar, scalar_++, w);
    return (w[0] & 1u) != 0;
  }
  std::vector<double> standard_normal(std::size_t n) {
    return normals(n, kKindNormal);
  }
  std::vector<double> normals(std::size_t n, uint32_t kind) {
    std::vector<double> z(n);
    for (std::size_t j = 0; 2 * j < n; ++j) {
      double z0, z1;
      philox_normal_pair(seed_, chain_, iter_, kind, static_cast<uint32_t>(j),
                         z0, z1);
      z[2 * j] = z0;
      if (2 * j + 1 < n) z[2 * j + 1] = z1;
    }
    return z;
  }

 private:
  uint32_t seed_, chain_, iter_, scalar_;
};

}  // namespace oracle
