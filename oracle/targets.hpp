// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// Target log densities with the reference's LogpGrad call shape
//   void(const VectorXd& x, double& logp, VectorXd& grad)   (concepts.hpp:258-262)
// restated on std::vector<double>.  std_normal follows
// examples/walnutpie_api.cpp:39-43 / tests/test_util.hpp:78-82; the diagonal
// Gaussian generalises examples/examples.cpp:20-31 (`ill_normal`) to an
// arbitrary precision vector; the funnel and the logistic regression are the
// SURVEY.md §8(d) definitions (not in the reference repo).
// All sums are plain left-to-right fp64 loops.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace oracle {

using Vec = std::vector<double>;

// Arithmetic policy of the accumulate sites a*b + c of the hot loops (leapfrog kicks and
// drift, the targets' sums of squares, kinetic energy, U-turn dots, the discounted sum
// of squares of OnlineMoments).  Default = the reference's baseline x86-64 build: product
// and sum round separately (this file is compiled with -ffp-contract=off).  "Fused" = one
// rounding, what the device kernels ship (walnuts_b200/csrc/chain_kernel.cuh, kFusedArith)
// and what a contracting build of the reference would do.  The fused policy also takes
// the two algebraic short cuts of the device estimators: the common weight of the two
// discounted Welford estimators cancels in var_draws / var_scores
// (adaptive_walnuts.hpp:89-94 -> sqrt(S_draw / S_score)) and the mean update multiplies
// by one reciprocal of that scalar weight (online_moments.hpp:187).  Process-wide switch;
// the reference-policy oracle is the one pinned bit for bit to the reference's headers.
inline bool& fused_arith() {
  static bool fused = false;
  return fused;
}
inline double madd(double a, double b, double c) {
  return fused_arith() ? std::fma(a, b, c) : a * b + c;
}

enum TargetKind : int { kStdNormal = 0, kDiagGaussian = 1, kFunnel = 2,
                        kLogistic = 3 };

struct StdNormal {
  void eval(const double* x, std::size_t n, double& lp, double* g) const {
    double s = 0.0;
    for (std::size_t i = 0; i < n; ++i) {
      s = madd(x[i], x[i], s);
      g[i] = -x[i];
    }
    lp = -0.5 * s;
  }
  void operator()(const Vec& x, double& lp, Vec& g) const {
    g.resize(x.size());
    eval(x.data(), x.size(), lp, g.data());
  }
};

// logp = -1/2 sum_d x_d^2 * prec_d ; grad_d = -(x_d * prec_d)
struct DiagGaussian {
  Vec prec;  // 1 / sigma_d^2
  void eval(const double* x, std::size_t n, double& lp, double* g) const {
    double s = 0.0;
    for (std::size_t i = 0; i < n; ++i) {
      double t = x[i] * prec[i];
      s = madd(x[i], t, s);
      g[i] = -t;
    }
    lp = -0.5 * s;
  }
  void operator()(const Vec& x, double& lp, Vec& g) const {
    g.resize(x.size());
    eval(x.data(), x.size(), lp, g.data());
  }
};

// Neal's funnel: v = x_0 ~ N(0, 3^2); x_i | v ~ N(0, e^v), i >= 1
// logp = -v^2/18 - (D-1)/2 v - 1/2 e^{-v} sum_i x_i^2
// The prior's constants enter as the rounded reciprocals 1/18 and 1/9 (one multiplication
// each): this target is specified here, not in the reference, and an fp64 division is
// 30 instructions per gradient on the device for the same density to 1 ulp.
struct Funnel {
  static constexpr double kInv18 = 1.0 / 18.0, kInv9 = 1.0 / 9.0;
  void operator()(const Vec& x, double& lp, Vec& g) const {
    g.resize(x.size());
    eval(x.data(), x.size(), lp, g.data());
  }
  void eval(const double* x, std::size_t D, double& lp, double* g) const {
    const double v = x[0];
    const double ev = std::exp(-v);
    double ss = 0.0;
    for (std::size_t i = 1; i < D; ++i) {
      ss = madd(x[i], x[i], ss);
      g[i] = -(x[i] * ev);
    }
    const double half_dm1 = 0.5 * static_cast<double>(D - 1);
    const double q = 0.5 * ev * ss;
    lp = -((v * v) * kInv18) - half_dm1 * v - q;
    g[0] = -(v * kInv9) - half_dm1 + q;
  }
};

// Bayesian logistic regression, prior theta ~ N(0, I):
// logp = sum_n [y_n z_n - softplus(z_n)] - 1/2 |theta|^2,  z = X theta
// grad = X^T (y - sigmoid(z)) - theta.     X row-major [N][D].
struct Logistic {
  std::size_t N = 0, D = 0;
  const double* X = nullptr;
  const double* y = nullptr;
  void operator()(const Vec& x, double& lp, Vec& g) const {
    g.resize(D);
    eval(x.data(), D, lp, g.data());
  }
  void eval(const double* x, std::size_t, double& lp, double* g) const {
    for (std::size_t d = 0; d < D; ++d) g[d] = 0.0;
    double ll = 0.0;
    for (std::size_t n = 0; n < N; ++n) {
      const double* row = X + n * D;
      double z = 0.0;
      for (std::size_t d = 0; d < D; ++d) z += row[d] * x[d];
      // softplus(z) = max(z,0) + log1p(exp(-|z|))
      double sp = (z > 0 ? z : 0.0) + std::log1p(std::exp(-std::fabs(z)));
      double sig = z >= 0 ? 1.0 / (1.0 + std::exp(-z))
                          : std::exp(z) / (1.0 + std::exp(z));
      ll += y[n] * z - sp;
      double r = y[n] - sig;
      for (std::size_t d = 0; d < D; ++d) g[d] += row[d] * r;
    }
    double ss = 0.0;
    for (std::size_t d = 0; d < D; ++d) {
      ss += x[d] * x[d];
      g[d] -= x[d];
    }
    lp = ll - 0.5 * ss;
  }
};

// the C form of the plug-in (walnutpy.cpp:127-132)
typedef int (*LOGP_CFUNC)(std::size_t n, const double* theta, double* grad,
                          double* lp, void* data);

struct CFuncTarget {
  LOGP_CFUNC fn = nullptr;
  void* data = nullptr;
  void operator()(const Vec& x, double& lp, Vec& g) const {
    g.resize(x.size());
    eval(x.data(), x.size(), lp, g.data());
  }
  void eval(const double* x, std::size_t n, double& lp, double* g) const {
    int rc = fn(n, x, g, &lp, data);
    if (rc != 0) {
      throw std::runtime_error("logp failed with code " + std::to_string(rc));
    }
  }
};

}  // namespace oracle
