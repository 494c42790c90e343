// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// Restatement of the posterior summaries of include/walnutpie/summary.hpp on
// ragged chains: pooled mean / variance (:371-405), biased autocovariance
// (:55-91, computed directly instead of by FFT — same quantity, tests pin both
// ways at tests/summary_test.cpp:681-691), R-hat (:594-619), Geyer ESS
// (:664-749) and MCSE (:765-769).
// Layout: draws are ROW-MAJOR [sum(lengths)][D], chains stacked in order.
// Pinned by the reference's golden values (tests/summary_test.cpp:661-677,
// :846-879, :1073-1083, :1182-1192) in tests/test_oracle_kat.py.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace oracle {

struct Chains {
  const double* draws;            // row-major [num_draws][D]
  std::size_t D;
  std::vector<std::size_t> len;   // per chain
  std::vector<std::size_t> start;

  Chains(const double* d, std::size_t num_draws, std::size_t dims,
         const std::vector<std::size_t>& lengths)
      : draws(d), D(dims), len(lengths), start(lengths.size()) {
    std::size_t total = 0;
    for (std::size_t m = 0; m < len.size(); ++m) {
      start[m] = total;
      total += len[m];
    }
    if (total != num_draws) {
      throw std::invalid_argument(
          "The number of rows in draws and sum of chain_sizes must be equal.");
    }
  }
  std::size_t num_chains() const { return len.size(); }
  std::size_t num_draws() const {
    return len.empty() ? 0 : start.back() + len.back();
  }
  std::size_t min_chain_size() const {
    return *std::min_element(len.begin(), len.end());
  }
  double at(std::size_t chain, std::size_t i, std::size_t d) const {
    return draws[(start[chain] + i) * D + d];
  }
};

inline double chain_mean(const Chains& c, std::size_t m, std::size_t d) {
  double s = 0.0;
  for (std::size_t i = 0; i < c.len[m]; ++i) s += c.at(m, i, d);
  return s / static_cast<double>(c.len[m]);
}

// detail::sample_variance (summary.hpp:93-105) of one chain column
inline double chain_var(const Chains& c, std::size_t m, std::size_t d,
                        double mean) {
  double s = 0.0;
  for (std::size_t i = 0; i < c.len[m]; ++i) {
    double x = c.at(m, i, d) - mean;
    s += x * x;
  }
  return s / static_cast<double>(c.len[m] - 1);  // signed in ref: (rows() - 1)
}

inline std::vector<double> pooled_mean(const Chains& c) {  // :371-378
  std::vector<double> mu(c.D, 0.0);
  for (std::size_t m = 0; m < c.num_chains(); ++m) {
    for (std::size_t d = 0; d < c.D; ++d) {
      double s = 0.0;
      for (std::size_t i = 0; i < c.len[m]; ++i) s += c.at(m, i, d);
      mu[d] += s;
    }
  }
  for (double& v : mu) v /= static_cast<double>(c.num_draws());
  return mu;
}

inline std::vector<double> pooled_variance(const Chains& c) {  // :397-405
  std::vector<double> mu = pooled_mean(c), ss(c.D, 0.0);
  for (std::size_t m = 0; m < c.num_chains(); ++m) {
    for (std::size_t d = 0; d < c.D; ++d) {
      double s = 0.0;
      for (std::size_t i = 0; i < c.len[m]; ++i) {
        double x = c.at(m, i, d) - mu[d];
        s += x * x;
      }
      ss[d] += s;
    }
  }
  for (double& v : ss) v /= static_cast<double>(c.num_draws() - 1);
  return ss;
}

// biased autocovariance of chain m, dimension d, all lags (summary.hpp:55-73)
inline std::vector<double> autocovariance(const Chains& c, std::size_t m,
                                          std::size_t d) {
  const std::size_t N = c.len[m];
  std::vector<double> y(N), ac(N);
  double mu = chain_mean(c, m, d);
  for (std::size_t i = 0; i < N; ++i) y[i] = c.at(m, i, d) - mu;
  for (std::size_t lag = 0; lag < N; ++lag) {
    double s = 0.0;
    for (std::size_t i = 0; i + lag < N; ++i) s += y[i] * y[i + lag];
    ac[lag] = s / static_cast<double>(N);
  }
  return ac;
}

// variance with ddof=1 across the K chain summaries (detail::sample_variance
// applied to the K x D matrix, summary.hpp:616, :685)
inline double across_var(const std::vector<double>& v) {
  double m = 0.0;
  for (double x : v) m += x;
  m /= static_cast<double>(v.size());
  double s = 0.0;
  for (double x : v) s += (x - m) * (x - m);
  return s / static_cast<double>(v.size() - 1);
}

inline std::vector<double> r_hat(const Chains& c) {  // :594-619
  if (c.num_chains() < 2) {
    throw std::invalid_argument("require at least two chains to compute R-hat");
  }
  for (std::size_t m = 0; m < c.num_chains(); ++m) {
    if (c.len[m] < 3) {
      throw std::invalid_argument("each chain must have at least 3 draws");
    }
  }
  const std::size_t M = c.num_chains();
  std::vector<double> out(c.D);
  std::vector<double> mu(M), s2(M);
  for (std::size_t d = 0; d < c.D; ++d) {
    double mean_s2 = 0.0;
    for (std::size_t m = 0; m < M; ++m) {
      mu[m] = chain_mean(c, m, d);
      s2[m] = chain_var(c, m, d, mu[m]);
      mean_s2 += s2[m];
    }
    mean_s2 /= static_cast<double>(M);
    out[d] = std::sqrt(1.0 + across_var(mu) / mean_s2);
  }
  return out;
}

inline std::vector<double> effective_sample_size(const Chains& c) {  // :664-749
  if (c.num_draws() < 3) {
    throw std::invalid_argument("chains must have at least 3 draws");
  }
  const std::size_t K = c.num_chains();
  const std::size_t N_total = c.num_draws();
  const long min_len = static_cast<long>(c.min_chain_size());
  std::vector<double> out(c.D);
  std::vector<double> mu(K), s2(K);
  for (std::size_t d = 0; d < c.D; ++d) {
    double W = 0.0;
    for (std::size_t k = 0; k < K; ++k) {
      mu[k] = chain_mean(c, k, d);
      s2[k] = chain_var(c, k, d, mu[k]);
      W += s2[k];
    }
    W /= static_cast<double>(K);
    double var_plus = W;
    if (K > 1) var_plus += across_var(mu);

    std::vector<std::vector<double>> acov(K);
    for (std::size_t k = 0; k < K; ++k) acov[k] = autocovariance(c, k, d);
    auto mean_acov_at_lag = [&](long t) {
      double sum = 0.0;
      for (std::size_t k = 0; k < K; ++k) sum += acov[k][t];
      return sum / static_cast<double>(K);
    };

    // the reference writes rho_hat_t(max_t + 1) unconditionally (:737); give
    // the scratch two spare slots so short chains stay in bounds
    std::vector<double> rho_hat_t(static_cast<std::size_t>(min_len) + 4, 0.0);
    double rho_hat_even = 1.0;
    rho_hat_t[0] = rho_hat_even;
    double rho_hat_odd = 1.0 - (W - mean_acov_at_lag(1)) / var_plus;
    rho_hat_t[1] = rho_hat_odd;
    long t = 1;
    while (t < min_len - 4 && (rho_hat_even + rho_hat_odd) > 0.0) {
      rho_hat_even = 1.0 - (W - mean_acov_at_lag(t + 1)) / var_plus;
      rho_hat_odd = 1.0 - (W - mean_acov_at_lag(t + 2)) / var_plus;
      if ((rho_hat_even + rho_hat_odd) >= 0.0) {
        rho_hat_t[t + 1] = rho_hat_even;
        rho_hat_t[t + 2] = rho_hat_odd;
      }
      if (rho_hat_t[t + 1] + rho_hat_t[t + 2] >
          rho_hat_t[t - 1] + rho_hat_t[t]) {
        rho_hat_t[t + 1] = (rho_hat_t[t - 1] + rho_hat_t[t]) / 2.0;
        rho_hat_t[t + 2] = rho_hat_t[t + 1];
      }
      t += 2;
    }
    long max_t = t;
    if (rho_hat_even > 0.0) rho_hat_t[max_t + 1] = rho_hat_even;
    double head = 0.0;
    for (long s = 0; s < max_t; ++s) head += rho_hat_t[s];
    double tau_hat = -1.0 + 2.0 * head + rho_hat_t[max_t + 1];
    tau_hat = std::max(tau_hat, 1.0 / std::log10(static_cast<double>(N_total)));
    out[d] = static_cast<double>(N_total) / tau_hat;
  }
  return out;
}

inline std::vector<double> mcse(const Chains& c) {  // :765-769
  std::vector<double> ess = effective_sample_size(c);
  std::vector<double> var = pooled_variance(c);
  std::vector<double> out(c.D);
  for (std::size_t d = 0; d < c.D; ++d) out[d] = std::sqrt(var[d]) / std::sqrt(ess[d]);
  return out;
}

}  // namespace oracle
