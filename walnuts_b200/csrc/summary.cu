// Posterior summaries on the device: per-dimension R-hat, effective sample
// size, MCSE, pooled mean and variance over ragged chains.
//
// Reference: include/walnutpie/summary.hpp — r_hat :594-619,
// effective_sample_size :664-749 (Geyer initial positive / monotone sequence on
// the chain-averaged BIASED autocovariance :55-73), monte_carlo_standard_error
// :765-769, mean :371-378, sample_variance :397-405.  The reference computes
// every lag with an FFT per (chain, dimension); the Geyer loop only ever reads
// the first few lags, so the kernels here compute lags directly in blocks of
// kLagBlock and extend only while some dimension's sequence is still positive.
#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.cuh"

namespace wb200 {

struct SummaryView {
  const double* x;  // device rows, stride ld
  int ld, D, K;
  const long long* start;  // [K]
  const long long* len;    // [K]
  long long n_total, min_len;
};

// mean and unbiased variance of every (chain, dimension) series
__global__ void chain_moments_kernel(SummaryView v, double* mu, double* s2) {
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  const int k = blockIdx.x;  // chains on the x dimension of the grid: no 65 535 limit
  if (d >= v.D) return;
  const double* base = v.x + v.start[k] * v.ld + d;
  const long long n = v.len[k];
  double s = 0.0;
  for (long long i = 0; i < n; ++i) s += base[i * v.ld];
  const double m = s / static_cast<double>(n);
  double q = 0.0;
  for (long long i = 0; i < n; ++i) {
    const double c = base[i * v.ld] - m;
    q += c * c;
  }
  mu[static_cast<long long>(k) * v.D + d] = m;
  s2[static_cast<long long>(k) * v.D + d] = q / static_cast<double>(n - 1);
}

// across-chain pieces: W = mean_k s2, B = var_{K-1}(mu_k), pooled mean / var
__global__ void across_kernel(SummaryView v, const double* mu, const double* s2,
                              double* W, double* B, double* pmean, double* pvar) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= v.D) return;
  double sw = 0.0, sm = 0.0, tot = 0.0;
  for (int k = 0; k < v.K; ++k) {
    sw += s2[static_cast<long long>(k) * v.D + d];
    const double m = mu[static_cast<long long>(k) * v.D + d];
    sm += m;
    tot += m * static_cast<double>(v.len[k]);
  }
  const double mbar = sm / v.K;
  double q = 0.0;
  for (int k = 0; k < v.K; ++k) {
    const double c = mu[static_cast<long long>(k) * v.D + d] - mbar;
    q += c * c;
  }
  W[d] = sw / v.K;
  B[d] = q / static_cast<double>(v.K - 1);  // NaN for one chain, like the reference
  const double pm = tot / static_cast<double>(v.n_total);
  // sum_i (x - pm)^2 = (n-1) s2 + n (mu - pm)^2 per chain
  double ss = 0.0;
  for (int k = 0; k < v.K; ++k) {
    const double n = static_cast<double>(v.len[k]);
    const double c = mu[static_cast<long long>(k) * v.D + d] - pm;
    ss += (n - 1.0) * s2[static_cast<long long>(k) * v.D + d] + n * c * c;
  }
  pmean[d] = pm;
  pvar[d] = ss / static_cast<double>(v.n_total - 1);
}

// sum over chains of the biased autocovariance at lags [lag0, lag0+nlag):
// acc[(t-lag0)*D + d] += (1/n_k) sum_i y_i y_{i+t}.  One block = one chain x
// TILE dimensions staged in shared memory.
template <int TILE>
__global__ void acov_kernel(SummaryView v, const double* mu, int lag0, int nlag,
                            double* acc) {
  extern __shared__ double tile[];  // [len][TILE]
  const int k = blockIdx.x;
  const int d0 = blockIdx.y * TILE;
  const long long n = v.len[k];
  const double* base = v.x + v.start[k] * v.ld;
  for (long long p = threadIdx.x; p < n * TILE; p += blockDim.x) {
    const long long i = p / TILE;
    const int j = static_cast<int>(p % TILE);
    const int d = d0 + j;
    tile[p] = d < v.D ? base[i * v.ld + d] - mu[static_cast<long long>(k) * v.D + d] : 0.0;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < nlag * TILE; p += blockDim.x) {
    const int t = lag0 + p / TILE;
    const int j = p % TILE;
    if (d0 + j >= v.D || t >= n) continue;
    double s = 0.0;
    for (long long i = 0; i + t < n; ++i) s += tile[i * TILE + j] * tile[(i + t) * TILE + j];
    atomicAdd(acc + static_cast<long long>(t - lag0) * v.D + d0 + j, s / static_cast<double>(n));
  }
}

// The same for chains too long for the shared-memory tile: operands come straight from
// global memory (L1 / L2 serve the re-reads across lags).
template <int TILE>
__global__ void acov_global_kernel(SummaryView v, const double* mu, int lag0, int nlag,
                                   double* acc) {
  const int k = blockIdx.x;
  const int d0 = blockIdx.y * TILE;
  const long long n = v.len[k];
  const double* base = v.x + v.start[k] * v.ld;
  for (int p = threadIdx.x; p < nlag * TILE; p += blockDim.x) {
    const int t = lag0 + p / TILE;
    const int j = p % TILE;
    const int d = d0 + j;
    if (d >= v.D || t >= n) continue;
    const double m = mu[static_cast<long long>(k) * v.D + d];
    double s = 0.0;
    for (long long i = 0; i + t < n; ++i) {
      s += (base[i * v.ld + d] - m) * (base[(i + t) * v.ld + d] - m);
    }
    atomicAdd(acc + static_cast<long long>(t - lag0) * v.D + d, s / static_cast<double>(n));
  }
}

// Geyer estimator, one thread per dimension (summary.hpp:700-748).
// macov[t*D + d] holds the chain SUM of autocovariances for lags < nlag.
// flag[d] = 1 when a lag >= nlag was needed (caller extends and reruns).
__global__ void geyer_kernel(SummaryView v, const double* macov, int nlag,
                             const double* W, const double* B, const double* pvar,
                             double* rho_buf, double* rhat, double* ess, double* mcse,
                             int* flag) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= v.D) return;
  const double w_d = W[d];
  const double vp_d = v.K > 1 ? w_d + B[d] : w_d;
  rhat[d] = sqrt(1.0 + B[d] / w_d);
  const long long min_len = v.min_len;
  double* rho_hat_t = rho_buf + d;  // stride D, min_len + 4 entries (zeroed)
  const long long S = v.D;
  bool need_more = false;
  auto acov_at = [&](long long t) -> double {
    if (t >= nlag) { need_more = true; return 0.0; }
    return macov[t * v.D + d] / static_cast<double>(v.K);
  };
  double rho_hat_even = 1.0;
  rho_hat_t[0] = rho_hat_even;
  double rho_hat_odd = 1.0 - (w_d - acov_at(1)) / vp_d;
  rho_hat_t[1 * S] = rho_hat_odd;
  long long t = 1;
  while (t < min_len - 4 && (rho_hat_even + rho_hat_odd) > 0.0) {
    rho_hat_even = 1.0 - (w_d - acov_at(t + 1)) / vp_d;
    rho_hat_odd = 1.0 - (w_d - acov_at(t + 2)) / vp_d;
    if (need_more) break;
    if ((rho_hat_even + rho_hat_odd) >= 0.0) {
      rho_hat_t[(t + 1) * S] = rho_hat_even;
      rho_hat_t[(t + 2) * S] = rho_hat_odd;
    }
    if (rho_hat_t[(t + 1) * S] + rho_hat_t[(t + 2) * S] >
        rho_hat_t[(t - 1) * S] + rho_hat_t[t * S]) {
      rho_hat_t[(t + 1) * S] = (rho_hat_t[(t - 1) * S] + rho_hat_t[t * S]) / 2.0;
      rho_hat_t[(t + 2) * S] = rho_hat_t[(t + 1) * S];
    }
    t += 2;
  }
  flag[d] = need_more ? 1 : 0;
  if (need_more) return;
  const long long max_t = t;
  if (rho_hat_even > 0.0) rho_hat_t[(max_t + 1) * S] = rho_hat_even;
  double head = 0.0;
  for (long long s = 0; s < max_t; ++s) head += rho_hat_t[s * S];
  double tau_hat = -1.0 + 2.0 * head + rho_hat_t[(max_t + 1) * S];
  tau_hat = fmax(tau_hat, 1.0 / log10(static_cast<double>(v.n_total)));
  const double e = static_cast<double>(v.n_total) / tau_hat;
  ess[d] = e;
  mcse[d] = sqrt(pvar[d]) / sqrt(e);
}

// {sum_k mu_k, sum_k mu_k^2, sum_k s2_k}[D] -- the all-reduce payload of a cross-GPU R-hat
__global__ void rhat_sums_kernel(SummaryView v, const double* mu, const double* s2,
                                 double* out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= v.D) return;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int k = 0; k < v.K; ++k) {
    const double m = mu[static_cast<long long>(k) * v.D + d];
    a += m;
    b += m * m;
    c += s2[static_cast<long long>(k) * v.D + d];
  }
  out[d] = a;
  out[v.D + d] = b;
  out[2 * v.D + d] = c;
}

void device_rhat_moments(const double* draws, int ld, int D,
                         const std::vector<long long>& start,
                         const std::vector<long long>& len, double* moments_host,
                         cudaStream_t stream) {
  const int K = static_cast<int>(len.size());
  if (K == 0) throw std::invalid_argument("chains cannot be empty.");
  long long n_total = 0, min_len = len[0];
  for (long long l : len) { n_total += l; min_len = std::min(min_len, l); }
  DeviceBuffer<long long> d_start, d_len;
  d_start.alloc(K); d_len.alloc(K);
  WB200_CUDA(cudaMemcpyAsync(d_start.ptr, start.data(), K * 8, cudaMemcpyHostToDevice, stream));
  WB200_CUDA(cudaMemcpyAsync(d_len.ptr, len.data(), K * 8, cudaMemcpyHostToDevice, stream));
  SummaryView v{draws, ld, D, K, d_start.ptr, d_len.ptr, n_total, min_len};
  DeviceBuffer<double> mu, s2, out;
  mu.alloc(static_cast<size_t>(K) * D); s2.alloc(static_cast<size_t>(K) * D);
  out.alloc(3 * static_cast<size_t>(D));
  const int tb = 128;
  chain_moments_kernel<<<dim3(K, (D + tb - 1) / tb), tb, 0, stream>>>(v, mu.ptr, s2.ptr);
  rhat_sums_kernel<<<(D + tb - 1) / tb, tb, 0, stream>>>(v, mu.ptr, s2.ptr, out.ptr);
  WB200_CUDA(cudaGetLastError());
  WB200_CUDA(cudaMemcpyAsync(moments_host, out.ptr, 3 * D * 8, cudaMemcpyDeviceToHost, stream));
  WB200_CUDA(cudaStreamSynchronize(stream));
  moments_host[3 * D] = static_cast<double>(K);
}

constexpr int kLagBlock = 32;

void device_summary(const double* draws, int ld, int D,
                    const std::vector<long long>& start,
                    const std::vector<long long>& len, double* rhat, double* ess,
                    double* mcse, double* mean, double* var, cudaStream_t stream) {
  const int K = static_cast<int>(len.size());
  if (K == 0) throw std::invalid_argument("chains cannot be empty.");
  long long n_total = 0, min_len = len[0], max_len = 0;
  for (long long l : len) {
    if (l <= 0) throw std::invalid_argument("chains must have at least one draw.");
    n_total += l;
    min_len = std::min(min_len, l);
    max_len = std::max(max_len, l);
  }
  const bool want_rhat = rhat != nullptr;
  const bool want_ess = ess != nullptr || mcse != nullptr;
  if (want_rhat) {  // summary.hpp:595-603
    if (K < 2) throw std::invalid_argument("require at least two chains to compute R-hat");
    for (long long l : len) {
      if (l < 3) throw std::invalid_argument("each chain must have at least 3 draws");
    }
  }
  if (want_ess && n_total < 3) {  // :665-667
    throw std::invalid_argument("chains must have at least 3 draws");
  }
  DeviceBuffer<long long> d_start, d_len;
  d_start.alloc(K); d_len.alloc(K);
  WB200_CUDA(cudaMemcpyAsync(d_start.ptr, start.data(), K * 8, cudaMemcpyHostToDevice, stream));
  WB200_CUDA(cudaMemcpyAsync(d_len.ptr, len.data(), K * 8, cudaMemcpyHostToDevice, stream));
  SummaryView v{draws, ld, D, K, d_start.ptr, d_len.ptr, n_total, min_len};
  DeviceBuffer<double> mu, s2, W, B, pm, pv, o_rhat, o_ess, o_mcse;
  mu.alloc(static_cast<size_t>(K) * D); s2.alloc(static_cast<size_t>(K) * D);
  W.alloc(D); B.alloc(D); pm.alloc(D); pv.alloc(D);
  o_rhat.alloc(D); o_ess.alloc(D); o_mcse.alloc(D);
  const int tb = 128;
  chain_moments_kernel<<<dim3(K, (D + tb - 1) / tb), tb, 0, stream>>>(v, mu.ptr, s2.ptr);
  WB200_CUDA(cudaGetLastError());
  across_kernel<<<(D + tb - 1) / tb, tb, 0, stream>>>(v, mu.ptr, s2.ptr, W.ptr, B.ptr,
                                                       pm.ptr, pv.ptr);
  WB200_CUDA(cudaGetLastError());
  if (want_ess || want_rhat) {
    // lags in blocks until every dimension's Geyer sequence has terminated
    constexpr int TILE = 8;
    const size_t smem = static_cast<size_t>(max_len) * TILE * sizeof(double);
    const bool staged = smem <= 200 * 1024;  // up to 3200 draws per chain
    if (staged) {
      WB200_CUDA(cudaFuncSetAttribute(acov_kernel<TILE>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    }
    DeviceBuffer<double> macov, rho_buf;
    DeviceBuffer<int> flag;
    flag.alloc(D);
    rho_buf.alloc(static_cast<size_t>(min_len + 4) * D);
    int nlag = 0;
    const int max_lag = static_cast<int>(std::max<long long>(min_len, 2));
    std::vector<int> h_flag(D);
    int target = std::min(kLagBlock, max_lag);
    DeviceBuffer<double> grown;
    while (true) {
      // extend macov to `target` lags, keeping what is already there
      grown.alloc(static_cast<size_t>(target) * D);
      WB200_CUDA(cudaMemsetAsync(grown.ptr, 0, static_cast<size_t>(target) * D * 8, stream));
      if (nlag > 0) {
        WB200_CUDA(cudaMemcpyAsync(grown.ptr, macov.ptr, static_cast<size_t>(nlag) * D * 8,
                                   cudaMemcpyDeviceToDevice, stream));
      }
      if (staged) {
        acov_kernel<TILE><<<dim3(K, (D + TILE - 1) / TILE), 256, smem, stream>>>(
            v, mu.ptr, nlag, target - nlag, grown.ptr + static_cast<size_t>(nlag) * D);
      } else {
        acov_global_kernel<TILE><<<dim3(K, (D + TILE - 1) / TILE), 256, 0, stream>>>(
            v, mu.ptr, nlag, target - nlag, grown.ptr + static_cast<size_t>(nlag) * D);
      }
      WB200_CUDA(cudaGetLastError());
      WB200_CUDA(cudaStreamSynchronize(stream));
      std::swap(macov.ptr, grown.ptr);
      std::swap(macov.count, grown.count);
      nlag = target;
      WB200_CUDA(cudaMemsetAsync(rho_buf.ptr, 0, rho_buf.count * 8, stream));
      geyer_kernel<<<(D + tb - 1) / tb, tb, 0, stream>>>(
          v, macov.ptr, nlag, W.ptr, B.ptr, pv.ptr, rho_buf.ptr, o_rhat.ptr, o_ess.ptr,
          o_mcse.ptr, flag.ptr);
      WB200_CUDA(cudaGetLastError());
      WB200_CUDA(cudaMemcpyAsync(h_flag.data(), flag.ptr, D * 4, cudaMemcpyDeviceToHost, stream));
      WB200_CUDA(cudaStreamSynchronize(stream));
      bool more = false;
      for (int f : h_flag) more = more || f;
      if (!more || nlag >= max_lag) break;
      target = std::min(max_lag, nlag * 4);
    }
  }
  auto fetch = [&](double* dst, const double* src) {
    if (dst) WB200_CUDA(cudaMemcpyAsync(dst, src, D * 8, cudaMemcpyDeviceToHost, stream));
  };
  fetch(rhat, o_rhat.ptr); fetch(ess, o_ess.ptr); fetch(mcse, o_mcse.ptr);
  fetch(mean, pm.ptr); fetch(var, pv.ptr);
  WB200_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace wb200

using namespace wb200;

// walnutpy.cpp:86-95 + :333-369, with the host matrix taken ROW-MAJOR as
// python/src/walnutpie/summary.py:30 actually passes it
static int host_summary(const double* draws, int num_draws, int num_params,
                        const int* lengths, int num_chains, double* out, int which,
                        WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    require_gpu();
    std::vector<long long> start(num_chains), len(num_chains);
    long long total = 0;
    for (int i = 0; i < num_chains; ++i) {
      start[i] = total;
      len[i] = lengths[i];
      total += lengths[i];
    }
    if (total != num_draws) {  // summary.hpp:264-267
      throw std::invalid_argument(
          "The number of rows in draws and sum of chain_sizes must be equal.");
    }
    DeviceBuffer<double> d;
    d.alloc(static_cast<size_t>(num_draws) * num_params);
    WB200_CUDA(cudaMemcpy(d.ptr, draws, d.count * 8, cudaMemcpyHostToDevice));
    device_summary(d.ptr, num_params, num_params, start, len,
                   which == 1 ? out : nullptr, which == 0 ? out : nullptr,
                   which == 2 ? out : nullptr, nullptr, nullptr, nullptr);
  });
}

extern "C" {

int walnutpie_ess(const double* draws, int num_draws, int num_params,
                  const int* lengths, int num_chains, double* out, WalnutpyError** err) {
  return host_summary(draws, num_draws, num_params, lengths, num_chains, out, 0, err);
}
int walnutpie_r_hat(const double* draws, int num_draws, int num_params,
                    const int* lengths, int num_chains, double* out, WalnutpyError** err) {
  return host_summary(draws, num_draws, num_params, lengths, num_chains, out, 1, err);
}
int walnutpie_mcse(const double* draws, int num_draws, int num_params,
                   const int* lengths, int num_chains, double* out, WalnutpyError** err) {
  return host_summary(draws, num_draws, num_params, lengths, num_chains, out, 2, err);
}

int wb200_device_summary(const double* draws_device, size_t num_chains,
                         long long capacity, long long first, long long count, int D,
                         int ld, double* rhat, double* ess, double* mcse, double* mean,
                         double* var, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    std::vector<long long> start(num_chains), len(num_chains, count);
    for (size_t c = 0; c < num_chains; ++c) {
      start[c] = static_cast<long long>(c) * capacity + first;
    }
    device_summary(draws_device, ld, D, start, len, rhat, ess, mcse, mean, var, nullptr);
  });
}

}  // extern "C"
