// Streaming posterior summaries: exact reference R-hat / ESS / MCSE without keeping
// (or shipping) the draws, and across any number of GPUs.
//
// Reference: include/walnutpie/summary.hpp -- r_hat :594-619, effective_sample_size
// :664-749, monte_carlo_standard_error :765-769, mean :371-378, sample_variance :397-405.
// Everything those functions need of a chain is linear in a few running sums, so each
// chain keeps, per dimension (y = x - ref, ref = the chain's first streamed draw):
//     n, S1 = sum y_i, P_t = sum_i y_i y_{i+t} (t < T), the first T and the last T values.
// From them  mu = ref + S1/n,  s^2 = (P_0 - n mu_y^2)/(n-1)  and the BIASED autocovariance
//     acov(t) = [P_t - mu_y (A_t + B_t) + (n - t) mu_y^2] / n,
//     A_t = S1 - (sum of the last t values),  B_t = S1 - (sum of the first t values)
// -- exactly (1/n) sum_i (x_i - mu)(x_{i+t} - mu) of summary.hpp:55-73.  The cross-chain
// part is two small reductions (the payload of an NCCL all-reduce when chains are
// sharded over GPUs): phase 1 {sum_k mu_k, sum_k n_k mu_k}[D] + {K, N, min_len}; phase 2,
// centred on the global means, {sum_k (mu_k - mbar)^2, sum_k s_k^2, pooled SS,
// sum_k acov_k(t) for t < T}[D].  The Geyer loop (:700-748) then runs on the combined
// sums.  T bounds the lags it may read; a dimension whose sequence is still positive at
// lag T is flagged `truncated` (its ESS is then an upper bound).
#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.cuh"
#include "stream.cuh"

namespace wb200 {

constexpr int kStreamTB = 128;  // dimensions per block

// Fold `rows` new draws of every chain (staging rows row0 .. row0+rows) into its sums.
// One thread owns one (chain, dimension) series: its T lag accumulators AND the window of
// its last T values stay in registers.  The window is held in order of recency
// (w[k] = the value k + 1 positions back), so every index below is a compile-time
// constant: rows are processed in groups of kStreamGroup whose members meet each other as
// ys[r - t] and older values as w[t - r - 1]; the window then shifts by the group size.
// (The first version kept the window in shared memory under its absolute position mod T:
// 31 shared loads per row and thread made the fold shared-memory bound -- 2.8 ms for 50
// rows of 4096 x 1000 series against 0.35 ms of fp64 FMAs.)  Accumulation order per lag is
// the row order, as before: bitwise the same sums.  `tail` keeps its layout in memory
// (slot = position mod T), which the summary kernels read.
constexpr int kStreamGroup = 8;

template <int T, int G>
__device__ __forceinline__ void stream_fold_rows(const double (&xs)[G], double r,
                                                 double (&acc)[T], double (&w)[T], double& s1,
                                                 long long i0, double* head_cd, long long hstride) {
  double ys[G];
#pragma unroll
  for (int q = 0; q < G; ++q) {
    const double y = xs[q] - r;
    ys[q] = y;
    s1 += y;
    acc[0] += y * y;
#pragma unroll
    for (int t = 1; t < T; ++t) {
      // (lags that reach back before the start of the chain are skipped, not multiplied
      // by the window's zeros: a non-finite draw must not leak into them)
      const double partner = t <= q ? ys[q - t] : w[t - q - 1];
      if (t <= i0 + q) acc[t] += y * partner;
    }
    if (i0 + q < T) head_cd[(i0 + q) * hstride] = y;
  }
  // shift by G: the group's values become the most recent ones
#pragma unroll
  for (int k = T - 1; k >= 0; --k) {
    w[k] = k < G ? ys[G - 1 - k] : w[k - G];
  }
}

template <int T>
__global__ void __launch_bounds__(kStreamTB)
stream_update_kernel(const double* draws, long long draw_cap, int ld, const long long* rows_c,
                     long long rows_uniform, long long row0, const long long* n_c, double* ref,
                     double* S1, double* P, double* head, double* tail) {
  const int tx = threadIdx.x;
  const int d = blockIdx.y * kStreamTB + tx;
  const int c = blockIdx.x;  // chains on the x dimension of the grid: no 65 535 limit
  if (d >= ld) return;
  const long long B = rows_c ? rows_c[c] : rows_uniform;
  if (B <= 0) return;
  const long long n0 = n_c[c];
  const long long cd = static_cast<long long>(c) * ld + d;
  const double* x = draws + (static_cast<long long>(c) * draw_cap + row0) * ld + d;
  double r;
  if (n0 == 0) {
    r = x[0];
    ref[cd] = r;
  } else {
    r = ref[cd];
  }
  double acc[T], w[T];
  const long long lanes = static_cast<long long>(c) * T * ld + d;  // + t * ld
#pragma unroll
  for (int t = 0; t < T; ++t) {
    acc[t] = n0 > 0 ? P[lanes + static_cast<long long>(t) * ld] : 0.0;
    // w[t] = the value at position n0 - 1 - t (slot (n0 - 1 - t) mod T of `tail`)
    const long long pos = n0 - 1 - t;
    w[t] = pos >= 0 ? tail[lanes + (pos & (T - 1)) * ld] : 0.0;
  }
  double s1 = n0 > 0 ? S1[cd] : 0.0;
  // the next group's rows are in flight while this group's products are formed
  long long j = 0;
  double cur[kStreamGroup], nxt[kStreamGroup];
  if (kStreamGroup <= B) {
#pragma unroll
    for (int q = 0; q < kStreamGroup; ++q) cur[q] = x[q * static_cast<long long>(ld)];
  }
  for (; j + kStreamGroup <= B; j += kStreamGroup) {
    if (j + 2 * kStreamGroup <= B) {
#pragma unroll
      for (int q = 0; q < kStreamGroup; ++q) {
        nxt[q] = x[(j + kStreamGroup + q) * static_cast<long long>(ld)];
      }
    }
    stream_fold_rows<T, kStreamGroup>(cur, r, acc, w, s1, n0 + j, head + lanes, ld);
#pragma unroll
    for (int q = 0; q < kStreamGroup; ++q) cur[q] = nxt[q];
  }
  for (; j < B; ++j) {
    const double one[1] = {x[j * static_cast<long long>(ld)]};
    stream_fold_rows<T, 1>(one, r, acc, w, s1, n0 + j, head + lanes, ld);
  }
  const long long n1 = n0 + B;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    P[lanes + static_cast<long long>(t) * ld] = acc[t];
    const long long pos = n1 - 1 - t;  // the T most recent positions, each to its slot
    if (pos >= 0) tail[lanes + (pos & (T - 1)) * ld] = w[t];
  }
  S1[cd] = s1;
}

__global__ void stream_advance_kernel(long long* n_c, const long long* rows_c,
                                      long long rows_uniform, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) n_c[c] += rows_c ? rows_c[c] : rows_uniform;
}

// per (chain, dimension): mean and unbiased variance (summary.hpp:676-684)
__global__ void stream_chain_stats_kernel(const long long* n_c, int ld, int D, int T,
                                          const double* ref, const double* S1,
                                          const double* P, double* mu, double* s2) {
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  const int c = blockIdx.x;
  if (d >= D) return;
  const long long n = n_c[c];
  if (n < 3) return;  // left out, as summary.hpp:595-603 would reject it
  const long long cd = static_cast<long long>(c) * ld + d;
  const double nn = static_cast<double>(n);
  const double my = S1[cd] / nn;
  mu[static_cast<long long>(c) * D + d] = ref[cd] + my;
  s2[static_cast<long long>(c) * D + d] =
      (P[(static_cast<long long>(c) * T) * ld + d] - nn * my * my) / (nn - 1.0);
}

// Cross-chain sums per dimension, in two deterministic stages: blocks of 32 dimensions x 8
// chain lanes sum a chunk of the chains each (lanes combined in lane order through shared
// memory) into part[chunk][v][D]; stream_combine_kernel adds the chunks in chunk order.
// (One thread per dimension walking all chains took 1.9 + 2.4 ms at c2 and most of the
// 30 ms summary phase at c3, where D = 100 leaves 100 threads for 16 384 chains.)
constexpr int kSumDims = 32, kSumLanes = 8;

template <int NV>
__device__ __forceinline__ void stream_block_sums(double (&v)[NV], int D, int d, double* part) {
  __shared__ double sh[kSumLanes][NV][kSumDims];
#pragma unroll
  for (int k = 0; k < NV; ++k) sh[threadIdx.y][k][threadIdx.x] = v[k];
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double a = sh[0][k][threadIdx.x];
#pragma unroll
      for (int l = 1; l < kSumLanes; ++l) a += sh[l][k][threadIdx.x];
      part[(static_cast<long long>(blockIdx.y) * NV + k) * D + d] = a;
    }
  }
}

__global__ void stream_combine_kernel(int chunks, int total, const double* part, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // total = NV * D
  if (i >= total) return;
  double a = 0.0;
  for (int b = 0; b < chunks; ++b) a += part[static_cast<long long>(b) * total + i];
  out[i] = a;
}

// phase 1: sums of mu_k and n_k mu_k over the chains with n >= 3
__global__ void __launch_bounds__(kSumDims * kSumLanes)
stream_phase1_kernel(const long long* n_c, int C, int D, int chains_per_chunk,
                     const double* mu, double* part) {
  const int d = blockIdx.x * kSumDims + threadIdx.x;
  const int c0 = blockIdx.y * chains_per_chunk;
  const int c1 = min(C, c0 + chains_per_chunk);
  double v[2] = {0.0, 0.0};
  if (d < D) {
    for (int c = c0 + threadIdx.y; c < c1; c += kSumLanes) {
      const long long n = n_c[c];
      if (n < 3) continue;
      const double m = mu[static_cast<long long>(c) * D + d];
      v[0] += m;
      v[1] += static_cast<double>(n) * m;
    }
  }
  stream_block_sums<2>(v, D, d, part);
}

// phase 2, centred: sum_k (mu_k - mbar)^2, sum_k s2_k,
// sum_k [(n_k - 1) s2_k + n_k (mu_k - pm)^2]
__global__ void __launch_bounds__(kSumDims * kSumLanes)
stream_phase2_kernel(const long long* n_c, int C, int D, int chains_per_chunk,
                     const double* mu, const double* s2, const double* mbar,
                     const double* pm, double* part) {
  const int d = blockIdx.x * kSumDims + threadIdx.x;
  const int c0 = blockIdx.y * chains_per_chunk;
  const int c1 = min(C, c0 + chains_per_chunk);
  double v[3] = {0.0, 0.0, 0.0};
  if (d < D) {
    const double mb = mbar[d], p = pm[d];
    for (int c = c0 + threadIdx.y; c < c1; c += kSumLanes) {
      const long long n = n_c[c];
      if (n < 3) continue;
      const double nn = static_cast<double>(n);
      const double m = mu[static_cast<long long>(c) * D + d];
      const double var = s2[static_cast<long long>(c) * D + d];
      v[0] += (m - mb) * (m - mb);
      v[1] += var;
      v[2] += (nn - 1.0) * var + nn * (m - p) * (m - p);
    }
  }
  stream_block_sums<3>(v, D, d, part);
}

// macov[t * D + d] += sum over this block's chains of acov_k(t)
template <int T>
__global__ void __launch_bounds__(kStreamTB)
stream_acov_kernel(const long long* n_c, int C, int chains_per_block, int ld, int D,
                   const double* S1, const double* P, const double* head, const double* tail,
                   double* macov) {
  const int d = blockIdx.y * kStreamTB + threadIdx.x;
  if (d >= D) return;
  double acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.0;
  const int c0 = blockIdx.x * chains_per_block;
  const int c1 = min(C, c0 + chains_per_block);
  for (int c = c0; c < c1; ++c) {
    const long long n = n_c[c];
    if (n < 3) continue;
    const double nn = static_cast<double>(n);
    const long long cd = static_cast<long long>(c) * ld + d;
    const double s1 = S1[cd];
    const double my = s1 / nn;
    double hsum = 0.0, lsum = 0.0;  // sums of the first / last t values
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (t < n) {
        const long long o = (static_cast<long long>(c) * T + t) * ld + d;
        const double a_t = s1 - lsum, b_t = s1 - hsum;
        acc[t] += (P[o] - my * (a_t + b_t) + (nn - t) * my * my) / nn;
        hsum += head[o];
        lsum += tail[(static_cast<long long>(c) * T + ((n - 1 - t) & (T - 1))) * ld + d];
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) atomicAdd(macov + static_cast<long long>(t) * D + d, acc[t]);
}

// W = mean_k s2, B = var_{K-1}(mu_k), pooled mean and variance from the reduced sums
__global__ void stream_across_kernel(int D, double K, double N, const double* r1,
                                     const double* r2, double* W, double* B, double* pm,
                                     double* pv) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  W[d] = r2[D + d] / K;
  B[d] = r2[d] / (K - 1.0);  // NaN for one chain, like the reference
  pm[d] = r1[D + d] / N;
  pv[d] = r2[2 * D + d] / (N - 1.0);
}

// Geyer estimator on the combined sums (summary.hpp:700-748); lags >= nlag are not
// available: the sequence is cut there and the dimension flagged.
__global__ void stream_geyer_kernel(int D, int K, long long n_total, long long min_len,
                                    const double* macov, int nlag, const double* W,
                                    const double* B, const double* pvar, double* rho_buf,
                                    double* rhat, double* ess, double* mcse, int* truncated) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const double w_d = W[d];
  const double vp_d = K > 1 ? w_d + B[d] : w_d;
  rhat[d] = sqrt(1.0 + B[d] / w_d);
  double* rho_hat_t = rho_buf + d;  // stride D, nlag + 4 entries (zeroed)
  const long long S = D;
  auto acov_at = [&](long long t) { return macov[t * D + d] / static_cast<double>(K); };
  double rho_hat_even = 1.0;
  rho_hat_t[0] = rho_hat_even;
  double rho_hat_odd = 1.0 - (w_d - acov_at(1)) / vp_d;
  rho_hat_t[1 * S] = rho_hat_odd;
  long long t = 1;
  int cut = 0;
  while (t < min_len - 4 && (rho_hat_even + rho_hat_odd) > 0.0) {
    if (t + 2 >= nlag) { cut = 1; break; }
    rho_hat_even = 1.0 - (w_d - acov_at(t + 1)) / vp_d;
    rho_hat_odd = 1.0 - (w_d - acov_at(t + 2)) / vp_d;
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
  truncated[d] = cut;
  const long long max_t = t;
  if (rho_hat_even > 0.0) rho_hat_t[(max_t + 1) * S] = rho_hat_even;
  double head = 0.0;
  for (long long s = 0; s < max_t; ++s) head += rho_hat_t[s * S];
  double tau_hat = -1.0 + 2.0 * head + rho_hat_t[(max_t + 1) * S];
  tau_hat = fmax(tau_hat, 1.0 / log10(static_cast<double>(n_total)));
  const double e = static_cast<double>(n_total) / tau_hat;
  ess[d] = e;
  mcse[d] = sqrt(pvar[d]) / sqrt(e);
}

// ---------------------------------------------------------------------------
#define WB200_STREAM_T(T_, MACRO)                                              \
  do {                                                                         \
    switch (T_) {                                                              \
      case 8: MACRO(8); break;                                                 \
      case 16: MACRO(16); break;                                               \
      case 32: MACRO(32); break;                                               \
      default: throw std::invalid_argument("max_lags must be 8, 16 or 32");    \
    }                                                                          \
  } while (0)

void stream_begin(wb200_session& s, int max_lags) {
  if (max_lags != 8 && max_lags != 16 && max_lags != 32) {
    throw std::invalid_argument("max_lags must be 8, 16 or 32");
  }
  if (s.draw_cap <= 0) {
    throw std::runtime_error("reserve the staging block of draws first "
                             "(wb200_session_reserve_draws)");
  }
  if (!s.acc) s.acc = new StreamState();
  StreamState& st = *s.acc;
  st.T = max_lags;
  const size_t CL = static_cast<size_t>(s.C) * s.ld;
  st.ref.alloc(CL); st.S1.alloc(CL);
  st.P.alloc(CL * st.T); st.head.alloc(CL * st.T); st.tail.alloc(CL * st.T);
  st.n.alloc(s.C); st.rows.alloc(s.C);
  WB200_CUDA(cudaMemsetAsync(st.n.ptr, 0, s.C * sizeof(long long), s.stream));
  WB200_CUDA(cudaMemsetAsync(st.head.ptr, 0, CL * st.T * 8, s.stream));
  WB200_CUDA(cudaMemsetAsync(st.tail.ptr, 0, CL * st.T * 8, s.stream));
  st.blocks = 0;
  s.rows_written = 0;
}

void stream_end(wb200_session& s) {
  delete s.acc;
  s.acc = nullptr;
}

// rows_c: device per-chain counts of staged rows (null: `rows_uniform` for every chain)
void stream_update(wb200_session& s, const long long* rows_c, long long rows_uniform) {
  StreamState& st = *s.acc;
  const dim3 grid(s.C, (s.ld + kStreamTB - 1) / kStreamTB);
#define WB200_STREAM_UPDATE(T_)                                                          \
  stream_update_kernel<T_><<<grid, kStreamTB, 0, s.stream>>>(                   \
      s.draws.ptr, s.draw_cap, s.ld, rows_c, rows_uniform, 0, st.n.ptr, st.ref.ptr,      \
      st.S1.ptr, st.P.ptr, st.head.ptr, st.tail.ptr)
  WB200_STREAM_T(st.T, WB200_STREAM_UPDATE);
  WB200_CUDA(cudaGetLastError());
  stream_advance_kernel<<<(s.C + 255) / 256, 256, 0, s.stream>>>(
      st.n.ptr, rows_c, rows_uniform, s.C);
  WB200_CUDA(cudaGetLastError());
  s.launches += 2;
  st.blocks += 1;
}

// fold the rows staged by iteration-quota launches (the same count for every chain)
void stream_flush(wb200_session& s) {
  if (s.rows_written > 0) stream_update(s, nullptr, s.rows_written);
  s.rows_written = 0;
}

// host copy of the per-chain counts -> {K, N, min_len} over chains with >= 3 draws
static void stream_counts(wb200_session& s, double* K, double* N, double* min_len) {
  StreamState& st = *s.acc;
  std::vector<long long> n(s.C);
  WB200_CUDA(cudaMemcpyAsync(n.data(), st.n.ptr, s.C * sizeof(long long),
                             cudaMemcpyDeviceToHost, s.stream));
  WB200_CUDA(cudaStreamSynchronize(s.stream));
  long long k = 0, tot = 0, mn = -1;
  for (long long v : n) {
    if (v < 3) continue;
    ++k;
    tot += v;
    mn = mn < 0 ? v : std::min(mn, v);
  }
  *K = static_cast<double>(k);
  *N = static_cast<double>(tot);
  *min_len = mn < 0 ? 0.0 : static_cast<double>(mn);
}

static void stream_chain_stats(wb200_session& s) {
  StreamState& st = *s.acc;
  const size_t CD = static_cast<size_t>(s.C) * s.D;
  if (st.mu.count != CD) { st.mu.alloc(CD); st.s2.alloc(CD); }
  const dim3 grid(s.C, (s.D + 127) / 128);
  stream_chain_stats_kernel<<<grid, 128, 0, s.stream>>>(
      st.n.ptr, s.ld, s.D, st.T, st.ref.ptr, st.S1.ptr, st.P.ptr, st.mu.ptr, st.s2.ptr);
  WB200_CUDA(cudaGetLastError());
  s.launches += 1;
}

void stream_phase1(wb200_session& s, double* out_host) {
  StreamState& st = *s.acc;
  const int D = s.D;
  stream_flush(s);
  stream_chain_stats(s);
  DeviceBuffer<double> out, part;
  out.alloc(2 * static_cast<size_t>(D));
  const int chunks = std::max(1, std::min(128, (s.C + 63) / 64));
  const int per_chunk = (s.C + chunks - 1) / chunks;
  part.alloc(static_cast<size_t>(chunks) * 2 * D);
  stream_phase1_kernel<<<dim3((D + kSumDims - 1) / kSumDims, chunks),
                         dim3(kSumDims, kSumLanes), 0, s.stream>>>(
      st.n.ptr, s.C, D, per_chunk, st.mu.ptr, part.ptr);
  WB200_CUDA(cudaGetLastError());
  stream_combine_kernel<<<(2 * D + 127) / 128, 128, 0, s.stream>>>(chunks, 2 * D, part.ptr,
                                                                    out.ptr);
  WB200_CUDA(cudaGetLastError());
  WB200_CUDA(cudaMemcpyAsync(out_host, out.ptr, 2 * D * 8, cudaMemcpyDeviceToHost,
                             s.stream));
  stream_counts(s, out_host + 2 * D, out_host + 2 * D + 1, out_host + 2 * D + 2);
  s.launches += 1;
}

void stream_phase2(wb200_session& s, const double* reduced1, double* out_host) {
  StreamState& st = *s.acc;
  const int D = s.D, T = st.T;
  const double K = reduced1[2 * D], N = reduced1[2 * D + 1];
  if (!(K >= 1) || !(N >= 3)) throw std::invalid_argument("chains must have at least 3 draws");
  std::vector<double> centre(2 * static_cast<size_t>(D));
  for (int d = 0; d < D; ++d) {
    centre[d] = reduced1[d] / K;        // mean of the chain means
    centre[D + d] = reduced1[D + d] / N;  // pooled mean
  }
  DeviceBuffer<double> cen, out;
  cen.alloc(centre.size());
  out.alloc(static_cast<size_t>(3 + T) * D);
  WB200_CUDA(cudaMemcpyAsync(cen.ptr, centre.data(), centre.size() * 8, cudaMemcpyHostToDevice,
                             s.stream));
  WB200_CUDA(cudaMemsetAsync(out.ptr, 0, out.count * 8, s.stream));
  if (st.mu.count == 0) stream_chain_stats(s);
  DeviceBuffer<double> part;
  const int chunks = std::max(1, std::min(128, (s.C + 63) / 64));
  const int per_chunk = (s.C + chunks - 1) / chunks;
  part.alloc(static_cast<size_t>(chunks) * 3 * D);
  stream_phase2_kernel<<<dim3((D + kSumDims - 1) / kSumDims, chunks),
                         dim3(kSumDims, kSumLanes), 0, s.stream>>>(
      st.n.ptr, s.C, D, per_chunk, st.mu.ptr, st.s2.ptr, cen.ptr, cen.ptr + D, part.ptr);
  WB200_CUDA(cudaGetLastError());
  stream_combine_kernel<<<(3 * D + 127) / 128, 128, 0, s.stream>>>(chunks, 3 * D, part.ptr,
                                                                    out.ptr);
  WB200_CUDA(cudaGetLastError());
  const int cpb = 32;
  const dim3 grid((s.C + cpb - 1) / cpb, (D + kStreamTB - 1) / kStreamTB);
#define WB200_STREAM_ACOV(T_)                                                            \
  stream_acov_kernel<T_><<<grid, kStreamTB, 0, s.stream>>>(                     \
      st.n.ptr, s.C, cpb, s.ld, D, st.S1.ptr, st.P.ptr, st.head.ptr, st.tail.ptr,        \
      out.ptr + 3 * static_cast<size_t>(D))
  WB200_STREAM_T(T, WB200_STREAM_ACOV);
  WB200_CUDA(cudaGetLastError());
  WB200_CUDA(cudaMemcpyAsync(out_host, out.ptr, out.count * 8, cudaMemcpyDeviceToHost,
                             s.stream));
  WB200_CUDA(cudaStreamSynchronize(s.stream));
  s.launches += 2;
}

void stream_finish(int D, int T, const double* reduced1, const double* reduced2, double* rhat,
                   double* ess, double* mcse, double* mean, double* var, int* truncated,
                   cudaStream_t stream) {
  const double K = reduced1[2 * D], N = reduced1[2 * D + 1], min_len = reduced1[2 * D + 2];
  if (rhat && K < 2) {  // summary.hpp:595-603
    throw std::invalid_argument("require at least two chains to compute R-hat");
  }
  if (N < 3) throw std::invalid_argument("chains must have at least 3 draws");
  DeviceBuffer<double> r1, r2, W, B, pm, pv, rho, o_rhat, o_ess, o_mcse;
  DeviceBuffer<int> cut;
  r1.alloc(2 * static_cast<size_t>(D));
  r2.alloc(static_cast<size_t>(3 + T) * D);
  W.alloc(D); B.alloc(D); pm.alloc(D); pv.alloc(D);
  o_rhat.alloc(D); o_ess.alloc(D); o_mcse.alloc(D); cut.alloc(D);
  rho.alloc(static_cast<size_t>(T + 4) * D);
  WB200_CUDA(cudaMemcpyAsync(r1.ptr, reduced1, r1.count * 8, cudaMemcpyHostToDevice, stream));
  WB200_CUDA(cudaMemcpyAsync(r2.ptr, reduced2, r2.count * 8, cudaMemcpyHostToDevice, stream));
  WB200_CUDA(cudaMemsetAsync(rho.ptr, 0, rho.count * 8, stream));
  stream_across_kernel<<<(D + 127) / 128, 128, 0, stream>>>(D, K, N, r1.ptr, r2.ptr, W.ptr,
                                                             B.ptr, pm.ptr, pv.ptr);
  WB200_CUDA(cudaGetLastError());
  stream_geyer_kernel<<<(D + 127) / 128, 128, 0, stream>>>(
      D, static_cast<int>(K), static_cast<long long>(N), static_cast<long long>(min_len),
      r2.ptr + 3 * static_cast<size_t>(D), T, W.ptr, B.ptr, pv.ptr, rho.ptr, o_rhat.ptr,
      o_ess.ptr, o_mcse.ptr, cut.ptr);
  WB200_CUDA(cudaGetLastError());
  auto fetch = [&](double* dst, const double* src) {
    if (dst) WB200_CUDA(cudaMemcpyAsync(dst, src, D * 8, cudaMemcpyDeviceToHost, stream));
  };
  fetch(rhat, o_rhat.ptr); fetch(ess, o_ess.ptr); fetch(mcse, o_mcse.ptr);
  fetch(mean, pm.ptr); fetch(var, pv.ptr);
  if (truncated) {
    WB200_CUDA(cudaMemcpyAsync(truncated, cut.ptr, D * 4, cudaMemcpyDeviceToHost, stream));
  }
  WB200_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace wb200

long long* wb200_session::acc_rows() { return acc ? acc->rows.ptr : nullptr; }

using namespace wb200;

extern "C" {

int wb200_session_stream_begin(wb200_session* s, int max_lags, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    stream_begin(*s, max_lags);
  });
}

int wb200_session_stream_phase1(wb200_session* s, double* out, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->acc) throw std::runtime_error("streaming summaries have not been started");
    stream_phase1(*s, out);
  });
}

int wb200_session_stream_phase2(wb200_session* s, const double* reduced1, double* out,
                                WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->acc) throw std::runtime_error("streaming summaries have not been started");
    stream_phase2(*s, reduced1, out);
  });
}

int wb200_stream_finish(int num_params, int max_lags, const double* reduced1,
                        const double* reduced2, double* rhat, double* ess, double* mcse,
                        double* mean, double* var, int* truncated, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    require_gpu();
    stream_finish(num_params, max_lags, reduced1, reduced2, rhat, ess, mcse, mean, var,
                  truncated, nullptr);
  });
}

int wb200_session_stream_summary(wb200_session* s, double* rhat, double* ess, double* mcse,
                                 double* mean, double* var, int* truncated,
                                 WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->acc) throw std::runtime_error("streaming summaries have not been started");
    const int D = s->D, T = s->acc->T;
    std::vector<double> r1(2 * static_cast<size_t>(D) + 3), r2(static_cast<size_t>(3 + T) * D);
    stream_phase1(*s, r1.data());
    stream_phase2(*s, r1.data(), r2.data());
    stream_finish(D, T, r1.data(), r2.data(), rhat, ess, mcse, mean, var, truncated,
                  s->stream);
  });
}

int wb200_session_stream_counts(wb200_session* s, long long* counts, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->acc) throw std::runtime_error("streaming summaries have not been started");
    stream_flush(*s);
    WB200_CUDA(cudaMemcpyAsync(counts, s->acc->n.ptr, s->C * sizeof(long long),
                               cudaMemcpyDeviceToHost, s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
  });
}

}  // extern "C"
