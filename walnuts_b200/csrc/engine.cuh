// Host-side session object of the device sampler (declarations shared by
// engine.cu, summary.cu and capi.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/walnuts_b200.h"
#include "chain_kernel.cuh"

// errors.hpp:26-36
struct WalnutpyError {
  std::string msg;
  WalnutpyErrorType type;
};

namespace wb200 {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define WB200_CUDA(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      throw ::wb200::CudaError(std::string("CUDA error: ") +                    \
                               cudaGetErrorString(_e) + " at " + __FILE__ +     \
                               ":" + std::to_string(__LINE__));                 \
    }                                                                           \
  } while (0)

// Device memory comes from the device's stream-ordered pool with an unlimited
// release threshold: a session's buffers (GBs of stored draws) go back to the
// pool when it is destroyed and the next session reuses them, instead of a
// cudaMalloc/cudaFree pair per one-shot call (cudaFree of multi-GB buffers was
// measured at 50-500 ms).  wb200_trim_memory() returns the pool to the driver.
inline void* pool_alloc(size_t bytes) {
  static bool configured[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess && dev >= 0 && dev < 64 && !configured[dev]) {
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e == cudaSuccess) {
      unsigned long long keep = ~0ull;
      e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    configured[dev] = true;
  }
  void* ptr = nullptr;
  if (e == cudaSuccess) e = cudaMallocAsync(&ptr, bytes, cudaStream_t{0});
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStream_t{0});
  if (e != cudaSuccess) {
    throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(e) +
                    " allocating " + std::to_string(bytes) + " bytes of device memory");
  }
  return ptr;
}
inline void pool_free(void* ptr) {
  // like cudaFree: nothing in flight may still use the buffer
  cudaDeviceSynchronize();
  cudaFreeAsync(ptr, cudaStream_t{0});
}

template <class T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t count = 0;
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void alloc(size_t n) {
    release();
    if (n) ptr = static_cast<T*>(pool_alloc(n * sizeof(T)));
    count = n;
  }
  void release() {
    if (ptr) pool_free(ptr);
    ptr = nullptr;
    count = 0;
  }
};

struct LaunchShape {
  int T, K, cta, chains_per_cta;
};
LaunchShape shape_for_dim(int D);
void occupancy_for(int kind, const LaunchShape& shape, int ld, int precision, int* adapt,
                   int* sample);

// errors.hpp:30-33: the user pressed Ctrl+C (interrupts.hpp)
struct InterruptException {};

// python/src/walnutpie/errors.hpp:42-72: exceptions -> error object + rc
template <class F>
int catch_exceptions(WalnutpyError** err, F&& f) {
  try {
    f();
    return 0;
  } catch (const InterruptException&) {
    if (err) *err = new WalnutpyError{"", wb200_interrupt};
  } catch (const std::invalid_argument& e) {
    if (err) *err = new WalnutpyError{e.what(), wb200_config};
  } catch (const std::exception& e) {
    if (err) *err = new WalnutpyError{e.what(), wb200_generic};
  } catch (...) {
    if (err) *err = new WalnutpyError{"Unknown error", wb200_generic};
  }
  return -1;
}

inline void require_gpu() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    throw std::runtime_error(
        "walnuts_b200: no CUDA device is visible; this library has no CPU path");
  }
}

// host [rows][D] <-> device [rows][ld]
inline void upload_rows(double* dst, int ld, const double* src, int D, size_t rows,
                        cudaStream_t st) {
  WB200_CUDA(cudaMemcpy2DAsync(dst, ld * sizeof(double), src, D * sizeof(double),
                               D * sizeof(double), rows, cudaMemcpyHostToDevice, st));
}
inline void download_rows(double* dst, int D, const double* src, int ld, size_t rows,
                          cudaStream_t st) {
  WB200_CUDA(cudaMemcpy2DAsync(dst, D * sizeof(double), src, ld * sizeof(double),
                               D * sizeof(double), rows, cudaMemcpyDeviceToHost, st));
}

}  // namespace wb200

namespace wb200 { struct TickEngine; struct StreamState; struct UserModule; }

struct wb200_session {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tm0 = nullptr, tm1 = nullptr;
  int kind = 0, D = 0, ld = 0;
  int precision = 0;  // 0: fp64; 1: fp32 integrator state inside transitions (chain kernel)
  size_t N = 0;
  int C = 0;
  uint32_t seed = 0, chain_offset = 0;
  WalnutTuning tuning{};
  wb200::LaunchShape shape{};
  int slots = 0, grid = 0, grid_adapt = 0;  // sampling / adaptive instance grids
  bool frozen = false, initialised = false;
  long long draw_cap = 0, rows_written = 0;
  bool trace = false;
  bool ragged = false;  // a free-running phase has stored draws: per-chain row counts
  unsigned long long launches = 0;
  // kind 4: batched density evaluations that failed inside a transition and were
  // replaced by logp = -inf, grad = 0 (util.hpp:336-346); {tick, return code} of the first
  unsigned long long logp_exceptions = 0;
  std::vector<std::pair<unsigned long long, int>> exception_log;

  wb200::DeviceBuffer<double> theta, inv_mass, est, tparam, scratch, draws,
      lp_out, step_out, im_out, red, adam_tab, sums;
  wb200::DeviceBuffer<int> depth_out;
  wb200::DeviceBuffer<wb200::ChainScalars> sc;
  wb200::DeviceBuffer<unsigned int> ticket;
  wb200::DeviceBuffer<int> order;                      // [C] ticket -> chain (LPT order)
  wb200::DeviceBuffer<unsigned long long> prev_evals;  // [C] grad_evals before the last launch
  wb200::DeviceBuffer<long long> chain_rows;  // [C] draw rows per chain (free-running chain engine)
  wb200::DeviceBuffer<long long> iter_stats;  // {min, max, sum} of per-chain iteration counts, total evals
  wb200::TickEngine* tick = nullptr;  // lock-step engine (logistic; WB200_ENGINE=tick)
  wb200::StreamState* acc = nullptr;  // streaming summary accumulators (stream.cu)
  std::shared_ptr<wb200::UserModule> user;  // kind 5: the run-time compiled kernels

  wb200::ChainParams params(int n_iter, int adapt, bool store);
  long long* acc_rows();  // device [C]: staged rows per chain of the streaming block
};

namespace wb200 {
// eval_budget > 0: free-running launch (every chain completes the transitions that fit
// into that many gradient evaluations, at most n_iter, never beyond iter_cap of the phase;
// rows = per-chain draw row counters)
void launch_chains(wb200_session& s, int n_iter, int adapt, bool store,
                   long long eval_budget = 0, long long iter_cap = 0,
                   long long* rows = nullptr);
void chain_iter_stats(wb200_session& s, bool sampling, long long* out4_host);
void launch_init(wb200_session& s, bool have_mass, bool have_steps,
                 bool have_positions, double init_radius);
void launch_freeze(wb200_session& s);
// lock-step tick engine (tick_engine.cu)
void tick_create(wb200_session& s, const WalnutModelDesc& model);
void tick_destroy(wb200_session& s);
void tick_init(wb200_session& s, bool have_mass, bool have_steps, bool have_positions,
               double init_radius);
void tick_run(wb200_session& s, int n_iter, int adapt, bool store);
// iter_cap > 0: no chain goes beyond that many iterations of the phase in total
void tick_run_ticks(wb200_session& s, int n_ticks, int adapt, bool store,
                    long long iter_cap = 0);
void tick_chain_rows(wb200_session& s, long long* rows_host);
// free-running + streaming: per-chain staged row counts -> rows_dev, then reset to 0
void tick_take_rows(wb200_session& s, long long* rows_dev);
void stream_end(wb200_session& s);
void stream_update(wb200_session& s, const long long* rows_c, long long rows_uniform);
void stream_flush(wb200_session& s);
void tick_abort_inflight(wb200_session& s);
unsigned long long tick_count(const wb200_session& s);
// user_orbit: the orbit kernel of a run-time compiled density (kind 5), else null
void launch_orbit(int kind, int precision, int D, int ld, int C, const double* tparam,
                  double* theta, double* rho, const double* inv_mass, double* grad,
                  double* logp, double* joint, double step, int num_steps,
                  cudaStream_t stream, void* user_orbit = nullptr);
void device_rhat_moments(const double* draws, int ld, int D,
                         const std::vector<long long>& start,
                         const std::vector<long long>& len, double* moments_host,
                         cudaStream_t stream);
// summaries over ragged chains: chain c = rows start[c] .. start[c]+len[c] of a
// device matrix with row stride ld; outputs are HOST arrays of D (nullable)
void device_summary(const double* draws, int ld, int D,
                    const std::vector<long long>& start,
                    const std::vector<long long>& len, double* rhat, double* ess,
                    double* mcse, double* mean, double* var, cudaStream_t stream);
}  // namespace wb200
