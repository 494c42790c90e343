// Streaming summary accumulators of a session (stream.cu).
#pragma once
#include "engine.cuh"

namespace wb200 {

struct StreamState {
  int T = 0;                                   // lags kept (0 .. T-1)
  DeviceBuffer<double> ref, S1;                // [C][ld]
  DeviceBuffer<double> P, head, tail;          // [C][T][ld]
  DeviceBuffer<long long> n, rows;             // [C] draws folded in; staged rows (ragged)
  DeviceBuffer<double> mu, s2;                 // [C][D] per-chain mean / variance (summary)
  unsigned long long blocks = 0;
};

void stream_begin(wb200_session& s, int max_lags);
void stream_end(wb200_session& s);
void stream_update(wb200_session& s, const long long* rows_c, long long rows_uniform);
void stream_flush(wb200_session& s);
void stream_phase1(wb200_session& s, double* out_host);
void stream_phase2(wb200_session& s, const double* reduced1, double* out_host);
void stream_finish(int D, int T, const double* reduced1, const double* reduced2, double* rhat,
                   double* ess, double* mcse, double* mean, double* var, int* truncated,
                   cudaStream_t stream);

}  // namespace wb200
