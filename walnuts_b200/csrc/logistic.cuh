// Batched logistic-regression log density / gradient on tcgen05 (see logistic.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace wb200 {

class LogisticGrad {
 public:
  // X host fp64 [N][D] row-major, y host fp64 [N] in {0,1}; C chains with row stride ld
  LogisticGrad(const double* X, const double* y, size_t N, int D, int C, int ld,
               cudaStream_t stream);
  // a second batch of C chains over the SAME data (X, X^T, X^T y stay shared on the device)
  LogisticGrad(const LogisticGrad& data_of, int C, cudaStream_t stream);
  ~LogisticGrad();
  LogisticGrad(const LogisticGrad&) = delete;
  LogisticGrad& operator=(const LogisticGrad&) = delete;
  // TH, G device fp64 [C][ld]; LP device fp64 [C]
  void evaluate(const double* TH, double* G, double* LP, cudaStream_t stream);
  int kernels_per_eval() const;
  double flops_per_eval() const;  // executed tensor-core flops of one batched evaluation
  struct Impl;

 private:
  Impl* impl_;
};

}  // namespace wb200
