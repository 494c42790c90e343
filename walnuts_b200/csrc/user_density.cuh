// Run-time compiled user densities (model kind 5, user_density.cu).
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "engine.cuh"

namespace wb200 {

// the kernels of one (source, launch shape, device): opaque driver-API handles
struct UserModule {
  void* mod = nullptr;
  void *adapt = nullptr, *sample = nullptr, *adapt_free = nullptr, *sample_free = nullptr;
  void *init = nullptr, *orbit = nullptr;
  std::string log;  // compiler output (warnings)
  ~UserModule();
};

// NVRTC: source -> sm_100a cubin of the six kernels for `shape`; throws
// std::invalid_argument with the compiler's log if the source does not compile
std::vector<char> user_compile(const char* source, const LaunchShape& shape,
                               std::vector<std::string>* lowered, std::string* log);
// compiled, loaded into `device`'s primary context and cached per (device, shape, source)
std::shared_ptr<UserModule> user_module(const char* source, const LaunchShape& shape,
                                        int device);
int user_blocks_per_sm(void* func, int cta, size_t dyn_smem);
// one kernel parameter (ChainParams / InitParams / OrbitParams), passed by value
void user_launch(void* func, int grid, int cta, size_t dyn_smem, cudaStream_t stream,
                 void* params);

}  // namespace wb200
