// Launch shapes of the chain-resident kernel, shared by the translation units that
// instantiate it (engine.cu: fp64 quota launches, init, orbit; engine_f32.cu: fp32 quota
// launches; engine_free.cu / engine_free_f32.cu: the free-running kernels) -- split so that
// the ~200 instantiations compile in parallel.
#pragma once
#include "engine.cuh"

namespace wb200 {

// (TARGET, T, K, CTA, min resident CTAs per SM -> register cap)
// (TARGET, T, K, CTA, resident CTAs per SM asked of the adaptive / the sampling instance ->
// register cap).  The sampling instance of the wide shapes fits 128 registers without
// spilling since the start state moved to shared memory; the adaptive one needs 168.
#ifndef WB200_MINB_32X2
#define WB200_MINB_32X2 4
#endif
#ifndef WB200_MINB_128X4
#define WB200_MINB_128X4 4
#endif
#ifndef WB200_MINB_128X4_ADAPT
#define WB200_MINB_128X4_ADAPT 3
#endif
#define WB200_FOR_SHAPE(S, MACRO, TARGET)                                      \
  do {                                                                         \
    if ((S).T == 32 && (S).K == 1) { MACRO(TARGET, 32, 1, 128, 4, 4); }        \
    else if ((S).T == 32 && (S).K == 2) { MACRO(TARGET, 32, 2, 128, WB200_MINB_32X2, WB200_MINB_32X2); } \
    else if ((S).T == 64 && (S).K == 2) { MACRO(TARGET, 64, 2, 64, 6, 6); }    \
    else if ((S).T == 64) { MACRO(TARGET, 64, 8, 64, 4, 4); }                  \
    else if ((S).T == 128 && (S).K == 2) { MACRO(TARGET, 128, 2, 128, 3, 3); } \
    else if ((S).T == 128 && (S).K == 4) { MACRO(TARGET, 128, 4, 128, WB200_MINB_128X4_ADAPT, WB200_MINB_128X4); } \
    else if ((S).T == 256 && (S).K == 2) { MACRO(TARGET, 256, 2, 256, 2, 2); } \
    else if ((S).T == 256 && (S).K == 4) { MACRO(TARGET, 256, 4, 256, 1, 1); } \
    else { MACRO(TARGET, 512, 4, 512, 1, 1); }                                 \
  } while (0)

// fp32 mode: half the register footprint per element, so more resident CTAs
#ifndef WB200_MINB_128X4_F32
#define WB200_MINB_128X4_F32 5
#endif
#define WB200_FOR_SHAPE_F32(S, MACRO, TARGET)                                  \
  do {                                                                         \
    if ((S).T == 32 && (S).K == 1) { MACRO(TARGET, 32, 1, 128, 4, 4); }        \
    else if ((S).T == 32 && (S).K == 2) { MACRO(TARGET, 32, 2, 128, 4, 4); }   \
    else if ((S).T == 64 && (S).K == 2) { MACRO(TARGET, 64, 2, 64, 6, 8); }    \
    else if ((S).T == 64) { MACRO(TARGET, 64, 8, 64, 4, 6); }                  \
    else if ((S).T == 128 && (S).K == 2) { MACRO(TARGET, 128, 2, 128, 4, 5); } \
    else if ((S).T == 128 && (S).K == 4) { MACRO(TARGET, 128, 4, 128, 4, WB200_MINB_128X4_F32); } \
    else if ((S).T == 256 && (S).K == 2) { MACRO(TARGET, 256, 2, 256, 2, 2); } \
    else if ((S).T == 256 && (S).K == 4) { MACRO(TARGET, 256, 4, 256, 1, 2); } \
    else { MACRO(TARGET, 512, 4, 512, 1, 1); }                                 \
  } while (0)

#define WB200_FOR_TARGET_F32(KIND, S, MACRO)                                   \
  do {                                                                         \
    switch (KIND) {                                                            \
      case kStdNormal: WB200_FOR_SHAPE_F32(S, MACRO, StdNormalTargetF); break; \
      case kDiagGaussian: WB200_FOR_SHAPE_F32(S, MACRO, DiagGaussianTargetF); break; \
      case kFunnel: WB200_FOR_SHAPE_F32(S, MACRO, FunnelTargetF); break;       \
      default: throw std::invalid_argument("model kind has no chain-resident " \
                                           "kernel");                          \
    }                                                                          \
  } while (0)

#define WB200_FOR_TARGET(KIND, S, MACRO)                                       \
  do {                                                                         \
    switch (KIND) {                                                            \
      case kStdNormal: WB200_FOR_SHAPE(S, MACRO, StdNormalTarget); break;      \
      case kDiagGaussian: WB200_FOR_SHAPE(S, MACRO, DiagGaussianTarget); break;\
      case kFunnel: WB200_FOR_SHAPE(S, MACRO, FunnelTarget); break;            \
      default: throw std::invalid_argument("model kind has no chain-resident " \
                                           "kernel");                          \
    }                                                                          \
  } while (0)


// dynamic shared memory of a chain-kernel CTA: the parked start states and sub-tree
// stacks of its resident chains
inline size_t chain_dyn_smem(const LaunchShape& shape, int ld) {
  return static_cast<size_t>(shape.chains_per_cta) * chain_smem_doubles(ld) * sizeof(double);
}

// the chain kernel of a session's (kind, shape, precision), quota or free-running
void launch_chain_f32(wb200_session& s, const ChainParams& p, size_t dyn_smem);
void launch_chain_free(wb200_session& s, const ChainParams& p, size_t dyn_smem);
void launch_chain_free_f32(wb200_session& s, const ChainParams& p, size_t dyn_smem);
void occupancy_f32(int kind, const LaunchShape& shape, size_t dyn_smem, int* adapt, int* sample);

template <class Kernel>
inline int blocks_per_sm(Kernel k, int cta, size_t dyn_smem) {
  int n = 0;
  if (dyn_smem > 48 * 1024) {
    WB200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(dyn_smem)));
  }
  WB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, cta, dyn_smem));
  return n > 1 ? n : 1;
}

}  // namespace wb200
