// fp32 quota launches and occupancy of the chain-resident kernel: a translation unit of its own so that
// the instantiations compile in parallel (engine_shapes.cuh).
#include "engine_shapes.cuh"

namespace wb200 {

#define WB200_OCC_F32(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)                  \
  do {                                                                         \
    *adapt = blocks_per_sm(                                                    \
        walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_A_, true, float>, CTA_, \
        dyn_smem);                                                             \
    *sample = blocks_per_sm(                                                   \
        walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_S_, false, float>, CTA_, \
        dyn_smem);                                                             \
  } while (0)
#define WB200_LAUNCH_CHAIN_F32(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)         \
  do {                                                                         \
    if (p.adapt) {                                                             \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_A_, true, float> \
          <<<s.grid_adapt, CTA_, dyn_smem, s.stream>>>(p);                     \
    } else {                                                                   \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_S_, false, float> \
          <<<s.grid, CTA_, dyn_smem, s.stream>>>(p);                           \
    }                                                                          \
  } while (0)

void occupancy_f32(int kind, const LaunchShape& shape, size_t dyn_smem, int* adapt,
                   int* sample) {
  WB200_FOR_TARGET_F32(kind, shape, WB200_OCC_F32);
}

void launch_chain_f32(wb200_session& s, const ChainParams& p, size_t dyn_smem) {
  WB200_FOR_TARGET_F32(s.kind, s.shape, WB200_LAUNCH_CHAIN_F32);
}

}  // namespace wb200
