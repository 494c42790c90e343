// Free-running fp32 launches of the chain-resident kernel: a translation unit of its own so that
// the instantiations compile in parallel (engine_shapes.cuh).
#include "engine_shapes.cuh"

namespace wb200 {

#define WB200_LAUNCH_CHAIN_FREE_F32(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)    \
  do {                                                                         \
    if (p.adapt) {                                                             \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_A_, true, float, true> \
          <<<s.grid_adapt, CTA_, dyn_smem, s.stream>>>(p);                     \
    } else {                                                                   \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_S_, false, float, true> \
          <<<s.grid, CTA_, dyn_smem, s.stream>>>(p);                           \
    }                                                                          \
  } while (0)

void launch_chain_free_f32(wb200_session& s, const ChainParams& p, size_t dyn_smem) {
  WB200_FOR_TARGET_F32(s.kind, s.shape, WB200_LAUNCH_CHAIN_FREE_F32);
}

}  // namespace wb200
