// Free-running fp64 launches of the chain-resident kernel: a translation unit of its own so that
// the instantiations compile in parallel (engine_shapes.cuh).
#include "engine_shapes.cuh"

namespace wb200 {

#define WB200_LAUNCH_CHAIN_FREE(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)        \
  do {                                                                         \
    if (p.adapt) {                                                             \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_A_, true, double, true> \
          <<<s.grid_adapt, CTA_, dyn_smem, s.stream>>>(p);                     \
    } else {                                                                   \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_S_, false, double, true> \
          <<<s.grid, CTA_, dyn_smem, s.stream>>>(p);                           \
    }                                                                          \
  } while (0)

void launch_chain_free(wb200_session& s, const ChainParams& p, size_t dyn_smem) {
  WB200_FOR_TARGET(s.kind, s.shape, WB200_LAUNCH_CHAIN_FREE);
}

}  // namespace wb200
