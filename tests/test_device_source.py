"""Model kind 5 (a density given as CUDA source, compiled at run time into the
chain-resident kernel), the part that needs no GPU: NVRTC compiles the source against the
kernel headers embedded in the library; errors surface as the reference's config errors
(errors.hpp:10-24 -> ValueError) carrying the compiler's log."""
import pytest

GAUSS = r'''
__device__ void wb200_logp_grad(int d, double x, const double* par, double& lp, double& g) {
  const double t = x * par[d];
  lp = -0.5 * x * t;
  g = -t;
}
'''


@pytest.mark.parametrize("D", [3, 100, 1000])   # one-warp, one-warp K=2, four-warp groups
def test_device_source_compiles_for_every_launch_shape_family(wb, D):
    assert "error" not in wb.models.compile_device_source(GAUSS, D)


def test_device_source_with_a_full_target(wb):
    src = r'''
    template <int T, int K, class Real>
    struct Shifted {
      double mu;
      __device__ void init(const wb200::ChainParams& p, int) { mu = p.tparam[0]; }
      __device__ void grad(const Real (&th)[K][2], Real (&g)[K][2], Real& lp_part,
                           wb200::Group<T>&) const {
        Real s = 0;
        for (int k = 0; k < K; ++k)
          for (int v = 0; v < 2; ++v) {
            const Real z = th[k][v] - static_cast<Real>(mu);
            s = wb200::madd(z, z, s);
            g[k][v] = -z;
          }
        lp_part = static_cast<Real>(-0.5) * s;
      }
    };
    #define WB200_USER_TARGET Shifted
    '''
    wb.models.compile_device_source(src, 64)


def test_device_source_errors_carry_the_compiler_log(wb):
    with pytest.raises(ValueError, match="does not compile") as e:
        wb.models.compile_device_source(
            "__device__ void wb200_logp_grad(int d, double x) { undefined_symbol; }", 10)
    assert "undefined_symbol" in str(e.value) and "device_source(1)" in str(e.value)
    with pytest.raises(ValueError, match="empty"):
        wb.models.compile_device_source("", 10)
    with pytest.raises(ValueError, match="num_params"):
        wb.models.compile_device_source(GAUSS, 0)
