"""The usage examples of README.md, run end to end (needs a GPU)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, walnuts_b200 as wb
fit = wb.walnuts_device(wb.models.diag_gaussian(np.linspace(0.5, 4.0, 100)),
                        num_chains=4096, seed=1, save_warmup=False)
print(wb.r_hat(fit[:16]).max(), wb.ess(fit[:16]).min())
import torch
fit = wb.walnuts_device(wb.models.torch_density(5, lambda th: -0.5 * (th ** 2).sum(dim=1)),
                        num_chains=64, seed=2)
src = '''__device__ void wb200_logp_grad(int d, double x, const double* par, double& lp, double& g) {
  const double z = x - par[d]; lp = -0.5 * z * z; g = -z; }'''
fit = wb.walnuts_device(wb.models.device_source(src, 100, params=np.linspace(-1.0, 1.0, 100)), num_chains=1024)
print(np.mean([np.asarray(f).mean(0) for f in fit], axis=0)[:5])
out = wb.walnuts_device_summary(wb.models.funnel(100), num_chains=16384, seed=4, devices=[0])
print(out["r_hat"].max(), out["sampling_iters"])
with wb.Session(wb.models.funnel(100), 16384, seed=3, max_step_halvings=8) as s:
    s.init(init_radius=1.0).reserve(200)
    s.warmup(300).freeze().sample(200).sync()
    print(s.summary(0, 200)["r_hat"].max())
print("readme ok")
