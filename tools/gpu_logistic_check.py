import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
import torch
import walnuts_b200 as wb
from walnuts_b200.sampler import logistic_logp_grad
from oracle.binding import load_oracle, Target
o = load_oracle()
def bf16_round(a):
    return torch.tensor(a, dtype=torch.float64).to(torch.bfloat16).to(torch.float64).numpy()
for (N, D, C) in [(256, 64, 128), (300, 16, 5), (1000, 200, 130), (4096, 512, 256)]:
    rng = np.random.default_rng(N)
    X = bf16_round(rng.normal(size=(N, D)))
    tstar = rng.normal(size=D) / np.sqrt(D)
    y = (rng.uniform(size=N) < 1/(1+np.exp(-X @ tstar))).astype(np.float64)
    theta = rng.normal(size=(C, D)) * 0.3
    lp, g, _ = logistic_logp_grad(X, y, theta)
    t = Target("logistic", D, X=X, y=y)
    worst_lp = worst_g = 0
    for c in range(min(C, 6)):
        olp, og = o.logp_grad(t, theta[c])
        worst_lp = max(worst_lp, abs(lp[c]-olp)/max(1, abs(olp)))
        worst_g = max(worst_g, np.max(np.abs(g[c]-og))/np.max(np.abs(og)))
    print(f"N={N} D={D} C={C}: rel err logp {worst_lp:.3e}  grad {worst_g:.3e}   lp[0]={lp[0]:.6f}", flush=True)
# timing at c4-like size (smaller chains)
N, D, C = 100000, 512, 2048
rng = np.random.default_rng(1)
X = bf16_round(rng.normal(size=(N, D))); y = (rng.uniform(size=N) < 0.5).astype(np.float64)
theta = rng.normal(size=(C, D)) * 0.05
lp, g, ms = logistic_logp_grad(X, y, theta, repeats=5)
fl = 4.0 * N * D * C
print(f"N={N} D={D} C={C}: {ms:.3f} ms per batched eval -> {fl/ms/1e9:.1f} algorithmic TFLOP/s, {C/ms*1e3:.3e} chain-grads/s", flush=True)
