"""Progress lines (handlers.hpp:38-48) and R-hat lines (:164-172) of the one-shot call in
free-running and uniform mode (needs a GPU)."""
import contextlib
import io
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import walnuts_b200 as wb  # noqa: E402

for mode in ("free", "uniform"):
    if mode == "uniform":
        os.environ["WB200_BLOCKS"] = "uniform"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
        fit = wb.walnuts_device(wb.models.std_normal(3), num_chains=4, seed=1, refresh=10,
                                min_warmup_iter=20, max_warmup_iter=60,
                                min_sampling_iter=20, max_sampling_iter=80)
    text = buf.getvalue()
    lines = [ln for ln in text.splitlines() if ln.strip()]
    warm = [ln for ln in lines if "(Warmup)" in ln]
    samp = [ln for ln in lines if "(Sampling)" in ln]
    rhat = [ln for ln in lines if "R-hat" in ln]
    print(mode, "lengths", [len(f) for f in fit], "| lines:", len(warm), "warm-up,",
          len(samp), "sampling,", len(rhat), "R-hat; first:", (warm or lines or ["-"])[0])
