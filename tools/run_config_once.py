#!/usr/bin/env python
"""A few eager ELBO fwd+bwd steps of one model config (for ncu launch lists): python tools/run_config_once.py mocap_gpode"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import gpode_oracle as O  # noqa: E402
from bench_configs import CONFIGS  # noqa: E402
from util import build_product_model  # noqa: E402
from gaussian_process_odes_b200 import builders  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mocap_gpode"
c = CONFIGS[name]
kw = {k: v for k, v in c.items() if k not in ("kind", "ts_dense_scale")}
p, ys, ts, draws, proj = O.make_problem(seed=121, **kw)
model = build_product_model(c["kind"], p, ys, c["S"], "rk4", ts_dense_scale=c.get("ts_dense_scale", 4),
                            proj=None if proj is None else proj.components)
ysd, tsd = ys.cuda(), ts.cuda()
for _ in range(4):
    model.zero_grad(set_to_none=True)
    if c["kind"] == "gpode":
        loss = builders.compute_loss_gpode(model, ysd, tsd)[0]
    else:
        loss = builders.compute_loss_shooting(model, ysd, tsd, num_samples=c["S_mc"])[0]
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss.detach()))
