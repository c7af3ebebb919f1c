"""Error table of the ELBO gradients over many seeds: CUDA vs float64 oracle vs the float32 oracle port (which is pinned
to the unmodified reference, oracle/pin_against_reference.py). Runs on the GPU box (the oracle runs on its host cores).

    python tools/seed_table.py [--case vdp_gpode_rk4] [--seeds 20] [--out profiles/r02_seed_table_vdp_gpode_rk4.md]

For every gradient tensor: relative error (max-abs / max-abs) of CUDA vs fp64 and of the float32 port vs fp64; the
acceptance line of round 2 is "CUDA no further from float64 than the reference's own float32 path".
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import gpode_oracle as O  # noqa: E402

CASES = {
    "vdp_gpode_rk4": ("gpode", dict(D=2, M=16, S=256, N=1, T=25), "rk4", dict(ts_dense_scale=4)),
    "vdp_shooting_rk4": ("shooting", dict(D=2, M=16, S=256, N=1, T=25, S_mc=5), "rk4", {}),
    "mocap_gpode_rk4": ("gpode", dict(D=5, M=100, S=256, N=2, T=20, D_obs=50, dt=0.01, ell0=1.25), "rk4",
                        dict(ts_dense_scale=2)),
    "mocap_shooting_rk4": ("shooting", dict(D=5, M=100, S=256, N=2, T=20, S_mc=3, D_obs=50, dt=0.01, ell0=1.25),
                           "rk4", {}),
}


from util import elbo_errors as run_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="vdp_gpode_rk4")
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--first-seed", type=int, default=121)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    kind, kw, solver, extra = CASES[a.case]
    torch.set_num_threads(os.cpu_count())
    allrows = []
    for s in range(a.first_seed, a.first_seed + a.seeds):
        allrows.append((s, run_case(kind, kw, solver, extra, s)))
    keys = list(allrows[0][1].keys())
    lines = ["# %s: relative error vs the float64 oracle, %d seeds (cuda | float32 reference port)" % (a.case, a.seeds),
             "", "| seed | " + " | ".join(keys) + " |", "|---|" + "---|" * len(keys)]
    worst = {k: [0.0, 0.0] for k in keys}
    n_worse = {k: 0 for k in keys}
    for s, rows in allrows:
        lines.append("| %d | " % s + " | ".join("%.1e / %.1e" % rows[k][:2] for k in keys) + " |")
        for k in keys:
            worst[k][0] = max(worst[k][0], rows[k][0])
            worst[k][1] = max(worst[k][1], rows[k][1])
            n_worse[k] += rows[k][0] > max(rows[k][1], 1e-5)
    lines.append("| **max** | " + " | ".join("%.1e / %.1e" % tuple(worst[k]) for k in keys) + " |")
    lines.append("| seeds where cuda > max(ref32, 1e-5) | " + " | ".join(str(n_worse[k]) for k in keys) + " |")
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(txt + "\n")


if __name__ == "__main__":
    main()
