"""Beamforming throughput (BASELINE.json configs[2] scaled): G(source.sample)/s and gather-FMA/s, with the
shared-memory-bandwidth roofline of SURVEY.md 8(d).  python tools/bench_beam.py [--sources K] [--samples N]

`run()` is also what bench.py calls for the "beamform" object of its JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seismic_bpmf_b200 import _lib, device as dev  # noqa: E402


def run(K=50000, N=432000, S=32, keep=32, random_moveouts=False, steps=3, warmup=2, device_index=0):
    """Synthetic backprojection: features (S, 2, N) in [0, 1], identity phase weights, a regular source grid
    under a surface network at constant velocity (neighbouring grid points have close moveouts, as in
    BPMF's travel-time tables) or random moveouts, all-ones row-normalised source weights (optionally only
    the `keep` closest stations, BPMF/template_search.py:779-798).  reduce="max", device-resident."""
    Cc, P = 2, 2
    torch.manual_seed(0)
    d = torch.device("cuda", device_index)
    feat = torch.rand((S, Cc, N), device=d).clamp_(0, 1)
    wp = torch.zeros((S, Cc, P), device=d)
    wp[:, 0, 0] = 1.0
    wp[:, 1, 1] = 1.0
    if random_moveouts:
        tp_ = torch.randint(0, 1501, (K, S), device=d).double()
    else:
        n1 = int(round(K ** (1 / 3))) + 1
        g = torch.stack(torch.meshgrid(*(torch.linspace(0, 1, n1, device=d),) * 3, indexing="ij"), -1).reshape(-1, 3)[:K]
        g = g * torch.tensor([100.0, 100.0, 30.0], device=d)
        sta = torch.cat([torch.rand((S, 2), device=d) * 100.0, torch.zeros((S, 1), device=d)], 1)
        tp_ = (g[:, None, :] - sta[None]).norm(dim=-1).double() / 6.0 * 50.0  # samples @ 50 Hz, vP = 6 km/s
    tau = torch.stack([tp_.floor(), (tp_ * 1.73).floor()], -1)
    tau = (tau - tau.amin(dim=(1, 2), keepdim=True)).to(torch.int32).contiguous()
    ws = torch.ones((K, S), device=d)
    if keep < S:
        order = tau[:, :, 0].argsort(dim=1)
        ws.scatter_(1, order[:, keep:], 0.0)
    ws = (ws / ws.sum(1, keepdim=True)).contiguous()
    plan = dev.BeamformPlan(K, S, Cc, P, N)
    lib = _lib.load()
    for _ in range(warmup):
        plan.run_max(feat, tau, wp, ws)
    torch.cuda.synchronize()
    lib.bpmf_set_kernel_timing(1)
    lib.bpmf_main_kernel_times((C.c_float * 256)(), 256)  # drain earlier records
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.run_max(feat, tau, wp, ws)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    buf = (C.c_float * 64)()
    nk = lib.bpmf_main_kernel_times(buf, 64)
    lib.bpmf_set_kernel_timing(0)
    kms = float(np.mean([buf[i] for i in range(nk)])) if nk else 1e3 * dt
    gathers = K * keep * P * N
    sm_clock = 1.965e9
    return {"workload": "beamform reduce=max: %d grid sources x %d stations x %d phases x %d samples%s%s" % (
                K, S, P, N, "" if keep == S else ", %d closest stations weighted" % keep,
                ", RANDOM moveouts" if random_moveouts else ""),
            "K": K, "N": N, "S": S, "keep": keep, "moveouts": "random" if random_moveouts else "grid",
            "tau_max": int(tau.max().item()), "ms_per_step": 1e3 * dt, "kernel_ms": kms,
            "G_source_samples_per_s": K * N / dt / 1e9, "G_gather_fma_per_s": gathers / (kms * 1e-3) / 1e9,
            "roofline": {"bound": "shared-memory bandwidth (4 B per gather-FMA)", "unit": "TB/s",
                         "achieved": 4 * gathers / (kms * 1e-3) / 1e12, "peak": 148 * 128 * sm_clock / 1e12,
                         "frac": 4 * gathers / (kms * 1e-3) / (148 * 128 * sm_clock),
                         "peak_source": "148 SM x 128 B/clk x 1965 MHz", "kernel": "beam_group_kernel"}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=50000)
    ap.add_argument("--samples", type=int, default=432000)
    ap.add_argument("--stations", type=int, default=32)
    ap.add_argument("--keep", type=int, default=32, help="stations with non-zero source weight per source")
    ap.add_argument("--random-moveouts", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(run(a.sources, a.samples, a.stations, a.keep, a.random_moveouts, a.steps)))
