"""Beamforming throughput (BASELINE.json configs[2] scaled): G(source.sample)/s and gather-FMA/s, with the
shared-memory-bandwidth roofline of SURVEY.md 8(d).  python tools/bench_beam.py [--sources K] [--samples N]"""
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

ap = argparse.ArgumentParser()
ap.add_argument("--sources", type=int, default=100000)
ap.add_argument("--samples", type=int, default=432000)
ap.add_argument("--stations", type=int, default=32)
ap.add_argument("--keep", type=int, default=32, help="stations with non-zero source weight per source")
ap.add_argument("--random-moveouts", action="store_true")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
K, N, S, Cc, P = args.sources, args.samples, args.stations, 2, 2
torch.manual_seed(0)
d = torch.device("cuda", 0)
feat = torch.rand((S, Cc, N), device=d).clamp_(0, 1)
wp = torch.zeros((S, Cc, P), device=d)
wp[:, 0, 0] = 1.0
wp[:, 1, 1] = 1.0
if args.random_moveouts:
    tp_ = torch.randint(0, 1501, (K, S), device=d).double()
else:  # a regular source grid under a surface network, constant velocity: neighbours have similar moveouts
    n1 = int(round(K ** (1 / 3))) + 1
    g = torch.stack(torch.meshgrid(*(torch.linspace(0, 1, n1, device=d),) * 3, indexing="ij"), -1).reshape(-1, 3)[:K]
    g = g * torch.tensor([100.0, 100.0, 30.0], device=d)
    sta = torch.cat([torch.rand((S, 2), device=d) * 100.0, torch.zeros((S, 1), device=d)], 1)
    tp_ = (g[:, None, :] - sta[None]).norm(dim=-1).double() / 6.0 * 50.0  # samples @ 50 Hz, vP = 6 km/s
tau = torch.stack([tp_.floor(), (tp_ * 1.73).floor()], -1)
tau = (tau - tau.amin(dim=(1, 2), keepdim=True)).to(torch.int32).contiguous()
ws = torch.ones((K, S), device=d)
if args.keep < S:
    order = tau[:, :, 0].argsort(dim=1)
    ws.scatter_(1, order[:, args.keep:], 0.0)
ws = (ws / ws.sum(1, keepdim=True)).contiguous()
plan = dev.BeamformPlan(K, S, Cc, P, N)
lib = _lib.load()
for _ in range(2):
    plan.run_max(feat, tau, wp, ws)
torch.cuda.synchronize()
lib.bpmf_set_kernel_timing(1)
t0 = time.perf_counter()
for _ in range(args.steps):
    plan.run_max(feat, tau, wp, ws)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / args.steps
buf = (C.c_float * 64)()
nk = lib.bpmf_main_kernel_times(buf, 64)
kms = float(np.mean([buf[i] for i in range(nk)]))
gathers = K * args.keep * P * N
sm_clock = 1.965e9
print(json.dumps({"K": K, "N": N, "S": S, "keep": args.keep, "moveouts": "random" if args.random_moveouts else "grid",
                  "tau_max": int(tau.max().item()), "ms_per_step": 1e3 * dt, "kernel_ms": kms,
                  "G_source_samples_per_s": K * N / dt / 1e9, "G_gather_fma_per_s": gathers / (kms * 1e-3) / 1e9,
                  "smem_roofline_frac_at_1965MHz": 4 * gathers / (kms * 1e-3) / (148 * 128 * sm_clock)}))
