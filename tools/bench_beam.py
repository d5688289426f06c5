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

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from clocks import ClockSampler  # noqa: E402


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
    lib.bpmf_main_kernel_times((C.c_float * 2048)(), 2048)  # drain earlier records
    sampler = ClockSampler(device_index).start()
    t0 = time.perf_counter()
    for _ in range(steps):
        plan.run_max(feat, tau, wp, ws)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    buf = (C.c_float * 64)()
    nk = lib.bpmf_main_kernel_times(buf, 64)
    lib.bpmf_set_kernel_timing(0)
    kms = float(np.mean([buf[i] for i in range(nk)])) if nk else 1e3 * dt
    gathers = K * keep * P * N
    sm_clock = 1.965e9
    sm_seen = (clocks.get("sm_mhz") or 1965.0) * 1e6
    return {"clocks": clocks, "workload": "beamform reduce=max: %d grid sources x %d stations x %d phases x %d samples%s%s" % (
                K, S, P, N, "" if keep == S else ", %d closest stations weighted" % keep,
                ", RANDOM moveouts" if random_moveouts else ""),
            "K": K, "N": N, "S": S, "keep": keep, "moveouts": "random" if random_moveouts else "grid",
            "tau_max": int(tau.max().item()), "ms_per_step": 1e3 * dt, "kernel_ms": kms,
            "G_source_samples_per_s": K * N / dt / 1e9, "G_gather_fma_per_s": gathers / (kms * 1e-3) / 1e9,
            "roofline": {"bound": "shared-memory bandwidth (4 B per gather-FMA)", "unit": "TB/s",
                         "achieved": 4 * gathers / (kms * 1e-3) / 1e12, "peak": 148 * 128 * sm_clock / 1e12,
                         "frac": 4 * gathers / (kms * 1e-3) / (148 * 128 * sm_clock),
                         "frac_at_sampled_clock": 4 * gathers / (kms * 1e-3) / (148 * 128 * sm_seen),
                         "peak_source": "148 SM x 128 B/clk x 1965 MHz (frac_at_sampled_clock: x the median SM clock "
                                        "nvidia-smi reported during the timed passes)",
                         "kernel": "beam_group_kernel"}}


def _grid_inputs_host(K, N, S, seed=0):
    """The same kind of workload as run(), as HOST arrays with the dtypes BPMF passes (int64 moveouts,
    BPMF/utils.py:1270): features (S,2,N) f32, moveouts (K,S,2) int64, weights_phases, weights_sources."""
    rng = np.random.default_rng(seed)
    feat = np.clip(np.abs(rng.standard_normal((S, 2, N), dtype=np.float32)), 0, 1)
    wp = np.zeros((S, 2, 2), np.float32)
    wp[:, 0, 0] = 1.0
    wp[:, 1, 1] = 1.0
    n1 = int(round(K ** (1 / 3))) + 1
    g = np.stack(np.meshgrid(*(np.linspace(0, 1, n1),) * 3, indexing="ij"), -1).reshape(-1, 3)[:K] * np.array([100.0, 100.0, 30.0])
    sta = np.concatenate([rng.uniform(0, 100, (S, 2)), np.zeros((S, 1))], axis=1)
    tp_ = np.linalg.norm(g[:, None, :] - sta[None], axis=-1) / 6.0 * 50.0
    tau = np.stack([np.floor(tp_), np.floor(tp_ * 1.73)], -1).astype(np.int64)
    tau -= tau.min(axis=(1, 2), keepdims=True)
    ws = np.full((K, S), 1.0 / S, np.float32)
    return feat, tau, wp, ws


def cpu_baseline(seconds=10.0, S=32):
    """oracle/beam_oracle.c (C99 + OpenMP, every host thread) on a bounded sample of the same workload; the
    number of sources is scaled from a calibration run so that the sample takes about `seconds`."""
    from oracle import cpu

    native = cpu.use_native()
    cores = os.cpu_count() or 1

    def one(K, N):
        f, tau, wp, ws = _grid_inputs_host(K, N, S, seed=1)
        t0 = time.perf_counter()
        cpu.beamform(f, tau, wp, ws, num_threads=cores)
        return time.perf_counter() - t0

    N = 43200
    dt = one(64, N)
    K = int(max(64, min(20000, 64 * seconds / max(dt, 1e-3))))
    dt = one(K, N)
    return {"value": K * N / dt / 1e9, "unit": "G source.samples/s", "G_gather_fma_per_s": K * S * 2 * N / dt / 1e9,
            "cores": cores, "kind": "port",
            "build": cpu.build_kind if native else "portable (-march=x86-64-v3)",
            "sample": "%d sources x %d stations x 2 phases x %d samples, %.1f s" % (K, S, N, dt)}


def e2e_host_api(K=50000, N=432000, S=32, steps=2):
    """The call BPMF makes (BPMF/template_search.py:549-558): beampower.beamform(features, int64 moveouts,
    weights_phases, weights_sources, device="gpu", reduce="max") with HOST arrays; H2D and D2H inside."""
    from seismic_bpmf_b200.beampower import beampower as bp

    f, tau, wp, ws = _grid_inputs_host(K, N, S, seed=2)
    bp.beamform(f, tau, wp, ws, device="gpu", reduce="max")  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        mb, am = bp.beamform(f, tau, wp, ws, device="gpu", reduce="max")
    dt = (time.perf_counter() - t0) / steps
    return {"value": K * N / dt / 1e9, "unit": "G source.samples/s", "ms_per_step": 1e3 * dt, "steps": steps,
            "h2d_bytes_per_step": int(f.nbytes + tau.nbytes // 2 + wp.nbytes + ws.nbytes),
            "d2h_bytes_per_step": int(mb.nbytes + am.nbytes),
            "note": "int64 moveouts are narrowed to int32 on the host before the copy (counted as int32 bytes)"}


def run_sharded(K=200000, N=432000, S=32, steps=3, warmup=2, rank=0, world=1, device_index=0):
    """Source-sharded reduce="max" (SURVEY 8e): strong scaling of a fixed K x N problem over `world` ranks,
    all-gather of (max, argmax) + G-way select (sharded.beamform_max_sharded).  Device-resident inputs;
    CUDA events, max over ranks."""
    import torch.distributed as dist

    from seismic_bpmf_b200 import sharded

    d = torch.device("cuda", device_index)
    f, tau, wp, ws = _grid_inputs_host(K, N, S, seed=3)
    f, wp, ws = (torch.from_numpy(a).to(d) for a in (f, wp, ws))
    tau = torch.from_numpy(tau.astype(np.int32)).to(d)
    for _ in range(warmup):
        mb, am = sharded.beamform_max_sharded(f, tau, wp, ws)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        mb, am = sharded.beamform_max_sharded(f, tau, wp, ws)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=d, dtype=torch.float64)
    chk = torch.stack([mb.double().sum(), am.double().sum()])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
    else:
        same = True
    ms = float(t.item())
    return {"workload": "beamform reduce=max, %d grid sources x %d stations x 2 phases x %d samples, sources "
                        "block-partitioned over %d rank(s), all-gather of (max, argmax)" % (K, S, N, world),
            "scaling": "strong", "n_gpus": world, "ms_per_step": ms, "G_source_samples_per_s": K * N / (ms * 1e-3) / 1e9,
            "checksum_maxbeam": float(chk[0].item()), "checksum_argmax": float(chk[1].item()),
            "result_identical_on_all_ranks": same}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=50000)
    ap.add_argument("--samples", type=int, default=432000)
    ap.add_argument("--stations", type=int, default=32)
    ap.add_argument("--keep", type=int, default=32, help="stations with non-zero source weight per source")
    ap.add_argument("--random-moveouts", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    print(json.dumps(run(a.sources, a.samples, a.stations, a.keep, a.random_moveouts, a.steps, a.warmup)))
