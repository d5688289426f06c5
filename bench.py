#!/usr/bin/env python
"""bench.py -- CC-sum throughput of the matched-filter hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N ...            # CPU oracle on the host cores
    torchrun ... bench.py --gpus N ...                       # N > 1: one rank per GPU

A "step" is one pass of the matched filter over one batch of synthetic input.  At N = 1 the
workload is BASELINE.json configs[1] (500 templates x 16 stations x 3 channels, 8 s templates, one
day @ 100 Hz, float32).  At N > 1 every rank processes that same per-GPU workload on its own shard
of templates (weak scaling: 500*N templates, traces replicated) and the per-shard CC slabs are
exchanged with NCCL all-gathers (chunked, because the gathered CC of configs[3] does not fit one
GPU -- SURVEY.md 8e).

metric = CC-sum GSamples/s = templates x channels x n_corr / seconds / 1e9.
  value : inputs resident in HBM when the timed region starts (device entry point, CUDA events).
  e2e   : same metric through the reference-facing call fast_matched_filter.matched_filter(...) with
          pinned HOST buffers, host<->device copies inside the timed region.
Extra objects on the same JSON line (each leg is untimed by the contract and can be switched off):
  parity      strided comparison with the CPU oracle over the whole day, every template tile (tests/_parity.py)
  variants    SURVEY 8(d)'s second weighting (10 smallest-moveout stations per template): value + parity
  cfg4        BASELINE.json configs[3], 5000 x 32 x 3 strong-scaled over the N ranks (tools/bench_cfg4.py)
  beamform    the second hot path: device figure + roofline, host-API e2e, CPU baseline; N > 1: source-sharded
  libc        the libc.so rows against the reference's own compiled libc.c (tools/bench_libc.py)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240726
CFG2 = dict(T=500, S=16, C=3, L=800, N=8640000, step=1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "fp16", "fp16x3"])
    ap.add_argument("--templates", type=int, default=CFG2["T"], help="templates per GPU")
    ap.add_argument("--samples", type=int, default=CFG2["N"], help="trace length in samples")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-beam", action="store_true", help="skip the beamforming figure")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the 5000-template strong-scaling leg")
    ap.add_argument("--no-libc", action="store_true", help="skip the libc.so rows")
    ap.add_argument("--no-variants", action="store_true", help="skip the second weighting variant")
    ap.add_argument("--no-parity", action="store_true", help="skip the strided oracle comparison")
    ap.add_argument("--cfg4-steps", type=int, default=2)
    ap.add_argument("--callers", action="store_true",
                    help="also time the SURVEY 8(f) rows (moveout tables, detection traces, inter-template CC) against "
                         "their CPU restatements on bounded samples (adds ~1 min; off by default)")
    ap.add_argument("--cpu-seconds", type=float, default=25.0)
    ap.add_argument("--batch", type=int, default=0, help="templates per kernel launch at N > 1 (0 = default)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tensor_tflops=p["bf16_tflops"],
                    tensor_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    sm_max_mhz=p.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, tensor_tflops=1590.0, tensor_tflops_sustained=1400.0, sm_max_mhz=1965.0,
                source="fallback")


sys.path.insert(0, os.path.join(ROOT, "tools"))
from clocks import ClockSampler  # noqa: E402


# ------------------------------------------------------------------ synthetic workload
def make_inputs_device(torch, T, S, Cc, L, N, seed, device):
    """Traces ~N(0,1) per-channel RMS-normalised (BPMF/similarity_search.py:181-185) with a 10 000
    sample zero gap on one station; templates ~N(0,1) demeaned + normalised (BPMF/dataset.py:4160-4165),
    1 % cut from the data (planted detections); int32 moveouts U{-100..3000}; weights all-ones
    row-normalised (BPMF/similarity_search.py:469-472).  SURVEY.md 8(d)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    data = torch.randn((S, Cc, N), generator=g, device=device, dtype=torch.float32)
    data[S // 2, :, N // 3: N // 3 + 10000] = 0.0
    std = data.std(dim=-1, keepdim=True)
    std[std == 0] = 1.0
    data /= std
    templates = torch.randn((T, S, Cc, L), generator=g, device=device, dtype=torch.float32)
    templates -= templates.mean(dim=-1, keepdim=True)
    templates /= templates.std(dim=-1, keepdim=True)
    mv_hi = int(min(3000, max(0, (N - L) // 4)))
    moveouts = torch.randint(-100, mv_hi + 1, (T, S, Cc), generator=g, device=device, dtype=torch.int32)
    planted = {}
    for t in range(0, T, 100):
        i0 = 5000 + 37 * t
        if i0 + mv_hi + L >= N // 3:
            break
        idx = (i0 + moveouts[t].long()).reshape(S, Cc, 1) + torch.arange(L, device=device).reshape(1, 1, L)
        templates[t] = torch.gather(data, 2, idx)
        planted[t] = i0
    weights = torch.full((T, S, Cc), 1.0 / (S * Cc), device=device, dtype=torch.float32)
    return templates.contiguous(), moveouts.contiguous(), weights, data.contiguous(), planted


def measured_traffic(kernel, T, S, Cc, L, N):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture of this exact
    configuration (profiles/r2_traffic.json); None when the configuration was not captured."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None
    for rec in json.load(open(path)):
        if rec["kernel"] == kernel and [rec["T"], rec["S"], rec["C"], rec["L"], rec["N"]] == [T, S, Cc, L, N]:
            return rec["dram_bytes_read"] + rec["dram_bytes_write"]
    return None


def algorithmic_bytes(T, S, Cc, L, N, n_corr):
    """compulsory traffic of one matched-filter pass."""
    return 4 * S * Cc * N + 4 * T * S * Cc * L + 8 * T * S * Cc + 4 * T * n_corr


# ------------------------------------------------------------------ CPU oracle arm
def cpu_sample_throughput(seconds, S, Cc, L, seed=SEED):
    """Time oracle/mf_oracle.c (C99 + OpenMP, all host threads) on a bounded sample of the same
    workload: S x C channels, L-sample templates, step 1; the number of templates and the trace
    length are scaled from a calibration run so the sample takes about `seconds`."""
    from oracle import cpu

    native = cpu.use_native()  # built on this host with the reference's flags, -march=native included
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(seed)

    def run(T, N):
        data = rng.standard_normal((S, Cc, N)).astype(np.float32)
        tp = rng.standard_normal((T, S, Cc, L)).astype(np.float32)
        mv = rng.integers(0, min(3000, N // 8) + 1, (T, S, Cc)).astype(np.int32)
        w = np.full((T, S, Cc), 1.0 / (S * Cc), np.float32)
        t0 = time.perf_counter()
        cc = cpu.matched_filter(tp, mv, w, data, 1, num_threads=cores)  # torchrun exports OMP_NUM_THREADS=1
        dt = time.perf_counter() - t0
        return T * S * Cc * cc.shape[1] / dt / 1e9, dt

    rate, dt = run(1, 60000)
    units = rate * 1e9 * seconds
    N = 400000
    T = int(max(1, min(64, units / (S * Cc * (N - L + 1)))))
    rate, dt = run(T, N)
    return rate, cores, "%d templates x %d stations x %d channels x %d samples (L=%d, step 1), %.1f s" % (
        T, S, Cc, N, L, dt), (cpu.build_kind if native else "portable (-march=x86-64-v3)")


def cpu_model():
    from oracle import build as ob

    return ob.cpu_model()


def run_reference(args, rank):
    if rank != 0:
        return
    T, S, Cc, L, N = args.templates, CFG2["S"], CFG2["C"], CFG2["L"], args.samples
    rates, sample = [], ""
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        rate, cores, sample, build = cpu_sample_throughput(per_step, S, Cc, L, seed=SEED + i)
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    n_corr = N - L + 1
    line = {
        "impl": "reference", "metric": "CC-sum GSamples/s (templates x traces x steps)", "value": value,
        "unit": "GSamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (T * S * Cc * n_corr / 1e9) / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "matched_filter cfg2: %d templates x %d stations x %d ch, L=%d, N=%d, step 1"
                   % (T, S, Cc, L, N), "note": "ms_per_step extrapolated from the bounded sample"},
        "cpu_baseline": {"value": value, "unit": "GSamples/s", "cores": cores, "kind": "port", "sample": sample,
                         "build": build, "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": "GSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ callers either side of the kernels
def callers_figures():
    """f2 / f3 / f4 through their public (reference-signature) entry points with HOST arrays, next to the CPU
    restatement (oracle/, single process, the box's host cores as NumPy / SciPy use them) on the same inputs."""
    from oracle import cpu, features_numpy as fo, moveouts_numpy as mo
    from seismic_bpmf_b200 import features as ft, intertemplate as itp, template_search as ts

    def timed(fn, reps=1):
        fn()  # warm-up (cuFFT plans, allocations)
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        return (time.perf_counter() - t0) / reps, out

    rng = np.random.default_rng(SEED)
    res = {}
    # f2: travel-time table of 250 000 grid sources x 32 stations x 2 phases (float32 seconds) -> int32 moveouts
    K, S, P = 250000, 32, 2
    tt = rng.uniform(0.0, 60.0, (K, S, P)).astype(np.float32)
    t_gpu, (mv, _) = timed(lambda: ts.moveouts_from_travel_times(tt, 50.0), 3)
    t0 = time.perf_counter()
    want, _ = mo.moveouts_from_travel_times(tt, 50.0)
    t_cpu = time.perf_counter() - t0
    res["moveouts"] = {"workload": "%d sources x %d stations x %d phases, float32 seconds -> samples, relative to first" % (K, S, P),
                       "gpu_ms": 1e3 * t_gpu, "cpu_numpy_ms": 1e3 * t_cpu, "bit_exact": bool(np.array_equal(mv, want)),
                       "h2d_bytes": int(tt.nbytes), "d2h_bytes": int(mv.nbytes)}
    # f3: saturated envelopes of 8 stations x 3 components x 1 day @ 25 Hz, and the threshold of a 1-day max-beam
    Sx, Cx, Nx = 8, 3, 2160000
    traces = (rng.standard_normal((Sx, Cx, Nx)) * 1e-7).astype(np.float32)
    traces[2, 1, 500000:700000] = 0.0
    t_gpu, (feat, avail) = timed(lambda: ft.saturated_envelopes(traces), 2)
    sub = traces[:1]  # the CPU restatement is timed on one station (3 traces) and scaled
    t0 = time.perf_counter()
    fw, aw = fo.saturated_envelopes(sub)
    t_cpu = (time.perf_counter() - t0) * Sx
    err = float(np.abs(feat[:1] - fw).max() / max(1.0, np.abs(fw).max()))
    res["saturated_envelopes"] = {"workload": "%d stations x %d components x %d samples" % (Sx, Cx, Nx), "gpu_ms": 1e3 * t_gpu,
                                  "cpu_scipy_ms_scaled_from_one_station": 1e3 * t_cpu, "max_rel_err_first_station": err,
                                  "availability_matches": bool(np.array_equal(avail[:1], aw))}
    nr = np.abs(rng.standard_normal(Nx)).astype(np.float32)
    t_gpu, thr = timed(lambda: ft.time_dependent_threshold(nr, 45000, overlap=0.75, CNR_threshold=10.0), 2)
    t0 = time.perf_counter()
    thr_w = fo.time_dependent_threshold(nr, 45000, overlap=0.75, CNR_threshold=10.0)
    t_cpu = time.perf_counter() - t0
    res["bp_time_dependent_threshold"] = {"workload": "%d samples, 30 min windows @ 25 Hz, overlap 0.75" % Nx, "gpu_ms": 1e3 * t_gpu,
                                          "cpu_scipy_ms": 1e3 * t_cpu,
                                          "max_rel_err": float(np.abs(thr - thr_w).max() / np.abs(thr_w).max())}
    # f4: inter-template CC of 400 templates x 10 stations x 3 components x 820 samples, max_lag 10
    T, Sg, Cg, Lw, lag = 400, 10, 3, 820, 10
    wav = rng.standard_normal((T, Sg, Cg, Lw)).astype(np.float32)
    sw = np.full((T, Sg, Cg), 1.0 / (Sg * Cg), np.float32)
    t_gpu, cc = timed(lambda: itp.intertemplate_cc(wav, sw, None, max_lag=lag), 3)
    t_loop, cc_loop = timed(lambda: itp.intertemplate_cc_reference_loop(wav, sw, None, max_lag=lag), 1)
    # one iteration of the reference's loop (dataset.py:4818-4830) with the CPU matched filter, scaled to T iterations
    t0 = time.perf_counter()
    cc1 = cpu.matched_filter(wav[..., lag: Lw - lag], np.zeros((T, Sg, Cg), np.int32), np.repeat(sw[:1], T, 0), wav[0], 1,
                             network_sum=False, num_threads=os.cpu_count() or 1)
    row0 = np.sum(sw[0] * np.max(cc1, axis=1), axis=(-1, -2))
    t_cpu = (time.perf_counter() - t0) * T
    cc_raw = itp.intertemplate_cc(wav, sw, None, max_lag=lag, symmetric=False)
    res["intertemplate_cc"] = {"workload": "%d templates x %d stations x %d components x %d samples, max_lag %d (%.1e pair-channel-lags)"
                               % (T, Sg, Cg, Lw, lag, T * T * Sg * Cg * (2 * lag + 1)),
                               "gpu_batched_ms": 1e3 * t_gpu, "gpu_loop_of_matched_filter_calls_ms": 1e3 * t_loop,
                               "cpu_oracle_loop_ms_scaled_from_one_iteration": 1e3 * t_cpu,
                               "max_abs_diff_batched_vs_gpu_loop": float(np.abs(cc - cc_loop).max()),
                               "max_abs_diff_row0_vs_cpu_oracle": float(np.abs(cc_raw[0] - row0).max())}
    return res


# ------------------------------------------------------------------ our arm
def ten_closest_weights(torch, moveouts, keep=10):
    """SURVEY 8(d), second variant: only the `keep` smallest-moveout stations of every template carry weight
    (mirrors N_SELECTED_STATIONS = 10, tutorial 8_template_matching.ipynb:90), rows normalised to sum 1."""
    T, S, Cc = moveouts.shape
    order = moveouts.min(dim=2).values.argsort(dim=1)
    w = torch.zeros((T, S, Cc), device=moveouts.device, dtype=torch.float32)
    w.scatter_(1, order[:, :keep].unsqueeze(-1).expand(T, keep, Cc), 1.0)
    return (w / w.sum(dim=(1, 2), keepdim=True)).contiguous()


def strided_parity(torch, plan, tp, mv, w, data_host, planted, N, L, n_windows=64, win_len=2048):
    """Untimed: the CC slab the timed passes left in HBM against the CPU oracle on runs of `win_len` outputs spread
    over the whole day, for three templates of every 16-template tile plus every planted template."""
    from oracle import cpu
    from tests import _parity

    n_corr = plan.n_corr
    T = tp.shape[0]
    sel = _parity.tile_templates(T, per_tile=3, extra=planted.keys())
    starts = _parity.window_starts(n_corr, n_windows, win_len, gap=(N // 3, N // 3 + 10000), planted=planted.values(), L=L)
    idx = torch.from_numpy(starts).to(plan.cc.device).reshape(-1, 1) + torch.arange(win_len, device=plan.cc.device)
    sel_d = torch.from_numpy(sel).to(plan.cc.device)
    got = plan.cc[sel_d][:, idx.reshape(-1)].reshape(len(sel), len(starts), win_len).cpu().numpy()
    t0 = time.perf_counter()
    want = cpu.matched_filter_windows(tp[sel_d].cpu().numpy(), mv[sel_d].cpu().numpy(), w[sel_d].cpu().numpy(), data_host,
                                      starts, win_len, num_threads=os.cpu_count() or 1)
    rec = _parity.compare(got, want)
    ok_plant = all(int(torch.argmax(plan.cc[t, :n_corr]).item()) == i0 for t, i0 in planted.items())
    rec.update({"planted_detections_recovered": bool(ok_plant), "oracle_seconds": time.perf_counter() - t0,
                "sample": "%d templates (3 of each of the %d tiles of 16 + the planted ones) x %d runs of %d outputs over "
                          "[0, n_corr): first / last outputs, segment boundaries, the zero gap, planted indices, seeded "
                          "random positions" % (len(sel), (T + 15) // 16, len(starts), win_len)})
    return rec


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from seismic_bpmf_b200 import _lib, device as dev, fast_matched_filter as fmf

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    lib = _lib.load()
    pk = peaks()
    T, S, Cc, L, N, step = args.templates, CFG2["S"], CFG2["C"], CFG2["L"], args.samples, 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def guarded(fn, *a, **k):
        """An extra figure must never break the contract line."""
        try:
            return fn(*a, **k)
        except Exception as exc:  # noqa: BLE001
            torch.cuda.synchronize()
            return {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    # rank r owns templates [r*T, (r+1)*T) of the global 500*N set: different seeds per rank
    tp, mv, w, data, planted = make_inputs_device(torch, T, S, Cc, L, N, SEED + rank, device)
    if world > 1:  # traces are replicated: take rank 0's
        dist.broadcast(data, src=0)
        planted = {}
    # N > 1: the per-rank pass runs in template batches so that the all-gather of one batch's CC slab
    # overlaps the kernels of the next one
    batch = T if world == 1 else (args.batch or (128 if world <= 2 else (64 if world <= 4 else 32)))
    plan = dev.MatchedFilterPlan(T, S, Cc, L, N, step=step, precision=args.precision, device=device, batch=batch)
    n_corr = plan.n_corr
    units = T * S * Cc * n_corr  # per rank, per step

    gather, gbufs, collective = None, None, "none"
    if world > 1:
        from seismic_bpmf_b200 import sharded

        try:  # copy-engine pushes over NVLink peer memory: overlap with the persistent kernel without taking SMs
            gather = sharded.PeerSlabGather(batch * plan.pitch, device)
            collective = ("all-gather of per-shard CC slabs, one per %d-template batch, pushed to peer memory over " % batch +
                          "NVLink by the copy engines (torch symmetric memory) while the next batch computes")
        except Exception as exc:  # noqa: BLE001 -- fall back to NCCL kernels
            if rank == 0:
                print("bench: symmetric memory unavailable (%s); using NCCL all-gather" % exc, file=sys.stderr)
            gbufs = [torch.empty(world * batch * plan.pitch, dtype=torch.float32, device=device) for _ in range(2)]
            collective = "NCCL all-gather of per-shard CC slabs, one per %d-template batch" % batch

    # batch schedule at N > 1: full batches, then a geometric tail (.., 32, 16, 8) so that the push of the LAST batch,
    # which nothing overlaps, is short
    schedule, t_next = [], 0
    while t_next < T:
        rows = min(batch, T - t_next)
        if world > 1 and T - t_next <= batch:
            rows = max(8, (T - t_next + 1) // 2) if T - t_next > 8 else T - t_next
        schedule.append((t_next, rows))
        t_next += rows

    def step_fn(weights=None):
        ww = w if weights is None else weights
        if world == 1:
            return plan.run(tp, mv, ww, data)
        plan.prepare(data)
        works = []
        for b, (t0, rows) in enumerate(schedule):
            plan.run_batch(tp, mv, ww, data, t0, rows)
            src = plan.cc[t0: t0 + rows].reshape(-1)
            if gather is not None:
                gather.push(b % 2, src)
            else:
                if b >= 2:
                    works[b - 2].wait()  # the gather buffer about to be reused is free again
                works.append(dist.all_gather_into_tensor(gbufs[b % 2][: world * rows * plan.pitch], src, async_op=True))
        if gather is not None:
            gather.wait()
        for wk in works:
            wk.wait()
        return plan.cc

    def timed(n_steps, n_warm, weights=None, with_clocks=False):
        """(ms per step [max over ranks], main-kernel ms per step on this rank, launches per step, clocks)."""
        for _ in range(n_warm):
            step_fn(weights)
        barrier()
        sampler = ClockSampler(local_rank)
        if with_clocks and rank == 0:
            sampler.start()
        lib.bpmf_set_kernel_timing(1)
        lib.bpmf_main_kernel_times((C.c_float * 2048)(), 2048)  # drain earlier records
        launches0 = lib.bpmf_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n_steps):
            step_fn(weights)
        ev1.record()
        barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        launches = int(lib.bpmf_launch_count() - launches0)
        buf = (C.c_float * 2048)()
        nk = lib.bpmf_main_kernel_times(buf, 2048)
        lib.bpmf_set_kernel_timing(0)
        # one record per kernel launch (= per template batch at N > 1): the step's kernel time is their SUM / steps
        complete = nk == n_steps * len(schedule)
        kernel_ms = float(sum(buf[i] for i in range(nk))) / n_steps if (nk and complete) else None
        clocks = sampler.stop() if (with_clocks and rank == 0) else None
        t_ms = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms.item()) / n_steps, kernel_ms, launches // max(1, n_steps), clocks

    ms_per_step, kernel_ms, launches, clocks = timed(args.steps, args.warmup, with_clocks=True)
    prec_name = {1: "fp32", 2: "fp16", 3: "fp16x3"}[plan.last_precision]
    value = world * units / (ms_per_step * 1e-3) / 1e9
    if clocks is not None and kernel_ms is None:
        clocks["note"] = "main-kernel event ring incomplete"

    # ---- parity (untimed): strided comparison with the oracle over the whole day, every template tile
    parity = None
    data_host = None
    if rank == 0 and not args.no_parity:
        data_host = data.cpu().numpy()
        parity = guarded(strided_parity, torch, plan, tp, mv, w, data_host, planted, N, L)

    # ---- exchange spot check (untimed): the last batch of every rank arrived in this rank's gather buffer
    if world > 1 and gather is not None:
        n_batches = len(schedule)
        t0, rows = schedule[-1]
        mine = plan.cc[t0: t0 + rows].double().sum().reshape(1)
        sums = torch.empty(world, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(sums, mine)
        torch.cuda.synchronize()
        got = gather.gathered((n_batches - 1) % 2, rows, plan.pitch).double().sum(dim=(1, 2))
        ok = bool(torch.allclose(got, sums, rtol=1e-9, atol=1e-6))
        if isinstance(parity, dict):
            parity["gathered_slabs_match"] = ok

    # ---- SURVEY 8(d) second weighting: the 10 smallest-moveout stations of every template (N = 1)
    variants = None
    if world == 1 and not args.no_variants:
        def second_variant():
            w2 = ten_closest_weights(torch, mv, keep=min(10, S))
            ms2, k2, _, _ = timed(3, 2, weights=w2)
            units2 = T * int((w2[0] != 0).sum().item()) * n_corr
            rec = {"weights": "10 smallest-moveout stations per template, row-normalised",
                   "weighted_channels_per_template": int((w2[0] != 0).sum().item()), "ms_per_step": ms2,
                   "value": units2 / (ms2 * 1e-3) / 1e9, "unit": "GSamples/s (weighted channels only, SURVEY 8d)",
                   "precision": {1: "fp32", 2: "fp16", 3: "fp16x3"}[plan.last_precision], "kernel_ms": k2}
            if k2:
                rec["tensor_frac"] = 2.0 * L * units2 / (k2 * 1e-3) / 1e12 / pk["tensor_tflops_sustained"]
            if data_host is not None:
                rec["parity"] = strided_parity(torch, plan, tp, mv, w2, data_host, planted, N, L, n_windows=24)
            return rec

        variants = {"ten_closest_stations": guarded(second_variant)}
        step_fn()  # leave the all-stations result in the slab for the legs below
        torch.cuda.synchronize()
    del data_host

    # ---- fused detection (the product path when the caller keeps the CC on the device): traces + templates from
    # pinned host memory, matched filter, rms threshold + declustering + compaction on the device, only
    # the detections come back -- what replaces similarity_search.py:526-666.  At N > 1 every rank does this for its
    # own templates and the detections are all-gathered.
    fused = None
    if not args.no_e2e:
        def fused_leg():
            from seismic_bpmf_b200 import sharded

            hp = [torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_) for t_ in (tp, mv, w, data)]
            gauss = torch.randn(500, generator=torch.Generator().manual_seed(1)).pin_memory()
            win = 180000  # 30 min @ 100 Hz, overlap 0.25 as BPMF's callers pass (similarity_search.py:552)
            det = dev.DetectBatches(1, T, n_corr, win // 2, int(0.75 * win), capacity=1 << 22, device=device)
            times = []
            for it in range(4):
                barrier()
                t0_ = time.perf_counter()
                dv = [h_.to(device, non_blocking=True) for h_ in hp]
                plan.run(*dv)
                det.run(0, plan.cc, gauss.to(device, non_blocking=True), 8.0, search_win=6000, mode="select")
                s_, i_, v_, th_ = det.collect(series_offsets=[rank * T])
                if world > 1:
                    s_, i_, v_, th_ = sharded.gather_detections(s_, i_, v_, th_, device=device)
                torch.cuda.synchronize()
                if it:
                    times.append(time.perf_counter() - t0_)
            t_f = torch.tensor([float(np.median(times))], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t_f, op=dist.ReduceOp.MAX)
            dt_f = float(t_f.item())
            hit = all(((s_ == t_) & (i_ == i0)).any() for t_, i0 in planted.items())
            return {"value": world * units / dt_f / 1e9, "unit": "GSamples/s", "ms_per_step": 1e3 * dt_f,
                    "detections": int(len(s_)), "planted_detections_found": bool(hit),
                    "h2d_bytes_per_step": int(sum(h_.numel() * h_.element_size() for h_ in hp)),
                    "d2h_bytes_per_step": int(16 * len(s_) + 4),
                    "what": "H2D inputs + matched filter + device threshold/declustering/compaction + D2H"
                            + (" + all-gather" if world > 1 else "") + " of detections"}

        fused = guarded(fused_leg)

    # ---- end-to-end arm: the reference-facing call with pinned host buffers.  At N > 1 every rank's host process
    # holds its own block of templates and receives its own (T/N, n_corr) block of CC.
    e2e = None
    if not args.no_e2e:
        h = {}
        host_kind = "pinned"
        try:
            for name, ten in (("tp", tp), ("mv", mv), ("w", w), ("data", data)):
                pinned = torch.empty(ten.shape, dtype=ten.dtype, pin_memory=True)
                pinned.copy_(ten)
                h[name] = pinned.numpy()
            out_pinned = torch.empty((T, n_corr), dtype=torch.float32, pin_memory=True)
            out_np = out_pinned.numpy()
        except RuntimeError:  # not enough lockable host memory on this box: pageable buffers, same bytes
            host_kind = "pageable"
            h = {name: ten.cpu().numpy() for name, ten in (("tp", tp), ("mv", mv), ("w", w), ("data", data))}
            out_np = np.empty((T, n_corr), dtype=np.float32)
        h2d = sum(h[k].nbytes for k in h)
        d2h = out_np.nbytes
        del plan  # the host entry point allocates its own device buffers
        gbufs = gather = None
        torch.cuda.empty_cache()
        fmf.matched_filter(h["tp"], h["mv"], h["w"], h["data"], step, arch="gpu", precision=args.precision,
                           check_zeros=False, out=out_np)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fmf.matched_filter(h["tp"], h["mv"], h["w"], h["data"], step, arch="gpu", precision=args.precision,
                               check_zeros=False, out=out_np)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        t_e = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * units / float(t_e.item()) / 1e9, "unit": "GSamples/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "ms_per_step": 1e3 * float(t_e.item()), "host_buffers": host_kind,
               "aggregate_d2h_gbs": world * d2h / float(t_e.item()) / 1e9}
        if world > 1:
            e2e["note"] = ("every rank's process downloads its own (T/N, n_corr) block into pinned host memory: the figure "
                           "is bounded by the host's aggregate ingest (aggregate_d2h_gbs), see e2e_fused_detection for the "
                           "path that returns detections instead of the CC slab")
        # BPMF's template-chunk loop (similarity_search.py:773-790): three successive calls on the SAME day of traces,
        # without and with fast_matched_filter.hold_traces (one upload + staging of the day instead of one per chunk)
        if world == 1:
            def chunk_loop():
                def three_calls(dd):
                    t0_ = time.perf_counter()
                    for _ in range(3):
                        fmf.matched_filter(h["tp"], h["mv"], h["w"], dd, step, arch="gpu", precision=args.precision,
                                           check_zeros=False, out=out_np)
                    return (time.perf_counter() - t0_) / 3
                plain = three_calls(h["data"])
                with fmf.hold_traces(h["data"]) as dd:
                    three_calls(dd)  # first call inside the block stages the traces
                    held = three_calls(dd)
                return {"ms_per_chunk_plain": 1e3 * plain, "ms_per_chunk_held": 1e3 * held,
                        "what": "3 chunks of %d templates on one day of traces through fast_matched_filter.matched_filter; held = "
                                "inside fast_matched_filter.hold_traces (no re-upload / re-staging of the %.2f GB day)"
                                % (T, h["data"].nbytes / 1e9)}
            e2e["template_chunk_loop"] = guarded(chunk_loop)
        bpmf_release = getattr(lib, "bpmf_release_cache", None)
        if bpmf_release:
            bpmf_release()
        del h, out_np
    else:
        del plan
        gbufs = gather = None
    del tp, mv, w, data
    torch.cuda.empty_cache()

    # ---- BASELINE.json configs[3]: 5000 x 32 x 3 strong-scaled over the ranks
    cfg4 = None
    if not args.no_cfg4:
        import bench_cfg4

        cfg4 = guarded(bench_cfg4.run, rank, world, device, steps=args.cfg4_steps, warmup=1, precision=args.precision,
                       T=(None if args.templates == CFG2["T"] else 10 * args.templates // 125 * 125 or 125),
                       samples=(None if args.samples == CFG2["N"] else args.samples))
        torch.cuda.empty_cache()

    # ---- second hot path: beamforming / backprojection, BASELINE.json configs[2] scaled to a few seconds (the full
    # 1e6-source x 1-day grid is ~1 minute per pass: profiles/): device figure + shared-memory roofline,
    # the host-API call BPMF makes, the CPU oracle; at N > 1 the source-sharded strong-scaling leg
    beam = None
    if not args.no_beam:
        import bench_beam

        if world == 1:
            beam = guarded(bench_beam.run, K=50000, N=432000, S=32, keep=32, steps=3, warmup=2, device_index=local_rank)
            if isinstance(beam, dict) and "error" not in beam:
                beam["ten_closest_stations"] = guarded(
                    lambda: {k: v for k, v in bench_beam.run(K=50000, N=432000, S=32, keep=10, steps=3, warmup=2,
                                                             device_index=local_rank).items()
                             if k in ("ms_per_step", "kernel_ms", "G_source_samples_per_s", "G_gather_fma_per_s", "roofline")})
                beam["e2e"] = guarded(bench_beam.e2e_host_api, K=50000, N=432000, S=32, steps=2)
                if not args.no_cpu_baseline:
                    beam["cpu_baseline"] = guarded(bench_beam.cpu_baseline, seconds=8.0, S=32)
        beam_sh = guarded(bench_beam.run_sharded, K=200000, N=432000, S=32, steps=3, warmup=2, rank=rank, world=world,
                          device_index=local_rank)
        if beam is None:
            beam = {}
        beam["sharded"] = beam_sh

    callers = None
    if world == 1 and args.callers:
        callers = guarded(callers_figures)

    if rank != 0:
        return
    libc = None
    if world == 1 and not args.no_libc:
        import bench_libc

        def libc_leg():
            with bench_libc._quiet():  # the reference (and the drop-in, like it) print progress with printf
                return bench_libc.run(device_index=local_rank)

        libc = guarded(libc_leg)

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        rate, cores, sample, build = cpu_sample_throughput(args.cpu_seconds, S, Cc, L)
        cpu_base = {"value": rate, "unit": "GSamples/s", "cores": cores, "kind": "port", "sample": sample,
                    "build": build, "cpu_model": cpu_model()}

    flops = 2.0 * L * units
    roof = None
    if kernel_ms:
        sec = kernel_ms * 1e-3
        tflops = flops / sec / 1e12
        if plan_is_tensor(prec_name):
            mult = 3.0 if prec_name == "fp16x3" else 1.0
            roof = {"bound": "tensor", "achieved": tflops * mult, "peak": pk["tensor_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": tflops * mult / pk["tensor_tflops_sustained"],
                    "traffic": None if os.environ.get("BPMF_TC_PAIR", "0") == "1" else measured_traffic("mf_tc_kernel", T, S, Cc, L, N),
                    "peak_source": pk["source"] + " (sustained bf16 cuBLAS; fp16 runs at the same tensor rate)",
                    "algorithmic_tflops": tflops, "kernel_ms": kernel_ms, "kernel": "mf_tc_kernel",
                    "kernel_ms_is": "sum of the main-kernel launches of one step on rank 0 (%d per step)" % len(schedule),
                    "frac_of_burst_peak": tflops * mult / pk["tensor_tflops"]}
        else:
            sm_mhz = (clocks or {}).get("sm_mhz") or pk["sm_max_mhz"]
            peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
            roof = {"bound": "fp32-fma", "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak,
                    "traffic": None, "peak_source": "148 SM x 128 lanes x 2 x measured SM clock (no tensor cores on this path)",
                    "kernel_ms": kernel_ms, "kernel": "mf_fp32_step1_kernel"}
        ab = algorithmic_bytes(T, S, Cc, L, N, n_corr)
        roof["hbm"] = {"algorithmic_bytes": ab, "achieved_gbs": ab / sec / 1e9, "peak_gbs": pk["hbm_gbs"],
                       "frac": ab / sec / 1e9 / pk["hbm_gbs"]}

    line = {
        "metric": "CC-sum GSamples/s (templates x traces x steps)", "value": value, "unit": "GSamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if prec_name == "fp32" else "f16",
        "data": "synthetic",
        "config": {"workload": "matched_filter cfg2 per GPU: %d templates x %d stations x %d ch, L=%d, N=%d, step 1"
                   % (T, S, Cc, L, N), "precision": prec_name, "templates_total": T * world,
                   "l2": "inputs (%.2f GB) and output (%.2f GB) exceed the 126 MB L2" % (
                       (S * Cc * N * 4 + T * S * Cc * L * 4) / 1e9, T * n_corr * 4 / 1e9),
                   "collective": collective,
                   "kernel": ("CTA pairs (tcgen05 cta_group::2)" if os.environ.get("BPMF_TC_PAIR", "0") == "1" else "one CTA per SM")
                   + (", shuffle-free epilogue (16x128b TMEM loads, red.v2)" if os.environ.get("BPMF_TC_EPI", "1") != "0"
                      else ", r1 epilogue (32x32b TMEM loads, butterfly transpose, red.v4)")},
        "roofline": roof, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": launches * args.steps, "clocks": clocks,
        "parity": parity, "variants": variants, "e2e_fused_detection": fused, "cfg4": cfg4, "beamform": beam, "libc": libc,
    }
    if callers is not None:
        line["callers"] = callers
    print(json.dumps(line))


def plan_is_tensor(prec_name):
    return prec_name in ("fp16", "fp16x3")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        # keep NCCL's own stdout chatter out of the JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        try:
            run_ours(args, rank, world, local_rank)
        finally:
            dist.destroy_process_group()
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
