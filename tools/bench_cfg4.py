"""BASELINE.json configs[3]: matched filter, 5000 templates x 32 stations x 3 channels, one day @ 100 Hz,
template-sharded over the ranks (STRONG scaling: the total work is fixed, rank r takes a block of templates).

Two exchange modes, both timed on the device (CUDA events, max over ranks):
  * "detections" -- the product default: each rank thresholds + declusters its own CC rows on the device
    (8 x rms in 30 min windows, overlap 0.25, 60 s search window: BPMF/similarity_search.py:548-666) and only
    the detections are all-gathered;
  * "cc_gather" -- the contract option BASELINE.json names: the per-shard CC slabs are all-gathered batch by
    batch through NVLink peer memory while the next batch computes (the 172.8 GB gathered slab does not fit one
    GPU, so the gather buffer is recycled every two batches).
The (T, n_corr) slab never exists in full anywhere: every rank holds `batch` templates of CC at a time.

`run()` is what bench.py calls for the "cfg4" object of its JSON line at every N.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG4 = dict(T=5000, S=32, C=3, L=800, N=8640000)
GROUP = 125  # templates are generated in groups keyed by their GLOBAL group id: the same 5000 templates at every N


def _templates(lo, hi, S, Cc, L, N, data, device, seed):
    """Templates [lo, hi) of the global set: noise, demeaned / normalised; template 0 of every group is cut from
    the traces at a known index (a planted detection at global template id g * GROUP, index 5000 + 37 * g)."""
    tps, mvs, planted = [], [], {}
    assert lo % GROUP == 0
    for g in range(lo // GROUP, (hi + GROUP - 1) // GROUP):
        gen = torch.Generator(device=device)
        gen.manual_seed(seed + 17 * g)
        n = min(GROUP, hi - g * GROUP)
        tp = torch.randn((n, S, Cc, L), generator=gen, device=device, dtype=torch.float32)
        tp -= tp.mean(dim=-1, keepdim=True)
        tp /= tp.std(dim=-1, keepdim=True)
        mv = torch.randint(-100, 3001, (n, S, Cc), generator=gen, device=device, dtype=torch.int32)
        i0 = 5000 + 37 * g
        idx = (i0 + mv[0].long()).reshape(S, Cc, 1) + torch.arange(L, device=device).reshape(1, 1, L)
        tp[0] = torch.gather(data, 2, idx)
        planted[g * GROUP] = i0
        tps.append(tp)
        mvs.append(mv)
    tp, mv = torch.cat(tps).contiguous(), torch.cat(mvs).contiguous()
    w = torch.full(mv.shape, 1.0 / (S * Cc), device=device, dtype=torch.float32)
    return tp, mv, w, planted


def run(rank, world, device, steps=2, warmup=1, precision="auto", seed=20240726, T=None, samples=None,
        modes=("detections", "cc_gather")):
    import torch.distributed as dist

    from seismic_bpmf_b200 import _lib, device as dev, sharded

    lib = _lib.load()
    T = CFG4["T"] if T is None else int(T)
    S, Cc, L = CFG4["S"], CFG4["C"], CFG4["L"]
    N = CFG4["N"] if samples is None else int(samples)
    per = (T // GROUP + world - 1) // world * GROUP  # block partition in whole groups
    lo, hi = min(T, rank * per), min(T, (rank + 1) * per)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)  # the traces are the same on every rank
    data = torch.randn((S, Cc, N), generator=gen, device=device, dtype=torch.float32)
    data[S // 2, :, N // 3: N // 3 + 10000] = 0.0
    data /= data.std(dim=-1, keepdim=True)
    tp, mv, w, planted = _templates(lo, hi, S, Cc, L, N, data, device, seed)
    n_local = hi - lo
    batch = 125 if world <= 2 else 64
    batch = min(batch, max(1, n_local))
    n_batches = (per + batch - 1) // batch  # every rank runs the same number of batches (barriers)
    plan = dev.MatchedFilterPlan(2 * batch, S, Cc, L, N, precision=precision, device=device, batch=batch)
    n_corr, pitch = plan.n_corr, plan.pitch
    units = T * S * Cc * n_corr  # whole job, per step
    window = 180000
    half, shift = window // 2, int(0.75 * window)
    gauss = torch.randn(500, generator=torch.Generator().manual_seed(1)).to(device)
    det = dev.DetectBatches(n_batches, batch, n_corr, half, shift, capacity=1 << 16, device=device)
    gather = None
    if world > 1 and "cc_gather" in modes:
        try:
            gather = sharded.PeerSlabGather(batch * pitch, device)
        except Exception as exc:  # noqa: BLE001
            gather = None
            gather_err = str(exc)[:160]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    slot_free = [torch.cuda.Event(), torch.cuda.Event()]

    def step(mode):
        plan.prepare(data)
        for b in range(n_batches):
            t0 = b * batch
            n_here = max(0, min(batch, n_local - t0))
            slot = (b % 2) * batch
            if mode == "cc_gather" and gather is not None and b >= 2:
                torch.cuda.current_stream().wait_event(slot_free[b % 2])  # the copies out of this slot are done
            if n_here > 0:
                plan.run_batch(tp, mv, w, data, t0, n_here, cc_row0=slot)
            if mode == "detections":
                if n_here > 0:
                    det.run(b, plan.cc[slot: slot + n_here], gauss, 8.0, search_win=6000, mode="select")
            elif gather is not None:
                gather.push(b % 2, plan.cc[slot: slot + n_here].reshape(-1) if n_here > 0 else plan.cc[:0].reshape(-1))
                slot_free[b % 2].record(gather.stream)
        if mode == "detections":
            # the exchange: counts, then the padded records -- bytes, not gigabytes
            cnt, ser, idx, val, thr = det.packed()
            if world > 1:
                m = torch.clamp(cnt.max().reshape(1).long(), max=det.capacity)
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
                k = int(m.item())
                if k > 0:
                    mine = torch.stack([ser[:, :k].float(), idx[:, :k].float(), val[:, :k], thr[:, :k]], 0).contiguous()
                    allr = torch.empty((world,) + tuple(mine.shape), dtype=torch.float32, device=device)
                    dist.all_gather_into_tensor(allr.view(-1), mine.view(-1))
                allc = torch.empty((world, n_batches), dtype=torch.int32, device=device)
                dist.all_gather_into_tensor(allc.view(-1), cnt)
        elif gather is not None:
            gather.wait()

    out = {"workload": "matched_filter cfg4: %d templates x %d stations x %d ch, L=%d, N=%d, step 1; templates "
                       "block-partitioned over %d rank(s) (%d per rank), %d templates of CC resident per rank at a time"
                       % (T, S, Cc, L, N, world, per, 2 * batch),
           "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup, "unit": "GSamples/s",
           "templates_per_rank": per, "batch": batch}
    for mode in modes:
        if mode == "cc_gather" and world > 1 and gather is None:
            out[mode] = {"error": "peer gather unavailable: " + gather_err}
            continue
        for _ in range(warmup):
            step(mode)
        barrier()
        lib.bpmf_set_kernel_timing(1)
        lib.bpmf_main_kernel_times((C.c_float * 2048)(), 2048)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(mode)
        e1.record()
        barrier()
        buf = (C.c_float * 2048)()
        nk = lib.bpmf_main_kernel_times(buf, 2048)
        lib.bpmf_set_kernel_timing(0)
        kernel_ms = sum(buf[i] for i in range(nk)) / steps if nk else None
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rec = {"ms_per_step": ms, "value": units / (ms * 1e-3) / 1e9, "main_kernel_ms_per_step_rank0": kernel_ms}
        if mode == "detections":
            s_, i_, v_, t_ = det.collect(series_offsets=[lo + b * batch for b in range(n_batches)])
            found = all(((s_ == t) & (i_ == i0)).any() for t, i0 in planted.items())
            ok = torch.tensor([1 if found else 0, len(s_)], device=device, dtype=torch.int64)
            if world > 1:
                allok = torch.empty((world, 2), dtype=torch.int64, device=device)
                dist.all_gather_into_tensor(allok.view(-1), ok)
                ok = torch.stack([allok[:, 0].min(), allok[:, 1].sum()])
            rec.update({"detections_total": int(ok[1].item()), "planted_detections_found_on_every_rank": bool(ok[0].item()),
                        "exchange": "all-gather of detection records (template, index, cc, threshold)"})
        else:
            rec["exchange"] = ("all-gather of %d-template CC slabs pushed to peer memory over NVLink by the copy engines, "
                               "overlapped with the next batch" % batch) if world > 1 else "none (single GPU)"
            if world > 1:
                rec["gathered_bytes_per_rank_per_step"] = int((world - 1) * per * pitch * 4)
        out[mode] = rec
    out["value"] = out[modes[0]].get("value")
    out["ms_per_step"] = out[modes[0]].get("ms_per_step")
    out["precision"] = {1: "fp32", 2: "fp16", 3: "fp16x3"}.get(plan.last_precision)
    del plan, det, gather, data, tp
    torch.cuda.empty_cache()
    return out
