"""BASELINE.json configs[4] -- a full BPMF day on the N GPUs of one box:

    detection traces -> backprojection of a 1e6-point source grid (SOURCE-sharded, all-gather of (max, argmax))
    -> threshold + picks -> templates cut at the picks -> matched filter of the ~2000 templates over the day
    (TEMPLATE-sharded, threshold + declustering on each shard, all-gather of the detections)

at 32 stations x 3 components, one day @ 100 Hz, with per-stage times (wall clock around a device synchronise and a
barrier, max over ranks).  The call order is BPMF's: template_search.py:508-572 (backproject) and :574-673 (find
detections), dataset.py:3452-3475 (template windows), similarity_search.py:726-807 (template-chunk loop) and :548-666
(detections).  Components N, E carry the S phase and Z the P phase (BPMF's default phase_on_comp, dataset.py:1713).

    torchrun --nproc-per-node 8 tools/full_day_sharded.py                  # BASELINE size
    python tools/full_day_sharded.py --sources 125000 --events 250         # one GPU, 1/8 of the grid and templates
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stations", type=int, default=32)
    ap.add_argument("--sources", type=int, default=1000000)
    ap.add_argument("--samples", type=int, default=8640000)
    ap.add_argument("--events", type=int, default=2000, help="planted events = templates to extract")
    ap.add_argument("--sr", type=float, default=100.0)
    ap.add_argument("--max-templates", type=int, default=2000)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    from seismic_bpmf_b200 import clib, features, sharded, template_search

    S, Cc, N, K, sr = a.stations, 3, a.samples, a.sources, a.sr
    phase_on_comp = [1, 1, 0]  # N, E -> S ; Z -> P
    times = {}

    def stage(name, t0):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times[name] = float(dt.item())
        return time.perf_counter()

    # ---------------------------------------------------------------- synthetic day (same on every rank), untimed
    g = torch.Generator(device=dev)
    g.manual_seed(20240726)
    n1 = int(round(K ** (1 / 3)))
    while n1 ** 3 < K:
        n1 += 1
    ax = torch.linspace(0.02, 0.98, n1, device=dev)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)[:K] * torch.tensor([100.0, 100.0, 30.0], device=dev)
    sta = torch.cat([torch.rand((S, 2), generator=g, device=dev) * 100.0, torch.zeros((S, 1), device=dev)], 1)
    tp_s = (grid[:, None, :] - sta[None]).norm(dim=-1) / 6.0  # P travel time, seconds
    tt = torch.stack([tp_s, tp_s * 1.73], -1).float().contiguous()  # (K, S, 2)
    traces = torch.randn((S, Cc, N), generator=g, device=dev, dtype=torch.float32)
    E = a.events
    wl = int(2 * sr)
    slot = (N - int(80 * sr)) // max(1, E)
    ev_src = torch.randint(0, K, (E,), generator=g, device=dev)
    ev_t0 = (torch.arange(E, device=dev) * slot + int(5 * sr) + torch.randint(0, max(1, slot // 4), (E,), generator=g, device=dev)).long()
    ev_amp = 6.0 + 4.0 * torch.rand((E,), generator=g, device=dev)
    samp = torch.sign(tt) * torch.floor(torch.abs(tt * np.float32(sr)) + np.float32(0.2))  # utils.sec_to_samp
    taper = torch.hann_window(wl, periodic=False, device=dev)
    for e0 in range(0, E, 100):  # wavelets: event- and channel-specific noise bursts, 2 s long
        e1 = min(E, e0 + 100)
        k = ev_src[e0:e1]
        rel = samp[k] - samp[k].amin(dim=(1, 2), keepdim=True)  # (e, S, 2) relative to the first arrival
        relc = rel[:, :, phase_on_comp].long()  # (e, S, C)
        wav = torch.randn((e1 - e0, S, Cc, wl), generator=g, device=dev) * taper
        wav = wav / wav.pow(2).mean(dim=-1, keepdim=True).sqrt() * ev_amp[e0:e1, None, None, None]
        idx = (ev_t0[e0:e1, None, None] + relc)[..., None] + torch.arange(wl, device=dev)  # (e, S, C, wl)
        flat = (torch.arange(S, device=dev)[None, :, None, None] * Cc + torch.arange(Cc, device=dev)[None, None, :, None]) * N + idx
        traces.view(-1).index_add_(0, flat.reshape(-1), wav.reshape(-1))
    del samp
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---------------------------------------------------------------- 1. detection traces
    t0 = time.perf_counter()
    t_all = t0
    feat, avail = features.saturated_envelopes_dev(traces)  # the day is already on the device
    t0 = stage("detection_traces_s", t0)
    # ---------------------------------------------------------------- 2. moveout table (int32 samples, relative to first)
    mv_bp, m2t = template_search.moveouts_from_travel_times_dev(tt, sr, relative_to_first=True)
    t0 = stage("moveout_table_s", t0)
    # ---------------------------------------------------------------- 3. backprojection, source-sharded
    wp = torch.zeros((S, Cc, 2), device=dev)
    wp[:, 0, 1] = 1.0
    wp[:, 1, 1] = 1.0
    wp[:, 2, 0] = 1.0
    ws = torch.full((K, S), 1.0 / S, device=dev)
    maxbeam, argmax = sharded.beamform_max_sharded(feat, mv_bp, wp, ws)
    t0 = stage("backprojection_s", t0)
    # ---------------------------------------------------------------- 4. threshold + picks (replicated, small)
    mb_h, am_h = maxbeam.cpu().numpy(), argmax.cpu().numpy()
    thr = features.time_dependent_threshold(mb_h, int(1800 * sr), overlap=0.75, CNR_threshold=10.0)
    search_win = int(mv_bp.max().item()) + 1
    picks = np.flatnonzero(clib.select_cc_indexes(mb_h, np.float32(thr), search_win))
    if len(picks) > a.max_templates:  # keep the strongest
        keep = np.argsort(mb_h[picks])[::-1][: a.max_templates]
        picks = np.sort(picks[keep])
    srcs = am_h[picks]
    t0 = stage("threshold_and_picks_s", t0)
    # ---------------------------------------------------------------- 5. templates at the picks
    L, offset = int(8 * sr), int(1 * sr)
    T = len(picks)
    data = traces / traces.std(dim=-1, keepdim=True)  # MatchedFilter.set_data, similarity_search.py:181-185
    picks_d = torch.from_numpy(picks).to(dev)
    srcs_d = torch.from_numpy(srcs.astype(np.int64)).to(dev)
    tmv = (mv_bp[srcs_d][:, :, phase_on_comp] - offset).to(torch.int32).contiguous()  # (T, S, C)
    templates = torch.zeros((T, S, Cc, L), device=dev, dtype=torch.float32)
    for a0 in range(0, T, 64):
        a1 = min(T, a0 + 64)
        j = (picks_d[a0:a1, None, None] + tmv[a0:a1].long())[..., None] + torch.arange(L, device=dev)
        ok = (j >= 0) & (j < N)
        templates[a0:a1] = torch.where(ok, torch.gather(data[None].expand(a1 - a0, S, Cc, N), 3, j.clamp(0, N - 1)),
                                       torch.zeros((), device=dev))
    w = torch.full((T, S, Cc), 1.0 / (S * Cc), device=dev)
    t0 = stage("template_extraction_s", t0)
    # ---------------------------------------------------------------- 6. matched filter + detection, template-sharded
    gauss = torch.randn(500, generator=torch.Generator().manual_seed(1)).to(dev)
    det = (np.zeros(0, np.int64),) * 4
    if T > 0:
        det = sharded.matched_filter_detect_sharded(templates, tmv, w, data, gauss, 8.0, int(1800 * sr), overlap=0.25,
                                                    search_win=L, batch=64)
    t0 = stage("matched_filter_and_detection_s", t0)
    times["total_s"] = time.perf_counter() - t_all

    if rank == 0:
        ev_first = ev_t0.cpu().numpy()
        hit = int(sum(np.any(np.abs(picks - i0) <= int(1.5 * sr)) for i0 in ev_first))
        s_, i_, v_, t_ = det
        self_det = int(sum(np.any((s_ == t) & (np.abs(i_ - picks[t]) <= 2)) for t in range(T)))
        n_corr = N - L + 1
        print(json.dumps({
            "workload": "full day: %d stations x %d components x %d samples @ %g Hz; %d-point source grid x 2 phases; "
                        "%d planted events; templates %d samples" % (S, Cc, N, sr, K, E, L),
            "n_gpus": world, "stages_s": {k: round(v, 4) for k, v in times.items()},
            "backprojection_G_source_samples_per_s": K * N / times["backprojection_s"] / 1e9,
            "matched_filter_GSamples_per_s": T * S * Cc * n_corr / max(times["matched_filter_and_detection_s"], 1e-9) / 1e9,
            "picks": int(len(picks)), "planted_events_picked": hit, "templates": int(T),
            "matched_filter_detections": int(len(s_)), "templates_that_detect_their_own_event": self_det,
            "sharding": "backprojection: sources block-partitioned, all-gather of (max, argmax); matched filter: templates "
                        "block-partitioned, threshold + declustering per shard, all-gather of detections"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
