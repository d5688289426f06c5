"""CPU, world_size 2, gloo: the multi-GPU host logic (partition, padded all-gather, arg-max merge)
with the oracle injected as the per-shard compute step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seismic_bpmf_b200 import sharded
from tests import _synth


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 8, 500, 5000, 1000003):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_max_argmax_ties_to_lowest_index():
    m = torch.tensor([[1.0, 2.0, 0.0], [1.0, 3.0, 0.0], [0.5, 3.0, 0.0]])
    a = torch.tensor([[7, 1, 0], [3, 9, 5], [1, 4, 2]], dtype=torch.int32)
    best, arg = sharded.merge_max_argmax(m, a)
    assert best.tolist() == [1.0, 3.0, 0.0]
    assert arg.tolist() == [3, 4, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu

    try:
        # ---- matched filter: 5 templates over 2 ranks (uneven shards -> padded gather)
        tp, mv, w, d, _ = _synth.mf_case(5, 3, 2, 40, 3000, seed=11)

        def mf_compute(t, m, ww, dd, step, normalize, precision):
            return torch.from_numpy(cpu.matched_filter(t.numpy(), m.numpy(), ww.numpy(), dd.numpy(), step,
                                                       normalize=normalize))

        got = sharded.matched_filter_sharded(torch.from_numpy(tp), torch.from_numpy(mv), torch.from_numpy(w),
                                             torch.from_numpy(d), step=2, compute=mf_compute)
        want = cpu.matched_filter(tp, mv, w, d, 2)
        ok_mf = np.array_equal(got.numpy(), want)
        local, (lo, hi) = sharded.matched_filter_sharded(torch.from_numpy(tp), torch.from_numpy(mv),
                                                         torch.from_numpy(w), torch.from_numpy(d), step=2,
                                                         compute=mf_compute, gather=False)
        ok_local = np.array_equal(local.numpy(), want[lo:hi])

        # ---- beamform: 7 sources over 2 ranks, duplicated source across the shard boundary (tie)
        f, tau, wp, ws = _synth.beam_case(7, 4, 2, 2, 1500, seed=12)
        tau[4], ws[4] = tau[2], ws[2]

        def beam_compute(ff, tt, wpp, wss, oob, off):
            mb, am = cpu.beamform(ff.numpy(), tt.numpy(), wpp.numpy(), wss.numpy(), out_of_bounds=oob)
            return torch.from_numpy(mb), torch.from_numpy(am + off)

        mb, am = sharded.beamform_max_sharded(torch.from_numpy(f), torch.from_numpy(tau), torch.from_numpy(wp),
                                              torch.from_numpy(ws), compute=beam_compute)
        mb_w, am_w = cpu.beamform(f, tau, wp, ws)
        ok_beam = np.array_equal(mb.numpy(), mb_w) and np.array_equal(am.numpy(), am_w)

        # reduce="none": the (K/G, N) blocks concatenated in source order (7 sources over 2 ranks: uneven shards)
        def beam_full_compute(ff, tt, wpp, wss, oob):
            return torch.from_numpy(cpu.beamform(ff.numpy(), tt.numpy(), wpp.numpy(), wss.numpy(), out_of_bounds=oob,
                                                 reduce="none"))

        full = sharded.beamform_none_sharded(torch.from_numpy(f), torch.from_numpy(tau), torch.from_numpy(wp),
                                             torch.from_numpy(ws), compute=beam_full_compute)
        ok_beam = ok_beam and np.array_equal(full.numpy(), cpu.beamform(f, tau, wp, ws, reduce="none"))
        # ---- shard-local detection + all-gather of the detections: 5 templates over 2 ranks
        from oracle import libc_numpy

        wn = np.random.default_rng(3).standard_normal(500).astype(np.float32)
        tp3, mv3, w3, d3, planted = _synth.mf_case(5, 3, 2, 40, 6000, seed=14)

        def mf1(t, m, ww, dd, step, normalize, precision):
            return torch.from_numpy(cpu.matched_filter(t.numpy(), m.numpy(), ww.numpy(), dd.numpy(), step))

        def detect(local):
            rows, idx, val, thr = [], [], [], []
            for r, cc in enumerate(local.numpy()):
                th, _ = libc_numpy.time_dependent_threshold(cc, wn, 6.0, 1000 // 2, int(0.75 * 1000))
                sel = libc_numpy.select_cc_indexes(cc, th, 50)
                k = np.flatnonzero(sel)
                rows.append(np.full(len(k), r)); idx.append(k); val.append(cc[k]); thr.append(th[k])
            return (np.concatenate(rows), np.concatenate(idx).astype(np.int32), np.concatenate(val),
                    np.concatenate(thr))

        args3 = [torch.from_numpy(a) for a in (tp3, mv3, w3, d3)]
        s_, i_, v_, t_ = sharded.matched_filter_detect_sharded(*args3, torch.from_numpy(wn), 6.0, 1000, overlap=0.25,
                                                              search_win=50, compute=mf1, detect=detect)
        ws_, wi_, wv_, wt_ = detect(mf1(*args3, 1, "short", "auto"))  # the same on one rank, all templates
        order = np.lexsort((wi_, ws_))
        ok_det = (np.array_equal(s_, ws_[order]) and np.array_equal(i_, wi_[order]) and np.array_equal(v_, wv_[order])
                  and all(((s_ == t) & (i_ == i0)).any() for t, i0 in planted.items()))
        ret[rank] = (ok_mf, ok_local, ok_beam and ok_det)
    finally:
        dist.destroy_process_group()


def test_sharded_paths_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        assert ret[rank] == (True, True, True), (rank, ret[rank])


def test_more_ranks_than_items_single_process():
    # world 1 degenerate paths (no process group)
    tp, mv, w, d, _ = _synth.mf_case(1, 2, 1, 16, 400, seed=13, zero_gap=False)
    from oracle import cpu

    got = sharded.matched_filter_sharded(
        torch.from_numpy(tp), torch.from_numpy(mv), torch.from_numpy(w), torch.from_numpy(d),
        compute=lambda t, m, ww, dd, s, n, p: torch.from_numpy(cpu.matched_filter(t.numpy(), m.numpy(), ww.numpy(), dd.numpy(), s)))
    assert np.array_equal(got.numpy(), cpu.matched_filter(tp, mv, w, d, 1))
