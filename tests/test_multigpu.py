"""GPU, >= 2 devices: the sharded paths over NCCL and the copy-engine peer gather (skipped on 1-GPU boxes)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import _synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import cpu
    from seismic_bpmf_b200 import sharded

    try:
        dev = torch.device("cuda", rank)
        tp, mv, w, d, _ = _synth.mf_case(37, 4, 3, 200, 30000, seed=41)
        got = sharded.matched_filter_sharded(*(torch.from_numpy(a).to(dev) for a in (tp, mv, w, d)), precision="fp32")
        want = cpu.matched_filter(tp, mv, w, d, 1)
        ok_mf = float(np.abs(got.cpu().numpy() - want).max()) < 2e-6

        # the same through many small batches with one rank lagging: its peer runs ahead and must not overwrite a
        # gather buffer this rank is still laying out (double-buffer ordering of PeerSlabGather)
        from seismic_bpmf_b200 import device as devmod

        if rank == 1:
            plain = devmod.MatchedFilterPlan.run_batch

            def slow(self, *a, **k):
                torch.cuda._sleep(int(4e7))  # ~20 ms in front of every batch of this rank
                return plain(self, *a, **k)

            devmod.MatchedFilterPlan.run_batch = slow
        got = sharded.matched_filter_sharded(*(torch.from_numpy(a).to(dev) for a in (tp, mv, w, d)), precision="fp32",
                                             batch=3)
        ok_mf = ok_mf and float(np.abs(got.cpu().numpy() - want).max()) < 2e-6
        if rank == 1:
            devmod.MatchedFilterPlan.run_batch = plain

        f, tau, wp, ws = _synth.beam_case(301, 6, 2, 2, 9000, seed=42)
        tau[200], ws[200] = tau[100], ws[100]  # a tie across the shard boundary
        mb, am = sharded.beamform_max_sharded(torch.from_numpy(f).to(dev), torch.from_numpy(tau.astype(np.int32)).to(dev),
                                              torch.from_numpy(wp).to(dev), torch.from_numpy(ws).to(dev))
        mb_w, am_w = cpu.beamform(f, tau, wp, ws)
        full = cpu.beamform(f, tau, wp, ws, reduce="none")
        srt = np.sort(full, axis=0)
        clear = srt[-1] - srt[-2] > 1e-5
        ok_beam = float(np.abs(mb.cpu().numpy() - mb_w).max()) < 1e-5 and \
            np.array_equal(am.cpu().numpy()[clear], am_w[clear])

        fullb = sharded.beamform_none_sharded(torch.from_numpy(f[:, :, :3000].copy()).to(dev),
                                              torch.from_numpy(tau.astype(np.int32)).to(dev),
                                              torch.from_numpy(wp).to(dev), torch.from_numpy(ws).to(dev))
        ok_beam = ok_beam and float(np.abs(fullb.cpu().numpy() - cpu.beamform(f[:, :, :3000].copy(), tau, wp, ws,
                                                                                reduce="none")).max()) < 1e-5

        # copy-engine peer gather: every rank ends with every rank's slab, batch after batch
        n = 5000
        g = sharded.PeerSlabGather(n, dev)
        ok_gather = True
        for b in range(4):
            slab = torch.full((n - b,), float(10 * b + rank), device=dev)
            g.push(b % 2, slab)
            g.wait()
            torch.cuda.synchronize()
            view = g.local[b % 2].view(world, n)
            for r in range(world):
                ok_gather &= bool((view[r, : n - b] == float(10 * b + r)).all().item())
            dist.barrier()
        ret[rank] = (ok_mf, ok_beam, ok_gather)
    finally:
        dist.destroy_process_group()


def test_sharded_paths_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for rank in range(world):
        assert ret[rank] == (True, True, True), (rank, ret[rank])
