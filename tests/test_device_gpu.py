"""GPU: device-resident entry points (torch tensors in, no host sync), the fused CC -> threshold
hand-off, and caract_GPU."""
import numpy as np
import pytest
import torch

from oracle import cpu, libc_numpy
from seismic_bpmf_b200 import _lib, device as dev
from tests import _synth

pytestmark = pytest.mark.gpu


def test_matched_filter_plan_matches_host_path():
    tp, mv, w, d, _ = _synth.mf_case(9, 4, 3, 200, 25000, seed=21)
    plan = dev.MatchedFilterPlan(9, 4, 3, 200, 25000, precision="fp32")
    cc = plan.run(*(torch.from_numpy(a).cuda() for a in (tp, mv, w, d)))
    torch.cuda.synchronize()
    assert cc.shape == (9, 25000 - 200 + 1)
    assert np.abs(cc.cpu().numpy() - cpu.matched_filter(tp, mv, w, d, 1)).max() < 2e-6
    # re-running into the same slab must not accumulate
    cc2 = plan.run(*(torch.from_numpy(a).cuda() for a in (tp, mv, w, d)))
    assert torch.equal(cc, cc2)


def test_cc_to_threshold_without_leaving_the_device():
    tp, mv, w, d, _ = _synth.mf_case(3, 3, 2, 100, 40000, seed=22)
    plan = dev.MatchedFilterPlan(3, 3, 2, 100, 40000, precision="fp32")
    cc = plan.run(*(torch.from_numpy(a).cuda() for a in (tp, mv, w, d)))
    cc_host = cc.cpu().numpy().copy()
    wn = np.random.default_rng(0).standard_normal(500).astype(np.float32)
    thr = dev.time_dependent_threshold_dev(plan.cc, torch.from_numpy(wn).cuda(), 8.0, 2000, 3000, n_valid=plan.n_corr)
    torch.cuda.synchronize()
    for t in range(3):
        want, _ = libc_numpy.time_dependent_threshold(cc_host[t], wn, 8.0, 2000, 3000)
        assert np.abs(thr[t].cpu().numpy() - want).max() < 5e-6


def test_beam_plan_and_merge():
    f, tau, wp, ws = _synth.beam_case(120, 6, 2, 2, 8000, seed=23)
    args = [torch.from_numpy(a).cuda() for a in (f, tau.astype(np.int32), wp, ws)]
    plan = dev.BeamformPlan(120, 6, 2, 2, 8000)
    mb, am = plan.run_max(*args)
    mb_w, am_w = cpu.beamform(f, tau, wp, ws)
    assert np.abs(mb.cpu().numpy() - mb_w).max() < 1e-5
    a = dev.BeamformPlan(50, 6, 2, 2, 8000)
    b = dev.BeamformPlan(70, 6, 2, 2, 8000)
    m1, a1 = a.run_max(args[0], args[1][:50].contiguous(), args[2], args[3][:50].contiguous())
    m2, a2 = b.run_max(args[0], args[1][50:].contiguous(), args[2], args[3][50:].contiguous(), source_offset=50)
    dev.beam_merge_(m1, a1, m2, a2)
    assert torch.equal(m1, mb) and torch.equal(a1, am)


def test_caract_gpu_and_launch_counter(capfd):
    lib = _lib.load()
    before = lib.bpmf_launch_count()
    lib.caract_GPU()
    assert "cuda devices found" in capfd.readouterr().out
    tp, mv, w, d, _ = _synth.mf_case(1, 1, 1, 16, 500, seed=1, zero_gap=False)
    from seismic_bpmf_b200 import fast_matched_filter as fmf

    fmf.matched_filter(tp, mv, w, d, 1, check_zeros=False)
    assert lib.bpmf_launch_count() > before
