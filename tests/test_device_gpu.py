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


def test_detections_leave_the_device_as_a_compact_list():
    """threshold + peak selection on the CC slab; only candidates are downloaded."""
    tp, mv, w, d, planted = _synth.mf_case(6, 3, 2, 100, 60000, seed=24)
    plan = dev.MatchedFilterPlan(6, 3, 2, 100, 60000, precision="fp32")
    cc = plan.run(*(torch.from_numpy(a).cuda() for a in (tp, mv, w, d)))
    cc_host = cc.cpu().numpy().copy()
    wn = np.random.default_rng(1).standard_normal(500).astype(np.float32)
    g = torch.from_numpy(wn).cuda()
    half, shift, num_dev, win = 3000, 4500, 6.0, 150
    for mode in ("above", "select"):
        s_, i_, v_, t_ = dev.detect_dev(plan.cc, g, num_dev, half, shift, search_win=win, mode=mode, n_valid=plan.n_corr)
        assert torch.equal(cc, plan.cc[:, : plan.n_corr]) and np.array_equal(cc.cpu().numpy(), cc_host)  # slab untouched
        want = []
        for t in range(6):
            thr, _ = libc_numpy.time_dependent_threshold(cc_host[t], wn, num_dev, half, shift)
            if mode == "above":
                idx = np.nonzero(cc_host[t] > thr)[0]
            else:
                idx = np.nonzero(libc_numpy.select_cc_indexes(cc_host[t], thr, win))[0]
            want += [(t, int(i)) for i in idx]
        got = list(zip(s_.tolist(), i_.tolist()))
        # thresholds agree to float32 round-off: allow candidates sitting exactly on the threshold to differ
        assert set(got) == set(want) or len(set(got) ^ set(want)) <= 2, (mode, len(got), len(want))
        assert np.allclose(v_, cc_host[s_, i_])
        for t, i0 in planted.items():
            assert (t, i0) in got
    with pytest.raises(_lib.BpmfError):
        dev.detect_dev(plan.cc, g, -100.0, half, shift, n_valid=plan.n_corr, capacity=16)  # everything is above


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


def test_prepare_traces_once_run_many_batches():
    """BPMF loops template chunks over one day of data (similarity_search.py:783-790): the trace staging is
    prepared once and reused by every batch / by a second template set, for each arithmetic."""
    tp, mv, w, d, _ = _synth.mf_case(45, 4, 3, 300, 30000, seed=25)
    want = cpu.matched_filter(tp, mv, w, d, 1)
    dev_args = [torch.from_numpy(a).cuda() for a in (tp, mv, w, d)]
    for prec, tol in (("fp32", 2e-6), ("fp16", 1e-4), ("fp16x3", 1e-5), ("auto", 1e-4)):
        plan = dev.MatchedFilterPlan(45, 4, 3, 300, 30000, precision=prec, batch=16)
        cc = plan.run(*dev_args)
        assert np.abs(cc.cpu().numpy() - want).max() < tol, prec
        # second pass on the same traces with permuted templates, without re-staging the traces
        perm = torch.arange(44, -1, -1).cuda()
        cc2 = plan.run(dev_args[0][perm].contiguous(), dev_args[1][perm].contiguous(),
                       dev_args[2][perm].contiguous(), dev_args[3], prepared=True)
        assert np.abs(cc2.cpu().numpy() - want[::-1]).max() < tol, prec


def test_host_path_reuses_cached_buffers():
    from seismic_bpmf_b200 import fast_matched_filter as fmf

    lib = _lib.load()
    tp, mv, w, d, _ = _synth.mf_case(150, 3, 2, 100, 20000, seed=26)  # > 128 templates: chunked on two streams
    want = cpu.matched_filter(tp, mv, w, d, 1)
    for _ in range(2):
        got = fmf.matched_filter(tp, mv, w, d, 1, precision="fp32", check_zeros=False)
        assert np.abs(got - want).max() < 2e-6
    lib.bpmf_release_cache()
    got = fmf.matched_filter(tp[:3], mv[:3], w[:3], d, 1, precision="fp32", check_zeros=False)
    assert np.abs(got - want[:3]).max() < 2e-6


def test_detect_gated_threshold_cap_and_kurtosis_gate():
    """The two remaining per-template steps of MatchedFilter.find_detections on the device: the threshold cap
    max_CC_threshold * sum(w) (similarity_search.py:629) and the kurtosis sanity check (:633-642)."""
    import torch
    from scipy.stats import kurtosis as sk

    from oracle import libc_numpy
    from seismic_bpmf_b200 import device as dev

    rng = np.random.default_rng(9)
    n, win = 60000, 6000
    x = (0.03 * rng.standard_normal((3, n))).astype(np.float32)
    x[0, 20000] = 0.5            # a clean detection
    x[1, 30000] = 0.5
    x[1, 5000:9000] *= 12.0      # a burst: kurtosis far above 20 -> the whole series is discarded
    x[2, 40000] = 0.2            # below the rms threshold of this noisy series, above the cap
    x[2, 10000:30000] *= 3.0
    wn = rng.standard_normal(500).astype(np.float32)
    d = torch.device("cuda")
    xs, g = torch.from_numpy(x).to(d), torch.from_numpy(wn).to(d)
    det = dev.DetectBatches(1, 3, n, win // 2, int(0.75 * win), capacity=4096, device=d)
    kout = torch.zeros(3, device=d)
    cap = torch.tensor([1.0, 1.0, 0.15], device=d)
    det.run(0, xs, g, 8.0, search_win=500, mode="select", thr_cap=cap, max_kurtosis=20.0, kurtosis_out=kout)
    s_, i_, v_, t_ = det.collect()
    want_k = sk(x.astype(np.float64), axis=1)
    assert np.abs(kout.cpu().numpy() - want_k).max() < 1e-3 * np.abs(want_k).max()
    assert want_k[1] > 20.0 > want_k[0]
    got = set(zip(s_.tolist(), i_.tolist()))
    assert (0, 20000) in got and (2, 40000) in got
    assert not any(s == 1 for s, _ in got)                     # gated out by its kurtosis
    for b in (0, 2):                                            # everything else follows the capped threshold
        thr, _ = libc_numpy.time_dependent_threshold(x[b], wn, 8.0, win // 2, int(0.75 * win))
        thr = np.minimum(float(cap[b]), thr)
        sel = np.flatnonzero(libc_numpy.select_cc_indexes(x[b], thr.astype(np.float32), 500))
        assert sorted(i for s, i in got if s == b) == sel.tolist()
