"""GPU parity at the sizes BASELINE.json names (configs[1] for the matched filter, configs[2]'s day length for the
beamformer), where the oracle cannot produce the whole result: strided runs of outputs over the whole day for
templates of every tile, whole-day series for two templates, and time slices of the day-long beam.

These are the cases the small-shape tests cannot reach: output indices beyond 2^23, the last segment / tile tail of
the tensor kernel, the zero gap at N/3 seen through every template tile, tiles 1..31 of a 500-template call."""
import numpy as np
import pytest

from oracle import cpu
from tests import _parity, _synth

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north star: |dCC| < 1e-4 float32


def _cfg2_inputs(T=500, S=16, C=3, L=800, N=8640000, seed=20240726):
    """SURVEY 8(d) synthetic day, built on the device (a numpy day of 48 channels takes a minute to draw)."""
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    data = torch.randn((S, C, N), generator=g, device="cuda", dtype=torch.float32)
    data[S // 2, :, N // 3: N // 3 + 10000] = 0.0
    data /= data.std(dim=-1, keepdim=True)
    tp = torch.randn((T, S, C, L), generator=g, device="cuda", dtype=torch.float32)
    tp -= tp.mean(dim=-1, keepdim=True)
    tp /= tp.std(dim=-1, keepdim=True)
    mv = torch.randint(-100, 3001, (T, S, C), generator=g, device="cuda", dtype=torch.int32)
    planted = {}
    for k, t in enumerate((3, T // 2 + 1, T - 1)):  # one near the start, one beyond index 2^23, one in the last segment
        i0 = (5000, (1 << 23) + 4321, N - L - 3000 - 777)[k]
        idx = (i0 + mv[t].long()).reshape(S, C, 1) + torch.arange(L, device="cuda").reshape(1, 1, L)
        tp[t] = torch.gather(data, 2, idx)
        planted[t] = i0
    w = torch.full((T, S, C), 1.0 / (S * C), device="cuda", dtype=torch.float32)
    return tp.contiguous(), mv.contiguous(), w, data.contiguous(), planted


@pytest.mark.parametrize("weighting", ["all_stations", "ten_closest"])
def test_cfg2_full_day_every_tile_vs_oracle(weighting):
    import torch

    from seismic_bpmf_b200 import device as dev

    T, S, C, L, N = 500, 16, 3, 800, 8640000
    tp, mv, w, data, planted = _cfg2_inputs(T, S, C, L, N)
    if weighting == "ten_closest":  # SURVEY 8(d) second variant
        order = mv.min(dim=2).values.argsort(dim=1)
        w = torch.zeros_like(w)
        w.scatter_(1, order[:, :10].unsqueeze(-1).expand(T, 10, C), 1.0)
        w = (w / w.sum(dim=(1, 2), keepdim=True)).contiguous()
    plan = dev.MatchedFilterPlan(T, S, C, L, N, precision="auto")
    cc = plan.run(tp, mv, w, data)
    torch.cuda.synchronize()
    n_corr = plan.n_corr
    assert n_corr == N - L + 1
    data_h = data.cpu().numpy()
    # (1) strided: 3 templates of each of the 32 tiles + planted, runs of 2048 outputs over the whole day
    sel = _parity.tile_templates(T, per_tile=3, extra=planted.keys())
    starts = _parity.window_starts(n_corr, 40, 2048, gap=(N // 3, N // 3 + 10000), planted=planted.values(), L=L)
    idx = torch.from_numpy(starts).cuda().reshape(-1, 1) + torch.arange(2048, device="cuda")
    sel_d = torch.from_numpy(sel).cuda()
    got = cc[sel_d][:, idx.reshape(-1)].reshape(len(sel), len(starts), 2048).cpu().numpy()
    want = cpu.matched_filter_windows(tp[sel_d].cpu().numpy(), mv[sel_d].cpu().numpy(), w[sel_d].cpu().numpy(), data_h,
                                      starts, 2048)
    rec = _parity.compare(got, want, TOL)
    assert rec["compared_outputs"] >= 7e6
    assert rec["ok"], rec
    # (2) two whole-day series: first template of tile 0, last template of the (partial) last tile
    whole = np.array([0, T - 1])
    want = cpu.matched_filter_windows(tp[whole].cpu().numpy(), mv[whole].cpu().numpy(), w[whole].cpu().numpy(), data_h,
                                      np.array([0]), n_corr)[:, 0]
    got = cc[torch.from_numpy(whole).cuda()].cpu().numpy()
    assert np.abs(got - want).max() < TOL
    # head / tail the template cannot reach: exactly 0.0 on both sides; inside, a chance cancellation to 0.0 (about one
    # output in 1e8, see _parity.zero_pattern) is not a structural zero
    structural, chance = _parity.zero_pattern(got, want)
    assert structural == 0 and chance <= 4, (structural, chance)
    m = mv[torch.from_numpy(whole).cuda()].cpu().numpy()
    used = w[torch.from_numpy(whole).cuda()].cpu().numpy() != 0  # the range is set by the WEIGHTED channels only
    for r in range(2):
        lo_r, hi_r = max(0, -int(m[r][used[r]].min())), N - L - int(m[r][used[r]].max())
        assert not got[r, :lo_r].any() and not got[r, hi_r + 1:].any()
        assert not want[r, :lo_r].any() and not want[r, hi_r + 1:].any()
    # (3) planted detections are the global maxima, at indices on both sides of 2^23
    for t, i0 in planted.items():
        assert int(torch.argmax(cc[t]).item()) == i0
        assert abs(float(cc[t, i0].item()) - float(w[t].sum().item())) < TOL
    # (4) every template saw the zero gap the same way the oracle did: nothing is NaN / inf anywhere
    assert bool(torch.isfinite(cc).all().item())


def test_beamform_day_length_1e5_sources_vs_oracle_slices():
    """K = 100 000 grid sources x 32 stations x 2 phases over N = 4 320 000 samples (one day @ 50 Hz) through the call
    BPMF makes (host arrays, int64 moveouts); compared with the oracle on three time slices: the first outputs, the
    middle of the day, and the last outputs (where "strict" bounds bite)."""
    from seismic_bpmf_b200.beampower import beampower as bp

    K, S, N, n_out = 100000, 32, 4320000, 600
    f, tau, wp, ws = _synth.beam_case(K, S, 2, 2, N, seed=77, coherent=True)
    mb, am = bp.beamform(f, tau, wp, ws, device="gpu", reduce="max")
    assert mb.shape == (N,) and am.dtype == np.int32
    reach = int(tau.max())
    for t0 in (0, N // 2 + 12345, N - n_out - reach):
        last = t0 + n_out + reach >= N
        sl = f[:, :, t0: min(N, t0 + n_out + reach)]
        mb_w, am_w = cpu.beamform(np.ascontiguousarray(sl), tau, wp, ws)
        n_cmp = sl.shape[-1] if last else n_out  # inside the day the slice's own tail is not the day's tail
        assert np.abs(mb[t0: t0 + n_cmp] - mb_w[:n_cmp]).max() < 1e-5
        same = am[t0: t0 + n_cmp] == am_w[:n_cmp]
        if not same.all():  # the arg-max may differ only where two sources tie within float32 round-off:
            bad = np.flatnonzero(~same)  # the source the GPU names must reach the maximum at that sample
            assert len(bad) < 50
            srcs = am[t0 + bad]
            full = cpu.beamform(np.ascontiguousarray(sl), tau[srcs], wp, ws[srcs], reduce="none")
            for r, t in enumerate(bad):
                assert abs(full[r, t] - mb[t0 + t]) < 1e-5
