"""CPU: the helpers the full-size parity checks rely on (tests/_parity.py) and the oracle's windowed form."""
import numpy as np

from oracle import cpu
from tests import _parity, _synth


def test_windowed_oracle_equals_full_oracle():
    tp, mv, w, d, _ = _synth.mf_case(6, 4, 3, 200, 60000, seed=3, sparse_weights=True)
    full = cpu.matched_filter(tp, mv, w, d, 1)
    n_corr = full.shape[1]
    starts = np.array([0, -50, 100, 19990, n_corr - 300, n_corr - 5, 40000], dtype=np.int64)
    got = cpu.matched_filter_windows(tp, mv, w, d, starts, 700)
    for wi, s0 in enumerate(starts):
        idx = np.arange(s0, s0 + 700)
        ok = (idx >= 0) & (idx < n_corr)
        ref = np.zeros((6, 700), np.float32)
        ref[:, ok] = full[:, idx[ok]]
        assert np.abs(got[:, wi] - ref).max() < 1e-7
    whole = cpu.matched_filter_windows(tp, mv, w, d, np.array([0]), n_corr)[:, 0]  # cut in 32768-output pieces inside
    assert np.abs(whole - full).max() < 1e-7 and np.array_equal(whole == 0, full == 0)


def test_window_starts_cover_the_special_places():
    n_corr, L, N = 8639201, 800, 8640000
    gap = (N // 3, N // 3 + 10000)
    planted = [5000, (1 << 23) + 4321]
    s = _parity.window_starts(n_corr, 64, 2048, gap=gap, planted=planted, L=L)
    assert len(s) >= 60 and s.min() == 0 and s.max() == n_corr - 2048 and (np.diff(s) > 0).all()
    covered = lambda i: bool(((s <= i) & (i < s + 2048)).any())  # noqa: E731
    assert all(covered(i) for i in planted)
    assert covered(gap[0] - 1) and covered(gap[0] + 1) and covered(gap[1] - 1)
    for seg in (_parity.SEGMENT, _parity.SEGMENT // 2):  # both sides of the first segment boundary (32- and 16-tile segments)
        assert covered(seg - 1) and covered(seg)
    assert covered((1 << 23) + 100) or covered((1 << 23) + 4321)


def test_tile_templates_take_every_tile():
    sel = _parity.tile_templates(500, per_tile=3, extra=(3, 499, 777))
    assert {t // 16 for t in sel} == set(range(32)) and 499 in sel and 777 not in sel
    assert (np.bincount(sel // 16) >= 3)[:31].all()


def test_zero_pattern_separates_chance_cancellations():
    want = np.array([[0.0, 1e-3, 2e-7, 0.5, 0.0]], np.float32)
    got = np.array([[0.0, 1e-3, 0.0, 0.0, 1e-7]], np.float32)
    structural, chance = _parity.zero_pattern(got, want)
    assert structural == 1 and chance == 2  # 0.5 -> 0 is structural; 2e-7 -> 0 and 0 -> 1e-7 are rounding noise
    rec = _parity.compare(got[None], want[None])
    assert not rec["ok"] and rec["zero_pattern_mismatches"] == 1 and rec["chance_zeros"] == 2
