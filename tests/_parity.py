"""Strided parity of a full-day matched-filter result against the CPU oracle (TEST INFRASTRUCTURE: used by
tests/ and by bench.py's untimed parity leg only).

The whole (T, n_corr) slab of BASELINE.json configs[1] is hours of CPU work, so the comparison runs on runs of
`win_len` consecutive outputs spread over the whole day for a few templates of EVERY tile of 16 templates the
tensor kernel forms: the first and the last outputs of the day, both sides of every 65 536-sample segment
boundary that is sampled, the zero gap of the synthetic traces, every planted detection, and seeded random
positions in between ("outputs hashed + max-abs-diff vs oracle on a sampled subset").
"""
import numpy as np

SEGMENT = 65536  # samples swept per resident template image (TC_SEG_TILES * 2048 in csrc/mf_tc.cu)


def window_starts(n_corr, n_windows, win_len, gap=None, planted=(), mv_range=(-100, 3000), L=800, seed=7):
    """Start indices of `n_windows` runs of `win_len` outputs in [0, n_corr)."""
    rng = np.random.default_rng(seed)
    must = [0, n_corr - win_len]
    for seg in (SEGMENT, SEGMENT // 2):  # both sides of a few segment boundaries (64- and 32-tile segments)
        n_seg = max(1, n_corr // seg)
        for k in sorted({1, n_seg // 2, n_seg - 1, n_seg}):
            must.append(k * seg - win_len // 2)
    if gap is not None:  # outputs whose windows enter / leave the zero gap, for every moveout in range
        g0, g1 = gap
        must += [g0 - mv_range[1] - L - win_len // 2, g0 - mv_range[1] // 2 - L, g0 - win_len // 2, (g0 + g1) // 2,
                 g1 - mv_range[1], g1 - win_len // 2]
    for i0 in planted:
        must.append(int(i0) - win_len // 2)
    must = [int(min(max(0, s), max(0, n_corr - win_len))) for s in must]
    n_rand = max(0, n_windows - len(must))
    rand = rng.integers(0, max(1, n_corr - win_len), size=n_rand).tolist()
    return np.array(sorted(set(must + rand)), dtype=np.int64)


def tile_templates(T, per_tile=3, tile=16, extra=()):
    """A few templates of every tile of `tile` consecutive templates (first, middle, last), plus `extra`."""
    sel = set(int(t) for t in extra if 0 <= int(t) < T)
    for t0 in range(0, T, tile):
        n = min(tile, T - t0)
        for k in sorted({0, n // 2, n - 1})[:per_tile]:
            sel.add(t0 + k)
    return np.array(sorted(sel), dtype=np.int64)


def compare(cc_rows, want, tol=1e-4):
    """cc_rows, want: (n_templates, n_windows, win_len).  Returns the parity record."""
    diff = np.abs(cc_rows.astype(np.float64) - want.astype(np.float64))
    worst = np.unravel_index(int(diff.argmax()), diff.shape)
    structural, chance = zero_pattern(cc_rows, want)
    return {"compared_outputs": int(diff.size), "max_abs_err_vs_oracle": float(diff.max()), "tolerance": tol,
            "mean_abs_err": float(diff.mean()), "exact_zeros_in_oracle": int((want == 0.0).sum()),
            "zero_pattern_mismatches": structural, "chance_zeros": chance, "worst_at": [int(x) for x in worst],
            "ok": bool(diff.max() < tol and structural == 0)}


def zero_pattern(got, want, noise=2e-5):
    """BPMF reads an exact 0.0 in a CC series as "no data here" (similarity_search.py:615, libc.c:561,609), so the
    zeros must coincide.  A float32 sum of ~48 terms of magnitude 1e-3 also lands EXACTLY on 0.0 by cancellation about
    once per 1e8 outputs (measured: 14-19 of the 4.3e9 outputs of configs[1], at either end of the comparison), with
    the other side within rounding noise of zero: those are counted apart (`chance`), everything else -- a zero
    against a value beyond `noise` -- is a structural mismatch."""
    z_got, z_want = got == 0.0, want == 0.0
    mism = z_got != z_want
    other = np.where(z_got, np.abs(want), np.abs(got))
    structural = int((mism & (other > noise)).sum())
    return structural, int(mism.sum()) - structural
