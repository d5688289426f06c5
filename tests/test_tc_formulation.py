"""CPU: a NumPy restatement of the tensor-path formulation of seismic_bpmf_b200/csrc/mf_tc.cu
(Hankel GEMM over 8-sample columns, 8 phase rows per template delayed by r + moveout mod 8, whole-column
shift floor(moveout / 8) + 1 with an 8-sample front pad, power-of-two fp16 scaling, rsqrt(energy) column
factor, stability guard) checked against the oracle.  It pins the index algebra the CUDA pre-pass and
epilogue implement, so that it can be reviewed and regression-tested without a GPU."""
import numpy as np
import pytest

from oracle import cpu
from tests import _synth

FRONT = 8


def _pow2_scale(amax):
    if not amax > 0:
        return 1.0
    _, ex = np.frexp(np.float32(amax))
    return float(np.ldexp(1.0, int(np.clip(14 - ex, -100, 100))))


def emulate_tensor_path(templates, moveouts, weights, data, split=False):
    T, S, C, L = templates.shape
    N = data.shape[-1]
    SC = S * C
    tp = templates.reshape(T, SC, L).astype(np.float32)
    mv = moveouts.reshape(T, SC).astype(np.int64)
    w = weights.reshape(T, SC).astype(np.float32)
    d = data.reshape(SC, N).astype(np.float32)
    n_corr = N - L + 1
    Kp = -(-(L + 14) // 16) * 16
    n_cols = (N - L + FRONT) // 8 + 1
    cc = np.zeros((T, n_corr), dtype=np.float64)
    # per-template valid output range
    lo, hi = np.zeros(T, int), np.full(T, -1, int)
    for t in range(T):
        used = w[t] != 0
        if used.any():
            lo[t] = max(0, -int(mv[t][used].min()))
            hi[t] = min(N - L - int(mv[t][used].max()), n_corr - 1)
    for ch in range(SC):
        sB = _pow2_scale(np.abs(d[ch]).max())
        staged = np.zeros(8 * n_cols + Kp + 64, dtype=np.float64)
        x = d[ch] * np.float32(sB)
        hi16 = x.astype(np.float16)
        staged[FRONT: FRONT + N] = hi16.astype(np.float64)
        if split:
            staged[FRONT: FRONT + N] += (x - hi16.astype(np.float32)).astype(np.float16).astype(np.float64)
        c2 = np.concatenate([[0.0], np.cumsum(d[ch].astype(np.float64) ** 2)])
        energy = c2[L:] - c2[:-L]  # window start j in [0, N - L]
        colinv = np.zeros(8 * n_cols + 64)
        colinv[FRONT: FRONT + n_corr] = np.where(energy > 0, 1.0 / np.sqrt(np.where(energy > 0, energy, 1.0)), 0.0)
        colinv = colinv.astype(np.float32)
        hank = np.lib.stride_tricks.sliding_window_view(staged, Kp)[::8][:n_cols]  # row jj = staged[8jj : 8jj + Kp]
        for t in range(T):
            if w[t, ch] == 0 or hi[t] < lo[t]:
                continue
            sst = np.float32((tp[t, ch].astype(np.float64) ** 2).sum())
            if not sst > 0:
                continue
            sA = _pow2_scale(np.abs(tp[t, ch]).max())
            y = tp[t, ch] * np.float32(sA)
            th = y.astype(np.float16)
            trow = th.astype(np.float64)
            if split:
                trow = trow + (y - th.astype(np.float32)).astype(np.float16).astype(np.float64)
            m = int(mv[t, ch])
            q0 = m // 8  # floor
            scale = np.float32(w[t, ch]) / np.float32(sA * sB * np.sqrt(sst))
            ithr = np.float32(np.sqrt(sst) * 1000.0)
            for ir in range(8):
                e = ir + (m - 8 * q0)
                assert 0 <= e < 15
                q = q0 + FRONT // 8
                a_row = np.zeros(Kp)
                a_row[e: e + L] = trow
                acc = (hank @ a_row).astype(np.float32)  # C[(t, ir), jj]
                inv = colinv[8 * np.arange(n_cols) + e]
                i = 8 * (np.arange(n_cols) - q) + ir
                ok = (i >= lo[t]) & (i <= hi[t])
                val = np.where(inv < ithr, acc * inv * scale, 0.0)
                cc[t, i[ok]] += val[ok]
    return cc.astype(np.float32)


@pytest.mark.parametrize("split,tol", [(False, 4e-4), (True, 5e-6)])
def test_tensor_formulation_matches_oracle(split, tol):
    for seed, (T, S, C, L, N) in enumerate([(3, 2, 2, 40, 700), (2, 1, 3, 64, 520), (2, 2, 1, 9, 300)]):
        tp, mv, w, d, planted = _synth.mf_case(T, S, C, L, N, seed=100 + seed, mv_lo=-37, mv_hi=90)
        want = cpu.matched_filter(tp, mv, w, d, 1)
        got = emulate_tensor_path(tp, mv, w, d, split=split)
        assert got.shape == want.shape
        assert np.abs(got - want).max() < tol, (seed, np.abs(got - want).max())
        assert np.array_equal(got == 0, want == 0)
        for t, i0 in planted.items():
            assert got[t].argmax() == i0


def test_front_pad_is_needed_for_delays_beyond_a_column():
    """Rows delayed by e >= 8 start their first window at column -1 without the 8-sample front pad: the
    regression the GPU tests caught (moveouts that are not multiples of 8, outputs near i = -moveout)."""
    tp, mv, w, d, _ = _synth.mf_case(1, 1, 1, 16, 200, seed=7, plant=False, zero_gap=False)
    mv[:] = -3  # e = ir + 5 reaches 12
    want = cpu.matched_filter(tp, mv, w, d, 1)
    got = emulate_tensor_path(tp, mv, w, d, split=True)
    assert np.abs(got - want).max() < 5e-6 and want[0, 3] != 0 and np.all(want[0, :3] == 0)
