"""Batched inter-template CC against the reference's own loop (BPMF/dataset.py:4784-4833) driven
with the CPU oracle's matched filter, and against the same loop on the GPU drop-in."""
import numpy as np
import pytest

from oracle import cpu
from seismic_bpmf_b200 import intertemplate as itp

pytestmark = pytest.mark.gpu


def _case(T, S, C, N, seed, n_sta=3):
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((max(1, T // 3), S, C, N)).astype(np.float32)
    x = base[rng.integers(0, base.shape[0], T)] + 0.6 * rng.standard_normal((T, S, C, N)).astype(np.float32)
    for t in range(T):  # families of similar templates, slightly misaligned
        x[t] = np.roll(x[t], int(rng.integers(-4, 5)), axis=-1)
    x[T // 2, 0, 0] = 0.0  # a dead channel
    w = np.zeros((T, S, C), np.float32)
    for t in range(T):
        w[t, rng.permutation(S)[:n_sta]] = 1.0
        w[t, :, rng.integers(0, C)] *= rng.integers(0, 2)  # channel availability
        if w[t].sum() == 0:
            w[t, 0, 0] = 1.0
        w[t] /= w[t].sum()
    mask = rng.random((T, T)) < 0.7
    np.fill_diagonal(mask, True)
    return x, w, mask


def _oracle_mf(templates, moveouts, weights, data, step, network_sum=False, check_zeros=False):
    return cpu.matched_filter(templates, moveouts, weights, data, step, network_sum=network_sum)


@pytest.mark.parametrize("T,S,C,N,max_lag", [(17, 4, 3, 120, 10), (40, 5, 2, 301, 7), (3, 2, 1, 64, 0), (33, 3, 3, 830, 12)])
def test_batched_matches_reference_loop(T, S, C, N, max_lag):
    x, w, mask = _case(T, S, C, N, seed=T + N)
    got = itp.intertemplate_cc(x, w, mask, max_lag=max_lag)
    want = itp.intertemplate_cc_reference_loop(x, w, mask, max_lag=max_lag, matched_filter=_oracle_mf)
    assert got.shape == (T, T) and got.dtype == np.float32
    assert np.abs(got - want).max() < 2e-5
    assert np.allclose(got, got.T)
    # self-similarity: the diagonal is sum(w) = 1 (lag max_lag is the template itself) unless a weighted channel is dead
    diag_ok = [t for t in range(T) if not (t == T // 2 and w[t, 0, 0] != 0)]
    assert np.abs(np.diag(got)[diag_ok] - 1.0).max() < 1e-5
    # pairs outside the mask in both directions are exactly zero
    off = ~(mask | mask.T)
    assert np.all(got[off] == 0.0)


def test_batched_matches_gpu_loop_and_no_mask():
    x, w, _ = _case(12, 3, 3, 200, seed=5)
    got = itp.intertemplate_cc(x, w, None, max_lag=10, symmetric=False)
    loop = itp.intertemplate_cc_reference_loop(x, w, None, max_lag=10, symmetric=False)
    assert np.abs(got - loop).max() < 2e-5
    assert itp.intertemplate_cc(x[:0], w[:0]).shape == (0, 0)
