"""float64 NumPy brute force of SURVEY Appendix A.1 / A.2 (TEST INFRASTRUCTURE).

Independent of mf_oracle.c / beam_oracle.c (different language, different arithmetic, no prefix
sums) so the C oracle can itself be checked.  Small cases only.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

GUARD = 1e-6


def matched_filter(templates, moveouts, weights, data, step, normalize="short", network_sum=True):
    tp = np.asarray(templates, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float32)
    d = np.asarray(data, dtype=np.float64)
    mv = np.asarray(moveouts)
    T, S, C, L = tp.shape
    if mv.ndim == 2:
        mv = np.repeat(mv[:, :, None], C, axis=2)
    N = d.shape[-1]
    n_corr = (N - L) // step + 1
    out = np.zeros((T, n_corr) if network_sum else (T, n_corr, S, C), dtype=np.float64)
    for t in range(T):
        used = [(s, c) for s in range(S) for c in range(C) if w[t, s, c] != 0]
        if not used:
            continue
        mvs = [int(mv[t, s, c]) for s, c in used]
        n_lo = 0 if min(mvs) >= 0 else -(min(mvs) // step)  # ceil(-min/step)
        last = N - L - max(mvs)
        if last < 0:
            continue
        n_hi = min(last // step, n_corr - 1)
        if n_hi < n_lo:
            continue
        idx = np.arange(n_lo, n_hi + 1) * step
        for s, c in used:
            m = int(mv[t, s, c])
            win = sliding_window_view(d[s, c], L)[idx + m]  # (n, L)
            tt = tp[t, s, c]
            num = win @ tt
            ssd = np.einsum("ij,ij->i", win, win)
            if normalize == "full":
                mean = win.mean(axis=1)
                num = num - mean * tt.sum()
                ssd = ssd - L * mean * mean
            den = np.float32(np.sum(tt * tt)).astype(np.float64) * ssd
            cc = np.where(den > GUARD, num / np.sqrt(np.where(den > GUARD, den, 1.0)), 0.0)
            if network_sum:
                out[t, n_lo:n_hi + 1] += np.float64(w[t, s, c]) * cc
            else:
                out[t, n_lo:n_hi + 1, s, c] = cc
    return out


def beamform(features, moveouts, weights_phases, weights_sources, out_of_bounds="strict",
             reduce="max"):
    f = np.asarray(features, dtype=np.float64)
    tau = np.asarray(moveouts).astype(np.int64)
    wp = np.asarray(weights_phases, dtype=np.float64)
    ws = np.asarray(weights_sources, dtype=np.float64)
    S, C, N = f.shape
    K, _, P = tau.shape
    U = np.einsum("scp,scn->spn", wp, f)
    beam = np.zeros((K, N))
    t = np.arange(N)
    for k in range(K):
        valid_all = np.ones(N, dtype=bool)
        acc = np.zeros(N)
        for s in range(S):
            if ws[k, s] == 0:
                continue
            for p in range(P):
                x = t + tau[k, s, p]
                ok = (x >= 0) & (x < N)
                valid_all &= ok
                acc[ok] += ws[k, s] * U[s, p, x[ok]]
        if out_of_bounds == "strict":
            acc[~valid_all] = 0.0
        beam[k] = acc
    if reduce == "none":
        return beam
    return beam.max(axis=0), beam.argmax(axis=0).astype(np.int32)
