"""Seeded synthetic inputs of the shapes SURVEY.md 8(d) prescribes (no real data needed)."""
import numpy as np

SEED = 20240726


def mf_case(T, S, C, L, N, seed=SEED, zero_gap=True, plant=True, mv_lo=-100, mv_hi=3000,
            normalize_weights=True, sparse_weights=False):
    """Traces ~N(0,1) RMS-normalised like BPMF/similarity_search.py:181-185, one station with a zero
    gap; templates ~N(0,1) demeaned/normalised like BPMF/dataset.py:4160-4165, a few cut from the
    data (planted detections); int32 moveouts in [mv_lo, mv_hi]; weights row-normalised
    (BPMF/similarity_search.py:469-472)."""
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((S, C, N)).astype(np.float32)
    if zero_gap and N > 4 * L:
        g0 = N // 3
        data[S // 2, :, g0: g0 + min(10000, N // 8)] = 0.0
    norm = data.std(axis=-1, keepdims=True)
    norm[norm == 0] = 1.0
    data /= norm
    templates = rng.standard_normal((T, S, C, L)).astype(np.float32)
    templates -= templates.mean(axis=-1, keepdims=True)
    templates /= templates.std(axis=-1, keepdims=True)
    mv_hi = int(min(mv_hi, max(0, (N - L) // 4)))
    mv_lo = int(max(mv_lo, -max(0, (N - L) // 8)))
    moveouts = rng.integers(mv_lo, mv_hi + 1, size=(T, S, C)).astype(np.int32)
    weights = np.ones((T, S, C), dtype=np.float32)
    if sparse_weights:
        keep = max(1, S // 2)
        order = np.argsort(moveouts.min(axis=2), axis=1)
        for t in range(T):
            weights[t, order[t, keep:], :] = 0.0
    planted = {}
    if plant:
        for t in range(0, T, max(1, T // 3)):
            lo_i, hi_i = max(0, -mv_lo), max(1, N - L - mv_hi - 1)
            if zero_gap and N > 4 * L:  # keep planted windows clear of the zero gap
                g1 = N // 3 + min(10000, N // 8)
                if N // 3 - L - mv_hi - 1 > lo_i:
                    hi_i = N // 3 - L - mv_hi - 1
                elif g1 - mv_lo < hi_i:
                    lo_i = g1 - mv_lo
            i0 = int(rng.integers(lo_i, max(lo_i + 1, hi_i)))
            for s in range(S):
                for c in range(C):
                    j = i0 + moveouts[t, s, c]
                    templates[t, s, c] = data[s, c, j: j + L]
            planted[t] = i0
    if normalize_weights:
        nrm = weights.sum(axis=(1, 2), keepdims=True)
        nrm[nrm == 0] = 1.0
        weights = weights / nrm
    return templates, moveouts, weights.astype(np.float32), data, planted


def beam_case(K, S, C, P, N, seed=SEED, tau_max=1500, sparse_sources=False, coherent=False):
    """Features |N(0,1)| clipped to [0,1]; identity-like phase weights (5_backprojection.ipynb:905);
    int64 moveouts relative to the first arrival (template_search.py:212-214); row-normalised source
    weights (template_search.py:890-893)."""
    rng = np.random.default_rng(seed)
    feat = np.clip(np.abs(rng.standard_normal((S, C, N))), 0, 1).astype(np.float32)
    wp = np.zeros((S, C, P), dtype=np.float32)
    for p in range(P):
        wp[:, p % C, p] = 1.0
    tau_max = int(min(tau_max, max(1, N // 4)))
    if coherent:
        src = rng.uniform(0, 1, (K, 3)) * np.array([1.0, 1.0, 0.3])
        src = src[np.lexsort((src[:, 2], src[:, 1], src[:, 0]))]
        sta = np.concatenate([rng.uniform(0, 1, (S, 2)), np.zeros((S, 1))], axis=1)
        dist = np.linalg.norm(src[:, None, :] - sta[None], axis=-1)
        tp = dist / dist.max() * tau_max / 1.73
    else:
        tp = rng.integers(0, int(tau_max / 1.73) + 1, size=(K, S)).astype(np.float64)
    tau = np.zeros((K, S, P), dtype=np.int64)
    for p in range(P):
        tau[:, :, p] = np.floor(tp * (1.0 if p == 0 else 1.73)).astype(np.int64)
    tau -= tau.min(axis=(1, 2), keepdims=True)
    ws = np.ones((K, S), dtype=np.float32)
    if sparse_sources:
        keep = max(1, S // 3)
        order = np.argsort(tau[:, :, 0], axis=1)
        for k in range(K):
            ws[k, order[k, keep:]] = 0.0
    ws /= ws.sum(axis=1, keepdims=True)
    return feat, tau, wp, ws


def day_case(S=8, K=343, N=60000, sr=50.0, n_events=3, seed=11):
    """A miniature "day": S stations x 2 components (component 0 records the P phase, component 1 the S phase) of
    unit-variance noise, a cubic grid of K sources under the network with P / S travel times (seconds), and
    `n_events` distinct sources that each fire three times -- once strongly (8 sigma), twice weakly (2.5 sigma) --
    with a source-and-channel specific 2 s wavelet.  Returns a dict with the traces, the travel-time table and the
    ground truth (source, first-arrival sample of every occurrence)."""
    rng = np.random.default_rng(seed)
    n1 = int(round(K ** (1 / 3)))
    K = n1 ** 3
    g = np.stack(np.meshgrid(*(np.linspace(0.1, 0.9, n1),) * 3, indexing="ij"), -1).reshape(-1, 3)
    g = g * np.array([60.0, 60.0, 20.0])
    sta = np.concatenate([rng.uniform(0, 60, (S, 2)), np.zeros((S, 1))], axis=1)
    dist = np.linalg.norm(g[:, None, :] - sta[None], axis=-1)
    tt = np.stack([dist / 6.0, dist / 6.0 * 1.73], axis=-1).astype(np.float32)  # (K, S, 2) seconds
    traces = rng.standard_normal((S, 2, N)).astype(np.float32)
    wl = int(2 * sr)
    taper = np.hanning(wl)
    events = rng.choice(K, n_events, replace=False)
    truth = []
    samp = np.int64(np.sign(tt) * np.int64(np.abs(tt * np.float32(sr)) + np.float32(0.2)))  # utils.sec_to_samp
    rel = samp - samp.min(axis=(1, 2), keepdims=True)
    slots = rng.permutation(np.arange(3000, N - 8000, 4500))[: 3 * n_events]
    for e, k in enumerate(events):
        wav = rng.standard_normal((S, 2, wl)) * taper
        wav /= np.sqrt((wav ** 2).mean(axis=-1, keepdims=True))
        for r, amp in enumerate((8.0, 2.5, 2.5)):
            i0 = int(slots[3 * e + r])
            for s in range(S):
                for c in range(2):
                    j = i0 + int(rel[k, s, c])
                    traces[s, c, j: j + wl] += (amp * wav[s, c]).astype(np.float32)
            truth.append((int(k), i0, amp))
    wp = np.zeros((S, 2, 2), dtype=np.float32)
    wp[:, 0, 0] = 1.0
    wp[:, 1, 1] = 1.0
    ws = np.full((K, S), 1.0 / S, dtype=np.float32)
    return dict(traces=traces, travel_times=tt, sr=sr, weights_phases=wp, weights_sources=ws, phase_on_comp=[0, 1],
                truth=truth, grid=g)
