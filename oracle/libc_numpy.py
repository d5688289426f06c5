"""NumPy restatement of the reference's BPMF/libc.c exports (TEST INFRASTRUCTURE).

Race-free ("intended") semantics of each function, each citing the libc.c lines it follows.  It is
pinned against tests/golden/libc_golden.npz, which was produced by the reference's own libc.c
(tests/golden/make_golden.py).  Float work is done in float64 and cast, so agreement with the
reference is to float32 round-off; integer outputs are exact.
"""
import numpy as np

GAUSSIAN_SAMPLE_LEN = 500  # libc.c:526 ; clib.py:284


def time_dependent_threshold(time_series, gaussian_sample, num_dev, half_window_samp, shift_samp):
    """libc.c:516-673.  Returns (threshold f32[n], filled_series f32[n])."""
    x = np.array(time_series, dtype=np.float32)
    n = x.size
    W = 2 * half_window_samp  # :527
    n_win = (n - (W - shift_samp)) // shift_samp  # :528
    n_glob = n // W  # :553
    head = x[: n_glob * W].astype(np.float64)
    nz = head != 0.0  # :561
    cnt = int(nz.sum())
    center = np.float32(head[nz].sum() / cnt) if cnt else np.float32(np.nan)  # :573
    dev = np.float32(np.sqrt(((head[nz] - np.float64(center)) ** 2).sum() / cnt)) if cnt else np.float32(np.nan)
    zero = x == 0.0  # :609
    g = np.asarray(gaussian_sample, dtype=np.float32)
    idx = np.arange(n) % GAUSSIAN_SAMPLE_LEN
    x[zero] = (center + g[idx[zero]] * dev).astype(np.float32)  # :610
    xd = x.astype(np.float64)
    c1 = np.concatenate([[0.0], np.cumsum(xd)])
    tw = np.empty(n_win, dtype=np.float32)
    for i in range(n_win):  # :618-627
        o = i * shift_samp
        seg = xd[o: o + W]
        m = np.float32((c1[o + W] - c1[o]) / W)
        s = np.float32(np.sqrt(((seg - np.float64(m)) ** 2).sum() / W))
        tw[i] = np.float32(m + np.float32(num_dev) * s)
    # "delay the jump" smoothing, libc.c:631-651, in float32 like the reference
    d0 = tw[1:] - tw[:-1]
    tw1 = tw.copy()
    neg = d0 < 0
    tw1[1:][neg] = tw[1:][neg] - d0[neg]
    d1 = tw1[1:] - tw1[:-1]
    tw2 = tw1.copy()
    pos = d1 > 0
    tw2[:-1][pos] = tw1[:-1][pos] + d1[pos]
    # step-wise expansion, libc.c:654-669 (interior index clamped: SURVEY A.3)
    i = np.arange(n)
    w = np.minimum(i // shift_samp, n_win - 1)
    w[i < shift_samp] = 0
    w[i >= n - shift_samp] = n_win - 1
    return tw2[w].astype(np.float32), x


def select_cc_indexes(cc, threshold, search_win):
    """libc.c:441-485 (sequential scan; exact)."""
    cc = np.asarray(cc, dtype=np.float32)
    thr = np.asarray(threshold, dtype=np.float32)
    n = cc.size
    sel = np.zeros(n, dtype=np.int32)
    for i in range(n):
        sel[i] = 1 if cc[i] > thr[i] else 0
        start = 0 if i <= search_win else i - search_win
        for j in range(start, i):
            if cc[j] > cc[i]:
                sel[i] = 0
                break
            sel[j] = 0
    return sel


def kurtosis(signal, W):
    """libc.c:11-53: sample-corrected kurtosis of the W samples preceding n; 0 where m2 <= 1e-6."""
    sig = np.asarray(signal, dtype=np.float64)
    S, C, N = sig.shape
    out = np.zeros((S, C, N), dtype=np.float32)
    Wf = float(W)
    for s in range(S):
        for c in range(C):
            win = np.lib.stride_tricks.sliding_window_view(sig[s, c], W)[: N - W]  # rows nn = n-W
            mean = win.mean(axis=1, keepdims=True)
            ds = win - mean
            m2 = (ds ** 2).mean(axis=1)
            m4 = (ds ** 4).mean(axis=1)
            ok = m2 > 1e-6
            k = np.zeros(N - W)
            k[ok] = 1.0 / ((Wf - 2) * (Wf - 3)) * ((Wf ** 2 - 1) * m4[ok] / m2[ok] ** 2 - 3 * (Wf - 1) ** 2)
            out[s, c, W:] = k
    return out


def _acc_sq(a, b):
    """dt2 += pow(a - b, 2) exactly as C evaluates it: float subtract, double square, float store."""
    return (a.astype(np.float32) - b.astype(np.float32)).astype(np.float64) ** 2


def find_similar_moveouts(moveouts, lon, lat, cell_lon, cell_lat, threshold, n_for_diff, variant):
    """libc.c:55-223 (variant 1, 'smallest') and libc.c:225-387 (variant 2, 'closest'); exact."""
    mv = np.asarray(moveouts, dtype=np.float32)
    K, S = mv.shape
    lon = np.asarray(lon, dtype=np.float32)
    lat = np.asarray(lat, dtype=np.float32)
    thr2 = np.float32(float(np.float32(n_for_diff)) * float(np.float32(threshold)) ** 2)  # :79,:249
    red = np.zeros(K, dtype=np.int32)

    def dt2_of(n1, cand):
        if variant == 2:
            order = _selection_argsort(mv[n1])[:n_for_diff]  # sort_indexes :389-410
            acc = np.zeros(cand.size, dtype=np.float32)
            for s in order:
                acc = (acc.astype(np.float64) + _acc_sq(np.full(cand.size, mv[n1, s]), mv[cand, s])).astype(np.float32)
            return acc
        sq = np.sort(_acc_sq(np.broadcast_to(mv[n1], (cand.size, S)), mv[cand]).astype(np.float32), axis=1)
        acc = np.zeros(cand.size, dtype=np.float32)
        for s in range(n_for_diff):
            acc = (acc + sq[:, s]).astype(np.float32)
        return acc

    def sweep(members):
        for n1 in members:
            if n1 >= K - 1 or red[n1]:  # libc.c:121,130 / :288,297
                continue
            cand = members[(members > n1)]
            cand = cand[red[cand] == 0]
            if cand.size == 0:
                continue
            hit = dt2_of(n1, cand) < thr2
            red[cand[hit]] = 1

    ncl, nct = len(cell_lon) - 1, len(cell_lat) - 1
    for i in range(ncl):
        for j in range(nct):
            inside = ((lon >= cell_lon[i]) & (lon < cell_lon[i + 1]) & (lat >= cell_lat[j]) & (lat < cell_lat[j + 1]))
            sweep(np.nonzero(inside)[0])
    sweep(np.arange(K))
    return red


def _selection_argsort(v):
    """sort_indexes, libc.c:389-410: selection sort on indexes (not stable: swaps)."""
    idx = list(range(len(v)))
    n = len(v)
    for i in range(n - 1):
        m = i
        for j in range(i + 1, n):
            if v[idx[j]] < v[idx[m]]:
                m = j
        idx[m], idx[i] = idx[i], idx[m]
    return np.array(idx)
