"""Generate tests/golden/libc_golden.npz from the REFERENCE's own code.

Runs only in the build container (needs /root/reference): the reference's BPMF/libc.c is compiled
to oracle/_ref/lib/libc.so by oracle/build.py and driven through the reference's own
BPMF/clib.py wrappers (stub-package loader, oracle/ref_loader.py), single-threaded so the racy
OpenMP reductions of libc.c:554-586 are deterministic.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    clib = ref_loader.load_reference_clib()
    rng = np.random.default_rng(20240726)
    g = {}

    # ---- time_dependent_threshold (libc.c:516-673 via clib.py:257-309) ----
    wn = rng.standard_normal(500).astype(np.float32)
    g["thr_white_noise"] = wn
    cases = [("a", 6000, 600, 0.25, 8.0), ("b", 9137, 500, 0.0, 5.0), ("c", 12000, 1000, 0.4, 8.0)]
    for tag, n, win, overlap, num_dev in cases:
        x = (0.05 * rng.standard_normal(n)).astype(np.float32)
        x[n // 3: n // 3 + 700] = 0.0  # data gap -> zero CC (similarity_search.py:615)
        x[n // 2] = 0.9  # a detection
        thr = clib.time_dependent_threshold(x, win, num_dev, overlap=overlap, white_noise=wn,
                                            num_threads=1)
        g[f"thr_{tag}_x"] = x
        g[f"thr_{tag}_params"] = np.array([win, overlap, num_dev], dtype=np.float64)
        g[f"thr_{tag}_out"] = thr

    # ---- select_cc_indexes (libc.c:441-485 via clib.py:224-255) ----
    for tag, n, win in [("a", 5000, 40), ("b", 3001, 500), ("c", 64, 100)]:
        cc = (0.1 * rng.standard_normal(n)).astype(np.float32)
        cc[rng.integers(0, n, 12)] += 0.6
        cc[n // 2: n // 2 + 5] = 0.7  # plateau: ties
        thr = (0.25 + 0.02 * rng.standard_normal(n)).astype(np.float32)
        sel = clib.select_cc_indexes(cc, thr, win)
        g[f"sel_{tag}_cc"] = cc
        g[f"sel_{tag}_thr"] = thr
        g[f"sel_{tag}_win"] = np.array([win])
        g[f"sel_{tag}_out"] = sel

    # ---- kurtosis (libc.c:11-53 via clib.py:87-101) ----
    sig = rng.standard_normal((2, 3, 1500)).astype(np.float32)
    sig[1, 2, 400:700] = 0.0  # m2 below 1e-6 -> output left at 0
    sig[0, 1, 800] = 25.0
    g["kurt_signal"] = sig
    g["kurt_W"] = np.array([100])
    g["kurt_out"] = clib.kurtosis(sig, 100)

    # ---- find_similar_sources (libc.c:55-387 via clib.py:104-221) ----
    K, S = 400, 9
    lon = rng.uniform(30.0, 31.0, K).astype(np.float32)
    lat = rng.uniform(40.0, 41.0, K).astype(np.float32)
    sta = rng.uniform(0, 1, (S, 2))
    pos = np.stack([(lon - 30.0), (lat - 40.0)], axis=1)
    tt = (np.linalg.norm(pos[:, None, :] - sta[None], axis=-1) * 100.0 / 6.0).astype(np.float32)
    tt -= tt.min(axis=1, keepdims=True)
    cell_lon = np.linspace(30.0, 31.0 + 1e-3, 4).astype(np.float32)
    cell_lat = np.linspace(40.0, 41.0 + 1e-3, 3).astype(np.float32)
    g["fsm_moveouts"], g["fsm_lon"], g["fsm_lat"] = tt, lon, lat
    g["fsm_cell_lon"], g["fsm_cell_lat"] = cell_lon, cell_lat
    for tag, method, nfd, thr in [("closest_all", "closest", None, 0.35),
                                  ("closest_5", "closest", 5, 0.25),
                                  ("smallest_5", "smallest", 5, 0.15)]:
        red = clib.find_similar_sources(tt, lon, lat, cell_lon, cell_lat, thr, num_threads=1,
                                        num_stations_for_diff=nfd, method=method)
        g[f"fsm_{tag}_params"] = np.array([thr, -1 if nfd is None else nfd], dtype=np.float64)
        g[f"fsm_{tag}_out"] = red
        print(tag, int(red.sum()), "of", K, "redundant", file=sys.stderr)

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libc_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes", file=sys.stderr)


if __name__ == "__main__":
    main()
