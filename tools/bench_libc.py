"""The libc.so rows timed against the REAL reference: `oracle/_ref/lib/libc.so` is BPMF/libc.c
compiled unchanged (oracle/build.py), called here through the same C signatures on the same inputs as
libbpmf_b200.so, with parity checked in the same run.

    python tools/bench_libc.py            # one JSON object on stdout

`run()` is what bench.py calls for the "libc" object of its JSON line (rank 0, N = 1 only).
Workloads follow BPMF's callers: one CC series of a day @ 100 Hz (n = 8 639 201), 30 min windows, overlap 0.25,
8 x rms (BPMF/similarity_search.py:548-560), declustering over 60 s; the grid-thinning sweep
(BPMF/template_search.py -> clib.find_similar_sources, libc.c:225-387) on a regular grid under 32 stations.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_FP, _IP, _Z = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_size_t


def _ref_lib():
    """The reference's own libc.c, compiled by oracle/build.py (portable x86-64-v3 build: the reference sources do
    not travel to the GPU box, so it cannot be rebuilt there with -march=native)."""
    path = os.path.join(ROOT, "oracle", "_ref", "lib", "libc.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.time_dependent_threshold.argtypes = [_FP, _FP, C.c_float, _Z, _Z, _Z, C.c_int, _FP]
    lib.time_dependent_threshold.restype = None
    lib.select_cc_indexes.argtypes = [_FP, _FP, _Z, _Z, _IP]
    lib.select_cc_indexes.restype = None
    lib.find_similar_moveouts2.argtypes = [_FP, _FP, _FP, _FP, _FP, C.c_float, _Z, _Z, _Z, _Z, _Z, C.c_int, _IP]
    lib.find_similar_moveouts2.restype = None
    return lib


def _p(a, t):
    return a.ctypes.data_as(t)


class _quiet:
    """The reference prints its progress with printf: keep fd 1 clean (bench.py prints ONE JSON line)."""

    def __enter__(self):
        sys.stdout.flush()
        self.libc = C.CDLL(None)
        self.libc.fflush(None)
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        self.libc.fflush(None)  # drain the C buffer into /dev/null BEFORE fd 1 is restored
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def _best(fn, reps):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def _grid_case(K, S, seed=5):
    """Sources on a regular 3-D grid under a surface network, P travel times in seconds (what
    clib.find_similar_sources receives), 2 x 2 longitude/latitude cells."""
    rng = np.random.default_rng(seed)
    n1 = int(np.ceil(K ** (1 / 3)))
    g = np.stack(np.meshgrid(*(np.linspace(0.0, 1.0, n1),) * 3, indexing="ij"), -1).reshape(-1, 3)[:K]
    sta = np.concatenate([rng.uniform(0, 1, (S, 2)), np.zeros((S, 1))], axis=1)
    tt = (np.linalg.norm((g[:, None, :] - sta[None]) * np.array([100.0, 100.0, 30.0]), axis=-1) / 6.0).astype(np.float32)
    lon, lat = g[:, 0].astype(np.float32), g[:, 1].astype(np.float32)
    cells = np.linspace(0.0, 1.001, 3).astype(np.float32)
    return tt, lon, lat, cells


def run(n=8639201, window=180000, overlap=0.25, num_dev=8.0, search_win=6000, K_ref=20000, K_gpu=100000, S=32,
        device_index=0):
    import torch

    from seismic_bpmf_b200 import _lib, clib, device as dev

    torch.cuda.set_device(device_index)
    ours = _lib.load()
    ref = _ref_lib()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(11)
    out = {"reference": "BPMF/libc.c compiled unchanged -> oracle/_ref/lib/libc.so (gcc -O3 -fopenmp -march=x86-64-v3)"
           if ref is not None else "unavailable (oracle/_ref/lib/libc.so missing)", "host_cores": cores}

    # ------------------------------------------------------------ time_dependent_threshold, one CC series of a day
    x = (0.04 * rng.standard_normal(n)).astype(np.float32)
    x[n // 3: n // 3 + 10000] = 0.0
    for i0 in (12345, n // 2, n - 5000):
        x[i0] = 0.9
    wn = rng.standard_normal(500).astype(np.float32)
    half, shift = window // 2, int((1.0 - overlap) * window)

    def ours_thr():
        return clib.time_dependent_threshold(x, window, num_dev, overlap=overlap, white_noise=wn)

    thr = ours_thr()
    rec = {"workload": "n = %d samples, %d-sample windows, overlap %.2f, %g x rms" % (n, window, overlap, num_dev),
           "ours_host_call_ms": 1e3 * _best(ours_thr, 3)}
    if ref is not None:
        def ref_thr(nt):
            xs, th = x.copy(), np.zeros(n, np.float32)
            with _quiet():
                t0 = time.perf_counter()
                ref.time_dependent_threshold(_p(xs, _FP), _p(wn, _FP), num_dev, n, half, shift, nt, _p(th, _FP))
                dt = time.perf_counter() - t0
            return dt, th

        t1, th1 = min((ref_thr(1) for _ in range(2)), key=lambda r: r[0])
        tn, _ = min((ref_thr(cores) for _ in range(2)), key=lambda r: r[0])
        rec.update({"reference_1_thread_ms": 1e3 * t1, "reference_%d_threads_ms" % cores: 1e3 * tn,
                    "max_abs_diff_vs_reference_1_thread": float(np.abs(thr - th1).max()),
                    "note": "the reference's multi-threaded result races (SURVEY section 5): parity is taken at 1 thread"})
    # batched device form on a (B, n) slab that is already resident (what the fused detection uses)
    B = 32
    slab = torch.from_numpy(x).cuda().repeat(B, 1).contiguous()
    g = torch.from_numpy(wn).cuda()
    for _ in range(2):
        dev.time_dependent_threshold_dev(slab.clone(), g, num_dev, half, shift)
    torch.cuda.synchronize()
    work = slab.clone()
    t0 = time.perf_counter()
    th_d = dev.time_dependent_threshold_dev(work, g, num_dev, half, shift)
    torch.cuda.synchronize()
    rec["ours_device_batched_ms_per_series"] = 1e3 * (time.perf_counter() - t0) / B
    rec["device_batch"] = B
    rec["max_abs_diff_device_vs_host_call"] = float((th_d[0].cpu() - torch.from_numpy(thr)).abs().max())
    del slab, work, th_d
    out["time_dependent_threshold"] = rec

    # ------------------------------------------------------------ select_cc_indexes on the same series
    def ours_sel():
        return clib.select_cc_indexes(x, thr, search_win)

    sel = ours_sel()
    rec = {"workload": "n = %d, search_win %d" % (n, search_win), "ours_host_call_ms": 1e3 * _best(ours_sel, 3),
           "selected": int(sel.sum())}
    if ref is not None:
        sel_r = np.zeros(n, np.int32)

        def ref_sel():
            sel_r[:] = 0
            with _quiet():
                ref.select_cc_indexes(_p(x, _FP), _p(thr, _FP), search_win, n, _p(sel_r, _IP))

        rec["reference_ms"] = 1e3 * _best(ref_sel, 2)
        rec["identical_to_reference"] = bool(np.array_equal(sel, sel_r.astype(bool)))
    out["select_cc_indexes"] = rec

    # ------------------------------------------------------------ find_similar_moveouts2 (grid thinning)
    rec = {}
    tt, lon, lat, cells = _grid_case(K_ref, S)
    thr_s = 0.05

    def ours_fsm(tt_=tt, lon_=lon, lat_=lat):
        return clib.find_similar_sources(tt_, lon_, lat_, cells, cells, thr_s, num_threads=cores,
                                         num_stations_for_diff=10, method="closest")

    red = ours_fsm()
    rec["K_%d" % K_ref] = {"ours_ms": 1e3 * _best(ours_fsm, 2), "redundant": int(red.sum())}
    if ref is not None:
        flat = np.ascontiguousarray(tt.reshape(-1))
        red_r = np.zeros(K_ref, np.int32)

        def ref_fsm(nt):
            red_r[:] = 0
            with _quiet():
                t0 = time.perf_counter()
                ref.find_similar_moveouts2(_p(flat, _FP), _p(lon, _FP), _p(lat, _FP), _p(cells, _FP), _p(cells, _FP),
                                           thr_s, K_ref, S, 2, 2, 10, nt, _p(red_r, _IP))
                dt = time.perf_counter() - t0
            return dt

        t1 = ref_fsm(1)
        same = bool(np.array_equal(red, red_r.astype(bool)))
        rec["K_%d" % K_ref].update({"reference_1_thread_ms": 1e3 * t1, "identical_to_reference_1_thread": same,
                                    "reference_%d_threads_ms" % cores: 1e3 * ref_fsm(cores)})
    tt2, lon2, lat2, _ = _grid_case(K_gpu, S)
    red2 = ours_fsm(tt2, lon2, lat2)
    rec["K_%d" % K_gpu] = {"ours_ms": 1e3 * _best(lambda: ours_fsm(tt2, lon2, lat2), 1), "redundant": int(red2.sum()),
                           "note": "the reference sweep is quadratic per cell: not timed at this size"}
    rec["workload"] = "regular grid under %d stations, 2 x 2 cells, threshold %.2f s, 10 closest stations" % (S, thr_s)
    out["find_similar_moveouts2"] = rec
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
