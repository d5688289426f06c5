"""Generate tests/golden/features_golden.npz from the REFERENCE's own code (build container only).

``import BPMF`` fails here (obspy/h5py missing), so the functions on this path are lifted out of
BPMF/template_search.py by name with ``ast`` and executed unchanged: ``envelope`` (:1599-1617),
``saturated_envelopes`` (:1525-1574), ``time_dependent_threshold`` (:1418-1487).  ``envelope_parallel``
(:1576-1597) only distributes ``envelope`` over a process pool; a loop over the lifted ``envelope`` stands in.

    python tests/golden/make_golden_features.py
"""
import ast
import os
import time
import types

import numpy as np
from scipy.stats import median_abs_deviation as scimad

REF = "/root/reference/BPMF"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift(path, names, namespace):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), namespace)
    return {n: namespace[n] for n in names}


def main():
    cfg = types.SimpleNamespace(N_DEV_BP_THRESHOLD=10.0)
    ns = {"np": np, "cfg": cfg, "scimad": scimad, "give_time": time.time}
    fns = lift(os.path.join(REF, "template_search.py"), {"envelope", "saturated_envelopes", "time_dependent_threshold"}, ns)
    ns["envelope_parallel"] = lambda traces: np.float32(
        [fns["envelope"](t) for t in traces.reshape(-1, traces.shape[-1])]).reshape(traces.shape)

    rng = np.random.default_rng(20240726)
    g = {}
    # traces: noise + transients, a gap, a mostly-missing trace, a dead (tiny MAD) trace, a huge spike (clipped)
    S, C, N = 3, 2, 6075  # odd, 3^5 * 5^2: like BPMF's day-long records (86 400 s x 25/50/100 Hz) it has only small prime factors
    x = rng.standard_normal((S, C, N)).astype(np.float32) * np.float32(1e-6)
    x[0, 0, 2000:2100] *= 30.0
    x[0, 1, 1000:1500] = 0.0
    x[1, 0, : N // 2 + 50] = 0.0
    x[1, 1] = 0.0
    x[1, 1, 100] = 1e-20
    x[2, 0, 3000] = 5.0
    g["traces"] = x
    g["envelopes"] = ns["envelope_parallel"](x)
    feat, avail = fns["saturated_envelopes"](x.copy())
    g["features"], g["availability"] = feat, avail
    feat2, avail2 = fns["saturated_envelopes"](x.copy(), anomaly_threshold=1e-7, max_dynamic_range=50.0)
    g["features_b"], g["availability_b"] = feat2, avail2
    xe = rng.standard_normal((1, 1, 4096)).astype(np.float32)  # even length
    g["traces_even"] = xe
    g["envelopes_even"] = ns["envelope_parallel"](xe)
    # backprojection threshold
    for tag, n, window, overlap, cnr in (("a", 20000, 2000, 0.75, 10.0), ("b", 9137, 1500, 0.5, 6.0), ("c", 5300, 1000, 0.0, 8.0)):
        nr = np.abs(rng.standard_normal(n)).astype(np.float32)
        nr[n // 3: n // 3 + 40] += 25.0
        nr[n // 2: n // 2 + 700] = 0.0
        g["thr_%s_x" % tag] = nr
        g["thr_%s_params" % tag] = np.array([window, overlap, cnr])
        g["thr_%s_out" % tag] = np.asarray(fns["time_dependent_threshold"](nr, window, overlap=overlap, CNR_threshold=cnr))
    np.savez_compressed(os.path.join(HERE, "features_golden.npz"), **g)
    print("wrote features_golden.npz:", {k: (v.shape, str(v.dtype)) for k, v in g.items()})


if __name__ == "__main__":
    main()
