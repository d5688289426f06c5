"""Generate tests/golden/moveouts_golden.npz from the REFERENCE's own code (build container only).

``import BPMF`` fails here (obspy/h5py missing), so the two functions on this path are lifted out of the
reference files by name with ``ast`` and executed unchanged:
  * ``sec_to_samp``                         BPMF/utils.py:1257-1270
  * ``TravelTimes.convert_to_samples`` and ``TravelTimes.get_travel_times_array``
                                            BPMF/template_search.py:140-219
against a minimal stand-in for the ``TravelTimes`` object (a pandas table of per-(station, phase) arrays, as
``TravelTimes.read`` builds it).

    python tests/golden/make_golden_moveouts.py
"""
import ast
import os
import types

import numpy as np
import pandas as pd

REF = "/root/reference/BPMF"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift(path, names, namespace):
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, namespace)
            out[node.name] = namespace[node.name]
    return out


def main():
    cfg = types.SimpleNamespace(SAMPLING_RATE_HZ=50.0)
    ns_utils = {"np": np, "cfg": cfg}
    sec_to_samp = lift(os.path.join(REF, "utils.py"), {"sec_to_samp"}, ns_utils)["sec_to_samp"]
    utils = types.SimpleNamespace(sec_to_samp=sec_to_samp)
    ns_ts = {"np": np, "pd": pd, "utils": utils, "cfg": cfg}
    fns = lift(os.path.join(REF, "template_search.py"), {"convert_to_samples", "get_travel_times_array"}, ns_ts)

    rng = np.random.default_rng(20240726)
    g = {}
    for tag, dtype, sr, K, S, P in (("f32", np.float32, 50.0, 300, 5, 2), ("f64", np.float64, 100.0, 211, 4, 2),
                                    ("f32b", np.float32, 25.0, 64, 3, 1)):
        stations = ["ST%02d" % i for i in range(S)]
        phases = ["P", "S"][:P]
        tt = rng.uniform(0.0, 60.0, (K, S, P))
        # values that sit on sample boundaries (k / sr and k / sr - 0.2 / sr), where the float rounding decides
        kk = rng.integers(1, 3000, (K, S, P))
        edge = rng.random((K, S, P))
        tt = np.where(edge < 0.3, kk / sr, tt)
        tt = np.where((edge >= 0.3) & (edge < 0.6), (kk - 0.2) / sr, tt)
        tt = np.where((edge >= 0.6) & (edge < 0.65), -tt, tt)  # negative times keep their sign
        tt = tt.astype(dtype)
        obj = types.SimpleNamespace()
        obj.travel_times = pd.DataFrame({ph: {sta: tt[:, s, p] for s, sta in enumerate(stations)}
                                         for p, ph in enumerate(phases)})
        obj.phases = phases
        obj.n_sources = K
        fns["convert_to_samples"](obj, sr)
        samples = fns["get_travel_times_array"](obj, units="samples", stations=stations, phases=phases)
        rel = fns["get_travel_times_array"](obj, units="samples", stations=stations, phases=phases,
                                            relative_to_first=True)
        g[tag + "_tt"] = tt
        g[tag + "_sr"] = np.array([sr])
        g[tag + "_samples"] = np.asarray(samples, dtype=np.int64)
        g[tag + "_relative"] = np.asarray(rel, dtype=np.int64)
        g[tag + "_moveout_to_tt"] = np.asarray(obj.moveout_to_tt, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "moveouts_golden.npz"), **g)
    print("wrote moveouts_golden.npz:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
