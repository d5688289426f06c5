"""Generate tests/golden/peaks_golden.npz from the REFERENCE's own code (build container only).

``utils._detect_peaks`` (BPMF/utils.py:2203-2352) is lifted by name with ``ast`` and executed unchanged; the peak
selection of ``Beamformer.find_detections`` (BPMF/template_search.py:603-628: peaks above the threshold, then "keep
the largest peak for grouped detection") is lifted as the statements of that method between the `# select peaks` and
`source_indexes = ...` lines, executed unchanged on a stand-in ``self``.

    python tests/golden/make_golden_peaks.py
"""
import ast
import os
import textwrap
import types

import numpy as np

REF = "/root/reference/BPMF"
HERE = os.path.dirname(os.path.abspath(__file__))


def lift(path, names, namespace):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), namespace)
    return {n: namespace[n] for n in names}


def lift_find_detections_core():
    """The lines of Beamformer.find_detections from '# select peaks' to 'source_indexes = ...' as a function of
    (self, detection_threshold, minimum_interevent_time [samples])."""
    src = open(os.path.join(REF, "template_search.py")).read().splitlines()
    a = next(i for i, ln in enumerate(src) if ln.strip() == "# select peaks")
    b = next(i for i, ln in enumerate(src) if i > a and ln.strip().startswith("source_indexes = self.maxbeam_sources"))
    body = textwrap.dedent("\n".join(src[a: b + 1]))
    fn = "def core(self, detection_threshold, minimum_interevent_time):\n" + textwrap.indent(body, "    ") + \
         "\n    return peak_indexes, source_indexes\n"
    return fn


def main():
    ns = {"np": np}
    detect = lift(os.path.join(REF, "utils.py"), {"_detect_peaks"}, ns)["_detect_peaks"]
    ns2 = {"np": np, "_detect_peaks": detect}
    exec(lift_find_detections_core(), ns2)
    core = ns2["core"]
    rng = np.random.default_rng(20240726)
    g = {}
    cases = {
        "a": (20000, 250, 0),    # smooth-ish beam, events
        "b": (9137, 60, 1),      # many exact ties (rounded values), flat peaks
        "c": (5000, 1, 0),       # mpd = 1: no suppression
        "d": (30011, 777, 2),    # NaN runs
    }
    for tag, (n, mpd, kind) in cases.items():
        x = np.abs(rng.standard_normal(n)).astype(np.float32)
        x = np.convolve(x, np.ones(7, np.float32) / 7, mode="same").astype(np.float32)
        for i0 in rng.integers(100, n - 100, 12):
            x[i0 - 20: i0 + 20] += (np.hanning(40) * rng.uniform(2, 9)).astype(np.float32)
        if kind == 1:
            x = np.round(x, 1).astype(np.float32)
        if kind == 2:
            x[4000:4100] = np.nan
            x[12000] = np.nan
        thr = (np.nanmedian(x) + 3.0 * np.nanstd(x)) * np.ones(n, np.float32)
        thr[: n // 2] *= 1.2
        srcs = rng.integers(0, 100000, n).astype(np.int32)
        g[f"{tag}_x"], g[f"{tag}_thr"], g[f"{tag}_src"], g[f"{tag}_mpd"] = x, thr.astype(np.float32), srcs, np.array([mpd])
        g[f"{tag}_peaks_all"] = detect(x, mpd=mpd).astype(np.int64)
        me = types.SimpleNamespace(maxbeam=x, maxbeam_sources=srcs)
        with np.errstate(invalid="ignore"):
            pk, si = core(me, thr, mpd)
        g[f"{tag}_peaks"], g[f"{tag}_sources"] = np.asarray(pk).astype(np.int64), np.asarray(si).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "peaks_golden.npz"), **g)
    print({k: (v.shape, str(v.dtype)) for k, v in g.items() if not k.endswith(("_x", "_thr", "_src"))})


if __name__ == "__main__":
    main()
