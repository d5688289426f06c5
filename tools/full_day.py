"""One synthetic day through the whole chain (seismic_bpmf_b200/workflow.py) at a realistic size on ONE GPU, with
per-stage wall-clock times: BASELINE.json configs[4] scaled to one B200 (the 8-GPU version shards the source grid
and the templates, seismic_bpmf_b200/sharded.py).  Host arrays in, host arrays out at every stage, like BPMF.

    python tools/full_day.py [--stations 16] [--sources 97336] [--samples 4320000] [--events 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seismic_bpmf_b200 import workflow  # noqa: E402
from tests import _synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stations", type=int, default=16)
    ap.add_argument("--sources", type=int, default=97336)
    ap.add_argument("--samples", type=int, default=4320000)
    ap.add_argument("--events", type=int, default=10)
    ap.add_argument("--fused", action="store_true", help="keep the CC slab on the device for the detection step")
    a = ap.parse_args()
    t0 = time.perf_counter()
    case = _synth.day_case(S=a.stations, K=a.sources, N=a.samples, n_events=a.events, seed=5)
    t_gen = time.perf_counter() - t0
    be = workflow.gpu_backend(fused_detection=a.fused)
    times = {}

    def timed(name, fn):
        def wrapped(*args, **kw):
            t = time.perf_counter()
            out = fn(*args, **kw)
            times[name] = times.get(name, 0.0) + time.perf_counter() - t
            return out
        return wrapped

    for name in ("saturated_envelopes", "bp_threshold", "moveouts_from_travel_times", "beamform", "matched_filter",
                 "mf_threshold", "select_cc_indexes", "intertemplate_cc", "matched_filter_detect"):
        if hasattr(be, name):
            setattr(be, name, timed(name, getattr(be, name)))
    # warm-up of the libraries (cuFFT plans, cached device buffers) on a small case
    small = _synth.day_case()
    workflow.full_day(small["traces"], small["travel_times"], small["sr"], small["weights_phases"],
                      small["weights_sources"], small["phase_on_comp"], backend=be)
    times.clear()
    t0 = time.perf_counter()
    res = workflow.full_day(case["traces"], case["travel_times"], case["sr"], case["weights_phases"],
                            case["weights_sources"], case["phase_on_comp"], backend=be, bp_window=90000, mf_window=90000)
    total = time.perf_counter() - t0
    truth = case["truth"]
    picks = res.backprojection.picks
    hit = sum(any(20 <= int(p) - i0 <= 80 for p in picks) for k, i0, amp in truth)
    n_det = int(sum(len(d) for d in res.detections))
    K, S, P = case["travel_times"].shape
    print(json.dumps({
        "workload": "%d stations x 2 components x %d samples @ %g Hz, %d grid sources x %d phases, %d planted occurrences"
                    % (S, a.samples, case["sr"], K, P, len(truth)),
        "synthetic_data_generation_s": t_gen, "total_s": total, "stages_s": {k: round(v, 4) for k, v in times.items()},
        "host_glue_s": round(total - sum(times.values()), 4), "picks": int(len(picks)), "planted_occurrences_picked": int(hit),
        "templates": int(len(res.origin)), "matched_filter_detections": n_det, "fused_detection": bool(a.fused)}))


if __name__ == "__main__":
    main()
