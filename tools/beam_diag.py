"""Where beam_group_kernel waits (diagnostic build only: beam.cu compiled with -DBEAM_DIAG, selected with
BPMF_B200_LIB).  Prints cycles per record for the producer (free-stage wait, parameter wait, total) and for
consumer warp 0 (full-stage wait, group epilogues, total).  python tools/beam_diag.py [--keep 10] [--stations 32]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_beam  # noqa: E402
from seismic_bpmf_b200 import _lib  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sources", type=int, default=20000)
    ap.add_argument("--samples", type=int, default=432000)
    ap.add_argument("--stations", type=int, default=32)
    ap.add_argument("--keep", type=int, default=32)
    a = ap.parse_args()
    lib = _lib.load()
    out = (C.c_ulonglong * 8)()
    bench_beam.run(a.sources, a.samples, a.stations, min(a.keep, a.stations), False, 1, 0)
    lib.bpmf_beam_diag(out, 1)
    r = bench_beam.run(a.sources, a.samples, a.stations, min(a.keep, a.stations), False, 1, 0)
    lib.bpmf_beam_diag(out, 1)
    v = [int(x) for x in out]
    rec, grp = max(v[6], 1), max(v[7], 1)
    print(json.dumps({"stations": a.stations, "keep": a.keep, "ms": r["ms_per_step"], "frac": r["roofline"]["frac"],
                      "records": v[6], "groups": v[7], "rows_per_group": v[6] / grp - 1,
                      "producer_clk_per_record": {"wait_free_stage": v[0] / rec, "wait_params": v[1] / rec, "total": v[2] / rec},
                      "consumer0_clk_per_record": {"wait_full_stage": v[3] / rec, "epilogue": v[4] / rec, "total": v[5] / rec},
                      "consumer0_epilogue_clk_per_group": v[4] / grp}))
