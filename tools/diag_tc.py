"""Diagnostic: tensor-path error vs oracle over template lengths / moveout alignment (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cpu
from seismic_bpmf_b200 import fast_matched_filter as fmf
from tests import _synth

for L in (400, 700, 786, 787, 794, 800, 801):
    for mvmode in ("mult8", "random"):
        tp, mv, w, d, _ = _synth.mf_case(20, 3, 2, L, 30000, seed=L)
        if mvmode == "mult8":
            mv = (mv // 8) * 8
        want = cpu.matched_filter(tp, mv, w, d, 1)
        out = []
        for prec in ("fp16", "fp16x3"):
            got = fmf.matched_filter(tp, mv, w, d, 1, precision=prec, check_zeros=False)
            err = np.abs(got - want)
            t_bad = int(err.max(axis=1).argmax())
            out.append("%s err=%.2e (t=%d, i=%d)" % (prec, err.max(), t_bad, int(err[t_bad].argmax())))
        print("L=%d %s: %s" % (L, mvmode, " | ".join(out)))
