"""Small driver for an ncu launch list of the kernels either side of the two hot paths: detection
traces, backprojection threshold, peak picking, gated detection on a CC slab.  One pass of each at day length."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seismic_bpmf_b200 import device as dev, features as ft, template_search as ts  # noqa: E402

rng = np.random.default_rng(1)
S, C, N = 8, 3, 2160000
traces = (rng.standard_normal((S, C, N)) * 1e-7).astype(np.float32)
for rep in range(2):
    feat, avail = ft.saturated_envelopes(traces)
    nr = np.abs(rng.standard_normal(N)).astype(np.float32)
    nr = np.convolve(nr, np.ones(25, np.float32) / 25, mode="same").astype(np.float32)
    thr = ft.time_dependent_threshold(nr, 45000, overlap=0.75, CNR_threshold=10.0)
    pk, src = ts.find_detections(nr, np.zeros(N, np.int32), thr, 250)
    x = torch.randn((64, 8639232), device="cuda") * 0.03
    g = torch.randn(500, device="cuda")
    det = dev.DetectBatches(1, 64, 8639201, 90000, 135000, capacity=1 << 16)
    det.run(0, x, g, 8.0, search_win=6000, mode="select", thr_cap=torch.ones(64, device="cuda"), max_kurtosis=20.0)
    det.collect()
print("peaks", len(pk))
