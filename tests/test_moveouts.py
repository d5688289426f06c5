"""Travel times -> moveouts: the NumPy restatement against vectors made with the reference's own
``sec_to_samp`` / ``get_travel_times_array`` (CPU), and the CUDA kernel against the same vectors, bit for bit."""
import os

import numpy as np
import pytest

from oracle import moveouts_numpy as mo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "moveouts_golden.npz")
TAGS = ["f32", "f64", "f32b"]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLD)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_vectors(golden, tag):
    tt, sr = golden[tag + "_tt"], float(golden[tag + "_sr"][0])
    samples, _ = mo.moveouts_from_travel_times(tt, sr, relative_to_first=False)
    assert np.array_equal(samples, golden[tag + "_samples"])
    rel, m2t = mo.moveouts_from_travel_times(tt, sr, relative_to_first=True)
    assert np.array_equal(rel, golden[tag + "_relative"])
    assert np.array_equal(m2t, golden[tag + "_moveout_to_tt"])
    assert rel.min(axis=(1, 2)).max() == 0  # every source's earliest arrival is at sample 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_kernel_matches_reference_vectors_bit_exact(golden, tag):
    from seismic_bpmf_b200 import template_search as ts

    tt, sr = golden[tag + "_tt"], float(golden[tag + "_sr"][0])
    mv, m2t = ts.moveouts_from_travel_times(tt, sr, relative_to_first=False)
    assert mv.dtype == np.int32 and np.array_equal(mv, golden[tag + "_samples"])
    assert np.all(m2t == 0)
    mv, m2t = ts.moveouts_from_travel_times(tt, sr, relative_to_first=True)
    assert np.array_equal(mv, golden[tag + "_relative"])
    assert np.array_equal(m2t, golden[tag + "_moveout_to_tt"])


@pytest.mark.gpu
def test_kernel_matches_oracle_large_and_table_cache():
    from seismic_bpmf_b200 import _lib, template_search as ts

    rng = np.random.default_rng(3)
    K, S, P = 20011, 32, 2
    for dtype, sr in ((np.float32, 50.0), (np.float64, 100.0)):
        tt = rng.uniform(0, 80, (K, S, P))
        kk = rng.integers(0, 4000, (K, S, P))
        tt = np.where(rng.random((K, S, P)) < 0.5, (kk - 0.2) / sr, tt).astype(dtype)
        got, m2t = ts.moveouts_from_travel_times(tt, sr)
        want, want_m2t = mo.moveouts_from_travel_times(tt, sr)
        assert np.array_equal(got, want) and np.array_equal(m2t, want_m2t)
    assert ts.moveouts_from_travel_times(np.zeros((0, 3, 2), np.float32), 50.0)[0].shape == (0, 3, 2)
    with pytest.raises(_lib.BpmfError):  # 1e9 s at 100 Hz does not fit int32 samples
        ts.moveouts_from_travel_times(np.full((2, 1, 1), 1e9, np.float64), 100.0, relative_to_first=False)
    # the cached table: same integers as the reference's property, computed once per (stations, phases) key
    stations, phases = ["A", "B", "C"], ["P", "S"]
    cols = {ph: {sta: rng.uniform(0, 30, 500).astype(np.float32) for sta in stations} for ph in phases}
    table = ts.MoveoutTable(cols, 50.0)
    mv = table.moveouts()
    tt = np.stack([np.stack([cols[ph][sta] for ph in phases], -1) for sta in stations], 1)
    assert np.array_equal(mv, mo.moveouts_from_travel_times(tt, 50.0)[0])
    assert table.moveouts() is mv
    sub = table.moveouts(stations=["C", "A"], phases=["S"])
    assert np.array_equal(sub, mo.moveouts_from_travel_times(tt[:, [2, 0]][:, :, [1]], 50.0)[0])
    dev = table.device_moveouts()
    assert dev.is_cuda and dev.dtype.is_floating_point is False and np.array_equal(dev.cpu().numpy(), mv)


@pytest.mark.gpu
def test_device_resident_table_matches_host_form():
    import torch

    from seismic_bpmf_b200 import template_search as ts

    rng = np.random.default_rng(8)
    for dtype in (np.float32, np.float64):
        tt = rng.uniform(0.0, 60.0, (5000, 7, 2)).astype(dtype)
        mv, m2t = ts.moveouts_from_travel_times(tt, 100.0)
        mv_d, m2t_d = ts.moveouts_from_travel_times_dev(torch.from_numpy(tt).cuda(), 100.0)
        assert np.array_equal(mv_d.cpu().numpy(), mv) and np.array_equal(m2t_d.cpu().numpy(), m2t)
