"""Detection traces: the NumPy/SciPy restatement against vectors made with the reference's own
``envelope`` / ``saturated_envelopes`` / ``time_dependent_threshold`` (CPU), and the CUDA library against the same
vectors: bit-exact for the order-statistics step on given envelopes, float32-FFT tolerance for the transforms."""
import os

import numpy as np
import pytest

from oracle import features_numpy as fo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "features_golden.npz")
ENV_RTOL = 2e-5   # |envelope error| / max|envelope|: float32 FFT of another library (cuFFT vs pocketfft)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLD)


def test_oracle_matches_reference_vectors(golden):
    assert np.array_equal(fo.envelope_parallel(golden["traces"]), golden["envelopes"])
    assert np.array_equal(fo.envelope_parallel(golden["traces_even"]), golden["envelopes_even"])
    f, a = fo.saturated_envelopes(golden["traces"])
    assert np.array_equal(f, golden["features"]) and np.array_equal(a, golden["availability"])
    f, a = fo.saturated_envelopes(golden["traces"], anomaly_threshold=1e-7, max_dynamic_range=50.0)
    assert np.array_equal(f, golden["features_b"]) and np.array_equal(a, golden["availability_b"])
    assert list(golden["availability"]) == [2, 1, 2]  # the dead trace (MAD below the anomaly threshold) is dropped
    for tag in "abc":
        window, overlap, cnr = (float(v) for v in golden["thr_%s_params" % tag])  # Python floats, as BPMF passes them
        thr = fo.time_dependent_threshold(golden["thr_%s_x" % tag], int(window), overlap=overlap, CNR_threshold=cnr)
        assert np.array_equal(thr, golden["thr_%s_out" % tag])


@pytest.mark.gpu
def test_envelope_vs_reference_vectors(golden):
    from seismic_bpmf_b200 import features as ft

    for key in ("", "_even"):
        x, want = golden["traces" + key], golden["envelopes" + key]
        got = ft.envelope_parallel(x)
        assert got.shape == want.shape and got.dtype == np.float32
        for row_g, row_w in zip(got.reshape(-1, x.shape[-1]), want.reshape(-1, x.shape[-1])):
            scale = max(float(row_w.max()), 1e-30)
            assert np.abs(row_g - row_w).max() <= ENV_RTOL * scale
    one = ft.envelope(golden["traces"][0, 0])
    assert np.abs(one - golden["envelopes"][0, 0]).max() <= ENV_RTOL * golden["envelopes"][0, 0].max()
    # a long odd record with small prime factors (one hour @ 50 Hz + 75 samples = 3 * 5^2 * 7^4)
    rng = np.random.default_rng(0)
    long = rng.standard_normal((1, 1, 180075)).astype(np.float32)
    got = ft.envelope_parallel(long)
    want = fo.envelope_parallel(long)
    assert np.abs(got - want).max() <= ENV_RTOL * want.max()
    # lengths with a large prime factor (6001 = 17 * 353, 180001 prime) make cuFFT fall back to Bluestein's
    # algorithm, whose float32 chirps are less accurate; BPMF's records never have such lengths
    for n in (6001, 180001):
        odd = rng.standard_normal((2, 1, n)).astype(np.float32)
        got, want = ft.envelope_parallel(odd), fo.envelope_parallel(odd)
        assert np.abs(got - want).max() <= 1e-3 * want.max()


@pytest.mark.gpu
def test_standardisation_bit_exact_on_given_envelopes(golden):
    from seismic_bpmf_b200 import features as ft

    env = golden["envelopes"]
    f, a = ft.standardize_envelopes(env)
    assert np.array_equal(f, golden["features"]) and np.array_equal(a, golden["availability"])
    f, a = ft.standardize_envelopes(env, anomaly_threshold=1e-7, max_dynamic_range=50.0)
    assert np.array_equal(f, golden["features_b"]) and np.array_equal(a, golden["availability_b"])
    # random envelopes with gaps, even and odd counts of present samples
    rng = np.random.default_rng(1)
    e = np.abs(rng.standard_normal((4, 3, 5000))).astype(np.float32)
    e[0, 0, :777] = 0.0
    e[1, 2, 10:2000] = 0.0
    e[2, 1, :2600] = 0.0
    e[3, 0, 4000] = 1e9
    f, a = ft.standardize_envelopes(e, max_dynamic_range=1e3)
    fw, aw = fo.standardize_envelopes(e, max_dynamic_range=1e3)
    assert np.array_equal(f, fw) and np.array_equal(a, aw)


@pytest.mark.gpu
def test_saturated_envelopes_pipeline(golden):
    from seismic_bpmf_b200 import features as ft

    f, a = ft.saturated_envelopes(golden["traces"])
    want = golden["features"]
    assert np.array_equal(a, golden["availability"])
    assert np.array_equal(f == 0, want == 0)  # missing samples and dropped traces are exactly zero
    # standardised values differ only through the float32 FFT: error relative to the dynamic range of each trace
    for fg, fw in zip(f.reshape(-1, f.shape[-1]), want.reshape(-1, want.shape[-1])):
        assert np.abs(fg - fw).max() <= 1e-3 * max(1.0, float(np.abs(fw).max()))
    f, a = ft.saturated_envelopes(golden["traces"], anomaly_threshold=1e-7, max_dynamic_range=50.0)
    assert np.array_equal(a, golden["availability_b"]) and f.max() <= 50.0


@pytest.mark.gpu
def test_backprojection_threshold_vs_reference_vectors(golden):
    from seismic_bpmf_b200 import features as ft

    for tag in "abc":
        window, overlap, cnr = (float(v) for v in golden["thr_%s_params" % tag])  # Python floats, as BPMF passes them
        thr = ft.time_dependent_threshold(golden["thr_%s_x" % tag], int(window), overlap=overlap, CNR_threshold=cnr)
        want = golden["thr_%s_out" % tag]
        assert thr.dtype == np.float64 and thr.shape == want.shape
        assert np.abs(thr - want).max() <= 1e-9 * np.abs(want).max()


@pytest.mark.gpu
def test_saturated_envelopes_device_form_matches_host_form():
    import torch

    from seismic_bpmf_b200 import features as ft

    rng = np.random.default_rng(4)
    x = (rng.standard_normal((3, 2, 60750)) * 1e-6).astype(np.float32)
    x[1, 0, 1000:9000] = 0.0
    want, avail = ft.saturated_envelopes(x)
    got, avail_d = ft.saturated_envelopes_dev(torch.from_numpy(x).cuda())
    assert np.array_equal(got.cpu().numpy(), want) and np.array_equal(avail_d.cpu().numpy(), avail)
