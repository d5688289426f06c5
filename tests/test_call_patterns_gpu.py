"""GPU: the reference's own call patterns, line for line, with our modules bound under the names BPMF
uses (`fmf`, `bp`, `clib`) -- what a maintainer gets after the import swap of INTEGRATION.md."""
import numpy as np
import pytest

from oracle import cpu, libc_numpy
from seismic_bpmf_b200 import beampower as bp
from seismic_bpmf_b200 import clib
from seismic_bpmf_b200 import fast_matched_filter as fmf
from tests import _synth

pytestmark = pytest.mark.gpu


def test_compute_cc_time_series_network_chunks():
    """BPMF/similarity_search.py:505-546: station chunks, results summed, NaN scrub, zero-weight templates
    skipped by the caller."""
    waveforms_arr, moveouts_arr, weights_arr, data_arr, _ = _synth.mf_case(12, 7, 3, 200, 40000, seed=51)
    weights_arr[5] = 0.0  # a template with zero weights only
    step, device, n_network_chunks = 1, "gpu", 3
    select_tts = np.arange(12, dtype=np.int32)
    invalid_weights = np.sum(weights_arr, axis=(1, 2)) == 0
    select_tts = select_tts[~invalid_weights]
    nt, ns, nc, Nsamp = waveforms_arr[select_tts, ...].shape
    CC_SUMS = []
    L = ns // n_network_chunks + 1
    for i in range(n_network_chunks):
        id1, id2 = i * L, min((i + 1) * L, ns)
        cc_sums = fmf.matched_filter(
            waveforms_arr[select_tts, id1:id2, :, :],
            moveouts_arr[select_tts, id1:id2, :],
            weights_arr[~invalid_weights, id1:id2, :],
            data_arr[id1:id2, ...],
            step,
            arch=device,
        )
        CC_SUMS.append(cc_sums)
    cc_sums = CC_SUMS[0]
    for i in range(1, n_network_chunks):
        cc_sums += CC_SUMS[i]  # in place: the returned arrays are writable
    cc_sums[np.isnan(cc_sums)] = 0.0
    want = cpu.matched_filter(waveforms_arr[select_tts], moveouts_arr[select_tts], weights_arr[select_tts], data_arr, 1)
    live = want != 0  # chunked bounds are per chunk (looser); compare where the full call is defined
    assert np.abs(cc_sums - want)[live].max() < 1e-4
    assert np.isfinite(cc_sums).all()


def test_find_detections_threshold_and_selection():
    """BPMF/similarity_search.py:600-666 on one CC series: rms threshold via clib, ceiling, declustering."""
    tp, mv, w, d, planted = _synth.mf_case(3, 4, 3, 100, 60000, seed=52)
    cc = fmf.matched_filter(tp, mv, w, d, 1, arch="gpu", check_zeros=False)
    wn = np.random.default_rng(2).standard_normal(500).astype(np.float32)
    for t, i0 in planted.items():
        cc_t = cc[t]
        threshold = clib.time_dependent_threshold(cc_t, 12000, 8.0, overlap=0.25, threshold_type="rms",
                                                  white_noise=wn, num_threads=1)
        want, _ = libc_numpy.time_dependent_threshold(cc_t, wn, 8.0, 6000, int(0.75 * 12000))
        assert np.abs(threshold - want).max() < 5e-6
        threshold = np.minimum(0.8 * w[t].sum(), threshold)  # similarity_search.py:629
        sel = clib.select_cc_indexes(cc_t, threshold, 500)
        assert sel[i0] and sel.sum() >= 1
        assert np.array_equal(sel, libc_numpy.select_cc_indexes(cc_t, threshold, 500).astype(bool))


def test_intertemplate_cc_pattern():
    """BPMF/dataset.py:4776-4830: one template's waveforms as data, all templates trimmed by max_lag as
    templates, zero moveouts, network_sum=False, then max over lags and a weighted sum."""
    rng = np.random.default_rng(53)
    n_templates, S, C, L, max_lag = 6, 4, 3, 128, 10
    waveforms_arr = rng.standard_normal((n_templates, S, C, L)).astype(np.float32)
    waveforms_arr[3] = waveforms_arr[0] + 0.1 * rng.standard_normal((S, C, L)).astype(np.float32)
    template_arr = waveforms_arr[..., max_lag:-max_lag]
    moveouts_arr = np.zeros(waveforms_arr.shape[:-1], dtype=np.int32)
    weights = np.ones(template_arr.shape[:-1], dtype=np.float32)
    weights /= np.sum(weights, axis=(1, 2), keepdims=True)
    keep = np.sum(weights != 0.0, axis=(1, 2)) > 0
    t = 0
    cc = fmf.matched_filter(template_arr[keep, ...], moveouts_arr[keep, ...], weights[keep, ...],
                            waveforms_arr[t, ...], 1, arch="gpu", network_sum=False, check_zeros=False)
    assert cc.shape == (n_templates, 2 * max_lag + 1, S, C)
    intertp_cc = np.sum(weights[keep, ...] * np.max(cc, axis=1), axis=(-1, -2))
    want = cpu.matched_filter(template_arr, moveouts_arr, weights, waveforms_arr[t], 1, network_sum=False)
    assert np.abs(cc - want).max() < 2e-6
    assert abs(intertp_cc[0] - 1.0) < 1e-5 and intertp_cc[3] > 0.9 and np.all(intertp_cc[[1, 2, 4, 5]] < 0.5)


def test_backproject_pattern():
    """BPMF/template_search.py:549-569 with the int64 relative moveouts of get_travel_times_array (:207-219)."""
    waveform_features, moveouts, weights_phases, weights_sources = _synth.beam_case(500, 8, 2, 2, 30000, seed=54,
                                                                                   sparse_sources=True, coherent=True)
    assert moveouts.dtype == np.int64
    for out_of_bounds in ("strict", "flexible"):
        maxbeam, maxbeam_sources = bp.beampower.beamform(
            waveform_features, moveouts, weights_phases, weights_sources,
            device="gpu", out_of_bounds=out_of_bounds, num_threads=None, reduce="max")
        beam = bp.beampower.beamform(
            waveform_features, moveouts, weights_phases, weights_sources,
            device="gpu", out_of_bounds=out_of_bounds, num_threads=None, reduce="none")
        mb_w, _ = cpu.beamform(waveform_features, moveouts, weights_phases, weights_sources, out_of_bounds=out_of_bounds)
        assert np.abs(maxbeam - mb_w).max() < 1e-5
        src_idx, time_idx = np.unravel_index(beam.argmax(), beam.shape)  # dataset.py:2194-2196
        assert maxbeam_sources[time_idx] == src_idx and maxbeam[time_idx] == beam[src_idx, time_idx]


def test_grid_thinning_pattern():
    """tutorial 4_travel_times.ipynb:1050: clib.find_similar_sources on relative travel times (seconds)."""
    rng = np.random.default_rng(55)
    K, S = 900, 8
    lon, lat = rng.uniform(30, 31, K), rng.uniform(40, 41, K)
    sta = rng.uniform(0, 1, (S, 2))
    tts = np.linalg.norm(np.stack([lon - 30, lat - 40], 1)[:, None] - sta[None], axis=-1) * 15.0
    tts -= tts.min(axis=1, keepdims=True)
    cell_lon, cell_lat = np.linspace(30, 31.001, 4), np.linspace(40, 41.001, 4)
    red = clib.find_similar_sources(tts, lon, lat, cell_lon, cell_lat, 0.1, num_threads=4, num_stations_for_diff=5)
    want = libc_numpy.find_similar_moveouts(tts, lon, lat, cell_lon, cell_lat, 0.1, 5, 2).astype(bool)
    assert red.dtype == bool and np.array_equal(red, want) and 0 < red.sum() < K


def test_template_chunk_loop_reuses_held_traces():
    """run_matched_filter_search's chunk loop (BPMF/similarity_search.py:773-790): several calls on the same day of
    traces.  Inside hold_traces the day is uploaded and staged once; results are identical to the plain calls, a
    change of the array after the block is seen again, and the launch counter shows the staging kernels were skipped."""
    from seismic_bpmf_b200 import _lib

    lib = _lib.load()
    tp, mv, w, d, _ = _synth.mf_case(48, 4, 3, 400, 60000, seed=21)
    plain = [fmf.matched_filter(tp[a:a + 16], mv[a:a + 16], w[a:a + 16], d, 1, check_zeros=False) for a in (0, 16, 32)]
    n0 = lib.bpmf_launch_count()
    fmf.matched_filter(tp[:16], mv[:16], w[:16], d, 1, check_zeros=False)
    per_call = lib.bpmf_launch_count() - n0
    with fmf.hold_traces(d) as dd:
        n0 = lib.bpmf_launch_count()
        held = [fmf.matched_filter(tp[a:a + 16], mv[a:a + 16], w[a:a + 16], dd, 1, check_zeros=False) for a in (0, 16, 32)]
        n_held = lib.bpmf_launch_count() - n0
    for a, b in zip(plain, held):  # same arithmetic; the order of the fp32 reductions differs from run to run
        assert np.abs(a - b).max() < 1e-6 and np.array_equal(a == 0, b == 0)
    assert n_held < 3 * per_call  # the trace-side kernels ran once, not three times
    d2 = d.copy()
    d2[0, 0, 1000:2000] *= 3.0
    after = fmf.matched_filter(tp[:16], mv[:16], w[:16], d2, 1, check_zeros=False)
    assert np.abs(after - cpu.matched_filter(tp[:16], mv[:16], w[:16], d2, 1)).max() < 1e-4
