"""Device-resident entry points: inputs and outputs are torch CUDA tensors, work is enqueued on
torch's current stream through the ``*_dev`` functions of the C-ABI (include/bpmf_b200.h), with no
host synchronisation.  PyTorch is plumbing only here (device memory + streams).
"""
import torch

from . import _lib


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s" % (name, dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def mf_n_corr(n_samples_data, n_samples_template, step):
    return int(_lib.load().bpmf_mf_n_corr(n_samples_data, n_samples_template, step))


def mf_cc_pitch(n_corr):
    """Row pitch (elements) used for device CC slabs: rows start on 256-byte boundaries."""
    return (n_corr + 63) // 64 * 64


class MatchedFilterPlan:
    """Pre-allocated workspaces + output slab for device-resident matched-filter calls.

    The trace-side staging (``prepare``) is separate from the per-template-batch run (``run_batch``) so that
    many template batches can be run against the same traces -- what BPMF does when it loops template chunks
    over one day of data (BPMF/similarity_search.py:783-790).  ``run`` does both.
    """

    def __init__(self, n_templates, n_stations, n_components, n_samples_template, n_samples_data,
                 step=1, normalize="short", network_sum=True, precision="auto", device=None, batch=None):
        self.lib = _lib.load()
        _lib.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.T, self.S, self.C = int(n_templates), int(n_stations), int(n_components)
        self.L, self.N, self.step = int(n_samples_template), int(n_samples_data), int(step)
        self.batch = self.T if batch is None else max(1, min(int(batch), self.T))
        self.normalize = 1 if normalize == "full" else 0
        self.network_sum = 1 if network_sum else 0
        self.precision = _lib.PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        self.n_corr = mf_n_corr(self.N, self.L, self.step)
        self.resolved_precision = int(self.lib.bpmf_mf_resolve_precision(
            self.L, self.step, self.normalize, self.network_sum, self.precision))
        if self.resolved_precision <= 0:
            raise _lib.BpmfError("precision %r is not available for this shape" % (precision,))
        self.last_precision = self.resolved_precision
        tb = int(self.lib.bpmf_mf_trace_workspace_bytes(self.L, self.N, self.S, self.C, self.step, self.normalize,
                                                        self.network_sum, self.precision))
        pb = int(self.lib.bpmf_mf_template_workspace_bytes(self.L, self.N, self.batch, self.S, self.C, self.step,
                                                           self.normalize, self.network_sum, self.precision))
        self.trace_ws = torch.empty(max(tb, 256), dtype=torch.uint8, device=self.device)
        self.template_ws = torch.empty(max(pb, 256), dtype=torch.uint8, device=self.device)
        if self.network_sum:
            self.pitch = mf_cc_pitch(self.n_corr)
            self.cc = torch.empty((self.T, self.pitch), dtype=torch.float32, device=self.device)
        else:
            self.pitch = 0
            self.cc = torch.empty((self.T, self.n_corr, self.S, self.C), dtype=torch.float32,
                                  device=self.device)

    def prepare(self, data):
        """Stage the traces (fp16 copies / window energies) for every later ``run_batch``."""
        _chk(data, torch.float32, "data")
        assert tuple(data.shape) == (self.S, self.C, self.N)
        rc = self.lib.bpmf_mf_prepare_traces_dev(data.data_ptr(), self.step, self.L, self.N, self.S, self.C,
                                                 self.normalize, self.network_sum, self.precision,
                                                 self.trace_ws.data_ptr(), self.trace_ws.numel(), _stream_ptr())
        _lib.check(rc, "mf_prepare_traces_dev")

    def run_batch(self, templates, moveouts, weights, data, t0=0, n=None, cc_row0=None):
        """Templates [t0, t0+n) of the inputs into rows [cc_row0, cc_row0+n) of the CC slab (default: the same
        rows [t0, t0+n); a caller that streams more templates than the slab holds passes its own slot)."""
        n = min(self.batch, templates.shape[0] - t0) if n is None else int(n)
        r0 = t0 if cc_row0 is None else int(cc_row0)
        assert 0 < n <= self.batch and t0 + n <= templates.shape[0] and r0 + n <= self.T
        tp, mv, w = templates[t0:t0 + n], moveouts[t0:t0 + n], weights[t0:t0 + n]
        rc = self.lib.bpmf_mf_run_dev(
            tp.data_ptr(), mv.data_ptr(), w.data_ptr(), data.data_ptr(), self.step, self.L, self.N, n, self.S,
            self.C, self.normalize, self.network_sum, self.precision, self.trace_ws.data_ptr(),
            self.trace_ws.numel(), self.template_ws.data_ptr(), self.template_ws.numel(),
            self.cc[r0:r0 + n].data_ptr(), self.pitch, _stream_ptr())
        _lib.check(rc, "mf_run_dev")
        self.last_precision = int(self.lib.bpmf_mf_last_precision())

    def run(self, templates, moveouts, weights, data, prepared=False):
        """Enqueue one matched-filter pass; returns the (T, n_corr[, S, C]) CC view on the device."""
        _chk(templates, torch.float32, "templates")
        _chk(moveouts, torch.int32, "moveouts")
        _chk(weights, torch.float32, "weights")
        _chk(data, torch.float32, "data")
        assert tuple(templates.shape) == (self.T, self.S, self.C, self.L)
        assert tuple(data.shape) == (self.S, self.C, self.N)
        assert tuple(moveouts.shape) == (self.T, self.S, self.C) == tuple(weights.shape)
        if not prepared:
            self.prepare(data)
        for t0 in range(0, self.T, self.batch):
            self.run_batch(templates, moveouts, weights, data, t0)
        return self.cc[:, : self.n_corr] if self.network_sum else self.cc


def matched_filter_dev(templates, moveouts, weights, data, step=1, normalize="short", network_sum=True,
                       precision="auto"):
    T, S, C, L = templates.shape
    plan = MatchedFilterPlan(T, S, C, L, data.shape[-1], step, normalize, network_sum, precision,
                             device=templates.device)
    return plan.run(templates, moveouts, weights, data)


class BeamformPlan:
    def __init__(self, n_sources, n_stations, n_components, n_phases, n_samples, device=None):
        self.lib = _lib.load()
        _lib.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.K, self.S, self.C, self.P, self.N = map(int, (n_sources, n_stations, n_components, n_phases, n_samples))
        ws = int(self.lib.bpmf_beam_workspace_bytes(self.K, self.S, self.C, self.P, self.N))
        self.workspace = torch.empty(max(ws, 256), dtype=torch.uint8, device=self.device)
        self.maxbeam = torch.empty(self.N, dtype=torch.float32, device=self.device)
        self.sources = torch.empty(self.N, dtype=torch.int32, device=self.device)

    def _check(self, features, moveouts, weights_phases, weights_sources):
        _chk(features, torch.float32, "features")
        _chk(moveouts, torch.int32, "moveouts")
        _chk(weights_phases, torch.float32, "weights_phases")
        _chk(weights_sources, torch.float32, "weights_sources")
        assert tuple(features.shape) == (self.S, self.C, self.N)
        assert tuple(moveouts.shape) == (self.K, self.S, self.P)
        assert tuple(weights_phases.shape) == (self.S, self.C, self.P)
        assert tuple(weights_sources.shape) == (self.K, self.S)

    def run_max(self, features, moveouts, weights_phases, weights_sources, out_of_bounds="strict",
                source_offset=0):
        self._check(features, moveouts, weights_phases, weights_sources)
        rc = self.lib.bpmf_beamform_max_dev(
            features.data_ptr(), moveouts.data_ptr(), weights_phases.data_ptr(),
            weights_sources.data_ptr(), self.K, self.S, self.C, self.P, self.N,
            1 if out_of_bounds == "flexible" else 0, int(source_offset), self.workspace.data_ptr(),
            self.workspace.numel(), self.maxbeam.data_ptr(), self.sources.data_ptr(), _stream_ptr())
        _lib.check(rc, "beamform_max_dev")
        return self.maxbeam, self.sources

    def run_full(self, features, moveouts, weights_phases, weights_sources, out_of_bounds="strict"):
        self._check(features, moveouts, weights_phases, weights_sources)
        beam = torch.empty((self.K, self.N), dtype=torch.float32, device=self.device)
        rc = self.lib.bpmf_beamform_full_dev(
            features.data_ptr(), moveouts.data_ptr(), weights_phases.data_ptr(),
            weights_sources.data_ptr(), self.K, self.S, self.C, self.P, self.N,
            1 if out_of_bounds == "flexible" else 0, self.workspace.data_ptr(), self.workspace.numel(),
            beam.data_ptr(), _stream_ptr())
        _lib.check(rc, "beamform_full_dev")
        return beam


def beam_merge_(max_a, arg_a, max_b, arg_b):
    """In place a = better(a, b) with ties to the lower source index."""
    rc = _lib.load().bpmf_beam_merge_dev(max_a.data_ptr(), arg_a.data_ptr(), max_b.data_ptr(),
                                         arg_b.data_ptr(), max_a.numel(), _stream_ptr())
    _lib.check(rc, "beam_merge_dev")
    return max_a, arg_a


def time_dependent_threshold_dev(series, gaussian_sample, num_dev, half_window_samp, shift_samp,
                                 n_valid=None):
    """Batched rms threshold on a (B, pitch) device slab (e.g. the CC slab of MatchedFilterPlan).

    ``series`` is modified in place (zeros are replaced, as libc.c:606-612 does).  Returns the
    (B, n) threshold tensor.
    """
    lib = _lib.load()
    assert series.is_cuda and series.dtype == torch.float32 and series.dim() == 2
    B = series.shape[0]
    n = int(series.shape[1] if n_valid is None else n_valid)
    pitch = series.stride(0)
    assert series.stride(1) == 1
    g = _chk(gaussian_sample, torch.float32, "gaussian_sample")
    assert g.numel() == 500
    ws = int(lib.bpmf_threshold_workspace_bytes(B, n, half_window_samp, shift_samp))
    workspace = torch.empty(max(ws, 256), dtype=torch.uint8, device=series.device)
    thr = torch.empty((B, n), dtype=torch.float32, device=series.device)
    rc = lib.bpmf_time_dependent_threshold_dev(series.data_ptr(), pitch, B, g.data_ptr(), float(num_dev),
                                               n, int(half_window_samp), int(shift_samp),
                                               workspace.data_ptr(), workspace.numel(), thr.data_ptr(),
                                               n, _stream_ptr())
    _lib.check(rc, "time_dependent_threshold_dev")
    return thr


def detect_dev(series, gaussian_sample, num_dev, half_window_samp, shift_samp, search_win=0, mode="above",
               n_valid=None, capacity=1 << 20):
    """Candidate detections from a (B, pitch) device slab of CC series without downloading it.

    mode="above": every sample with cc > rms threshold (what BPMF/similarity_search.py:232-233 starts from);
    mode="select": additionally the declustering of ``clib.select_cc_indexes`` with ``search_win``.
    Returns numpy arrays (series, index, cc, threshold) sorted by (series, index).  Raises if more than
    ``capacity`` candidates were found.
    """
    import numpy as np

    lib = _lib.load()
    assert series.is_cuda and series.dtype == torch.float32 and series.dim() == 2 and series.stride(1) == 1
    B = series.shape[0]
    n = int(series.shape[1] if n_valid is None else n_valid)
    g = _chk(gaussian_sample, torch.float32, "gaussian_sample")
    assert g.numel() == 500
    dev_ = series.device
    ws = int(lib.bpmf_threshold_workspace_bytes(B, n, half_window_samp, shift_samp))
    workspace = torch.empty(max(ws, 256), dtype=torch.uint8, device=dev_)
    count = torch.zeros(1, dtype=torch.int32, device=dev_)
    o_series = torch.empty(capacity, dtype=torch.int32, device=dev_)
    o_index = torch.empty(capacity, dtype=torch.int32, device=dev_)
    o_value = torch.empty(capacity, dtype=torch.float32, device=dev_)
    o_thr = torch.empty(capacity, dtype=torch.float32, device=dev_)
    rc = lib.bpmf_detect_dev(series.data_ptr(), series.stride(0), B, g.data_ptr(), float(num_dev), n,
                             int(half_window_samp), int(shift_samp), int(search_win), 1 if mode == "select" else 0,
                             workspace.data_ptr(), workspace.numel(), count.data_ptr(), o_series.data_ptr(),
                             o_index.data_ptr(), o_value.data_ptr(), o_thr.data_ptr(), capacity, _stream_ptr())
    _lib.check(rc, "detect_dev")
    k = int(count.item()) & 0xFFFFFFFF
    if k > capacity:
        raise _lib.BpmfError("detect_dev: %d candidates exceed capacity %d" % (k, capacity))
    s_, i_ = o_series[:k].cpu().numpy(), o_index[:k].cpu().numpy()
    v_, t_ = o_value[:k].cpu().numpy(), o_thr[:k].cpu().numpy()
    order = np.lexsort((i_, s_))
    return s_[order], i_[order], v_[order], t_[order]


class DetectBatches:
    """Threshold + declustering + compaction of many CC batches without a host synchronisation in between:
    every batch gets its own counter and output region, the candidates are read back once (``collect``).

    What replaces, for a template-sharded day, the per-template ``time_dependent_threshold`` /
    ``select_cc_indexes`` calls of BPMF/similarity_search.py:548-666 when the CC slab stays on the device."""

    def __init__(self, n_batches, n_series_per_batch, n, half_window_samp, shift_samp, capacity=1 << 18, device=None):
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_batches, self.B, self.n = int(n_batches), int(n_series_per_batch), int(n)
        self.half, self.shift, self.capacity = int(half_window_samp), int(shift_samp), int(capacity)
        ws = int(self.lib.bpmf_threshold_workspace_bytes(self.B, self.n, self.half, self.shift))
        self.workspace = torch.empty(max(ws, 256), dtype=torch.uint8, device=self.device)
        self.count = torch.zeros(self.n_batches, dtype=torch.int32, device=self.device)
        self.series = torch.empty((self.n_batches, self.capacity), dtype=torch.int32, device=self.device)
        self.index = torch.empty((self.n_batches, self.capacity), dtype=torch.int32, device=self.device)
        self.value = torch.empty((self.n_batches, self.capacity), dtype=torch.float32, device=self.device)
        self.thr = torch.empty((self.n_batches, self.capacity), dtype=torch.float32, device=self.device)

    def run(self, k, series, gaussian_sample, num_dev, search_win=0, mode="select", thr_cap=None, max_kurtosis=None,
            kurtosis_out=None):
        """Enqueue the detection of batch ``k`` on ``series`` (B' <= B rows of a (.., pitch) device slab).
        ``thr_cap`` (B' floats on the device): threshold = min(cap, threshold) -- ``max_CC_threshold * sum(weights)``
        of similarity_search.py:629; ``max_kurtosis``: rows whose excess kurtosis exceeds it detect nothing
        (similarity_search.py:633-642); ``kurtosis_out`` (B' floats) receives the kurtosis values."""
        assert series.is_cuda and series.dtype == torch.float32 and series.dim() == 2 and series.stride(1) == 1
        assert series.shape[0] <= self.B and 0 <= k < self.n_batches
        g = _chk(gaussian_sample, torch.float32, "gaussian_sample")
        if thr_cap is not None:
            thr_cap = _chk(thr_cap, torch.float32, "thr_cap")
            assert thr_cap.numel() == series.shape[0]
        rc = self.lib.bpmf_detect_gated_dev(
            series.data_ptr(), series.stride(0), series.shape[0], g.data_ptr(), float(num_dev), self.n, self.half,
            self.shift, int(search_win), 1 if mode == "select" else 0, thr_cap.data_ptr() if thr_cap is not None else None,
            0 if max_kurtosis is None else 1, 0.0 if max_kurtosis is None else float(max_kurtosis),
            kurtosis_out.data_ptr() if kurtosis_out is not None else None, self.workspace.data_ptr(),
            self.workspace.numel(), self.count[k:].data_ptr(), self.series[k].data_ptr(), self.index[k].data_ptr(),
            self.value[k].data_ptr(), self.thr[k].data_ptr(), self.capacity, _stream_ptr())
        _lib.check(rc, "detect_gated_dev")

    def packed(self):
        """Device tensors (count (n_batches,), series, index, value, thr (n_batches, capacity)); no sync."""
        return self.count, self.series, self.index, self.value, self.thr

    def collect(self, series_offsets=None):
        """Synchronise and return numpy (series, index, value, threshold) of every batch, series numbered
        ``series_offsets[k] + local row`` (default k * B), sorted by (series, index)."""
        import numpy as np

        counts = self.count.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        if (counts > self.capacity).any():
            raise _lib.BpmfError("DetectBatches: %d candidates exceed capacity %d" % (int(counts.max()), self.capacity))
        out = [[], [], [], []]
        for k, c in enumerate(counts):
            c = int(c)
            if c == 0:
                continue
            off = k * self.B if series_offsets is None else int(series_offsets[k])
            out[0].append(self.series[k, :c].cpu().numpy().astype(np.int64) + off)
            out[1].append(self.index[k, :c].cpu().numpy())
            out[2].append(self.value[k, :c].cpu().numpy())
            out[3].append(self.thr[k, :c].cpu().numpy())
        if not out[0]:
            return (np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(0, np.float32))
        s_, i_, v_, t_ = (np.concatenate(x) for x in out)
        order = np.lexsort((i_, s_))
        return s_[order], i_[order], v_[order], t_[order]
