"""Drop-in for ``fast_matched_filter.matched_filter`` as BPMF calls it.

Reference call sites: BPMF/similarity_search.py:526-533 (positional ``templates, moveouts, weights,
data, step`` + ``arch=``) and BPMF/dataset.py:4818-4827 (additionally ``network_sum=False,
check_zeros=False``).  Semantics: SURVEY.md Appendix A.1.  The arithmetic runs on the current CUDA
device through ``bpmf_matched_filter`` (include/bpmf_b200.h); ``arch`` is accepted for signature
compatibility -- ``"cpu"`` is NOT a CPU fallback here, every value computes on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def matched_filter(templates, moveouts, weights, data, step, arch="gpu", check_zeros="first",
                   normalize="short", network_sum=True, precision="auto", out=None):
    """Sliding-window normalised cross-correlation, weighted and summed over stations/channels.

    Parameters
    ----------
    templates : (n_templates, n_stations, n_components, n_samples_template) float32
    moveouts : (n_templates, n_stations, n_components) or (n_templates, n_stations) int
        Window start offsets in samples; may be negative (BPMF/dataset.py:3452-3475).
    weights : (n_templates, n_stations, n_components) float32; zero-weight channels are skipped.
    data : (n_stations, n_components, n_samples_data) float32
    step : int, samples between consecutive correlations.
    arch : accepted for compatibility ('cpu', 'gpu', 'precise'); always runs on the GPU.
    check_zeros : 'first' | 'all' | False -- report templates with exactly-zero CCs
        (message text as in tutorial 8_template_matching.ipynb:373).
    normalize : 'short' (default, no window-mean removal) | 'full'.
    network_sum : True -> (n_templates, n_corr); False -> (n_templates, n_corr, n_stations,
        n_components) un-weighted per-channel CCs (BPMF/dataset.py:4828-4830).
    precision : 'auto' | 'fp32' | 'fp16' | 'fp16x3' (extension).
    out : optional pre-allocated float32 output (e.g. pinned memory).

    Returns
    -------
    cc_sums : writable float32 ndarray owned by the caller (BPMF mutates it in place,
        BPMF/similarity_search.py:538,540).  Never contains NaN from zero-energy windows.
    """
    lib = _lib.load()
    _lib.require_device()
    templates = _as(templates, np.float32)
    weights = _as(weights, np.float32)
    data = _as(data, np.float32)
    if templates.ndim != 4 or data.ndim != 3 or weights.ndim != 3:
        raise ValueError("expected templates (T,S,C,L), weights (T,S,C), data (S,C,N)")
    T, S, Cc, L = templates.shape
    if data.shape[:2] != (S, Cc) or weights.shape != (T, S, Cc):
        raise ValueError("inconsistent shapes: templates %s weights %s data %s"
                         % (templates.shape, weights.shape, data.shape))
    N = data.shape[-1]
    moveouts = np.asarray(moveouts)
    if moveouts.ndim == 2:  # broadcast over components
        moveouts = np.repeat(moveouts[:, :, None], Cc, axis=2)
    if moveouts.shape != (T, S, Cc):
        raise ValueError("moveouts shape %s does not match templates" % (moveouts.shape,))
    moveouts = _as(moveouts, np.int32)
    step = int(step)
    if step < 1:
        raise ValueError("step must be >= 1")
    if normalize not in ("short", "full"):
        raise ValueError("normalize must be 'short' or 'full'")
    if N < L:
        raise ValueError("data shorter than templates")
    n_corr = int(lib.bpmf_mf_n_corr(N, L, step))
    shape = (T, n_corr) if network_sum else (T, n_corr, S, Cc)
    if out is None:
        out = np.zeros(shape, dtype=np.float32)
    else:
        if out.shape != shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape %s" % (shape,))
    prec = _lib.PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    rc = lib.bpmf_matched_filter(_ptr(templates, C.c_float), _ptr(moveouts, C.c_int),
                                 _ptr(weights, C.c_float), _ptr(data, C.c_float), step, L, N, T, S,
                                 Cc, 1 if normalize == "full" else 0, 1 if network_sum else 0, prec,
                                 _ptr(out, C.c_float))
    _lib.check(rc, "matched_filter")
    if check_zeros and network_sum:
        _report_zeros(out, moveouts, weights, step, check_zeros)
    return out


class hold_traces:
    """``with hold_traces(data) as data: ...`` -- every ``matched_filter`` call made on ``data`` inside the block
    reuses ONE upload + staging of the traces (BPMF loops template chunks over one day of data, one call per chunk,
    BPMF/similarity_search.py:773-790; without the hold each call moves the 1.7-3.3 GB day again).  The block
    promises that ``data`` is not modified meanwhile.  ``data`` must be C-contiguous float32 -- the array returned
    by ``__enter__`` is (a copy is made otherwise) and is the one to pass to ``matched_filter``."""

    def __init__(self, data):
        self.data = _as(data, np.float32)

    def __enter__(self):
        lib = _lib.load()
        _lib.require_device()
        S, Cc, N = self.data.shape
        _lib.check(lib.bpmf_mf_hold_traces(_ptr(self.data, C.c_float), N, S, Cc), "hold_traces")
        return self.data

    def __exit__(self, *exc):
        _lib.load().bpmf_mf_release_traces()
        return False


def _report_zeros(cc_sums, moveouts, weights, step, mode):
    """Upstream prints how many CCs are exactly zero per template (8_template_matching.ipynb:373)."""
    n_templates = cc_sums.shape[0] if mode == "all" else min(1, cc_sums.shape[0])
    for t in range(n_templates):
        used = weights[t] != 0
        if not used.any():
            continue
        mv = moveouts[t][used]
        lo = int(-(int(min(mv.min(), 0)) // step))
        hi = cc_sums.shape[1] - int(max(int(mv.max()), 0) // step) - 1
        if hi <= lo:
            continue
        n_zeros = int(np.count_nonzero(cc_sums[t, lo:hi] == 0.0))
        if n_zeros > 10:
            print("{} correlation computations were skipped on the {:d}-th template. Can be caused by "
                  "zeros in data, or too low amplitudes (try to increase the gain).".format(n_zeros, t))
