"""seismic_bpmf_b200 -- B200 (sm_100a) native hot path for the Seismic_BPMF workflow.

Drop-in surfaces (same call signatures as the reference uses):

* :func:`seismic_bpmf_b200.fast_matched_filter.matched_filter`
  (``fmf.matched_filter`` at BPMF/similarity_search.py:526-533, BPMF/dataset.py:4818-4827)
* :func:`seismic_bpmf_b200.beampower.beampower.beamform`
  (``bp.beampower.beamform`` at BPMF/template_search.py:549-569)
* :mod:`seismic_bpmf_b200.clib` (mirror of BPMF/clib.py; the shared object itself can also be
  installed as ``BPMF/lib/libc.so``)

Everything computes on the GPU through ``lib/libbpmf_b200.so`` (hand-written CUDA, C-ABI in
``include/bpmf_b200.h``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["fast_matched_filter", "beampower", "clib", "device", "sharded"]
__version__ = "0.1.0"
