"""Drop-in for the ``beampower`` package surface BPMF uses (``beampower.beampower.beamform``)."""
from . import beampower  # noqa: F401
from .beampower import beamform, prestack_traces  # noqa: F401
