"""mps_born_machines_b200 — B200-native engine for the MPS Born machine hot path
(two-site DMRG-style training sweep + batched ancestral sampler of eigen-carmona/mps_born_machines)."""
from ._capi import BornEngineError, ZeroPsiError, EXPORTS, LIB_PATH  # noqa: F401
from .engine import BornEngine, sweep_schedule, uniforms_reference  # noqa: F401

__all__ = ['BornEngine', 'BornEngineError', 'ZeroPsiError', 'sweep_schedule', 'uniforms_reference']
