"""summersph_b200 — B200-native (sm_100a CUDA) drop-in for the per-step hot path of SUMMERSPH.

Public surface mirrors the reference's operator names (SUMMER_SPH.f90 / "SUMMER_SPH - Variable.f90"):
`read_data_from_file`, `read_params_from_file`, `make_save`, `simulate`, plus the engine context.
There is no CPU fallback: creating an engine without the CUDA library or a CUDA device raises.
"""
from ._abi import (SphParams, SphCounts, default_params, MODE_FIXED_H, MODE_VARIABLE_H, FLAG_SOFT_USES_HI, FLAG_SINK_MERGE_SPIN,
                   EVAL_TREE, EVAL_DENSITY, EVAL_GRAVITY, EVAL_SINKS, EVAL_SPH, EVAL_ALL)
from .state import Bodies, Sinks

__all__ = ["SphParams", "SphCounts", "default_params", "MODE_FIXED_H", "MODE_VARIABLE_H", "FLAG_SOFT_USES_HI", "FLAG_SINK_MERGE_SPIN",
           "EVAL_TREE", "EVAL_DENSITY", "EVAL_GRAVITY", "EVAL_SINKS", "EVAL_SPH", "EVAL_ALL", "Bodies", "Sinks"]
