"""lowrankintegrators.jl_b200 — B200-native engine for the per-step DLRA hot path of FHoltorf/LowRankIntegrators.jl.
Host-side mirror of the reference API (api.py) over the C ABI of libdlra.so (include/dlra.h)."""
from . import _lib
from .api import (DLRIntegrator, DLRSolution, DualLieTrotter, GreedyIntegrator, MatrixDataProblem, MatrixDEProblem,
                  MatrixHybridProblem, normal_component,
                  PrimalLieTrotter, ProjectorSplitting, RankAdaptiveUnconventionalAlgorithm, Strang, SubStepper,
                  SVDLikeRepresentation, TwoFactorRepresentation, UnconventionalAlgorithm, init, solve, step,
                  truncate_to_tolerance, truncated_svd, truncated_svd_device, update_sol)
from .engine import Engine, colmajor_device, empty_colmajor
from .rhs import BurgersRHS, FactoredRHS, LinearRHS, SylvesterSumRHS
from .distributed import attach_engine, comm_from_torch, row_shard

step_ = step  # `step!`
__all__ = [n for n in dir() if not n.startswith("_")]
