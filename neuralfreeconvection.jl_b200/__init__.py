"""neuralfreeconvection.jl_b200 -- B200-native (sm_100a) hot path of NeuralFreeConvection.jl.

The directory name contains a dot, so it is imported through the root-level `nfc_b200` alias module
(`import nfc_b200`), which loads this package with importlib.
"""
from . import _binding, api, distributed, formats, synthetic
from ._binding import Handle, NfcError, load, LIB_PATH, EXPORTS, measure_fp32_peak, comm_unique_id
from .api import (ADAM, Chain, ColumnDataset, Conv, ConvectiveAdjustmentNDE, ConvectiveAdjustmentNDEParameters, Dense,
                  FreeConvectionNDE, FreeConvectionNDEParameters, Tsit5, RKC2, ROCK4, ZeroMeanUnitVarianceScaling, destructure,
                  named_chain, compute_nde_solution_history, load_history, oceananigans_convective_adjustment, oceananigans_convective_adjustment_with_neural_network, solve_nde,
                  train_neural_differential_equation)

__all__ = ["measure_fp32_peak", "comm_unique_id", "Handle", "NfcError", "load", "LIB_PATH", "EXPORTS", "ADAM", "Chain", "ColumnDataset", "Conv",
           "ConvectiveAdjustmentNDE", "ConvectiveAdjustmentNDEParameters", "Dense", "FreeConvectionNDE",
           "FreeConvectionNDEParameters", "Tsit5", "RKC2", "ROCK4", "ZeroMeanUnitVarianceScaling", "destructure", "named_chain",
           "solve_nde", "train_neural_differential_equation", "oceananigans_convective_adjustment",
           "oceananigans_convective_adjustment_with_neural_network", "compute_nde_solution_history", "load_history", "synthetic", "api", "_binding", "distributed", "formats"]
