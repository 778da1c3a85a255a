"""montecarlooptionspricer_b200 -- B200-native (sm_100a) Monte-Carlo hot path of bcosm/MonteCarloOptionsPricer.

Philox normals -> GBM / rough-volatility path simulation -> Longstaff-Schwartz backward induction, as hand-written
CUDA behind a C ABI (include/mcp_b200.h).  This Python package is only the host-side mirror of the reference's
plugin interface plus marshalling; importing it does not require a GPU, calling it does (no CPU fallback).
"""
from . import _capi as capi
from ._capi import (MCP_BASIS_LAGUERRE, MCP_BASIS_MONOMIAL, MCP_BASIS_STANDARDISED, MCP_F32, MCP_F64, McpError)
from .engine import Engine, LsmOutput, PathSet
from .pricers import (LSM, AsymptoticAnalysis, BranchingProcesses, MartingaleOptimization, RoughVolatility,
                      default_engine)

__all__ = ["capi", "Engine", "PathSet", "LsmOutput", "LSM", "RoughVolatility", "MartingaleOptimization", "BranchingProcesses", "AsymptoticAnalysis",
           "default_engine", "McpError", "MCP_F32", "MCP_F64",
           "MCP_BASIS_MONOMIAL", "MCP_BASIS_LAGUERRE", "MCP_BASIS_STANDARDISED"]
