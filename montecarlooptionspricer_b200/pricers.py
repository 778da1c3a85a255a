"""Python mirror of the reference's pricing-method plugin interface for the hot path.

Same class names, method names, argument order/meaning and error behaviour as the reference headers, so the
parity tests read like tests of the reference:

    RoughVolatility().GenerateStockPricePaths(historical_prices, forward_steps, path_num)
        include/models/RoughVolatility.h:15-19; "Historical prices vector too small." (RoughVolatility.cpp:317-319)
    LSM().PredictOptionPrice(pricePaths, r, strike, maturity, dt, isCall, polyOrder)
        include/models/LSMPricer.h:8-14; "LSM::PredictOptionPrice: Empty pricePaths." (LSMPricer.cpp:28-30)
    MartingaleOptimization().PredictOptionPrice(..., polyOrder, maxIterations=5)
        include/models/MartingaleOptimizationPricer.h:10-18; errors MartingaleOptimizationPricer.cpp:31-36
    BranchingProcesses().PredictOptionPrice(..., numBranches, exerciseTimes)
        include/models/BranchingProcessPricer.h:8-16; errors BranchingProcessPricer.cpp:22-30
    AsymptoticAnalysis().PredictOptionPrice(..., sigma, dividend)
        include/models/AsymptoticAnalysisPricer.h:8-15; 0.0 on empty input, throws on sigma <= 0 (:48-53)

std::runtime_error -> RuntimeError with the reference's message.  Every call goes through the C ABI
(libmcp_b200.so) onto the GPU; nothing here computes.  The native C++ mirror of the same classes is
montecarlooptionspricer_b200/host/mcp_plugins.hpp.
"""
from __future__ import annotations

import itertools
import os
import threading

import numpy as np

from . import _capi as capi
from .engine import Engine

_tls = threading.local()


def default_engine(device: int = 0) -> Engine:
    """One engine per host thread (the reference instantiates its pricers per OpenMP thread,
    src/core/PredictionGen.cpp:566-570)."""
    eng = getattr(_tls, "engine", None)
    if eng is None or eng._h is None or eng.device != device:
        eng = Engine(device)
        _tls.engine = eng
    return eng


def _as_rows(pricePaths):
    if pricePaths is None or len(pricePaths) == 0 or len(pricePaths[0]) == 0:
        return None
    return np.ascontiguousarray(pricePaths, dtype=np.float64)


def _rethrow(e: capi.McpError):
    """Status codes that stand for a std::runtime_error of the reference carry its message verbatim."""
    if e.code in (capi.MCP_ERR_EMPTY_PATHS, capi.MCP_ERR_DOMAIN):
        raise RuntimeError(e.msg) from e
    raise e


class _Plugin:
    def __init__(self, engine: Engine | None = None):
        self._engine = engine

    def _eng(self) -> Engine:
        return self._engine or default_engine()


class RoughVolatility(_Plugin):
    """Rough-volatility path generator (reference: include/models/RoughVolatility.h, src/models/RoughVolatility.cpp).

    The reference seeds std::mt19937 from std::random_device on every call (:239-241), i.e. it is not
    reproducible; here the Philox seed defaults to OS entropy and successive calls advance the path counter,
    `seed=` makes a run reproducible."""

    _calls = itertools.count()

    def __init__(self, engine: Engine | None = None, seed: int | None = None):
        super().__init__(engine)
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self._next_path = 0

    def GenerateStockPricePaths(self, historical_prices, forward_steps, path_num):
        hist = np.ascontiguousarray(historical_prices, dtype=np.float64)
        if hist.size < 2:
            raise RuntimeError("Historical prices vector too small.")
        try:
            out = self._eng().generate_stock_price_paths(hist, int(forward_steps), int(path_num), seed=self._seed,
                                                         path_offset=self._next_path)
        except capi.McpError as e:
            _rethrow(e)
        self._next_path += max(int(path_num), 0)
        return out


class LSM(_Plugin):
    """Longstaff-Schwartz pricer (reference: include/models/LSMPricer.h, src/models/LSMPricer.cpp)."""

    def PredictOptionPrice(self, pricePaths, r, strike, maturity, dt, isCall, polyOrder) -> float:
        rows = _as_rows(pricePaths)
        if rows is None:
            raise RuntimeError("LSM::PredictOptionPrice: Empty pricePaths.")
        try:
            return self._eng().lsm_price_host_rows(rows, r, strike, maturity, dt, isCall, polyOrder)
        except capi.McpError as e:
            _rethrow(e)


class MartingaleOptimization(_Plugin):
    """reference: include/models/MartingaleOptimizationPricer.h, src/models/MartingaleOptimizationPricer.cpp"""

    def PredictOptionPrice(self, pricePaths, r, strike, maturity, dt, isCall, polyOrder, maxIterations=5) -> float:
        rows = _as_rows(pricePaths)
        if rows is None:
            raise RuntimeError("MartingaleOptimization: Empty pricePaths.")
        if maxIterations <= 0:
            raise RuntimeError("MartingaleOptimization: maxIterations must be positive.")
        eng = self._eng()
        try:
            ps = eng.upload_paths(rows, dtype=capi.MCP_F64)
            try:
                return eng.martingale_price(ps, r, strike, maturity, dt, isCall, polyOrder, maxIterations)
            finally:
                ps.close()
        except capi.McpError as e:
            _rethrow(e)


class BranchingProcesses(_Plugin):
    """reference: include/models/BranchingProcessPricer.h, src/models/BranchingProcessPricer.cpp.  The reference
    resamples paths with a std::mt19937 seeded from std::random_device and shared (racily) between OpenMP threads
    (:84-85, :108): its upper bound is random; here the resampling stream is Philox keyed by `seed`."""

    def __init__(self, engine: Engine | None = None, seed: int | None = None):
        super().__init__(engine)
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)

    def PredictOptionPrice(self, pricePaths, r, strike, maturity, dt, isCall, numBranches, exerciseTimes) -> float:
        rows = _as_rows(pricePaths)
        if rows is None:
            raise RuntimeError("BranchingProcesses: Empty pricePaths.")
        if exerciseTimes is None or len(exerciseTimes) == 0:
            raise RuntimeError("BranchingProcesses: No exercise times.")
        if strike <= 0.0:
            raise RuntimeError("BranchingProcesses: Strike must be positive.")
        eng = self._eng()
        try:
            ps = eng.upload_paths(rows, dtype=capi.MCP_F64)
            try:
                return eng.branching_price(ps, r, strike, maturity, dt, isCall, numBranches, exerciseTimes, seed=self._seed)
            finally:
                ps.close()
        except capi.McpError as e:
            _rethrow(e)


class AsymptoticAnalysis(_Plugin):
    """reference: include/models/AsymptoticAnalysisPricer.h, src/models/AsymptoticAnalysisPricer.cpp"""

    def PredictOptionPrice(self, pricePaths, r, strike, maturity, dt, isCall, sigma, dividend) -> float:
        rows = _as_rows(pricePaths)
        if rows is None:
            return 0.0  # :48-50
        if sigma <= 0.0:
            raise RuntimeError("AsymptoticAnalysis: Volatility must be positive.")
        eng = self._eng()
        try:
            ps = eng.upload_paths(rows, dtype=capi.MCP_F64)
            try:
                return eng.asymptotic_price(ps, r, strike, maturity, dt, isCall, sigma, dividend)
            finally:
                ps.close()
        except capi.McpError as e:
            _rethrow(e)
