"""Python mirror of the reference's pricing-method plugin interface for the hot path.

Same class names, method names, argument order/meaning and error behaviour as the reference headers, so the
parity tests read like tests of the reference:

    LSM().PredictOptionPrice(pricePaths, r, strike, maturity, dt, isCall, polyOrder)
        include/models/LSMPricer.h:8-14; throws std::runtime_error("LSM::PredictOptionPrice: Empty pricePaths.")
        on empty input (src/models/LSMPricer.cpp:28-30) -> RuntimeError here.

Every call goes through the C ABI (libmcp_b200.so) onto the GPU.  The native C++ mirror of the same classes is
montecarlooptionspricer_b200/host/mcp_plugins.hpp.
"""
from __future__ import annotations

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


class LSM:
    """Longstaff-Schwartz pricer (reference: include/models/LSMPricer.h, src/models/LSMPricer.cpp)."""

    def __init__(self, engine: Engine | None = None):
        self._engine = engine

    def PredictOptionPrice(self, pricePaths, r, strike, maturity, dt, isCall, polyOrder) -> float:
        rows = _as_rows(pricePaths)
        if rows is None:
            raise RuntimeError("LSM::PredictOptionPrice: Empty pricePaths.")
        eng = self._engine or default_engine()
        try:
            return eng.lsm_price_host_rows(rows, r, strike, maturity, dt, isCall, polyOrder)
        except capi.McpError as e:
            if e.code == capi.MCP_ERR_EMPTY_PATHS:
                raise RuntimeError("LSM::PredictOptionPrice: Empty pricePaths.") from e
            raise
