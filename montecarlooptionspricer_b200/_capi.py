"""ctypes binding of include/mcp_b200.h (libmcp_b200.so).  No compute happens in Python: every call below
lands in the CUDA library, and importing fails loudly when the library has not been built."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmcp_b200.so")

# status codes (include/mcp_b200.h: enum mcp_status)
MCP_OK, MCP_ERR_INVALID, MCP_ERR_CUDA, MCP_ERR_NCCL, MCP_ERR_NOMEM = 0, -1, -2, -3, -4
MCP_ERR_EMPTY_PATHS, MCP_ERR_UNSUPPORTED, MCP_ERR_DOMAIN = -5, -6, -7
MCP_F32, MCP_F64 = 0, 1
MCP_BASIS_MONOMIAL, MCP_BASIS_LAGUERRE, MCP_BASIS_STANDARDISED = 0, 1, 2


class RbergomiParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("S0", "r", "xi", "H", "eta", "rho", "dt")]


class GbmParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("S0", "r", "sigma", "dt")]


class LsmParams(C.Structure):
    _fields_ = [("r", C.c_double), ("strike", C.c_double), ("maturity", C.c_double), ("dt", C.c_double),
                ("is_call", C.c_int), ("poly_order", C.c_int), ("basis", C.c_int), ("carry", C.c_int)]


class LsmResult(C.Structure):
    _fields_ = [("price", C.c_double), ("std_error", C.c_double), ("sum_v0", C.c_double), ("sum_sq_dev", C.c_double),
                ("n_paths_global", C.c_int64), ("elapsed_ms", C.c_float), ("n_kernel_launches", C.c_int)]


class Row(C.Structure):
    _fields_ = [("model", RbergomiParams), ("n_steps", C.c_int), ("is_call", C.c_int), ("r", C.c_double), ("strike", C.c_double),
                ("maturity", C.c_double), ("dt", C.c_double), ("sigma", C.c_double), ("dividend", C.c_double)]


class RowResult(C.Structure):
    _fields_ = [("asymptotic", C.c_double), ("branching", C.c_double), ("lsm", C.c_double), ("martingale", C.c_double),
                ("lsm_std_error", C.c_double)]


class DualResult(C.Structure):
    _fields_ = [("lower", C.c_double), ("lower_se", C.c_double), ("upper", C.c_double), ("upper_se", C.c_double),
                ("n_outer_global", C.c_int64), ("policy_ms", C.c_float), ("outer_ms", C.c_float), ("nested_ms", C.c_float)]


class Profile(C.Structure):
    _fields_ = [("gen_kernel_ms", C.c_float), ("sweep_kernels_ms", C.c_float), ("n_sweep_launches", C.c_int),
                ("lsm_total_ms", C.c_float), ("n_sweep_steps", C.c_int)]


_vp = C.c_void_p
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)

# name -> (restype, argtypes): every symbol include/mcp_b200.h declares
SIGNATURES = {
    "mcp_abi_version": (C.c_int, []),
    "mcp_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "mcp_destroy": (C.c_int, [_vp]),
    "mcp_last_error": (C.c_char_p, [_vp]),
    "mcp_set_stream": (C.c_int, [_vp, _vp]),
    "mcp_synchronize": (C.c_int, [_vp]),
    "mcp_device_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mcp_launch_count": (C.c_uint64, [_vp]),
    "mcp_copy_counters": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mcp_set_profiling": (C.c_int, [_vp, C.c_int]),
    "mcp_get_profile": (C.c_int, [_vp, C.POINTER(Profile)]),
    "mcp_comm_unique_id": (C.c_int, [_vp]),
    "mcp_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "mcp_comm_uses_peer_memory": (C.c_int, [_vp]),
    "mcp_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mcp_pathset_create": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, C.POINTER(_vp)]),
    "mcp_pathset_destroy": (C.c_int, [_vp]),
    "mcp_pathset_info": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int), C.POINTER(_vp)]),
    "mcp_pathset_upload_f64": (C.c_int, [_vp, _dp, C.c_int64]),
    "mcp_pathset_upload_rows_f64": (C.c_int, [_vp, C.POINTER(_dp)]),
    "mcp_pathset_download_f64": (C.c_int, [_vp, _dp, C.c_int64]),
    "mcp_pathset_download_rows_f64": (C.c_int, [_vp, C.POINTER(_dp)]),
    "mcp_pathset_download_timemajor_f32": (C.c_int, [_vp, _fp, C.c_int64]),
    "mcp_gen_rbergomi": (C.c_int, [_vp, _vp, C.POINTER(RbergomiParams), C.c_uint64, C.c_uint64, _fp, _fp]),
    "mcp_gen_gbm": (C.c_int, [_vp, _vp, C.POINTER(GbmParams), C.c_uint64, C.c_uint64, _fp, _fp]),
    "mcp_rbergomi_host_tables": (C.c_int, [C.c_int, C.POINTER(RbergomiParams), _fp, _fp, _fp]),
    "mcp_philox_raw": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32,
                                 C.POINTER(C.c_uint32)]),
    "mcp_lsm_price": (C.c_int, [_vp, _vp, C.POINTER(LsmParams), C.POINTER(LsmResult), _dp, _ip, _dp]),
    "mcp_lsm_price_multi": (C.c_int, [_vp, _vp, C.POINTER(LsmParams), _dp, C.c_int, C.POINTER(LsmResult)]),
    "mcp_lsm_policy_value": (C.c_int, [_vp, _vp, C.POINTER(LsmParams), _dp, C.POINTER(LsmResult), _dp]),
    "mcp_lsm_price_host_rows": (C.c_int, [_vp, C.POINTER(_dp), C.c_int64, C.c_int, C.c_double, C.c_double,
                                          C.c_double, C.c_double, C.c_int, C.c_int, _dp]),
    "mcp_price_rbergomi_lsm": (C.c_int, [_vp, C.POINTER(RbergomiParams), C.POINTER(LsmParams), C.c_int64, C.c_int,
                                         C.c_uint64, C.c_uint64, C.POINTER(LsmResult), _fp]),
    "mcp_price_surface_rbergomi_lsm": (C.c_int, [_vp, C.POINTER(RbergomiParams), C.POINTER(LsmParams), _dp, C.c_int, _dp, C.c_int,
                                                 C.c_int, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, _dp, _dp, _fp, _fp]),
    "mcp_gbm_nested_dual": (C.c_int, [_vp, C.POINTER(GbmParams), C.c_double, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                      C.c_uint64, C.c_uint64, C.POINTER(DualResult)]),
    "mcp_price_rows": (C.c_int, [_vp, C.POINTER(Row), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                 C.POINTER(RowResult), _fp, _fp]),
    "mcp_estimate_rbergomi_params": (C.c_int, [_dp, C.c_int64, C.POINTER(RbergomiParams)]),
    "mcp_generate_stock_price_paths": (C.c_int, [_vp, _dp, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                                 C.POINTER(_dp)]),
    "mcp_asymptotic_price": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                       C.c_double, _dp]),
    "mcp_martingale_price": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                       C.c_int, _dp, _dp, _dp]),
    "mcp_branching_price": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                      C.POINTER(C.c_int), C.c_int, C.c_uint64, C.c_uint64, _ip, _dp, _dp, _dp]),
}

_lib = None


class McpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[mcp_b200 {code}] {msg}")
        self.code = code
        self.msg = msg


def lib() -> C.CDLL:
    """Load libmcp_b200.so (built by montecarlooptionspricer_b200.build).  There is no fallback."""
    global _lib
    if _lib is None:
        path = os.environ.get("MCP_B200_LIB") or LIB_PATH   # MCP_B200_LIB: load another build of the SAME library (libmcp_b200_dbg.so)
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: the CUDA extension has not been built "
                "(run `python -m montecarlooptionspricer_b200.build`). This package has no CPU fallback.")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
