"""Host-side handles over the C ABI: Engine (mcp_ctx) and PathSet (mcp_pathset).

Thin by design -- argument marshalling only.  numpy arrays are HOST buffers handed to the C ABI, which does
its own host<->device copies; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _capi as capi
from ._capi import GbmParams, LsmParams, LsmResult, McpError, RbergomiParams


@dataclass
class LsmOutput:
    price: float
    std_error: float
    sum_v0: float
    sum_sq_dev: float
    n_paths_global: int
    elapsed_ms: float
    n_kernel_launches: int
    coeffs: Optional[np.ndarray] = None
    first_exercise: Optional[np.ndarray] = None
    v0: Optional[np.ndarray] = None


_MODEL_KEYS = ("S0", "r", "xi", "H", "eta", "rho", "dt")
ROW_DTYPE = np.dtype([("model", np.float64, (7,)), ("n_steps", np.int32), ("is_call", np.int32), ("r", np.float64),
                      ("strike", np.float64), ("maturity", np.float64), ("dt", np.float64), ("sigma", np.float64),
                      ("dividend", np.float64)], align=True)  # byte-for-byte mcp_row (include/mcp_b200.h)
assert ROW_DTYPE.itemsize == C.sizeof(capi.Row) and C.sizeof(capi.RowResult) == 5 * 8


def rows_to_array(rows) -> np.ndarray:
    """mcp_row[] as a numpy structured array (ROW_DTYPE).  A ROW_DTYPE array passes through untouched -- callers with
    many rows should build that directly; an iterable of dicts (keys as in Engine.price_rows) is converted column-wise."""
    if isinstance(rows, np.ndarray) and rows.dtype == ROW_DTYPE:
        return np.ascontiguousarray(rows)
    rows = list(rows)
    arr = np.zeros(len(rows), dtype=ROW_DTYPE)
    if rows:
        arr["model"] = [[row["model"][k] for k in _MODEL_KEYS] for row in rows]
        arr["n_steps"] = [int(row["n_steps"]) for row in rows]
        arr["is_call"] = [1 if row["is_call"] else 0 for row in rows]
        for k in ("r", "strike", "maturity", "dt", "sigma", "dividend"):
            arr[k] = [row[k] for row in rows]
    return arr


class Engine:
    """One engine per host thread / per GPU rank (mcp_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._L = capi.lib()
        h = C.c_void_p()
        rc = self._L.mcp_create(device, C.byref(h))
        if rc != 0:
            raise McpError(rc, self._L.mcp_last_error(None).decode())
        self._h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    # -- plumbing ---------------------------------------------------------------------------------
    def _chk(self, rc: int):
        if rc != 0:
            raise McpError(rc, self._L.mcp_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.mcp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream: Optional[int]):
        self._chk(self._L.mcp_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._chk(self._L.mcp_synchronize(self._h))

    def device_info(self) -> dict:
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        fr, to = C.c_size_t(), C.c_size_t()
        self._chk(self._L.mcp_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(fr), C.byref(to)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), free_bytes=fr.value, total_bytes=to.value)

    @property
    def launch_count(self) -> int:
        return int(self._L.mcp_launch_count(self._h))

    def copy_counters(self) -> tuple:
        """(host->device bytes, device->host bytes) this engine has copied so far, counted inside the library."""
        h, d = C.c_uint64(), C.c_uint64()
        self._chk(self._L.mcp_copy_counters(self._h, C.byref(h), C.byref(d)))
        return int(h.value), int(d.value)

    def set_profiling(self, on: bool):
        self._chk(self._L.mcp_set_profiling(self._h, int(on)))

    def profile(self) -> dict:
        p = capi.Profile()
        self._chk(self._L.mcp_get_profile(self._h, C.byref(p)))
        return dict(gen_kernel_ms=p.gen_kernel_ms, sweep_kernels_ms=p.sweep_kernels_ms,
                    n_sweep_launches=p.n_sweep_launches, lsm_total_ms=p.lsm_total_ms, n_sweep_steps=p.n_sweep_steps)

    # -- multi-GPU --------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        L = capi.lib()
        buf = C.create_string_buffer(128)
        rc = L.mcp_comm_unique_id(buf)
        if rc != 0:
            raise McpError(rc, L.mcp_last_error(None).decode())
        return buf.raw

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        self._chk(self._L.mcp_comm_init(self._h, rank, nranks, buf))

    def comm_uses_peer_memory(self) -> bool:
        return bool(self._L.mcp_comm_uses_peer_memory(self._h))

    def comm_info(self):
        r, n = C.c_int(), C.c_int()
        self._chk(self._L.mcp_comm_info(self._h, C.byref(r), C.byref(n)))
        return r.value, n.value

    # -- path sets --------------------------------------------------------------------------------
    def pathset(self, n_paths: int, n_steps: int, dtype: int = capi.MCP_F32) -> "PathSet":
        return PathSet(self, n_paths, n_steps, dtype)

    def upload_paths(self, paths: np.ndarray, dtype: int = capi.MCP_F32) -> "PathSet":
        """paths: host [n_paths][n_steps+1] (the reference's layout)."""
        paths = np.ascontiguousarray(paths, dtype=np.float64)
        if paths.ndim != 2 or paths.shape[0] == 0 or paths.shape[1] == 0:
            raise McpError(capi.MCP_ERR_EMPTY_PATHS, "Empty pricePaths.")
        ps = PathSet(self, paths.shape[0], paths.shape[1] - 1, dtype)
        ps.upload(paths)
        return ps

    # -- generators -------------------------------------------------------------------------------
    def gen_rbergomi(self, ps: "PathSet", S0, r, xi, H, eta, rho, dt, seed: int = 0, path_offset: int = 0,
                     injected: Optional[np.ndarray] = None, dump: bool = False) -> Optional[np.ndarray]:
        prm = RbergomiParams(S0, r, xi, H, eta, rho, dt)
        n = ps.n_steps
        inj = None
        if injected is not None:
            inj = np.ascontiguousarray(injected, dtype=np.float32)
            assert inj.shape == (ps.n_paths, 4 * n), f"injected draws must be [{ps.n_paths}][{4 * n}]"
        out = np.empty((ps.n_paths, 4 * n), dtype=np.float32) if dump else None
        self._chk(self._L.mcp_gen_rbergomi(
            self._h, ps._h, C.byref(prm), seed, path_offset,
            inj.ctypes.data_as(capi._fp) if inj is not None else None,
            out.ctypes.data_as(capi._fp) if out is not None else None))
        return out

    def gen_gbm(self, ps: "PathSet", S0, r, sigma, dt, seed: int = 0, path_offset: int = 0,
                injected: Optional[np.ndarray] = None, dump: bool = False) -> Optional[np.ndarray]:
        prm = GbmParams(S0, r, sigma, dt)
        n = ps.n_steps
        inj = None
        if injected is not None:
            inj = np.ascontiguousarray(injected, dtype=np.float32)
            assert inj.shape == (ps.n_paths, n), f"injected draws must be [{ps.n_paths}][{n}]"
        out = np.empty((ps.n_paths, n), dtype=np.float32) if dump else None
        self._chk(self._L.mcp_gen_gbm(
            self._h, ps._h, C.byref(prm), seed, path_offset,
            inj.ctypes.data_as(capi._fp) if inj is not None else None,
            out.ctypes.data_as(capi._fp) if out is not None else None))
        return out

    def philox_raw(self, seed: int, first: int, count: int, c2: int = 0, c3: int = 0) -> np.ndarray:
        out = np.empty((count, 4), dtype=np.uint32)
        self._chk(self._L.mcp_philox_raw(self._h, seed, first, count, c2, c3, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    # -- LSM --------------------------------------------------------------------------------------
    def lsm_price(self, ps: "PathSet", r, strike, maturity, dt, is_call, poly_order, basis: int = capi.MCP_BASIS_MONOMIAL,
                  carry: int = capi.MCP_F64, want_coeffs: bool = False, want_first_exercise: bool = False,
                  want_v0: bool = False) -> LsmOutput:
        prm = LsmParams(r, strike, maturity, dt, int(bool(is_call)), poly_order, basis, carry)
        res = LsmResult()
        co = np.zeros((max(ps.n_steps, 1), poly_order + (3 if basis == capi.MCP_BASIS_STANDARDISED else 1))) if want_coeffs else None
        fe = np.zeros(ps.n_paths, dtype=np.int32) if want_first_exercise else None
        v0 = np.zeros(ps.n_paths) if want_v0 else None
        self._chk(self._L.mcp_lsm_price(
            self._h, ps._h, C.byref(prm), C.byref(res),
            co.ctypes.data_as(capi._dp) if co is not None else None,
            fe.ctypes.data_as(capi._ip) if fe is not None else None,
            v0.ctypes.data_as(capi._dp) if v0 is not None else None))
        return LsmOutput(res.price, res.std_error, res.sum_v0, res.sum_sq_dev, res.n_paths_global, res.elapsed_ms,
                         res.n_kernel_launches, co, fe, v0)

    def lsm_policy_value(self, ps: "PathSet", coeffs_std: np.ndarray, r, strike, maturity, dt, is_call, poly_order):
        """Out-of-sample value of a fitted exercise policy (mcp_lsm_policy_value): `coeffs_std` is the MCP_BASIS_STANDARDISED
        table of lsm_price(..., want_coeffs=True) on an INDEPENDENT path set.  Returns (LsmOutput, mean stopping column)."""
        co = np.ascontiguousarray(coeffs_std, dtype=np.float64)
        if co.shape != (max(ps.n_steps, 1), poly_order + 3):
            raise ValueError(f"coeffs_std must be [{ps.n_steps}][{poly_order + 3}] (MCP_BASIS_STANDARDISED rows)")
        prm = LsmParams(r, strike, maturity, dt, int(bool(is_call)), poly_order, capi.MCP_BASIS_STANDARDISED, capi.MCP_F64)
        res, stop = LsmResult(), C.c_double()
        self._chk(self._L.mcp_lsm_policy_value(self._h, ps._h, C.byref(prm), co.ctypes.data_as(capi._dp), C.byref(res), C.byref(stop)))
        return (LsmOutput(res.price, res.std_error, res.sum_v0, res.sum_sq_dev, res.n_paths_global, res.elapsed_ms, res.n_kernel_launches),
                stop.value)

    def lsm_price_multi(self, ps: "PathSet", strikes, r, maturity, dt, is_call, poly_order, carry: int = capi.MCP_F32):
        """Several strikes on the same path set (one sweep for the whole ladder in throughput mode)."""
        ks = np.ascontiguousarray(strikes, dtype=np.float64)
        prm = LsmParams(r, 0.0, maturity, dt, int(bool(is_call)), poly_order, capi.MCP_BASIS_MONOMIAL, carry)
        res = (LsmResult * max(ks.size, 1))()
        self._chk(self._L.mcp_lsm_price_multi(self._h, ps._h, C.byref(prm), ks.ctypes.data_as(capi._dp), ks.size, res))
        return [LsmOutput(x.price, x.std_error, x.sum_v0, x.sum_sq_dev, x.n_paths_global, x.elapsed_ms, x.n_kernel_launches) for x in res[:ks.size]]

    def lsm_price_host_rows(self, paths: np.ndarray, r, strike, maturity, dt, is_call, poly_order) -> float:
        """The exact reference call shape: host [N][M] doubles in, the mean out (kept fp64 on the device)."""
        paths = np.ascontiguousarray(paths, dtype=np.float64)
        if paths.ndim != 2 or paths.size == 0:
            rows, N, M = None, 0, 0
        else:
            N, M = paths.shape
            # vector<vector<double>>-shaped argument: an array of row pointers (built in numpy: 1e5 rows cost microseconds, not
            # the third of a second a Python loop over ctypes objects takes)
            self._row_ptrs = (paths.ctypes.data + np.arange(N, dtype=np.uint64) * np.uint64(paths.strides[0])).astype(np.uintp)
            rows = self._row_ptrs.ctypes.data_as(C.POINTER(capi._dp))
        px = C.c_double()
        self._chk(self._L.mcp_lsm_price_host_rows(self._h, rows, N, M, r, strike, maturity, dt, int(bool(is_call)),
                                                  poly_order, C.byref(px)))
        return px.value

    def price_surface_rbergomi_lsm(self, model: dict, strikes, maturities, n_paths: int, r: float, is_call: bool = False,
                                   poly_order: int = 3, carry: int = capi.MCP_F32, steps_per_year: int = 252, seed: int = 0,
                                   path_offset: int = 0, mat_first: int = 0, mat_stride: int = 1):
        """Strike x maturity surface (BASELINE config 5).  Returns (prices[n_mat][n_strikes], std_errors, gen_ms, lsm_ms);
        entries of maturities this call does not own (mat_first / mat_stride) are NaN."""
        m = RbergomiParams(model["S0"], model["r"], model["xi"], model["H"], model["eta"], model["rho"], model["dt"])
        q = LsmParams(r, 0.0, 0.0, model["dt"], int(bool(is_call)), poly_order, capi.MCP_BASIS_MONOMIAL, carry)
        ks = np.ascontiguousarray(strikes, dtype=np.float64)
        ts = np.ascontiguousarray(maturities, dtype=np.float64)
        px = np.full((ts.size, ks.size), np.nan)
        se = np.full((ts.size, ks.size), np.nan)
        g, l = C.c_float(), C.c_float()
        self._chk(self._L.mcp_price_surface_rbergomi_lsm(
            self._h, C.byref(m), C.byref(q), ks.ctypes.data_as(capi._dp), ks.size, ts.ctypes.data_as(capi._dp), ts.size,
            steps_per_year, n_paths, seed, path_offset, mat_first, mat_stride, px.ctypes.data_as(capi._dp),
            se.ctypes.data_as(capi._dp), C.byref(g), C.byref(l)))
        return px, se, g.value, l.value

    def price_rows(self, rows, n_paths: int = 250, poly_order: int = 2, num_branches: int = 10, max_iterations: int = 5,
                   seed: int = 0, path_offset: int = 0):
        """Batched row driver (mcp_price_rows).  rows: iterable of dicts with keys model (dict of S0, r, xi, H, eta, rho,
        dt), n_steps, is_call, r, strike, maturity, dt, sigma, dividend.  Returns (array [n_rows][5] = asymptotic,
        branching, lsm, martingale, lsm_std_error; gen_ms; price_ms)."""
        arr = rows_to_array(rows)
        n = int(arr.shape[0])
        res = np.zeros((max(n, 1), 5), dtype=np.float64)  # mcp_row_result = five doubles
        g, p = C.c_float(), C.c_float()
        self._chk(self._L.mcp_price_rows(self._h, arr.ctypes.data_as(C.POINTER(capi.Row)), n, n_paths, poly_order, num_branches,
                                         max_iterations, seed, path_offset, res.ctypes.data_as(C.POINTER(capi.RowResult)),
                                         C.byref(g), C.byref(p)))
        return res[:n], g.value, p.value

    def gbm_nested_dual(self, S0, r, sigma, dt, strike, is_call, n_steps, poly_order, n_policy_paths, n_outer, n_inner,
                        seed: int = 0, path_offset: int = 0) -> dict:
        """BASELINE config 4: Andersen-Broadie nested-simulation upper bound + policy lower bound under GBM."""
        prm = GbmParams(S0, r, sigma, dt)
        res = capi.DualResult()
        self._chk(self._L.mcp_gbm_nested_dual(self._h, C.byref(prm), strike, int(bool(is_call)), n_steps, poly_order,
                                              n_policy_paths, n_outer, n_inner, seed, path_offset, C.byref(res)))
        return {k: getattr(res, k) for k, _ in capi.DualResult._fields_}

    # -- the other three plugins ------------------------------------------------------
    def asymptotic_price(self, ps: "PathSet", r, strike, maturity, dt, is_call, sigma, dividend) -> float:
        px = C.c_double()
        self._chk(self._L.mcp_asymptotic_price(self._h, ps._h if ps is not None else None, r, strike, maturity, dt,
                                               int(bool(is_call)), sigma, dividend, C.byref(px)))
        return px.value

    def martingale_price(self, ps: "PathSet", r, strike, maturity, dt, is_call, poly_order, max_iterations: int = 5,
                         want_bounds: bool = False):
        px, lo, up = C.c_double(), C.c_double(), C.c_double()
        self._chk(self._L.mcp_martingale_price(self._h, ps._h if ps is not None else None, r, strike, maturity, dt,
                                               int(bool(is_call)), poly_order, max_iterations, C.byref(px), C.byref(lo),
                                               C.byref(up)))
        return (px.value, lo.value, up.value) if want_bounds else px.value

    def branching_price(self, ps: "PathSet", r, strike, maturity, dt, is_call, num_branches, exercise_times,
                        seed: int = 0, path_offset: int = 0, injected_rp: Optional[np.ndarray] = None,
                        want_bounds: bool = False):
        ex = np.ascontiguousarray(exercise_times, dtype=np.int32)
        inj = None
        if injected_rp is not None:
            inj = np.ascontiguousarray(injected_rp, dtype=np.int32)
            assert inj.ndim == 3 and inj.shape[1] == ps.n_paths and inj.shape[2] == num_branches
        px, lo, up = C.c_double(), C.c_double(), C.c_double()
        self._chk(self._L.mcp_branching_price(
            self._h, ps._h if ps is not None else None, r, strike, maturity, dt, int(bool(is_call)), num_branches,
            ex.ctypes.data_as(C.POINTER(C.c_int)) if ex.size else None, int(ex.size), seed, path_offset,
            inj.ctypes.data_as(capi._ip) if inj is not None else None, C.byref(px), C.byref(lo), C.byref(up)))
        return (px.value, lo.value, up.value) if want_bounds else px.value

    # -- exact-signature generator (host estimators + device generation) --------------------------
    @staticmethod
    def estimate_rbergomi_params(hist) -> dict:
        L = capi.lib()
        h = np.ascontiguousarray(hist, dtype=np.float64)
        out = RbergomiParams()
        rc = L.mcp_estimate_rbergomi_params(h.ctypes.data_as(capi._dp), h.size, C.byref(out))
        if rc != 0:
            raise McpError(rc, L.mcp_last_error(None).decode())
        return {k: getattr(out, k) for k in ("S0", "r", "xi", "H", "eta", "rho", "dt")}

    @staticmethod
    def rbergomi_host_tables(n_steps: int, model: dict) -> dict:
        """The generator's host-built constant tables (pure host code; known-answer tests): phis complex [M'],
        comp2 [M'], sw [M'] with M' = nextPow2(n_steps)."""
        L = capi.lib()
        prm = RbergomiParams(model["S0"], model["r"], model["xi"], model["H"], model["eta"], model["rho"], model["dt"])
        Mp = 1
        while Mp < n_steps:
            Mp <<= 1
        phis, comp2, sw = np.zeros(2 * Mp, np.float32), np.zeros(Mp, np.float32), np.zeros(Mp, np.float32)
        rc = L.mcp_rbergomi_host_tables(n_steps, C.byref(prm), phis.ctypes.data_as(capi._fp), comp2.ctypes.data_as(capi._fp),
                                        sw.ctypes.data_as(capi._fp))
        if rc < 0:
            raise McpError(rc, "mcp_rbergomi_host_tables: invalid arguments")
        assert rc == Mp
        return dict(Mp=Mp, phis=phis[0::2].astype(np.float64) + 1j * phis[1::2].astype(np.float64), comp2=comp2, sw=sw)

    def generate_stock_price_paths(self, hist, forward_steps: int, path_num: int, seed: int = 0,
                                   path_offset: int = 0) -> np.ndarray:
        h = np.ascontiguousarray(hist, dtype=np.float64)
        out = np.zeros((max(path_num, 0), max(forward_steps, 0) + 1), dtype=np.float64)
        ptrs = (out.ctypes.data + np.arange(max(path_num, 1), dtype=np.uint64) * np.uint64(out.strides[0] if path_num > 0 else 0)).astype(np.uintp)
        rows = ptrs.ctypes.data_as(C.POINTER(capi._dp))
        self._chk(self._L.mcp_generate_stock_price_paths(self._h, h.ctypes.data_as(capi._dp), h.size, forward_steps,
                                                         path_num, seed, path_offset, rows))
        return out

    def price_rbergomi_lsm(self, model: dict, lsm: dict, n_paths: int, n_steps: int, seed: int = 0,
                           path_offset: int = 0):
        """Generate (native Philox) + LSM on the device: parameters in, a result out."""
        m = RbergomiParams(model["S0"], model["r"], model["xi"], model["H"], model["eta"], model["rho"], model["dt"])
        q = LsmParams(lsm["r"], lsm["strike"], lsm["maturity"], lsm["dt"], int(bool(lsm["is_call"])), lsm["poly_order"],
                      lsm.get("basis", capi.MCP_BASIS_MONOMIAL), lsm.get("carry", capi.MCP_F32))
        res = LsmResult()
        gen_ms = C.c_float()
        self._chk(self._L.mcp_price_rbergomi_lsm(self._h, C.byref(m), C.byref(q), n_paths, n_steps, seed, path_offset,
                                                 C.byref(res), C.byref(gen_ms)))
        out = LsmOutput(res.price, res.std_error, res.sum_v0, res.sum_sq_dev, res.n_paths_global, res.elapsed_ms,
                        res.n_kernel_launches)
        return out, gen_ms.value


class PathSet:
    """Device-resident time-major slab S[(n_steps+1)][ld] (mcp_pathset)."""

    def __init__(self, eng: Engine, n_paths: int, n_steps: int, dtype: int = capi.MCP_F32):
        self._eng = eng
        self._L = eng._L
        h = C.c_void_p()
        eng._chk(self._L.mcp_pathset_create(eng._h, n_paths, n_steps, dtype, C.byref(h)))
        self._h = h
        self.n_paths, self.n_steps, self.dtype = n_paths, n_steps, dtype

    def close(self):
        if getattr(self, "_h", None) and getattr(self._eng, "_h", None):
            self._L.mcp_pathset_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        n, s, ld, dt, p = C.c_int64(), C.c_int(), C.c_int64(), C.c_int(), C.c_void_p()
        self._eng._chk(self._L.mcp_pathset_info(self._h, C.byref(n), C.byref(s), C.byref(ld), C.byref(dt), C.byref(p)))
        return dict(n_paths=n.value, n_steps=s.value, ld=ld.value, dtype=dt.value, device_ptr=p.value)

    def upload(self, paths: np.ndarray):
        paths = np.ascontiguousarray(paths, dtype=np.float64)
        assert paths.shape == (self.n_paths, self.n_steps + 1)
        self._eng._chk(self._L.mcp_pathset_upload_f64(self._h, paths.ctypes.data_as(capi._dp), paths.shape[1]))

    def download(self) -> np.ndarray:
        out = np.empty((self.n_paths, self.n_steps + 1), dtype=np.float64)
        self._eng._chk(self._L.mcp_pathset_download_f64(self._h, out.ctypes.data_as(capi._dp), out.shape[1]))
        return out

    def download_timemajor(self) -> np.ndarray:
        out = np.empty((self.n_steps + 1, self.n_paths), dtype=np.float32)
        self._eng._chk(self._L.mcp_pathset_download_timemajor_f32(self._h, out.ctypes.data_as(capi._fp), out.shape[1]))
        return out
