"""oracle/oracle.py -- TEST INFRASTRUCTURE: ctypes doors into the two CPU checkers.

  * ``port``  -> oracle/libmcp_oracle.so  (plain-C restatement, oracle/port/mcp_oracle.c)
  * ``ref``   -> oracle/_ref/libmcp_ref.so (the reference's own translation units, oracle/ref_api.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (montecarlooptionspricer_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "libmcp_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libmcp_ref.so")

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_long)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile. `ref` only succeeds where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------------------------- port
class _Port:
    def __init__(self):
        if not os.path.exists(_PORT_SO):
            build(ref=False)
        L = C.CDLL(_PORT_SO)
        L.orc_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.orc_philox4x32_10.restype = None
        L.orc_box_muller.argtypes = [C.c_uint32, C.c_uint32, _dp, _dp]
        L.orc_box_muller.restype = None
        L.orc_rbergomi_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_long, C.c_int, C.c_double, _dp]
        L.orc_rbergomi_draws.restype = None
        L.orc_gbm_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_long, C.c_int, _dp]
        L.orc_gbm_draws.restype = None
        L.orc_rbergomi_phi.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp]
        L.orc_rbergomi_phi.restype = C.c_int
        L.orc_rbergomi_paths.argtypes = [C.c_double] * 7 + [C.c_int, C.c_long, _dp, _dp, _dp, _dp]
        L.orc_rbergomi_paths.restype = C.c_int
        L.orc_gbm_paths.argtypes = [C.c_double] * 4 + [C.c_int, C.c_long, _dp, _dp]
        L.orc_gbm_paths.restype = C.c_int
        L.orc_lsm.argtypes = [_dp, C.c_long, C.c_long, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                              C.c_int, _dp, _dp, _dp, _ip, _u8p, _dp, _dp, _lp]
        L.orc_lsm.restype = C.c_int
        L.orc_lsm_timemajor_f32.argtypes = [_fp, C.c_long, C.c_long, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _ip, _dp, _dp, _lp]
        L.orc_lsm_timemajor_f32.restype = C.c_int
        common = [_dp, C.c_long, C.c_long, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_asymptotic.argtypes = common + [C.c_double, C.c_double, _dp]
        L.orc_martingale.argtypes = common + [C.c_int, C.c_int, _dp, _dp, _dp]
        L.orc_branching.argtypes = common + [C.c_int, _ip, C.c_int, _ip, _dp, _dp, _dp]
        L.orc_estimate_params.argtypes = [_dp, C.c_long, _dp]
        self.L = L

    def philox(self, ctr, key):
        c = np.asarray(ctr, dtype=np.uint32).copy()
        k = np.asarray(key, dtype=np.uint32).copy()
        o = np.zeros(4, dtype=np.uint32)
        self.L.orc_philox4x32_10(_ptr(c, _u32p), _ptr(k, _u32p), _ptr(o, _u32p))
        return o

    def box_muller(self, a, b):
        z0, z1 = C.c_double(), C.c_double()
        self.L.orc_box_muller(int(a), int(b), C.byref(z0), C.byref(z1))
        return z0.value, z1.value

    def rbergomi_draws(self, seed, path0, n_paths, n, rho):
        """Native-stream normals in the reference's 4-slot order (W slots carry rho*W and sqrt(1-rho^2)*W)."""
        d = np.empty((n_paths, 4 * n), dtype=np.float64)
        self.L.orc_rbergomi_draws(seed, path0, n_paths, n, rho, _ptr(d, _dp))
        return d

    def gbm_draws(self, seed, path0, n_paths, n):
        d = np.empty((n_paths, n), dtype=np.float64)
        self.L.orc_gbm_draws(seed, path0, n_paths, n, _ptr(d, _dp))
        return d

    def rbergomi_phi(self, n, H, dt):
        M = 1
        while M < n + 1:
            M <<= 1
        re, im = np.zeros(M), np.zeros(M)
        m = self.L.orc_rbergomi_phi(n, H, dt, _ptr(re, _dp), _ptr(im, _dp))
        assert m == M
        return re + 1j * im

    def rbergomi_paths(self, S0, r, xi, H, eta, rho, dt, n, draws, want_xv=False):
        draws = _f64(draws)
        P = draws.shape[0]
        assert draws.shape[1] == 4 * n
        out = np.empty((P, n + 1))
        X = np.empty((P, n)) if want_xv else None
        v = np.empty((P, n)) if want_xv else None
        rc = self.L.orc_rbergomi_paths(S0, r, xi, H, eta, rho, dt, n, P, _ptr(draws, _dp), _ptr(out, _dp),
                                       _ptr(X, _dp), _ptr(v, _dp))
        assert rc == 0
        return (out, X, v) if want_xv else out

    def gbm_paths(self, S0, r, sigma, dt, n, draws):
        draws = _f64(draws)
        P = draws.shape[0]
        assert draws.shape[1] == n
        out = np.empty((P, n + 1))
        rc = self.L.orc_gbm_paths(S0, r, sigma, dt, n, P, _ptr(draws, _dp), _ptr(out, _dp))
        assert rc == 0
        return out

    def lsm(self, paths, r, K, T, dt, is_call, p, want_mask=False):
        """paths [N][M] (path-major, doubles).  Returns a dict of the instrumented outputs."""
        paths = _f64(paths)
        N, M = paths.shape
        price, se, gap = C.c_double(), C.c_double(), C.c_double()
        coeffs = np.zeros((M - 1, p + 1))
        first = np.zeros(N, dtype=np.int32)
        mask = np.zeros((N, M), dtype=np.uint8) if want_mask else None
        V0 = np.zeros(N)
        nitm = np.zeros(max(M - 1, 1), dtype=np.int64)
        rc = self.L.orc_lsm(_ptr(paths, _dp), N, M, r, K, T, dt, int(is_call), p, C.byref(price), C.byref(se),
                            _ptr(coeffs, _dp), _ptr(first, _ip), _ptr(mask, _u8p), _ptr(V0, _dp), C.byref(gap),
                            _ptr(nitm, _lp))
        if rc != 0:
            raise RuntimeError(f"orc_lsm rc={rc}")
        return dict(price=price.value, stderr=se.value, coeffs=coeffs, first_ex=first, ex_mask=mask, V0=V0,
                    min_gap=gap.value, n_itm=nitm)

    def asymptotic(self, paths, r, K, T, dt, is_call, sigma, dividend):
        paths = _f64(paths)
        px = C.c_double()
        rc = self.L.orc_asymptotic(_ptr(paths, _dp), paths.shape[0], paths.shape[1], r, K, T, dt, int(is_call), sigma,
                                   dividend, C.byref(px))
        if rc != 0:
            raise RuntimeError(f"orc_asymptotic rc={rc}")
        return px.value

    def martingale(self, paths, r, K, T, dt, is_call, p, max_iter=5):
        paths = _f64(paths)
        px, lo, up = C.c_double(), C.c_double(), C.c_double()
        rc = self.L.orc_martingale(_ptr(paths, _dp), paths.shape[0], paths.shape[1], r, K, T, dt, int(is_call), p,
                                   max_iter, C.byref(px), C.byref(lo), C.byref(up))
        if rc != 0:
            raise RuntimeError(f"orc_martingale rc={rc}")
        return dict(price=px.value, primal=lo.value, dual=up.value)

    def branching(self, paths, r, K, T, dt, is_call, num_branches, exercise_times, rp):
        """rp: int32 [n_visited_dates][N][num_branches] resampled path indices (injected)."""
        paths = _f64(paths)
        ex = np.ascontiguousarray(exercise_times, dtype=np.int32)
        q = np.ascontiguousarray(rp, dtype=np.int32)
        px, lo, up = C.c_double(), C.c_double(), C.c_double()
        rc = self.L.orc_branching(_ptr(paths, _dp), paths.shape[0], paths.shape[1], r, K, T, dt, int(is_call),
                                  num_branches, _ptr(ex, _ip), ex.size, _ptr(q, _ip), C.byref(px), C.byref(lo),
                                  C.byref(up))
        if rc != 0:
            raise RuntimeError(f"orc_branching rc={rc}")
        return dict(price=px.value, lower=lo.value, upper=up.value)

    def estimate_params(self, hist):
        hist = _f64(hist)
        out = np.zeros(7)
        rc = self.L.orc_estimate_params(_ptr(hist, _dp), hist.size, _ptr(out, _dp))
        if rc != 0:
            raise RuntimeError(f"orc_estimate_params rc={rc}")
        return dict(zip(("S0", "r", "xi", "H", "eta", "rho", "dt"), out))

    def lsm_timemajor_f32(self, slab, r, K, T, dt, is_call, p):
        """slab [M][N] float32 (the GPU's own layout).  Prices exactly those values widened to double."""
        slab = np.ascontiguousarray(slab, dtype=np.float32)
        M, N = slab.shape
        price, se, gap = C.c_double(), C.c_double(), C.c_double()
        coeffs = np.zeros((M - 1, p + 1))
        first = np.zeros(N, dtype=np.int32)
        V0 = np.zeros(N)
        nitm = np.zeros(max(M - 1, 1), dtype=np.int64)
        rc = self.L.orc_lsm_timemajor_f32(_ptr(slab, _fp), N, M, r, K, T, dt, int(is_call), p, C.byref(price),
                                          C.byref(se), _ptr(coeffs, _dp), _ptr(first, _ip), _ptr(V0, _dp),
                                          C.byref(gap), _ptr(nitm, _lp))
        if rc != 0:
            raise RuntimeError(f"orc_lsm_timemajor_f32 rc={rc}")
        return dict(price=price.value, stderr=se.value, coeffs=coeffs, first_ex=first, V0=V0, min_gap=gap.value,
                    n_itm=nitm)


# ------------------------------------------------------------------------------------------------ ref
class _Ref:
    def __init__(self):
        if not os.path.exists(_REF_SO):
            raise FileNotFoundError(
                f"{_REF_SO} missing: run `make -C oracle ref` in the authoring container (needs /root/reference)")
        L = C.CDLL(_REF_SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_omp_max_threads.restype = C.c_int
        L.ref_generate_paths.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_size_t, _dp,
                                         C.POINTER(C.c_size_t)]
        L.ref_estimate_params.argtypes = [_dp, C.c_int, _dp]
        L.ref_rbergomi_phi.argtypes = [C.c_int, C.c_double, C.c_double, _dp, C.c_int, C.POINTER(C.c_int)]
        L.ref_rbergomi_paths.argtypes = [C.c_double] * 7 + [C.c_int, C.c_long, _dp, _dp, _dp, _dp]
        common = [_dp, C.c_long, C.c_long, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ref_lsm_price.argtypes = common + [C.c_int, _dp]
        L.ref_martingale_price.argtypes = common + [C.c_int, C.c_int, _dp]
        L.ref_branching_price.argtypes = common + [C.c_int, _ip, C.c_int, _ip, C.c_long, C.POINTER(C.c_long), _dp]
        L.ref_asymptotic_price.argtypes = common + [C.c_double, C.c_double, _dp]
        L.ref_bench_rows.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                     C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        self.L = L

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    def omp_max_threads(self):
        return int(self.L.ref_omp_max_threads())

    def generate_paths(self, hist, steps, n_paths, draws=None):
        hist = _f64(hist)
        out = np.empty((n_paths, steps + 1))
        used = C.c_size_t(0)
        d = _f64(draws).ravel() if draws is not None else None
        self._chk(self.L.ref_generate_paths(_ptr(hist, _dp), hist.size, steps, n_paths, _ptr(d, _dp),
                                            d.size if d is not None else 0, _ptr(out, _dp), C.byref(used)))
        return out, used.value

    def estimate_params(self, hist):
        hist = _f64(hist)
        out = np.zeros(5)
        self._chk(self.L.ref_estimate_params(_ptr(hist, _dp), hist.size, _ptr(out, _dp)))
        return dict(xi=out[0], H=out[1], eta=out[2], rho=out[3], S0=out[4])

    def rbergomi_phi(self, n, H, dt):
        buf = np.zeros(2 * 4096)
        M = C.c_int(0)
        self._chk(self.L.ref_rbergomi_phi(n, H, dt, _ptr(buf, _dp), 4096, C.byref(M)))
        return buf[0:2 * M.value:2] + 1j * buf[1:2 * M.value:2]

    def rbergomi_paths(self, S0, r, xi, H, eta, rho, dt, n, draws, want_xv=False):
        draws = _f64(draws)
        P = draws.shape[0]
        assert draws.shape[1] == 4 * n
        out = np.empty((P, n + 1))
        X = np.empty((P, n)) if want_xv else None
        v = np.empty((P, n)) if want_xv else None
        self._chk(self.L.ref_rbergomi_paths(S0, r, xi, H, eta, rho, dt, n, P, _ptr(draws, _dp), _ptr(out, _dp),
                                            _ptr(X, _dp), _ptr(v, _dp)))
        return (out, X, v) if want_xv else out

    def lsm_price(self, paths, r, K, T, dt, is_call, p):
        paths = _f64(paths)
        px = C.c_double()
        self._chk(self.L.ref_lsm_price(_ptr(paths, _dp), paths.shape[0], paths.shape[1], r, K, T, dt, int(is_call),
                                       p, C.byref(px)))
        return px.value

    def martingale_price(self, paths, r, K, T, dt, is_call, p, max_iter=5):
        paths = _f64(paths)
        px = C.c_double()
        self._chk(self.L.ref_martingale_price(_ptr(paths, _dp), paths.shape[0], paths.shape[1], r, K, T, dt,
                                              int(is_call), p, max_iter, C.byref(px)))
        return px.value

    def branching_price(self, paths, r, K, T, dt, is_call, num_branches, exercise_times, rp=None, want_used=False):
        """rp: injected resampling indices in the reference's consumption order [path][date with continuation][branch]."""
        paths = _f64(paths)
        ex = np.ascontiguousarray(exercise_times, dtype=np.int32)
        px, used = C.c_double(), C.c_long(0)
        q = np.ascontiguousarray(rp, dtype=np.int32).ravel() if rp is not None else None
        self._chk(self.L.ref_branching_price(_ptr(paths, _dp), paths.shape[0], paths.shape[1], r, K, T, dt,
                                             int(is_call), num_branches, _ptr(ex, _ip), ex.size, _ptr(q, _ip),
                                             q.size if q is not None else 0, C.byref(used), C.byref(px)))
        return (px.value, used.value) if want_used else px.value

    def asymptotic_price(self, paths, r, K, T, dt, is_call, sigma, dividend):
        paths = _f64(paths)
        px = C.c_double()
        self._chk(self.L.ref_asymptotic_price(_ptr(paths, _dp), paths.shape[0], paths.shape[1], r, K, T, dt,
                                              int(is_call), sigma, dividend, C.byref(px)))
        return px.value

    def bench_rows(self, hist, n_contracts, n_paths, n_steps, r, strike, is_call, p, threads=0):
        hist = _f64(hist)
        prices = np.zeros(n_contracts)
        sec, g, l = C.c_double(), C.c_double(), C.c_double()
        self._chk(self.L.ref_bench_rows(_ptr(hist, _dp), hist.size, n_contracts, n_paths, n_steps, r, strike,
                                        int(is_call), p, threads, _ptr(prices, _dp), C.byref(sec), C.byref(g),
                                        C.byref(l)))
        return dict(seconds=sec.value, gen_seconds_sum=g.value, lsm_seconds_sum=l.value, prices=prices)


_port = None
_ref = None


def port() -> _Port:
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref() -> _Ref:
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


# ------------------------------------------------------------------------------- numpy cross-checks
def np_rbergomi_paths(S0, r, xi, H, eta, rho, dt, n, draws):
    """numpy restatement -- an independent third implementation for cross-checks."""
    draws = _f64(draws)
    t = np.arange(n + 1) * dt
    lam = 0.5 * t ** (2 * H)
    M = 1
    while M < n + 1:
        M <<= 1
    Mp = 1
    while Mp < n:
        Mp <<= 1
    pad = np.zeros(M)
    pad[: n + 1] = lam
    phi = M * np.fft.ifft(pad)
    Z = draws[:, 0:2 * n:2] + 1j * draws[:, 1:2 * n:2]
    A = np.zeros((draws.shape[0], Mp), dtype=complex)
    A[:, :n] = phi[:n] * Z
    X = np.sqrt(2 * H) * eta * np.real(np.fft.fft(A, axis=1))[:, :n] / Mp
    v = xi * np.exp(X - 0.5 * eta ** 2 * t[:n] ** (2 * H))
    W1, W2 = draws[:, 2 * n:3 * n], draws[:, 3 * n:4 * n]
    dW = rho * (np.sqrt(dt) * W1) + np.sqrt(1 - rho ** 2) * (np.sqrt(dt) * W2)
    incr = (r - 0.5 * v) * dt + np.sqrt(np.maximum(0.0, v)) * dW
    S = np.empty((draws.shape[0], n + 1))
    S[:, 0] = S0
    S[:, 1:] = S0 * np.exp(np.cumsum(incr, axis=1))
    return S


def np_lsm(paths, r, K, T, dt, is_call, p):
    """numpy restatement of the reference LSM, lstsq with Eigen's rank rule."""
    paths = _f64(paths)
    N, M = paths.shape
    pay = (lambda S: np.maximum(0.0, S - K)) if is_call else (lambda S: np.maximum(0.0, K - S))
    disc = np.exp(-r * dt)
    V = pay(paths[:, M - 1])
    for j in range(M - 2, -1, -1):
        if j * dt > T:
            V = V * disc
            continue
        S = paths[:, j]
        im = pay(S)
        itm = im > 1e-14
        Vn = np.zeros(N)
        if itm.any():
            A = np.vander(S[itm], p + 1, increasing=True)
            c = np.linalg.lstsq(A, V[itm] * disc, rcond=min(A.shape) * np.finfo(float).eps)[0]
            Vn[itm] = np.maximum(im[itm], A @ c)
        otm = im < 1e-14
        Vn[otm] = V[otm] * disc
        V = Vn
    return float(V.mean())
