"""numpy statement of the "pair" native stream of the 256-point generator (csrc/gen_rbergomi_pair.cuh): one complex
transform drives two paths.  Test infrastructure only (used by test_stream_law.py on the CPU and test_gpu_paths.py)."""
import numpy as np


def spectrum(port, n, H, eta, dt):
    """phi_m scaled like the reference's X = sqrt(2H) eta Re(DFT(phi (.) Z)) / M' (RoughVolatility.cpp:264-292), and the
    symmetrised power spectrum w_m = (|phi_m|^2 [m<n] + |phi_{M'-m}|^2 [M'-m<n]) / 2."""
    Mp = 1
    while Mp < n:
        Mp <<= 1
    phis = port.rbergomi_phi(n, H, dt)[:n] * np.sqrt(2 * H) * eta / Mp
    c2 = np.zeros(Mp)
    c2[:n] = np.abs(phis) ** 2
    w = 0.5 * (c2 + c2[(-np.arange(Mp)) % Mp])
    return phis, w, Mp


def pair_X(G, w, n):
    """G [F][M'] standard complex normals -> X of the Re paths and of the Im paths, each [F][n]."""
    Y = np.fft.fft(np.sqrt(w) * G, axis=-1)
    return Y.real[:, :n], Y.imag[:, :n]


def equivalent_reference_draws(G, phis, w, n):
    """Z_A, Z_B [F][n] with Re DFT(phi (.) Z_A)[:n] = X_A and likewise B: frequencies m >= n fold onto M' - m."""
    Mp = w.size
    a = np.sqrt(w) * G
    uA = a[:, :n].copy()
    uB = -1j * a[:, :n]
    m = np.arange(1, Mp - n + 1)
    m = m[m < n]
    uA[:, m] += np.conj(a[:, Mp - m])
    uB[:, m] += 1j * np.conj(a[:, Mp - m])
    return uA / phis, uB / phis


def reference_X(Z, phis, Mp):
    """The reference's formula on n complex draws per path."""
    n = phis.size
    A = np.zeros((Z.shape[0], Mp), dtype=complex)
    A[:, :n] = phis * Z
    return np.fft.fft(A, axis=-1).real[:, :n]


def spec_G(port, seed, fids, Mp=256):
    """Stream spec: G_m of transform f = Box-Muller(x0, x1) (even m) / (x2, x3) (odd m) of Philox ctr (f_lo, f_hi, m>>1, 4)."""
    key = (seed & 0xFFFFFFFF, seed >> 32)
    G = np.empty((len(fids), Mp), dtype=complex)
    for i, f in enumerate(fids):
        for j in range(Mp // 2):
            x = port.philox((f & 0xFFFFFFFF, f >> 32, j, 4), key)
            z0, z1 = port.box_muller(x[0], x[1])
            G[i, 2 * j] = z0 + 1j * z1
            z0, z1 = port.box_muller(x[2], x[3])
            G[i, 2 * j + 1] = z0 + 1j * z1
    return G


def spec_W(port, seed, gids, n):
    """W_k of path g: lane (k>>4)&3 of Philox ctr (g_lo, g_hi, 4 (k&15) + (k>>6), 6); lanes = BM(x0,x1) then BM(x2,x3)."""
    key = (seed & 0xFFFFFFFF, seed >> 32)
    W = np.empty((len(gids), n))
    for i, g in enumerate(gids):
        for c in range(64):
            x = port.philox((g & 0xFFFFFFFF, g >> 32, c, 6), key)
            lanes = port.box_muller(x[0], x[1]) + port.box_muller(x[2], x[3])
            for t in range(4):
                k = (c >> 2) + 16 * (4 * (c & 3) + t)
                if k < n:
                    W[i, k] = lanes[t]
    return W


def spec_draws(port, seed, gids, n, H, eta, rho, dt):
    """Reference-order draws [len(gids)][4n] that replay path g of the pair stream through the reference / oracle."""
    phis, w, Mp = spectrum(port, n, H, eta, dt)
    assert Mp == 256
    gids = [int(g) for g in gids]
    fids = sorted({32 * (g >> 6) + (g & 31) for g in gids})
    G = spec_G(port, seed, fids, Mp)
    ZA, ZB = equivalent_reference_draws(G, phis, w, n)
    W = spec_W(port, seed, gids, n)
    d = np.empty((len(gids), 4 * n))
    for i, g in enumerate(gids):
        f = fids.index(32 * (g >> 6) + (g & 31))
        Z = ZB[f] if (g & 32) else ZA[f]
        d[i, 0:2 * n:2] = Z.real
        d[i, 1:2 * n:2] = Z.imag
        d[i, 2 * n:3 * n] = rho * W[i]
        d[i, 3 * n:] = np.sqrt(1 - rho * rho) * W[i]
    return d
