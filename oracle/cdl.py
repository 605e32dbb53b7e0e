"""Oracle for K11 (CDL channel): float64 NumPy restatement (test infrastructure only; PARITY UNPINNED).

The reference uses the closed-source toolbox object nrCDLChannel (configured at
+parameters/+channelModels/+communication/cdl.m:48-88) followed by nrChannelEstimate; neither is in the
repository.  This module restates TR 38.901 7.7.1 / 7.5 (fixed cluster angles of Tables 7.7.1-1/-3/-4, ray offsets of
Table 7.5-3, XPR matrix, 38.901 element pattern, polarisation model 2, array phases) and the frequency-domain
synthesis H[k,l,u,s] = sum_n e^{-2 pi j f_k tau_n} sum_{m in n} g_m[u,s] e^{2 pi j nu_m t_l}.  Ray coupling and
initial phases come from splitmix64 (documented; MATLAB's mt19937 stream cannot be reproduced), so parity with the
toolbox is statistical only; parity between this file and csrc/cdl.cu is exact to rounding.
"""
from __future__ import annotations

import numpy as np

MASK = (1 << 64) - 1

ALPHA = np.array([0.0447, -0.0447, 0.1413, -0.1413, 0.2492, -0.2492, 0.3715, -0.3715, 0.5129, -0.5129,
                  0.6797, -0.6797, 0.8844, -0.8844, 1.1481, -1.1481, 1.5195, -1.5195, 2.1551, -2.1551])

# TR 38.901 Tables 7.7.1-1 (CDL-A), 7.7.1-3 (CDL-C), 7.7.1-4 (CDL-D): delay, power dB, AOD, AOA, ZOD, ZOA
CDL_A = np.array([
    [0.0000, -13.4, -178.1, 51.3, 50.2, 125.4], [0.3819, 0, -4.2, -152.7, 93.2, 91.3], [0.4025, -2.2, -4.2, -152.7, 93.2, 91.3],
    [0.5868, -4, -4.2, -152.7, 93.2, 91.3], [0.4610, -6, 90.2, 76.6, 122, 94], [0.5375, -8.2, 90.2, 76.6, 122, 94],
    [0.6708, -9.9, 90.2, 76.6, 122, 94], [0.5750, -10.5, 121.5, -1.8, 150.2, 47.1], [0.7618, -7.5, -81.7, -41.9, 55.2, 56],
    [1.5375, -15.9, 158.4, 94.2, 26.4, 30.1], [1.8978, -6.6, -83, 51.9, 126.4, 58.8], [2.2242, -16.7, 134.8, -115.9, 171.6, 26],
    [2.1718, -12.4, -153, 26.6, 151.4, 49.2], [2.4942, -15.2, -172, 76.6, 157.2, 143.1], [2.5119, -10.8, -129.9, -7, 47.2, 117.4],
    [3.0582, -11.3, -136, -23, 40.4, 122.7], [4.0810, -12.7, 165.4, -47.2, 43.3, 123.2], [4.4579, -16.2, 148.4, 110.4, 161.8, 32.6],
    [4.5695, -18.3, 132.7, 144.5, 10.8, 27.2], [4.7966, -18.9, -118.6, 155.3, 16.7, 15.2], [5.0066, -16.6, -154.1, 102, 171.7, 146],
    [5.3043, -19.9, 126.5, -151.8, 22.7, 150.7], [9.6586, -29.7, -56.2, 55.2, 144.9, 156.1]])
CDL_C = np.array([
    [0, -4.4, -46.6, -101, 97.2, 87.6], [0.2099, -1.2, -22.8, 120, 98.6, 72.1], [0.2219, -3.5, -22.8, 120, 98.6, 72.1],
    [0.2329, -5.2, -22.8, 120, 98.6, 72.1], [0.2176, -2.5, -40.7, -127.5, 100.6, 70.1], [0.6366, 0, 0.3, 170.4, 99.2, 75.3],
    [0.6448, -2.2, 0.3, 170.4, 99.2, 75.3], [0.6560, -3.9, 0.3, 170.4, 99.2, 75.3], [0.6584, -7.4, 73.1, 55.4, 105.2, 67.4],
    [0.7935, -7.1, -64.5, 66.5, 95.3, 63.8], [0.8213, -10.7, 80.2, -48.1, 106.1, 71.4], [0.9336, -11.1, -97.1, 46.9, 93.5, 60.5],
    [1.2285, -5.1, -55.3, 68.1, 103.7, 90.6], [1.3083, -6.8, -64.3, -68.7, 104.2, 60.1], [2.1704, -8.7, -78.5, 81.5, 93.0, 61.0],
    [2.7105, -13.2, 102.7, 30.7, 104.2, 100.7], [4.2589, -13.9, 99.2, -16.4, 94.9, 62.3], [4.6003, -13.9, 88.8, 3.8, 93.1, 66.7],
    [5.4902, -15.8, -101.9, -13.7, 92.2, 52.9], [5.6077, -17.1, 92.2, 9.7, 106.7, 61.8], [6.3065, -16, 93.3, 5.6, 93.0, 51.9],
    [6.6374, -15.7, 106.6, 0.7, 92.9, 61.7], [7.0427, -21.6, 119.5, -21.9, 105.2, 58], [8.6523, -22.8, -123.8, 33.6, 107.8, 57]])
CDL_D = np.array([
    [0, -13.5, 0, -180, 98.5, 81.5], [0.035, -18.8, 89.2, 89.2, 85.5, 86.9], [0.612, -21, 89.2, 89.2, 85.5, 86.9],
    [1.363, -22.8, 89.2, 89.2, 85.5, 86.9], [1.405, -17.9, 13, 163, 97.5, 79.4], [1.804, -20.1, 13, 163, 97.5, 79.4],
    [2.596, -21.9, 13, 163, 97.5, 79.4], [1.775, -22.9, 34.6, -137, 98.5, 78.2], [4.042, -27.8, -64.5, 74.5, 88.4, 73.6],
    [7.937, -23.6, -32.9, 127.7, 91.3, 78.3], [9.424, -24.8, 52.6, -119.6, 103.8, 87], [9.708, -30.0, -132.1, -9.1, 80.3, 70.6],
    [12.525, -27.7, 77.2, -83.8, 86.5, 72.9]])
# TR 38.901 Tables 7.7.1-2 (CDL-B) and 7.7.1-5 (CDL-E; first row = Rayleigh part of the LOS cluster, specular path -0.03 dB)
CDL_B = np.array([
    [0.0000, 0, 9.3, -173.3, 105.8, 78.9], [0.1072, -2.2, 9.3, -173.3, 105.8, 78.9], [0.2155, -4, 9.3, -173.3, 105.8, 78.9],
    [0.2095, -3.2, -34.1, 125.5, 115.3, 63.3], [0.2870, -9.8, -65.4, -88.0, 119.3, 59.9], [0.2986, -1.2, -11.4, 155.1, 103.2, 67.5],
    [0.3752, -3.4, -11.4, 155.1, 103.2, 67.5], [0.5055, -5.2, -11.4, 155.1, 103.2, 67.5], [0.3681, -7.6, -67.2, -89.8, 118.2, 82.6],
    [0.3697, -3, 52.5, 132.1, 102.0, 66.3], [0.5700, -8.9, -72, -83.6, 100.4, 61.6], [0.5283, -9, 74.3, 95.3, 98.3, 58.0],
    [1.1021, -4.8, -52.2, 103.7, 103.4, 78.2], [1.2756, -5.7, -50.5, -87.8, 102.5, 82.0], [1.5474, -7.5, 61.4, -92.5, 101.4, 62.4],
    [1.7842, -1.9, 30.6, -139.1, 103.0, 78.0], [2.0169, -7.6, -72.5, -90.6, 100.0, 60.9], [2.8294, -12.2, -90.6, 58.6, 115.2, 82.9],
    [3.0219, -9.8, -77.6, -79.0, 100.5, 60.8], [3.6187, -11.4, -82.6, 65.8, 119.6, 57.3], [4.1067, -14.9, -103.6, 52.7, 118.7, 59.9],
    [4.2790, -9.2, 75.6, 88.7, 117.8, 60.1], [4.7834, -11.3, -77.6, -60.4, 115.7, 62.3]])
CDL_E = np.array([
    [0, -22.03, 0, -180, 99.6, 80.4], [0.5133, -15.8, 57.5, 18.2, 104.2, 80.4], [0.5440, -18.1, 57.5, 18.2, 104.2, 80.4],
    [0.5630, -19.8, 57.5, 18.2, 104.2, 80.4], [0.5440, -22.9, -20.1, 101.8, 99.4, 80.8], [0.7112, -22.4, 16.2, 112.9, 100.8, 86.3],
    [1.9092, -18.6, 9.3, -155.5, 98.8, 82.7], [1.9293, -20.8, 9.3, -155.5, 98.8, 82.7], [1.9589, -22.6, 9.3, -155.5, 98.8, 82.7],
    [2.6426, -22.3, 19, -143.3, 100.8, 82.9], [3.7136, -25.6, 32.7, -94.7, 96.4, 88], [5.4524, -20.2, 0.5, 147, 98.9, 81],
    [12.0034, -29.8, 55.9, -36.2, 95.6, 88.6], [20.6519, -29.2, 57.6, -26, 104.6, 78.3]])
PROFILES = {0: (CDL_A, (5, 11, 3, 3), 10.0, None), 1: (CDL_B, (10, 22, 3, 7), 8.0, None), 2: (CDL_C, (2, 15, 3, 7), 7.0, None),
            3: (CDL_D, (5, 8, 3, 3), 11.0, -0.2), 4: (CDL_E, (5, 11, 3, 7), 8.0, -0.03)}


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def uniform(self):
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)


def pattern38901(theta, phi):
    ph = (phi + 180.0) % 360.0 - 180.0
    av = -min(12.0 * ((theta - 90.0) / 65.0) ** 2, 30.0)
    ah = -min(12.0 * (ph / 65.0) ** 2, 30.0)
    return 10.0 ** ((-min(-(av + ah), 30.0) + 8.0) / 10.0)


def build_rays(profile, delay_spread, max_doppler, tx_size, rx_size, tx_pat, rx_pat, seed):
    """-> dict(tau[nCl], power[nCl], nu[nRays], cluster[nRays], g[nRays, nRx, nTx])."""
    rows, (casd, casa, czsd, czsa), xpr_db, los_db = PROFILES[profile]
    n_cl, M = rows.shape[0], 20
    n_tx, n_rx = int(np.prod(tx_size)), int(np.prod(rx_size))
    pw = 10.0 ** (rows[:, 1] / 10.0)
    plos = 10.0 ** (los_db / 10.0) if los_db is not None else 0.0
    tot = pw.sum() + plos
    pw, plos = pw / tot, plos / tot
    kappa = 10.0 ** (xpr_db / 10.0)
    deg = np.pi / 180.0

    def elem(size, e):
        m, n = e % size[0], (e // size[0]) % size[1]
        return 0.5 * n, 0.5 * m, e // (size[0] * size[1])

    def field(pat, npol, pol, th, ph):
        a = np.sqrt(pattern38901(th, ph)) if pat else 1.0
        zeta = (45.0 if pol == 0 else -45.0) if npol == 2 else 0.0
        return a * np.cos(zeta * deg), a * np.sin(zeta * deg)

    n_rays = n_cl * M + (1 if los_db is not None else 0)
    g = np.zeros((n_rays, n_rx, n_tx), complex)
    nu = np.zeros(n_rays)
    cl = np.zeros(n_rays, dtype=int)
    rng = SplitMix64(seed)

    def ray(idx, c, amp, aod, aoa, zod, zoa, X):
        cl[idx] = c
        nu[idx] = max_doppler * np.sin(zoa * deg) * np.cos(aoa * deg)
        rxv = (np.sin(zoa * deg) * np.cos(aoa * deg), np.sin(zoa * deg) * np.sin(aoa * deg), np.cos(zoa * deg))
        txv = (np.sin(zod * deg) * np.cos(aod * deg), np.sin(zod * deg) * np.sin(aod * deg), np.cos(zod * deg))
        for u in range(n_rx):
            yu, zu, pu = elem(rx_size, u)
            frt, frp = field(rx_pat, rx_size[2], pu, zoa, aoa)
            phr = 2 * np.pi * (rxv[1] * yu + rxv[2] * zu)
            for s in range(n_tx):
                ys, zs, ps = elem(tx_size, s)
                ftt, ftp = field(tx_pat, tx_size[2], ps, zod, aod)
                pht = 2 * np.pi * (txv[1] * ys + txv[2] * zs)
                pol = frt * (X[0] * ftt + X[1] * ftp) + frp * (X[2] * ftt + X[3] * ftp)
                g[idx, u, s] = amp * pol * np.exp(1j * (phr + pht)) / np.sqrt(n_rx)

    for n in range(n_cl):
        perm = []
        for _ in range(3):
            p = list(range(M))
            for m in range(M - 1, 0, -1):
                j = int(rng.uniform() * (m + 1))
                p[m], p[j] = p[j], p[m]
            perm.append(p)
        for m in range(M):
            X = [np.exp(1j * (2.0 * rng.uniform() - 1.0) * np.pi) for _ in range(4)]
            X[1] *= np.sqrt(1.0 / kappa)
            X[2] *= np.sqrt(1.0 / kappa)
            r = rows[n]
            ray(n * M + m, n, np.sqrt(pw[n] / M), r[2] + casd * ALPHA[m], r[3] + casa * ALPHA[perm[0][m]],
                r[4] + czsd * ALPHA[perm[1][m]], r[5] + czsa * ALPHA[perm[2][m]], X)
    if los_db is not None:
        r = rows[0]
        ray(n_cl * M, 0, np.sqrt(plos), r[2], r[3], r[4], r[5], [1.0, 0.0, 0.0, -1.0])
    return {"tau": rows[:, 0] * delay_spread, "power": pw, "plos": plos, "nu": nu, "cluster": cl, "g": g}


def frequency_response(rays, K, scs_hz, t):
    """H[K, L, nRx, nTx] from ray tables (float64)."""
    tau, nu, cl, g = rays["tau"], rays["nu"], rays["cluster"], rays["g"]
    t = np.asarray(t, float)
    n_cl = tau.size
    C = np.zeros((n_cl, t.size) + g.shape[1:], complex)
    for m in range(nu.size):
        C[cl[m]] += np.exp(2j * np.pi * nu[m] * t)[:, None, None] * g[m][None]
    f = (np.arange(K) - K // 2) * scs_hz
    E = np.exp(-2j * np.pi * f[:, None] * tau[None, :])
    return np.einsum("kn,nlus->klus", E, C)
