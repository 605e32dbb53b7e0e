"""Developer timings of the SURVEY 8(f) "next" rows (CUDA events on torch's current stream, medians; not the contract bench):
LoS / blockage decisions over the cached OSM city, MVDR / beamscan / MUSIC DoA scans, channel estimation, OFDM modulation."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
P = importlib.import_module(PKG)
_lib = P._lib
W = P.workloads
ctx = _lib.get_context(0)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3   # us


print(torch.cuda.get_device_name(0))
# ---- row 4: LoS decisions, config-5 scale: 19 gNB x (100 UE + 20 targets) links over the cached city (666 walls) ----------
blk = importlib.import_module(PKG + ".networkTopology.blockages")
z = np.load(os.path.join("tests", "golden", "osm_city.npz"))
off = z["fp_off"]
buildings = [(z["fp_flat"][:, off[i]:off[i + 1]], float(z["heights"][i])) for i in range(off.size - 1)]
city = blk.city(buildings)
rng = np.random.default_rng(0)
n = 19 * 120
lo, hi = z["ue"].min(axis=0), z["ue"].max(axis=0)
ue = rng.uniform(lo, hi, (n, 3))
ant = rng.uniform(lo, hi, (n, 3))
ue_d = torch.from_numpy(np.ascontiguousarray(ue.T.copy().T)).cuda()      # [n][3] == MATLAB [3 x n]
ant_d = torch.from_numpy(np.ascontiguousarray(ant)).cuda()
out_d = torch.empty(n, dtype=torch.int32, device="cuda")
ctx.use_torch_stream()
f = lambda: _lib.check(ctx.lib.isac_city_check_los_dev(city.handle, n, _lib.ptr(ue_d), _lib.ptr(ant_d), n, _lib.ptr(out_d)), ctx.handle)
t = timed(f)
print(f"LoS: {n} links x {city.nWalls} walls: {t:.1f} us per batch = {n * city.nWalls / t * 1e-3:.1f} G link-wall tests/s")
# ---- row 3: DoA scanners on a 16-element ULA covariance (host entry: includes the H2D of Ra and the D2H of the spectrum) ----
doa = P.sensing.estimation.doaEstimation
rp = {"antennaType": {"type": "ula", "nV": 8, "p": 2, "d": 0.5}, "azimuthScanScale": 360, "azimuthScanGranularity": 1,
      "elevationScanScale": 180, "elevationScanGranularity": 1}
X = rng.standard_normal((16, 500)) + 1j * rng.standard_normal((16, 500))
Ra = X @ X.conj().T / 500
for name in ("music", "mvdrBF", "digitalBF"):
    fn = getattr(doa, name)
    t0 = time.time()
    for _ in range(20):
        fn(3, rp, Ra)
    print(f"DoA {name}: {(time.time() - t0) / 20 * 1e6:.0f} us per call (host wall, 16-element ULA, 361-point scan)")
# ---- row 1: channel estimation, 32 UEs, 4-port CSI-RS row 5, 273 PRB, 8 rx ------------------------------------------------
def csirs_row5_layout(n_rb, k0=1, l0=0, seed=0, L=14):
    """4-port CSI-RS row 5 (TS 38.211 Table 7.4.1.5.3-1): FD-CDM2, CDM group j at (k0 + {0,1}, l0 + j); QPSK base sequence."""
    K = 12 * n_rb
    g = np.random.default_rng(seed)
    ind = np.zeros((2 * n_rb, 4), dtype=np.int64)
    sym = np.zeros((2 * n_rb, 4), dtype=np.complex128)
    for j in range(2):
        r = (g.integers(0, 2, (n_rb, 2)) * 2 - 1 + 1j * (g.integers(0, 2, (n_rb, 2)) * 2 - 1)) / np.sqrt(2)
        for q in range(2):
            k = (12 * np.arange(n_rb)[:, None] + k0 + np.arange(2)[None, :]).reshape(-1)
            ind[:, 2 * j + q] = 1 + k + K * (l0 + j) + K * L * (2 * j + q)
            sym[:, 2 * j + q] = (r * np.array([1.0, 1.0 if q == 0 else -1.0])[None, :]).reshape(-1)
    return ind, sym, (2, 1)


ph = P.communication.phyLayer
ind, sym, cdm = csirs_row5_layout(273, 1, 0, seed=1)
ce = ph.ChannelEstimator(3276, 14, 8, 4, ind, sym, cdm, max_batch=32)
rxg = torch.view_as_complex(torch.randn(32, 8, 14, 3276, 2, device="cuda"))
Hout = torch.empty((32, 4, 8, 14, 3276), dtype=torch.complex64, device="cuda")
t = timed(lambda: ce.run_dev(rxg, 32, Hout, sync=False))
by = 32 * 3276 * 14 * 8 * 4 * 8
print(f"channel estimate: 32 UEs -> Hest {by / 1e6:.0f} MB in {t:.1f} us = {by / t * 1e-3:.0f} GB/s written")
# ---- row 2: OFDM modulation of one cfg2 cell frame ---------------------------------------------------------------------------
cell, car, wave = W.cell_config("cfg2")
grid, txw = W.sensing_tx("cfg2", 1)
nSc, nSym, nTx = grid.shape
num = W.ofdm_numerology(int(car["NRBsDL"]), float(car["SubcarrierSpacing"]))
cp = np.ascontiguousarray(num["CyclicPrefixLengths"], dtype=np.int32)
g_d = torch.from_numpy(np.ascontiguousarray(grid.astype(np.complex64).transpose(2, 1, 0))).cuda()
w_d = torch.empty((nTx, txw.shape[0]), dtype=torch.complex64, device="cuda")
T = C.c_int64()
f = lambda: _lib.check(ctx.lib.isac_ofdm_modulate_dev(ctx.handle, _lib.ptr(g_d), nSc, nSym, nTx, int(num["Nfft"]), int(cp.size),
                                                      cp.ctypes.data, 1.0, _lib.ptr(w_d), C.byref(T)), ctx.handle)
t = timed(f)
by = 8 * nSc * nSym * nTx + 8 * txw.shape[0] * nTx
print(f"OFDM modulate: {by / 1e6:.1f} MB algorithmic in {t:.1f} us = {by / t * 1e-3:.0f} GB/s")
