"""Short driver for ncu captures of the echo, covariance and channel-estimation kernels (cfg2 shapes)."""
import ctypes as C
import importlib
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
P = importlib.import_module(PKG)
_lib = P._lib
W = P.workloads


def csirs_row5_layout(n_rb, k0=1, l0=0, seed=0, L=14):
    """4-port CSI-RS row 5 (TS 38.211 Table 7.4.1.5.3-1): FD-CDM2, CDM group j at (k0 + {0,1}, l0 + j); QPSK base sequence."""
    K = 12 * n_rb
    rng = np.random.default_rng(seed)
    ind = np.zeros((2 * n_rb, 4), dtype=np.int64)
    sym = np.zeros((2 * n_rb, 4), dtype=np.complex128)
    for j in range(2):
        r = (rng.integers(0, 2, (n_rb, 2)) * 2 - 1 + 1j * (rng.integers(0, 2, (n_rb, 2)) * 2 - 1)) / np.sqrt(2)
        for q in range(2):
            k = (12 * np.arange(n_rb)[:, None] + k0 + np.arange(2)[None, :]).reshape(-1)
            ind[:, 2 * j + q] = 1 + k + K * (l0 + j) + K * L * (2 * j + q)
            sym[:, 2 * j + q] = (r * np.array([1.0, 1.0 if q == 0 else -1.0])[None, :]).reshape(-1)
    return ind, sym, (2, 1)


cell, car, wave = W.cell_config("cfg2")
rp = P.sensing.radarParams(cell, car, wave)
cf = P.sensing.detection.cfar2D(rp)
grid, txw = W.sensing_tx("cfg2", 1)
nSc, nSym, nTx = grid.shape
T = txw.shape[0]
echo = importlib.import_module(PKG + ".sensing._echo")
est = importlib.import_module(PKG + ".sensing.estimation")
ctx = _lib.get_context(0)
tx_d = torch.from_numpy(np.ascontiguousarray(txw.astype(np.complex64).T)).cuda()
g_d = torch.from_numpy(np.ascontiguousarray(grid.astype(np.complex64).transpose(2, 1, 0))).cuda()[None].contiguous()
rx_d = torch.empty_like(g_d)
eargs = echo._EchoArgs(T, nTx, rp, cell["targetLoSConditions"], car, nSym)
plan = est.SensePlan(rp, cf, (nSc, nSym, nTx), max_batch=1, device=0)
n = C.c_int32()
ctx.use_torch_stream()
for i in range(3):
    _lib.check(ctx.lib.isac_mono_static_sensing_dev(ctx.handle, C.byref(eargs.cfg), _lib.ptr(tx_d), None, _lib.NOISE_PHILOX, i,
                                                    _lib.ptr(rx_d[0]), C.byref(n)), ctx.handle)
    plan.run_dev(rx_d, g_d, 1)
# channel estimation: 4-port CSI-RS row 5, 273 PRB, 8 rx antennas, 32 UEs
ph = P.communication.phyLayer
ind, sym, cdm = csirs_row5_layout(273, 1, 0, seed=1)
ce = ph.ChannelEstimator(3276, 14, 8, 4, ind, sym, cdm, max_batch=32)
rxg = torch.view_as_complex(torch.randn(32, 8, 14, 3276, 2, device="cuda"))
for i in range(3):
    ce.run_dev(rxg, 32)
torch.cuda.synchronize()
