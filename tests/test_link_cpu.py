"""Link-budget oracle (oracle/link.py) against hand-derived known answers from TR 38.901 Table 7.4.1-1, and the host-only
entry points of the library (thermal noise power, DFT fallback matrix) against the oracle -- no GPU needed."""
import importlib
import math

import numpy as np

from oracle import link as L

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"


def test_uma_los_below_and_above_breakpoint():
    fc, bs = 3.5e9, [0.0, 0.0, 25.0]
    d_bp = 4 * 24.0 * 0.5 * fc / L.LIGHTSPEED                         # h' = h - 1 m
    ue = [100.0, 0.0, 1.5]
    d3 = math.hypot(100.0, 23.5)
    assert 100.0 < d_bp
    assert math.isclose(L.path_loss("UMa", fc, 1, bs, ue), 28.0 + 22 * math.log10(d3) + 20 * math.log10(3.5), rel_tol=1e-12)
    ue = [900.0, 0.0, 1.5]
    d3 = math.hypot(900.0, 23.5)
    ref = 28.0 + 40 * math.log10(d3) + 20 * math.log10(3.5) - 9 * math.log10(d_bp ** 2 + 23.5 ** 2)
    assert 900.0 > d_bp and math.isclose(L.path_loss("UMa", fc, 1, bs, ue), ref, rel_tol=1e-12)


def test_nlos_is_never_below_los_and_matches_formula():
    fc, bs, ue = 3.5e9, [0.0, 0.0, 25.0], [300.0, 40.0, 1.5]
    d3 = float(np.linalg.norm(np.subtract(ue, bs)))
    for scn in ("UMa", "UMi", "RMa", "InH"):
        assert L.path_loss(scn, fc, 0, bs, ue) >= L.path_loss(scn, fc, 1, bs, ue)
    assert math.isclose(L.path_loss("UMa", fc, 0, bs, ue), 13.54 + 39.08 * math.log10(d3) + 20 * math.log10(3.5), rel_tol=1e-12)
    assert math.isclose(L.path_loss("UMi", fc, 0, bs, ue), 35.3 * math.log10(d3) + 22.4 + 21.3 * math.log10(3.5), rel_tol=1e-12)


def test_fspl_and_identical_positions():
    fc = 3.5e9
    assert math.isclose(L.path_loss("fspl", fc, 1, [0, 0, 0], [1000.0, 0, 0]), 20 * math.log10(4 * math.pi * 1000 * fc / L.LIGHTSPEED))
    assert L.path_loss("fspl", fc, 1, [0, 0, 0], [1e-3, 0, 0]) == 0.0                 # fspl clips at 0 dB
    assert L.path_loss("UMa", fc, 1, [1, 2, 3], [1, 2, 3]) == 0.0                     # config5GNRModels.m:32-33


def test_thermal_noise_and_dft_matrix_entry_points():
    pl = importlib.import_module(PKG + ".communication.pathlossModels")
    nt = pl.thermalNoisePower(6.0, 290.0, 122.88e6)
    assert math.isclose(nt, L.thermal_noise_power(6.0, 290.0, 122.88e6), rel_tol=1e-14)
    assert math.isclose(nt, 1.380649e-23 * 290 * 10 ** 0.6 * 122.88e6, rel_tol=1e-12)   # T = 290 K: k T NF fs
    for n_tx, n_rx in ((4, 2), (2, 4), (8, 8), (16, 2), (1, 1)):
        H = pl.dftChannelMatrix(n_tx, n_rx)
        ref = L.dft_channel_matrix(n_tx, n_rx)
        assert H.shape == ref.shape and np.abs(H - ref).max() < 1e-14
        assert math.isclose(np.linalg.norm(H, 2), 1.0, rel_tol=1e-12)                 # H / norm(H) (uePhy.m:738)
