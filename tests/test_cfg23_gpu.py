"""BASELINE configs 2 and 3 at full size for the COMM half.

cfg2: 8 UEs, 8 CSI-RS ports (2,2), 8 receive antennas, 273 PRB, 16-PRB subbands: the fused RI/PMI/CQI report of a whole cell in
      one batch against the vectorised float64 oracle (indices exact).
cfg3: 32 CSI-RS ports (4,4) -- the 64-element array virtualised onto the largest Type-I port count (dlPMISelect.m:623-626) --
      273 PRB, 4 UEs of a cell: the per-PRB SINR grid (SINRPerRE at the 273 CSI-RS REs x layers x candidates) within 1e-5 relative
      and the PMI it selects, for ranks 1 and 2.
The loop-faithful oracle would take minutes at these sizes; the vectorised restatement (same arithmetic through stacked
inverses, oracle/comm.py::sinr_per_re_vectorized) is checked against it at small sizes in tests/test_oracle_cpu.py."""
import importlib

import numpy as np
import pytest

from oracle import comm as C

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu
TABLE = np.array([-3.46, 1.54, 6.54, 11.05, 13.54, 16.04, 17.54, 20.04, 22.04, 24.43, 26.93, 27.43, 29.43, 32.43, 35.43])


def _channel(rng, K, R, P, taps=4):
    k = np.arange(K)[:, None, None]
    H = sum((rng.standard_normal((R, P)) + 1j * rng.standard_normal((R, P)))[None] * np.exp(-2j * np.pi * k * t * 7 / 4096.0) / (1 + t)
            for t in range(taps))
    return np.repeat(H[:, None], 14, axis=1).astype(np.complex64)          # [K, 14, R, P]


def _cfg(n_ports, panel, nrb, sb):
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": n_ports, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"NSizeBWP": nrb, "NStartBWP": 0, "PanelDimensions": panel, "CodebookMode": 1, "PMIMode": "Subband", "CQIMode": "Subband",
          "SubbandSize": sb}
    ocfg = C.report_config(n_ports, panel, nrb, 0, 1, "Subband", "Subband", sb)
    re_k, re_l = C.csirs_first_port_res(nrb, 1, 0)
    return carrier, csirs, rc, ocfg, re_k, re_l


def test_cfg2_cell_report_273prb_8x8(gpu):
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    carrier, csirs, rc, ocfg, re_k, re_l = _cfg(8, (2, 2), 273, 16)
    rng = np.random.default_rng(202)
    B = 8
    Hs = [_channel(rng, 273 * 12, 8, 8) for _ in range(B)]
    nvar = 10 ** (-np.array([5.0, 10.0, 15.0, 20.0, 25.0, 8.0, 12.0, 18.0]) / 10)
    RI, pm, cqi = ph.csiReport(carrier, csirs, rc, np.stack(Hs, axis=-1), nvar, TABLE, rankCap=4)
    for b in range(B):
        rank, pmo, cqo = C.csi_report_vectorized(ocfg, re_k, re_l, Hs[b], nvar[b], TABLE, rank_cap=4)
        assert RI[b] == rank, (b, RI[b], rank)
        assert np.array_equal(pm["i1"][:, b], pmo["i1"]), (b, pm["i1"][:, b], pmo["i1"])
        assert np.array_equal(pm["i2"][:, b], pmo["i2"], equal_nan=True), b
        d = cqo[1:] - cqo[:1]                      # the vectorised oracle returns absolute subband CQIs: map them to the
        off = np.where(np.isnan(d), np.nan, np.where(d == 0, 0, np.where(d == 1, 1, np.where(d >= 2, 2, 3))))   # differential
        cqo = np.vstack([cqo[:1], off])            # report format of cqiSelect.m:656-677 (TS 38.214 Table 5.2.2.1-1)
        assert np.array_equal(cqi[:, : cqo.shape[1], b], cqo, equal_nan=True), (b, cqi[:, 0, b], cqo[:, 0])
    print("cfg2 report: RI", RI, "wideband CQI", cqi[0, 0, :])


@pytest.mark.parametrize("nu", [1, 2])
def test_cfg3_sinr_grid_32_ports_273prb(gpu, nu):
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    carrier, csirs, rc, ocfg, re_k, re_l = _cfg(32, (4, 4), 273, 16)
    rng = np.random.default_rng(300 + nu)
    B, R = 4, 4
    Hs = [_channel(rng, 273 * 12, R, 32) for _ in range(B)]
    nvar = np.array([0.1, 0.03, 0.3, 0.01])
    pm, info = ph.dlPMISelect(carrier, csirs, rc, nu, np.stack(Hs, axis=-1), nvar)
    for b in range(B):
        So, _ = C.sinr_per_re_vectorized(ocfg, re_k, re_l, nu, Hs[b], nvar[b])
        Sg = info["SINRPerRE"][..., b]
        assert So.shape == Sg.shape and np.array_equal(np.isnan(So), np.isnan(Sg))
        m = ~np.isnan(So)
        err = (np.abs(Sg[m] - So[m]) / np.abs(So[m])).max()
        print(f"cfg3 UE {b} nu={nu}: per-PRB SINR grid {Sg.shape}, max rel err {err:.2e}")
        assert err <= 1e-5
        total = C.matlab_round4(np.nansum(So, axis=(0, 1)))                       # dlPMISelect.m:444-449
        i1 = pm["i1"][:, b].astype(int) - 1
        assert np.isclose(total[:, i1[0], i1[1], i1[2]].max(), total.max(), rtol=0, atol=1e-4)
