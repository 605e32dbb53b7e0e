"""Size-independent properties of the CUDA path at BASELINE config 2's full size (no oracle run needed at these sizes):

  * echo synthesis is linear in the target set (basicRadarChannel.m:64 sums the per-target echoes): grid(A u B) = grid(A) + grid(B)
  * a Doppler phase ramp on the receive grid whose frequency is a whole number of Doppler bins rotates the range-Doppler
    power map along the Doppler axis (fft2D.m:44-46)
  * CA-CFAR decisions do not change when the power map is scaled by a power of two (threshold and cell scale together, exactly)
  * LMMSE SINRs are invariant under a unitary transform of the receive antennas, so SINRPerRE and the report are unchanged
    (precodedSINR.m:14-17: only H'H enters)
Tolerances: 1e-5 relative for fp32 arrays, 1e-9 for the float64 SINR path, indices exact."""
import importlib

import numpy as np
import pytest

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(gpu):
    return importlib.import_module(PKG)


def _cfg2(P):
    W = P.workloads
    cell, car, wave = W.cell_config("cfg2")
    rp = P.sensing.radarParams(cell, car, wave)
    grid, txw = W.sensing_tx("cfg2", 3)
    return W, cell, car, wave, rp, grid.astype(np.complex64), txw.astype(np.complex64)


def test_echo_is_linear_in_the_target_set(P):
    W, cell, car, wave, rp, grid, txw = _cfg2(P)
    n = len(cell["targetLoSConditions"])
    assert n == 4
    full = P.sensing.monoStaticSensing(txw, grid.shape, car, rp, np.ones(n, dtype=int))
    a = P.sensing.monoStaticSensing(txw, grid.shape, car, rp, np.array([1, 0, 1, 0]))      # NLoS targets reflect nothing
    b = P.sensing.monoStaticSensing(txw, grid.shape, car, rp, np.array([0, 1, 0, 1]))      # (basicRadarChannel.m:57-58)
    rms = np.sqrt(np.mean(np.abs(full) ** 2))
    err = np.abs(a + b - full).max() / rms
    print("echo linearity err / rms", err)
    assert rms > 0 and err <= 1e-5


def test_doppler_ramp_rotates_the_map(P):
    import torch
    W, cell, car, wave, rp, grid, txw = _cfg2(P)
    cf = P.sensing.detection.cfar2D(rp)
    rdm = importlib.import_module(PKG + ".sensing._rdm")
    nSc, nSym, nA = grid.shape
    F, q0 = rp["nFFT"], 32
    assert nSym == 168 and F == 256 and (q0 * nSym) % F == 0          # the ramp closes over the rotated symbol axis
    rx = P.sensing.monoStaticSensing(txw, grid.shape, car, rp, cell["targetLoSConditions"], seed=5)
    ramp = np.exp(2j * np.pi * q0 * np.arange(nSym) / F).astype(np.complex64)
    rx2 = (rx * ramp[None, :, None]).astype(np.complex64)
    plan = rdm.RangeDopplerPlan(nSc, nSym, nA, rp["nIFFT"], F, cf["rngIdx"], cf["dopIdx"], rp["Pfa"])
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda()
    plan.run_dev(to_dev(rx), to_dev(grid), 1)
    P0 = plan.power(1)[..., 0].astype(np.float64)                       # [nIFFT, nFFT, nAnts]
    plan.run_dev(to_dev(rx2), to_dev(grid), 1)
    P1 = plan.power(1)[..., 0].astype(np.float64)
    err = np.abs(P1 - np.roll(P0, q0, axis=1)).max() / P0.max()
    print("Doppler rotation err / peak", err)
    assert err <= 1e-5
    plan.close()


def test_cfar_is_scale_invariant(P):
    import torch
    W, cell, car, wave, rp, grid, txw = _cfg2(P)
    cf = P.sensing.detection.cfar2D(rp)
    rdm = importlib.import_module(PKG + ".sensing._rdm")
    nSc, nSym, nA = grid.shape
    rx = P.sensing.monoStaticSensing(txw, grid.shape, car, rp, cell["targetLoSConditions"], seed=9)
    plan = rdm.RangeDopplerPlan(nSc, nSym, nA, rp["nIFFT"], rp["nFFT"], cf["rngIdx"], cf["dopIdx"], rp["Pfa"])
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda()
    pw = torch.empty((1, nA, rp["nFFT"], rp["nIFFT"]), dtype=torch.float32, device="cuda")
    plan.run_dev(to_dev(rx), to_dev(grid), 1, pw)
    cnt0, det0 = plan.detections(1)
    assert cnt0.sum() > 0
    for scale in (4.0, 2.0 ** -20):
        scaled = pw * scale                                   # exact in fp32 (power of two)
        plan.cfar_dev(scaled, 1)
        cnt, det = plan.detections(1)
        assert np.array_equal(cnt, cnt0)
        for r in range(nA):
            assert np.array_equal(det[0][r][0], det0[0][r][0])
            assert np.array_equal(det[0][r][1], (det0[0][r][1] * np.float32(scale)).astype(np.float32))
    plan.close()


def test_sinr_is_invariant_under_receive_side_unitaries(P):
    ph = P.communication.phyLayer
    rng = np.random.default_rng(12)
    nrb, R, Pn = 273, 8, 8
    K = 12 * nrb
    k = np.arange(K)[:, None, None]
    H = sum((rng.standard_normal((R, Pn)) + 1j * rng.standard_normal((R, Pn)))[None] * np.exp(-2j * np.pi * k * t * 5 / 4096.0) / (1 + t)
            for t in range(4))
    Q, _ = np.linalg.qr(rng.standard_normal((R, R)) + 1j * rng.standard_normal((R, R)))
    H1 = np.einsum("ab,kbp->kap", Q, H)
    Hs = [np.repeat(x[:, None], 14, axis=1) for x in (H, H1)]            # [K, 14, R, P]
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": Pn, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"NSizeBWP": nrb, "NStartBWP": 0, "PanelDimensions": (2, 2), "CodebookMode": 1, "PMIMode": "Subband", "CQIMode": "Subband",
          "SubbandSize": 16}
    # H is handed over in complex64: compare the float64 SINR arrays at the tolerance of that input rounding
    outs = [ph.dlPMISelect(carrier, csirs, rc, 4, h.astype(np.complex64), 0.05) for h in Hs]
    S0, S1 = outs[0][1]["SINRPerRE"], outs[1][1]["SINRPerRE"]
    m = ~np.isnan(S0)
    assert np.array_equal(m, ~np.isnan(S1))
    err = (np.abs(S1[m] - S0[m]) / np.abs(S0[m])).max()
    print("SINRPerRE change under a receive-side unitary", err)
    assert err <= 1e-5
    table = P.communication.setupSINRtoCQIMappingTable()["downlinkSINR90pc"]
    r0 = ph.csiReport(carrier, csirs, rc, Hs[0].astype(np.complex64), 0.05, table)
    r1 = ph.csiReport(carrier, csirs, rc, Hs[1].astype(np.complex64), 0.05, table)
    assert r0[0] == r1[0]
    assert np.array_equal(r0[2], r1[2], equal_nan=True)                  # CQI
