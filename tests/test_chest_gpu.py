"""Channel estimation on the GPU (csrc/chest.cu through the C ABI) against the float64 oracle and the committed fixture.

Tolerance: Hest within 1e-5 of the channel's RMS (fp32 arithmetic on O(1) data; observed ~2e-7), nVar within 1e-4 relative
(float64 reduction of fp32 despread estimates)."""
import importlib
import importlib.util
import os

import numpy as np
import pytest

from oracle import chest as OCH

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)


def _ph():
    return importlib.import_module(PKG + ".communication.phyLayer")


def _check(He, Ho, tol=1e-5):
    rms = np.sqrt(np.mean(np.abs(Ho) ** 2))
    err = np.abs(He - Ho).max() / rms
    assert err < tol, err
    return err


def test_fixture(gpu):
    g = np.load(os.path.join(HERE, "golden", "chest_small.npz"))
    K, L, R, P, ind, sym, cdm, H, rx, nv = MG.chest_case()
    He, nve = _ph().nrChannelEstimate(rx, ind, sym, P, cdm)
    _check(He, g["Hest"], 2e-6)
    assert np.isclose(nve, float(g["nVar"]), rtol=1e-4)
    Ha, _ = _ph().nrChannelEstimate(rx, ind, sym, P, cdm, (3, 1))
    _check(Ha, g["Hest_avg"], 2e-6)


@pytest.mark.parametrize("nrb,R,cdm_on", [(24, 2, True), (52, 4, True), (273, 8, True), (25, 1, False)])
def test_csirs_row5_against_oracle(gpu, nrb, R, cdm_on):
    K, L, P = 12 * nrb, 14, 4
    ind, sym, cdm = OCH.csirs_row5_layout(nrb, 1, 0, seed=nrb)
    rng = np.random.default_rng(nrb)
    k = np.arange(K)[:, None, None, None]
    H = sum((rng.standard_normal((R, P)) + 1j * rng.standard_normal((R, P)))[None, None] * np.exp(-2j * np.pi * k * t / 512.0)
            for t in range(4)) * np.ones((1, L, 1, 1))
    noise = 0.05 * (rng.standard_normal((K, L, R)) + 1j * rng.standard_normal((K, L, R)))
    rx = OCH.apply_channel(H, ind, sym, noise)
    cd = cdm if cdm_on else (1, 1)
    Ho, nvo = OCH.channel_estimate(rx, ind, sym, P, cd)
    He, nve = _ph().nrChannelEstimate(rx, ind, sym, P, cd)
    _check(He, Ho)
    assert np.isclose(nve, nvo, rtol=1e-4)


def test_two_symbol_comb_td_cdm_and_window(gpu):
    """SRS-like comb on two symbols, TD-CDM2, 'AveragingWindow' [3 1] (gNBPhy.m:1030 passes [0 7]; 0 = none here)."""
    K, L, R, P = 288, 14, 4, 2
    rng = np.random.default_rng(11)
    ind, sym = [], []
    for p in range(P):
        ks = np.arange(p, K, 4)
        for j, l in enumerate((8, 9)):
            ind.append(1 + ks + K * l + K * L * p)
            cover = 1.0 if (p == 0 or j == 0) else -1.0
            sym.append(cover * np.exp(2j * np.pi * rng.random(ks.size)))
    ind, sym = np.concatenate(ind), np.concatenate(sym)
    k = np.arange(K)[:, None, None, None]
    H = (rng.standard_normal((R, P)) + 1j * rng.standard_normal((R, P)))[None, None] * np.exp(-2j * np.pi * k * 5 / 512.0) * np.ones((1, L, 1, 1))
    rx = OCH.apply_channel(H, ind, sym, 0.02 * (rng.standard_normal((K, L, R)) + 1j * rng.standard_normal((K, L, R))))
    for win in ((0, 0), (3, 1)):
        Ho, nvo = OCH.channel_estimate(rx, ind, sym, P, (1, 2), win)
        He, nve = _ph().nrChannelEstimate(rx, ind, sym, P, (1, 2), win)
        _check(He, Ho)
        assert np.isclose(nve, nvo, rtol=1e-4)


def test_batched_estimate_feeds_the_csi_report(gpu):
    """rxGrid -> Hest stays on the device and goes straight into csiReport (uePhy.m:897-907); indices must equal the
    report computed from the oracle's channel estimate."""
    import torch
    P_ = importlib.import_module(PKG)
    ph = _ph()
    nrb, R, P, B = 52, 4, 4, 3
    K, L = 12 * nrb, 14
    ind, sym, cdm = OCH.csirs_row5_layout(nrb, 1, 0, seed=9)
    rng = np.random.default_rng(9)
    k = np.arange(K)[:, None, None, None]
    rxs, Hos, nvs = [], [], []
    for b in range(B):
        H = sum((rng.standard_normal((R, P)) + 1j * rng.standard_normal((R, P)))[None, None] * np.exp(-2j * np.pi * k * t / 256.0)
                for t in range(3)) * np.ones((1, L, 1, 1))
        rx = OCH.apply_channel(H, ind, sym, 0.03 * (rng.standard_normal((K, L, R)) + 1j * rng.standard_normal((K, L, R))))
        Ho, nvo = OCH.channel_estimate(rx, ind, sym, P, cdm)
        rxs.append(rx), Hos.append(Ho), nvs.append(nvo)
    est = ph.ChannelEstimator(K, L, R, P, ind, sym, cdm, max_batch=B)
    rx_d = torch.from_numpy(np.ascontiguousarray(np.stack(rxs).astype(np.complex64).transpose(0, 3, 2, 1))).cuda()
    Hd, nvar = est.run_dev(rx_d, B)
    for b in range(B):
        _check(Hd[b].permute(3, 2, 1, 0).cpu().numpy(), Hos[b])
    assert np.allclose(nvar, nvs, rtol=1e-4)
    assert np.allclose(est.nvar(B), nvar)
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": 4, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"NSizeBWP": nrb, "NStartBWP": 0, "PanelDimensions": (2, 1), "CodebookMode": 1, "PMIMode": "Subband",
          "CQIMode": "Subband", "SubbandSize": 4}
    table = P_.communication.setupSINRtoCQIMappingTable()["downlinkSINR90pc"]
    dev = ph.csiReport(carrier, csirs, rc, Hd, nvar, table)
    ref = ph.csiReport(carrier, csirs, rc, np.stack(Hos, axis=-1), np.array(nvs), table)
    for a, b in zip(dev, ref):
        if isinstance(a, dict):
            assert all(np.array_equal(a[q], b[q], equal_nan=True) for q in a)
        else:
            assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_invalid_layouts_are_rejected(gpu):
    ph = _ph()
    _lib = importlib.import_module(PKG + "._lib")
    with pytest.raises(_lib.IsacError):
        ph.ChannelEstimator(12, 14, 1, 1, [1, 2, 13], [1, 1, 1])                 # not a product grid
    with pytest.raises(_lib.IsacError):
        ph.ChannelEstimator(12, 14, 1, 2, [1, 2], [1, 1])                        # port 2 has no reference REs
    with pytest.raises(_lib.IsacError):
        ph.ChannelEstimator(12, 14, 1, 1, [1, 2], [1, 1], AveragingWindow=(2, 1))   # even window
