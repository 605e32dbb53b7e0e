"""The CUDA path (through the package API / C ABI) against the committed fixtures of tests/golden/ - no live oracle
run on the checked quantities.  Tolerances as in the parity tests: detection / index outputs exact, fp32 maps 1e-5 of
the peak, float64 SINR sums 1e-5 relative (observed ~1e-13)."""
import importlib
import importlib.util
import os

import numpy as np
import pytest

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)


def _load(name):
    return np.load(os.path.join(HERE, "golden", name))


def test_sensing_chain_against_fixture(gpu):
    P = importlib.import_module(PKG)
    g = _load("sensing_tiny.npz")
    W = P.workloads
    cell, car, wave = W.cell_config("tiny")
    grid, txw = W.sensing_tx("tiny", 1)
    noise = W.std_normal_complex(txw.shape, 2)
    prm = P.sensing.radarParams(cell, car, wave)
    rx = P.sensing.monoStaticSensing(txw.astype(np.complex64), grid.shape, car, prm, cell["targetLoSConditions"],
                                     noise=noise.astype(np.complex64))
    err = np.abs(rx[::7, ::5, :] - g["echo_grid_sample"]).max() / float(g["echo_grid_rms"])
    print("echo grid err / rms", err)
    assert err <= 1e-5
    cf = P.sensing.detection.cfar2D(prm)
    got = P.sensing.estimation.fft2D(prm, cf, rx.astype(np.complex64), grid.astype(np.complex64))
    # the GPU echo grid differs from the float64 one in the last fp32 bits, so the estimates (quantised to bins) must
    # still be identical to the fixture
    assert np.array_equal(got["rngEst"], g["rngEst"])
    assert np.array_equal(got["velEst"], g["velEst"])
    assert np.array_equal(got["aziEst"], g["aziEst"])


def test_rdm_and_cfar_against_fixture(gpu):
    P = importlib.import_module(PKG)
    g = _load("sensing_tiny.npz")
    cell, car, wave, rp, cf, grid, txw, noise, rx = MG.sensing_case()
    rdm = importlib.import_module(PKG + ".sensing._rdm")
    import torch
    plan = rdm.RangeDopplerPlan(grid.shape[0], grid.shape[1], grid.shape[2], rp["nIFFT"], rp["nFFT"], cf["rngIdx"], cf["dopIdx"], rp["Pfa"])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.complex64).transpose(2, 1, 0))).cuda()
    plan.run_dev(dev(rx), dev(grid), 1)
    Pw = plan.power(1)[..., 0].astype(np.float64)
    assert np.abs(Pw[::9, ::3, :] - g["power_sample"]).max() <= 1e-5 * float(g["power_peak"])
    assert np.allclose(Pw.sum(axis=(0, 1)), g["power_sum"], rtol=1e-5)
    cnt, dets = plan.detections(1)
    assert np.array_equal(np.array([d[0].shape[1] for d in dets[0]]), g["n_det"])
    assert np.array_equal(np.concatenate([d[0][0] for d in dets[0]]), g["det_rows"])
    assert np.array_equal(np.concatenate([d[0][1] for d in dets[0]]), g["det_cols"])
    assert abs(plan.alpha - float(g["cfar_alpha"])) <= 1e-12 * plan.alpha
    plan.close()


def test_csi_against_fixture(gpu):
    PH = importlib.import_module(PKG + ".communication.phyLayer")
    g = _load("comm_small.npz")
    for tag, (P, panel, nrb, R, seed) in MG.COMM_CASES.items():
        ocfg, re_k, re_l, H, nv = MG.comm_case(P, panel, nrb, R, seed)
        carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
        csirs = {"NumCSIRSPorts": P, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
        rc = {"NSizeBWP": nrb, "NStartBWP": 0, "CodebookMode": 1, "PMIMode": "Subband", "CQIMode": "Subband", "SubbandSize": 4,
              "PanelDimensions": panel}
        for nu in (1, 2, min(R, P)):
            pm, info = PH.dlPMISelect(carrier, csirs, rc, nu, H, nv)
            assert np.array_equal(pm["i1"], g[f"{tag}_nu{nu}_i1"]), (tag, nu)
            assert np.array_equal(pm["i2"], g[f"{tag}_nu{nu}_i2"], equal_nan=True), (tag, nu)
            s = np.nansum(info["SINRPerSubband"], axis=(0, 1))
            assert np.abs(s - g[f"{tag}_nu{nu}_sinr_sb_sum"]).max() <= 1e-5 * np.abs(g[f"{tag}_nu{nu}_sinr_sb_sum"]).max()
            smp = info["SINRPerRE"][::5, :, ...].reshape(-1)[::97]
            ref = g[f"{tag}_nu{nu}_sinr_re_sample"]
            m = ~np.isnan(ref)
            assert np.array_equal(m, ~np.isnan(smp)) and np.abs(smp[m] - ref[m]).max() <= 1e-5 * np.abs(ref[m]).max()
        ri, _ = PH.riSelect(carrier, csirs, rc, H, nv)
        assert ri == float(g[f"{tag}_ri"])
        cqi, _, _ = PH.cqiSelect(carrier, csirs, rc, int(ri), H, nv, g["cqi_table"])
        assert np.array_equal(cqi, g[f"{tag}_cqi"], equal_nan=True)
    rng = np.random.default_rng(21)
    K = 12 * 24
    hest = np.zeros((K, 14, 8, 4), dtype=np.complex64)
    sc = np.arange(1, K, 4)
    hest[sc, 13] = ((rng.standard_normal((sc.size, 8, 4)) + 1j * rng.standard_normal((sc.size, 8, 4))) / np.sqrt(2)).astype(np.complex64)
    pmi, sinr, idx = PH.pmiSelect(2, hest, 0.05, 4)
    assert np.array_equal(pmi, g["ul_pmi"], equal_nan=True) and np.array_equal(idx, g["ul_idx"])
    m = ~np.isnan(g["ul_sinr"])
    assert np.abs(sinr[m] - g["ul_sinr"][m]).max() <= 1e-5 * np.abs(g["ul_sinr"][m]).max()


def test_cdl_against_fixture(gpu):
    cm = importlib.import_module(PKG + ".communication.channelModels")
    g = _load("cdl_c.npz")
    ch = cm.CDLChannel("CDL-C", TransmitAntennaArraySize=(1, 4, 2), ReceiveAntennaArraySize=(1, 2, 2), Seed=73)
    rays = ch.rays()
    assert np.allclose(rays["tau"], g["tau"], rtol=1e-14) and np.allclose(rays["nu"], g["nu"], rtol=1e-12, atol=1e-12)
    H = ch.generate(24 * 12, 30e3, np.arange(14) * 35.7e-6).cpu().numpy().transpose(3, 2, 1, 0)
    err = np.abs(H[::17, ::3] - g["H_sample"]).max() / np.sqrt(float(g["H_power"]))
    print("CDL H err / rms", err)
    assert err <= 1e-5
    ch.close()
