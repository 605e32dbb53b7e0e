"""The MEX gateways of matlab/mex EXECUTED on the GPU through a functional mock of MATLAB's mx/mex API (tests/mexmock.py):
each gateway is fed the struct its .m shim builds and its outputs are compared with the Python mirror of the same
reference function (which the parity tests check against the oracle).  This exercises the marshalling code itself --
field names, dimension handling, 1-based indices, NaN conventions, device-buffer helpers -- which no MATLAB here can run."""
import importlib

import numpy as np
import pytest

import mexmock as M
from oracle import sensing as S

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(gpu):
    return importlib.import_module(PKG)


def _doa_fields(rp):
    ant = rp["antennaType"]
    upa = ant["type"] == "upa"
    return {"isUpa": float(upa), "nAnts": 0.0 if upa else float(ant["nV"] * ant["p"]), "nX": float(ant["nV"]) if upa else 0.0,
            "nY": float(ant.get("nH", 0)) if upa else 0.0, "aGran": float(rp["azimuthScanGranularity"]),
            "aMax": float(rp["azimuthScanScale"]), "eGran": float(rp["elevationScanGranularity"]),
            "eMax": float(rp["elevationScanScale"])}


def test_ul_pmi_gateway(P):
    ph = P.communication.phyLayer
    rng = np.random.default_rng(3)
    K, R, Pn, nu, band = 24 * 12, 4, 4, 2, 4
    hest = np.zeros((K, 1, R, Pn), np.complex64)
    sc = np.arange(1, K, 4)
    hest[sc, 0] = ((rng.standard_normal((sc.size, R, Pn)) + 1j * rng.standard_normal((sc.size, R, Pn))) / np.sqrt(2)).astype(np.complex64)
    pmi, sinr, sb = M.call("isac_ul_pmi_mex", 3, float(nu), hest, 0.05, float(band))
    rp, rs, rsb = ph.pmiSelect(nu, hest, 0.05, band)
    assert np.array_equal(pmi.ravel(), np.asarray(rp).ravel(), equal_nan=True)
    assert np.array_equal(sinr, rs, equal_nan=True) and np.array_equal(sb, rsb)
    none = M.call("isac_ul_pmi_mex", 3, float(nu), np.zeros_like(hest), 0.05, float(band))   # no estimate -> scalar NaNs
    assert all(x.shape == (1, 1) and np.isnan(x[0, 0]) for x in none)


def test_prg_precode_gateway(P):
    ph = P.communication.phyLayer
    rng = np.random.default_rng(8)
    nrb, L, nu, Pn, nprg = 24, 14, 2, 8, 6
    K = 12 * nrb
    pos = rng.choice(K * L, size=500, replace=False)
    portind = np.stack([pos + 1 + K * L * j for j in range(nu)], axis=1).astype(np.int32)
    portsym = (rng.standard_normal((500, nu)) + 1j * rng.standard_normal((500, nu))).astype(np.complex64)
    F = (rng.standard_normal((nu, Pn, nprg)) + 1j * rng.standard_normal((nu, Pn, nprg))).astype(np.complex64)
    sym, ind = M.call("isac_prg_precode_mex", 2, np.array([K, L, Pn], float), 0.0, portsym, portind, F)
    rs, ri = ph.prgPrecode((K, L, Pn), 0, portsym, portind, F)
    assert np.array_equal(sym, rs) and np.array_equal(ind, ri)


def test_doa_gateway(P):
    rng = np.random.default_rng(5)
    n, N = 16, 2000
    rp = {"antennaType": {"type": "ula", "nV": 8, "p": 2, "d": 0.5}, "azimuthScanScale": 360, "azimuthScanGranularity": 1,
          "elevationScanScale": 180, "elevationScanGranularity": 1}
    A = np.exp(-2j * np.pi * np.arange(n)[:, None] * 0.5 * S.sind(np.array([-35.0, 12.0, 48.0]))[None, :])
    X = A @ (rng.standard_normal((3, N)) + 1j * rng.standard_normal((3, N))) + 0.3 * (rng.standard_normal((n, N)) + 1j * rng.standard_normal((n, N)))
    Ra = X @ X.conj().T / N
    for method, fn in ((0, "music"), (1, "mvdrBF"), (2, "digitalBF")):
        L, azi, PdB = M.call("isac_doa_mex", 3, _doa_fields(rp), float(method), 3.0, Ra)
        ref = getattr(P.sensing.estimation.doaEstimation, fn)(3, rp, Ra, return_spectrum=True)
        azr, spec = (ref[1], ref[3]) if fn == "music" else (ref[0], ref[2])
        assert int(L[0, 0]) == 3 and np.array_equal(azi.ravel(), azr) and np.array_equal(PdB.ravel(), np.asarray(spec).ravel())
    L, azi, _ = M.call("isac_doa_mex", 3, _doa_fields(rp), 0.0, np.zeros((0, 0)), Ra)      # numDets = []: eigen-gap rule
    Lr, azr, _ = P.sensing.estimation.doaEstimation.music(None, rp, Ra)
    assert int(L[0, 0]) == Lr and np.array_equal(azi.ravel(), azr)
    with pytest.raises(M.MexError) as e:                                                      # zero sources: findpeaks errors
        M.call("isac_doa_mex", 3, _doa_fields(rp), 0.0, 0.0, Ra)
    assert e.value.identifier == "isac:doa:status7"


def test_fft2d_and_mono_static_gateways(P):
    W = P.workloads
    cell, car, wave = W.cell_config("tiny")
    rp = P.sensing.radarParams(cell, car, wave)
    cf = P.sensing.detection.cfar2D(rp)
    grid, txw = W.sensing_tx("tiny", 1)
    noise = W.std_normal_complex(txw.shape, 2).astype(np.complex64)
    txw32, grid32 = txw.astype(np.complex64), grid.astype(np.complex64)
    num = W.ofdm_numerology(int(car["NRBsDL"]), float(car["SubcarrierSpacing"]))
    nT = int(rp["nTargets"])
    ecfg = {"fc": float(rp["fc"]), "fs": float(rp["fs"]), "N0": float(rp["N0"]), "range": np.asarray(rp["range"], float).reshape(nT),
            "velocity": np.asarray(rp["velocity"], float).reshape(nT),
            "largeScaleFading": np.asarray(rp["largeScaleFading"], float).reshape(nT),
            "steeringVec": np.asarray(rp["RxSteeringVec"], np.complex128).reshape(txw.shape[1], nT),
            "los": np.asarray(cell["targetLoSConditions"], np.int32).reshape(nT), "nfft": float(num["Nfft"]),
            "nSc": float(12 * int(car["NRBsDL"])), "nSymTx": float(grid.shape[1]),
            "cpLengths": np.asarray(num["CyclicPrefixLengths"], np.int32)}
    echo = M.call("isac_mono_static_mex", 1, ecfg, txw32, np.array([7], np.uint64), noise)
    ref = P.sensing.monoStaticSensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"], noise=noise)
    assert echo.dtype == np.complex64 and echo.shape == ref.shape and np.array_equal(echo, ref)
    fcfg = {"nIFFT": float(rp["nIFFT"]), "nFFT": float(rp["nFFT"]), "rRes": float(rp["rRes"]), "vRes": float(rp["vRes"]),
            "cutRows": np.array([cf["CUTIdx"][0].min(), cf["CUTIdx"][0].max()], float).reshape(1, 2),
            "cutCols": np.array([cf["CUTIdx"][1].min(), cf["CUTIdx"][1].max()], float).reshape(1, 2), "Pfa": float(rp["Pfa"])}
    fcfg.update(_doa_fields(rp))
    est = M.call("isac_fft2d_mex", 1, fcfg, echo, grid32)
    r = P.sensing.estimation.fft2D(rp, cf, echo, grid32)
    for k in ("rngEst", "velEst", "aziEst"):
        assert np.array_equal(est[k].ravel(), r[k]), k
    assert est["eleEst"].size == r["eleEst"].size and np.all(np.isnan(est["eleEst"]))


def test_music2d_gateway(P):
    rng = np.random.default_rng(21)
    nSc, nSym, nA, scs, fc = 96, 40, 4, 30.0, 3.5e9
    lam = S.LIGHTSPEED / fc
    Tsri = 1 / (scs * 1e3) + 5e-6
    rp = {"fc": fc, "Tsri": Tsri, "cfarEstZone": np.array([[50.0, 300.0], [-50.0, 50.0]]),
          "antennaType": {"type": "ula", "nV": 2, "p": 2, "d": 0.5}, "azimuthScanScale": 360, "azimuthScanGranularity": 1,
          "elevationScanScale": 180, "elevationScanGranularity": 1}
    tx = np.exp(2j * np.pi * rng.random((nSc, nSym, nA)))
    k, l = np.arange(nSc)[:, None], np.arange(nSym)[None, :]
    H = sum(a * np.exp(-2j * np.pi * scs * 1e3 * 2 * r * k / S.LIGHTSPEED) * np.exp(2j * np.pi * Tsri * 2 * v * l / lam)
            for r, v, a in ((120.0, 10.0, 1.0), (210.5, -22.5, 0.7)))
    rx = np.stack([(H * np.exp(-2j * np.pi * a * 0.5 * S.sind(25.0))) * tx[:, :, a] for a in range(nA)], axis=2)
    rx = (rx + 0.05 * (rng.standard_normal(rx.shape) + 1j * rng.standard_normal(rx.shape))).astype(np.complex64)
    tx = tx.astype(np.complex64)
    cfg = {"scsHz": scs * 1e3, "fc": fc, "Tsri": Tsri, "rMax": 300.0, "vZone": 50.0}
    cfg.update(_doa_fields(rp))
    est = M.call("isac_music2d_mex", 1, cfg, rx, tx)
    ref = P.sensing.estimation.music2D(rp, {"scs": scs}, rx, tx)
    for key in ("rngEst", "velEst", "aziEst", "PrmusicdB", "PvmusicdB"):
        assert np.array_equal(est[key].ravel(), np.asarray(ref[key]).ravel()), key


def test_dl_pmi_and_csi_report_gateways(P):
    ph = P.communication.phyLayer
    rng = np.random.default_rng(30)
    nrb, R, Pn, sbs = 24, 4, 8, 4
    K = 12 * nrb
    H = ((rng.standard_normal((K, 14, R, Pn)) + 1j * rng.standard_normal((K, 14, R, Pn))) / np.sqrt(2)).astype(np.complex64)
    H = (H + np.roll(H, 1, axis=0)).astype(np.complex64)
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": Pn, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"NSizeBWP": nrb, "NStartBWP": 0, "PanelDimensions": (2, 2), "CodebookMode": 1, "PMIMode": "Subband", "CQIMode": "Subband",
          "SubbandSize": sbs}
    pm, info = ph.dlPMISelect(carrier, csirs, rc, 2, H, 0.1)
    cfg = {"nPorts": float(Pn), "N1": 2.0, "N2": 2.0, "O1": 4.0, "O2": 4.0, "codebookMode": 1.0, "nSizeBWP": float(nrb),
           "nStartBWP": 0.0, "subbandSize": float(sbs), "pmiSubband": 1.0, "cqiSubband": 1.0, "K": float(K), "L": 14.0,
           "subsetRestriction": np.ones(64, np.uint8), "i2Restriction": np.ones(16, np.uint8), "riRestriction": np.ones(8, np.uint8),
           "reK": info["reK"].astype(np.int32), "reL": info["reL"].astype(np.int32)}
    i1, i2, S_re, S_sb, Wm, reK, reL = M.call("isac_dl_pmi_mex", 7, cfg, 2.0, H, 0.1)
    assert np.array_equal(i1.ravel(), pm["i1"]) and np.array_equal(i2.ravel(), pm["i2"], equal_nan=True)
    assert S_re.shape == info["SINRPerRE"].shape and np.array_equal(S_re, info["SINRPerRE"], equal_nan=True)
    assert np.array_equal(S_sb, info["SINRPerSubband"], equal_nan=True)
    assert np.array_equal(Wm, info["W"]) and np.array_equal(reK.ravel(), info["reK"]) and np.array_equal(reL.ravel(), info["reL"])
    table = np.asarray(P.communication.setupSINRtoCQIMappingTable()["downlinkSINR90pc"], float)
    rk, pmr, cq = ph.csiReport(carrier, csirs, rc, H, 0.1, table, rankCap=4)
    RI, i1, i2, CQI = M.call("isac_csi_report_mex", 4, cfg, H, 0.1, table, 4.0, 0.0)         # mode 0: fused report
    assert RI[0, 0] == rk and np.array_equal(i1.ravel(), pmr["i1"]) and np.array_equal(i2.ravel(), pmr["i2"], equal_nan=True)
    assert np.array_equal(CQI, cq, equal_nan=True)
    ri_ref, pm_ri = ph.riSelect(carrier, csirs, rc, H, 0.1)
    RI, i1, i2, _ = M.call("isac_csi_report_mex", 4, cfg, H, 0.1, table, 0.0, 1.0)           # mode 1: riSelect
    assert RI[0, 0] == ri_ref and np.array_equal(i1.ravel(), pm_ri["i1"]) and np.array_equal(i2.ravel(), pm_ri["i2"], equal_nan=True)
    cq_ref, pm_cq, _ = ph.cqiSelect(carrier, csirs, rc, 3, H, 0.1, table)
    _, i1, i2, CQI = M.call("isac_csi_report_mex", 4, cfg, H, 0.1, table, 0.0, 2.0, 3.0)     # mode 2: cqiSelect at rank 3
    assert np.array_equal(i1.ravel(), pm_cq["i1"]) and np.array_equal(i2.ravel(), pm_cq["i2"], equal_nan=True)
    assert np.array_equal(CQI[:, : cq_ref.shape[1]], cq_ref, equal_nan=True)


def test_dl_pmi_gateway_multi_panel(P):
    """Type1MultiPanel through the gateway: flattened outputs + mpDims, un-flattened here as the .m shim does."""
    ph = P.communication.phyLayer
    rng = np.random.default_rng(31)
    nrb, R, panel, nu = 24, 4, (2, 2, 1), 2
    Pn, K = 8, 12 * nrb
    H = ((rng.standard_normal((K, 14, R, Pn)) + 1j * rng.standard_normal((K, 14, R, Pn))) / np.sqrt(2)).astype(np.complex64)
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": Pn, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"CodebookType": "Type1MultiPanel", "PanelDimensions": panel, "CodebookMode": 2, "NSizeBWP": nrb, "NStartBWP": 0,
          "PMIMode": "Subband", "CQIMode": "Wideband", "SubbandSize": 8}
    pm, info = ph.dlPMISelect(carrier, csirs, rc, nu, H, 0.1)
    cfg = {"nPanels": 2.0, "nPorts": float(Pn), "N1": 2.0, "N2": 1.0, "O1": 4.0, "O2": 1.0, "codebookMode": 2.0, "nSizeBWP": float(nrb),
           "nStartBWP": 0.0, "subbandSize": 8.0, "pmiSubband": 1.0, "cqiSubband": 0.0, "K": float(K), "L": 14.0,
           "subsetRestriction": np.ones(8, np.uint8), "i2Restriction": np.ones(16, np.uint8), "riRestriction": np.ones(8, np.uint8),
           "reK": info["reK"].astype(np.int32), "reL": info["reL"].astype(np.int32)}
    i1, i2, S_re, S_sb, Wm, reK, reL, mp = M.call("isac_dl_pmi_mex", 8, cfg, float(nu), H, 0.1)
    mp = [int(x) for x in mp.ravel()]
    full = tuple(mp[0:3]) + S_re.shape[3:5] + tuple(mp[3:7])
    assert S_re.ndim == 6 and np.prod(full) == np.prod(S_re.shape[2:])
    assert np.array_equal(S_re.reshape(S_re.shape[:2] + full, order="F"), info["SINRPerRE"], equal_nan=True)
    assert np.array_equal(S_sb.reshape(S_sb.shape[:2] + full, order="F"), info["SINRPerSubband"], equal_nan=True)
    assert np.array_equal(Wm.reshape(Wm.shape[:2] + full, order="F"), info["W"])
    i1 = i1.ravel()
    i1u = [i1[0], i1[1]] + [x + 1 for x in np.unravel_index(int(i1[2]) - 1, tuple(mp[3:7]), order="F")]
    assert np.array_equal(i1u, pm["i1"])
    for sb, v in enumerate(i2.ravel()):
        assert np.array_equal([x + 1 for x in np.unravel_index(int(v) - 1, tuple(mp[0:3]), order="F")], pm["i2"][:, sb])


def test_radar_channel_gateway(P):
    """sensing.channelModels.basicRadarChannel through its gateway, incl. the all-NLoS error the reference turns into an
    empty waveform (basicRadarChannel.m:59-64)."""
    W = P.workloads
    cell, car, wave = W.cell_config("tiny")
    rp = P.sensing.radarParams(cell, car, wave)
    _, txw = W.sensing_tx("tiny", 1)
    noise = W.std_normal_complex(txw.shape, 2).astype(np.complex64)
    txw32 = txw.astype(np.complex64)
    nT = int(rp["nTargets"])
    cfg = {"fc": float(rp["fc"]), "fs": float(rp["fs"]), "N0": float(rp["N0"]), "range": np.asarray(rp["range"], float).reshape(nT),
           "velocity": np.asarray(rp["velocity"], float).reshape(nT),
           "largeScaleFading": np.asarray(rp["largeScaleFading"], float).reshape(nT),
           "steeringVec": np.asarray(rp["RxSteeringVec"], np.complex128).reshape(txw.shape[1], nT),
           "los": np.asarray(cell["targetLoSConditions"], np.int32).reshape(nT)}
    rx = M.call("isac_radar_channel_mex", 1, cfg, txw32, np.array([0], np.uint64), noise)
    ref = P.sensing.channelModels.basicRadarChannel(txw32, rp, cell["targetLoSConditions"], noise=noise)
    assert rx.dtype == np.complex64 and rx.shape == ref.shape and np.array_equal(rx, ref)
    with pytest.raises(M.MexError) as e:
        M.call("isac_radar_channel_mex", 1, dict(cfg, los=np.zeros(nT, np.int32)), txw32, np.array([0], np.uint64), noise)
    assert e.value.identifier == "isac:basicRadarChannel:status6"


def test_precoded_sinr_and_codebook_gateways(P):
    ph = P.communication.phyLayer
    rng = np.random.default_rng(12)
    R, Pn, nu, B = 4, 4, 2, 37
    H = (rng.standard_normal((R, Pn, B)) + 1j * rng.standard_normal((R, Pn, B))) / np.sqrt(2)
    Wc = ph.puschCodebook(nu, Pn)[:, :, 3]
    got = M.call("isac_precoded_sinr_mex", 1, H, 0.3, Wc)
    ref = np.asarray(ph.precodedSINR(H, 0.3, Wc)).ravel()
    assert got.shape == (B, 1) and np.array_equal(got.ravel(), ref)
    one = M.call("isac_precoded_sinr_mex", 1, H[:, :, 0], 0.3, Wc)          # the reference's per-RE call (pmiSelect.m:52)
    assert one.shape == (1, 1) and one[0, 0] == ref[0]
    # gNB-side codebook copy (pmiType1SinglePanelCodebook.m), 16 ports rank 3: the copy's missing i13 index is reproduced
    rc = {"PanelDimensions": (4, 2), "OverSamplingFactors": (4, 4), "CodebookMode": 1}
    cfg = {"nPorts": 16.0, "N1": 4.0, "N2": 2.0, "O1": 4.0, "O2": 4.0, "codebookMode": 1.0, "nSizeBWP": 1.0, "nStartBWP": 0.0,
           "subbandSize": 0.0, "pmiSubband": 0.0, "cqiSubband": 0.0, "K": 12.0, "L": 14.0, "subsetRestriction": np.zeros(0, np.uint8),
           "i2Restriction": np.zeros(0, np.uint8), "riRestriction": np.ones(8, np.uint8), "reK": np.zeros(0, np.int32),
           "reL": np.zeros(0, np.int32)}
    for nl in (1, 3):
        Wg = M.call("isac_codebook_mex", 1, cfg, float(nl), 1.0)
        Wr = P.communication.pmiType1SinglePanelCodebook(rc, nl)
        assert Wg.shape == Wr.shape and np.array_equal(Wg, Wr)
    with pytest.raises(M.MexError) as e:
        M.call("isac_codebook_mex", 1, dict(cfg, N1=3.0), 1.0, 1.0)           # 2*N1*N2 != nPorts
    assert e.value.identifier == "isac:pmiType1SinglePanelCodebook:config"


def test_cdl_gateway(P):
    """communication.channelModels.cdlChannelMatrix: the gateway returns the same H as the Python mirror's channel object."""
    cm = importlib.import_module(PKG + ".communication.channelModels")
    K, scs = 24 * 12, 30e3
    num = P.workloads.ofdm_numerology(24, 30)
    st = np.ascontiguousarray(P.workloads.symbol_starts(num, 14) / num["SampleRate"], dtype=np.float64)
    cfg = {"profile": 2.0, "delaySpread": 300e-9, "fc": 3.5e9, "maxDoppler": 5.0, "txSize": np.array([1, 4, 2], np.int32),
           "rxSize": np.array([1, 1, 2], np.int32), "txPattern38901": 1.0, "rxPattern38901": 0.0, "seed": 73.0}
    H = M.call("isac_cdl_mex", 1, cfg, float(K), scs, st.reshape(1, -1), 0.002)
    ch = cm.CDLChannel("CDL-C", TransmitAntennaArraySize=(1, 4, 2), ReceiveAntennaArraySize=(1, 1, 2), Seed=73)
    ref = ch.generate(K, scs, st, 0.002).cpu().numpy().transpose(3, 2, 1, 0)      # [nTx][nRx][L][K] -> [K x L x nRx x nTx]
    assert H.dtype == np.complex64 and H.shape == ref.shape == (K, 14, 2, 8) and np.array_equal(H, ref)
    H2 = M.call("isac_cdl_mex", 1, cfg, float(K), scs, st.reshape(1, -1), 0.002)  # second call: cached channel, same realisation
    assert np.array_equal(H, H2)


def test_cqi_gateway_info_output_and_plan_reuse(P):
    """Mode 2 of the report gateway returns SINRPerSubbandPerCW for CQIInfo (cqiSelect.m:685); repeated calls with one
    configuration reuse the cached plan and give identical results."""
    ph = P.communication.phyLayer
    rng = np.random.default_rng(33)
    nrb, R, Pn, sbs = 24, 4, 8, 4
    K = 12 * nrb
    H = ((rng.standard_normal((K, 14, R, Pn)) + 1j * rng.standard_normal((K, 14, R, Pn))) / np.sqrt(2)).astype(np.complex64)
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": Pn, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"NSizeBWP": nrb, "NStartBWP": 0, "PanelDimensions": (2, 2), "CodebookMode": 1, "PMIMode": "Subband", "CQIMode": "Subband",
          "SubbandSize": sbs}
    _, info = ph.dlPMISelect(carrier, csirs, rc, 1, H, 0.1)
    cfg = {"nPorts": float(Pn), "N1": 2.0, "N2": 2.0, "O1": 4.0, "O2": 4.0, "codebookMode": 1.0, "nSizeBWP": float(nrb),
           "nStartBWP": 0.0, "subbandSize": float(sbs), "pmiSubband": 1.0, "cqiSubband": 1.0, "K": float(K), "L": 14.0,
           "subsetRestriction": np.ones(64, np.uint8), "i2Restriction": np.ones(16, np.uint8), "riRestriction": np.ones(8, np.uint8),
           "reK": info["reK"].astype(np.int32), "reL": info["reL"].astype(np.int32)}
    table = np.asarray(P.communication.setupSINRtoCQIMappingTable()["downlinkSINR90pc"], float)
    cq_ref, pm_ref, ci = ph.cqiSelect(carrier, csirs, rc, 2, H, 0.1, table)
    outs = [M.call("isac_csi_report_mex", 5, cfg, H, 0.1, table, 0.0, 2.0, 2.0) for _ in range(3)]
    for _, i1, i2, CQI, sbcw in outs:
        assert np.array_equal(i1.ravel(), pm_ref["i1"]) and np.array_equal(i2.ravel(), pm_ref["i2"], equal_nan=True)
        assert np.array_equal(CQI[:, :1], cq_ref, equal_nan=True)
        assert sbcw.shape == (nrb // sbs + 1, 2)
        assert np.array_equal(sbcw[:, :1], np.asarray(ci["SINRPerSubbandPerCW"]).reshape(-1, 1), equal_nan=True)
