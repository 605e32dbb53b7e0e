"""GPU parity of the COMM half behind the package API (csrc/comm.cu) against the float64 oracle:
dlPMISelect / riSelect / cqiSelect / fused CSI report, UL pmiSelect, prgPrecode.

Tolerances: SINR arrays within 1e-5 relative (they are computed in float64 from the same complex64 H, so the
observed error is ~1e-12); index outputs (PMI / RI / CQI / TPMI) exact, with the tie-aware comparator of
SURVEY.md 7 (hard part 6) for PMI: a different index is accepted only if its rounded metric equals the maximum.
"""
import importlib

import numpy as np
import pytest

from oracle import comm as C

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["pair", "direct"])
def PH(gpu, request):
    """Every test runs with both SINR kernels: the Gram-pair form (default) and the direct H*W form (fallback)."""
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    ph.setSinrKernel(request.param == "direct")
    yield ph
    ph.setSinrKernel(False)


def _setup(n_ports, panel, nrb, n_rx, seed, mode=1, pmi_mode="Subband", cqi_mode="Subband", sb=4, nstart=0, snr_db=10.0):
    rng = np.random.default_rng(seed)
    K = 12 * nrb
    H = ((rng.standard_normal((K, 14, n_rx, n_ports)) + 1j * rng.standard_normal((K, 14, n_rx, n_ports))) / np.sqrt(2)).astype(np.complex64)
    # smooth the channel a little across frequency so subbands differ but are not white
    H = (H + np.roll(H, 1, axis=0) + np.roll(H, 2, axis=0)).astype(np.complex64)
    n_var = float(10 ** (-snr_db / 10))
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": n_ports, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"NSizeBWP": nrb - nstart, "NStartBWP": nstart, "CodebookMode": mode, "PMIMode": pmi_mode, "CQIMode": cqi_mode, "SubbandSize": sb}
    if n_ports > 2:
        rc["PanelDimensions"] = panel
    ocfg = C.report_config(n_ports, panel, nrb - nstart, nstart, mode, pmi_mode, cqi_mode, sb)
    re_k, re_l = C.csirs_first_port_res(nrb, 1, 0)
    keep = (re_k >= nstart * 12 + 1) & (re_k <= nrb * 12)
    re_k, re_l = re_k[keep] - nstart * 12, re_l[keep]
    return carrier, csirs, rc, ocfg, re_k, re_l, H, n_var


def _pmi_equivalent(pm_g, pm_o, info_o, nu):
    """Exact match, or (tie-aware) equal rounded metrics."""
    S = info_o["SINRPerRE"]
    total = C.matlab_round4(np.nansum(S, axis=(0, 1)))
    i1g, i1o = pm_g["i1"].astype(int), pm_o["i1"].astype(int)
    if not np.array_equal(i1g, i1o):
        # wideband tie: the GPU's (i1, some i2) must reach the same rounded maximum
        assert np.isclose(total[:, i1g[0] - 1, i1g[1] - 1, i1g[2] - 1].max(), total.max(), rtol=0, atol=1e-4), (i1g, i1o)
        return False
    for sb in range(len(pm_o["i2"])):
        a, b = pm_g["i2"][sb], pm_o["i2"][sb]
        if np.isnan(b):
            assert np.isnan(a)
            continue
        if a != b:
            t = C.matlab_round4(np.nansum(info_o["SINRPerSubband"][sb][:, :, i1o[0] - 1, i1o[1] - 1, i1o[2] - 1], axis=0))
            assert abs(t[int(a) - 1] - t.max()) <= 1e-4, (sb, a, b)
    return True


CASES = [
    # ports, panel, nrb, nRx, layers, mode
    (4, (2, 1), 52, 2, 1, 1), (4, (2, 1), 52, 2, 2, 1), (4, (2, 1), 24, 4, 4, 1), (4, (2, 1), 24, 4, 2, 2),
    (8, (2, 2), 24, 8, 1, 1), (8, (2, 2), 24, 8, 3, 1), (8, (2, 2), 24, 8, 8, 1), (8, (4, 1), 24, 8, 5, 1),
    (8, (2, 2), 24, 4, 2, 2), (16, (4, 2), 24, 4, 3, 1), (16, (4, 2), 24, 4, 4, 1), (32, (4, 4), 24, 8, 1, 1),
    (32, (4, 4), 24, 8, 7, 1), (2, None, 24, 2, 2, 1), (2, None, 24, 2, 1, 1),
]


@pytest.mark.parametrize("n_ports,panel,nrb,n_rx,nu,mode", CASES)
def test_dl_pmi_select_matches_oracle(PH, n_ports, panel, nrb, n_rx, nu, mode):
    carrier, csirs, rc, ocfg, re_k, re_l, H, n_var = _setup(n_ports, panel, nrb, n_rx, 100 + n_ports + nu, mode)
    pm_o, info_o = C.dl_pmi_select(ocfg, re_k, re_l, nu, H, n_var)
    pm_g, info_g = PH.dlPMISelect(carrier, csirs, rc, nu, H, n_var)
    assert np.array_equal(info_g["reK"], re_k) and np.array_equal(info_g["reL"], re_l)
    So, Sg = info_o["SINRPerRE"], info_g["SINRPerRE"]
    assert So.shape == Sg.shape
    assert np.array_equal(np.isnan(So), np.isnan(Sg))
    m = ~np.isnan(So)
    err = np.abs(Sg[m] - So[m]) / np.abs(So[m])
    print(f"P={n_ports} nu={nu} mode={mode}: SINRPerRE max rel err {err.max():.2e} over {m.sum()} values; i1 {pm_g['i1']} i2 {pm_g['i2'][:4]}")
    assert err.max() <= 1e-5
    Bo, Bg = info_o["SINRPerSubband"], info_g["SINRPerSubband"]
    mb = ~np.isnan(Bo)
    assert np.array_equal(mb, ~np.isnan(Bg))
    assert (np.abs(Bg[mb] - Bo[mb]) / np.abs(Bo[mb])).max() <= 1e-5
    assert np.abs(info_g["W"] - info_o["W"]).max() <= 1e-14
    _pmi_equivalent(pm_g, pm_o, info_o, nu)


def test_dl_pmi_bwp_offset_wideband_and_batch(PH):
    """BWP that starts inside the carrier (first/last subband shorter), wideband PMI mode, batched UEs."""
    carrier, csirs, rc, ocfg, re_k, re_l, H, n_var = _setup(8, (2, 2), 52, 4, 7, 1, "Subband", "Subband", 8, nstart=3)
    pm_o, info_o = C.dl_pmi_select(ocfg, re_k, re_l, 2, H, n_var)
    pm_g, info_g = PH.dlPMISelect(carrier, csirs, rc, 2, H, n_var)
    assert len(pm_g["i2"]) == 7 == len(pm_o["i2"])
    _pmi_equivalent(pm_g, pm_o, info_o, 2)
    rc_w = dict(rc, PMIMode="Wideband", CQIMode="Wideband")
    ocfg_w = C.report_config(8, (2, 2), 49, 3, 1, "Wideband", "Wideband", None)
    pm_o, info_o = C.dl_pmi_select(ocfg_w, re_k, re_l, 2, H, n_var)
    pm_g, info_g = PH.dlPMISelect(carrier, csirs, rc_w, 2, H, n_var)
    assert len(pm_g["i2"]) == 1
    _pmi_equivalent(pm_g, pm_o, info_o, 2)
    # batch of 3 UEs == 3 single calls
    rng = np.random.default_rng(3)
    Hb = np.stack([H, H[::-1].copy(), (H * np.exp(1j)).astype(np.complex64)], axis=4)
    nv = np.array([n_var, 2 * n_var, 0.5 * n_var])
    pm_b, info_b = PH.dlPMISelect(carrier, csirs, rc, 2, Hb, nv)
    for b in range(3):
        pm_1, info_1 = PH.dlPMISelect(carrier, csirs, rc, 2, Hb[..., b], nv[b])
        assert np.array_equal(pm_b["i1"][:, b], pm_1["i1"]) and np.array_equal(pm_b["i2"][:, b], pm_1["i2"], equal_nan=True)
        assert np.array_equal(info_b["SINRPerRE"][..., b], info_1["SINRPerRE"], equal_nan=True)


def test_restricted_codebook_and_default_nvar(PH):
    carrier, csirs, rc, ocfg, re_k, re_l, H, n_var = _setup(8, (2, 2), 24, 4, 9)
    rng = np.random.default_rng(1)
    csr = (rng.random(64) > 0.5).astype(int)
    i2r = np.ones(16, dtype=int)
    i2r[[1, 2]] = 0
    rc2 = dict(rc, CodebookSubsetRestriction=csr, i2Restriction=i2r)
    ocfg2 = C.report_config(8, (2, 2), 24, 0, 1, "Subband", "Subband", 4, csr, i2r)
    pm_o, info_o = C.dl_pmi_select(ocfg2, re_k, re_l, 1, H)          # default nVar 1e-10 (dlPMISelect.m:318-321)
    pm_g, info_g = PH.dlPMISelect(carrier, csirs, rc2, 1, H)
    assert np.array_equal(np.isnan(info_o["SINRPerRE"]), np.isnan(info_g["SINRPerRE"]))
    m = ~np.isnan(info_o["SINRPerRE"])
    assert (np.abs(info_g["SINRPerRE"][m] - info_o["SINRPerRE"][m]) / info_o["SINRPerRE"][m]).max() <= 1e-5
    _pmi_equivalent(pm_g, pm_o, info_o, 1)
    assert int(pm_g["i2"][0]) in (1, 4)
    # everything restricted -> NaN report (dlPMISelect.m:362-379)
    rc3 = dict(rc, CodebookSubsetRestriction=np.zeros(64, dtype=int))
    pm_g, _ = PH.dlPMISelect(carrier, csirs, rc3, 1, H)
    assert np.all(np.isnan(pm_g["i1"])) and np.all(np.isnan(pm_g["i2"]))


@pytest.mark.parametrize("n_ports,panel,n_rx,snr", [(4, (2, 1), 2, 12.0), (8, (2, 2), 8, 25.0), (8, (2, 2), 4, 0.0)])
def test_ri_cqi_and_fused_report(PH, n_ports, panel, n_rx, snr):
    carrier, csirs, rc, ocfg, re_k, re_l, H, n_var = _setup(n_ports, panel, 24, n_rx, 50 + n_ports, snr_db=snr)
    table = np.array([-3.46, 1.54, 6.54, 11.05, 13.54, 16.04, 17.54, 20.04, 22.04, 24.43, 26.93, 27.43, 29.43, 32.43, 35.43])
    ri_o, pm_o = C.ri_select(ocfg, re_k, re_l, H, n_var)
    ri_g, pm_g = PH.riSelect(carrier, csirs, rc, H, n_var)
    print("RI", ri_g, "oracle", ri_o, "i1", pm_g["i1"])
    assert ri_g == ri_o
    assert np.array_equal(pm_g["i1"], pm_o["i1"]) and np.array_equal(pm_g["i2"], pm_o["i2"], equal_nan=True)
    rank = int(min(ri_o, 4))
    cqi_o, pmc_o, inf_o, _ = C.cqi_select(ocfg, re_k, re_l, rank, H, n_var, table)
    cqi_g, pmc_g, inf_g = PH.cqiSelect(carrier, csirs, rc, rank, H, n_var, table)
    print("CQI", cqi_g.ravel()[:8], "oracle", cqi_o.ravel()[:8])
    assert np.array_equal(cqi_g, cqi_o, equal_nan=True)
    assert np.array_equal(pmc_g["i1"], pmc_o["i1"]) and np.array_equal(pmc_g["i2"], pmc_o["i2"], equal_nan=True)
    a, b = inf_g["SINRPerSubbandPerCW"], inf_o["SINRPerSubbandPerCW"]
    assert (np.abs(a - b) / np.abs(b)).max() <= 1e-5
    rk, pmf, cqf = PH.csiReport(carrier, csirs, rc, H, n_var, table, rankCap=4)
    assert rk == rank
    assert np.array_equal(pmf["i1"], pmc_o["i1"]) and np.array_equal(pmf["i2"], pmc_o["i2"], equal_nan=True)
    assert np.array_equal(cqf[:, : cqi_o.shape[1]], cqi_o, equal_nan=True)
    # the same report in two halves, with unrelated work enqueued on the stream in between (isac_csi_report_enqueue_dev/_finish)
    import torch
    pend = PH.csiReportEnqueue(carrier, csirs, rc, H, n_var, table, rankCap=4)
    junk = torch.randn(1 << 22, device="cuda").cumsum(0)
    rk2, pmf2, cqf2 = pend.finish()
    assert rk2 == rk and np.array_equal(pmf2["i1"], pmf["i1"]) and np.array_equal(pmf2["i2"], pmf["i2"], equal_nan=True)
    assert np.array_equal(cqf2, cqf, equal_nan=True) and bool(torch.isfinite(junk[-1]))
    with pytest.raises(Exception):
        pend.finish()                      # nothing pending any more
    # 8-layer CQI has two codewords
    if n_rx == 8:
        cqi_o8, _, _, _ = C.cqi_select(ocfg, re_k, re_l, 8, H, n_var, table)
        cqi_g8, _, _ = PH.cqiSelect(carrier, csirs, rc, 8, H, n_var, table)
        assert cqi_g8.shape[1] == 2 and np.array_equal(cqi_g8, cqi_o8, equal_nan=True)
    # wideband PMI + subband CQI path (cqiSelect.m:586-596)
    rc_w = dict(rc, PMIMode="Wideband")
    ocfg_w = C.report_config(n_ports, panel, 24, 0, 1, "Wideband", "Subband", 4)
    cqi_o, pmw_o, _, _ = C.cqi_select(ocfg_w, re_k, re_l, 1, H, n_var, table)
    cqi_g, pmw_g, _ = PH.cqiSelect(carrier, csirs, rc_w, 1, H, n_var, table)
    assert np.array_equal(cqi_g, cqi_o, equal_nan=True)


@pytest.mark.parametrize("nu,P,R", [(2, 2, 16), (1, 2, 4), (1, 4, 4), (2, 4, 8), (3, 4, 4), (4, 4, 4)])
def test_ul_pmi_select(PH, nu, P, R):
    rng = np.random.default_rng(nu * 10 + P)
    nrb, band = 24, 4
    K = 12 * nrb
    hest = np.zeros((K, 14, R, P), dtype=np.complex64)
    sc = np.arange(1, K, 4)                        # comb-4 SRS in the last symbol (setupSRS.m:11-18)
    hest[sc, 13] = ((rng.standard_normal((sc.size, R, P)) + 1j * rng.standard_normal((sc.size, R, P))) / np.sqrt(2)).astype(np.complex64)
    hest[: 12 * 4] = 0                             # first subband without estimates -> NaN
    pmi_o, sinr_o, idx_o = C.pmi_select(nu, hest, 0.05, band)
    pmi_g, sinr_g, idx_g = PH.pmiSelect(nu, hest, 0.05, band)
    print("UL pmi", pmi_g)
    assert np.array_equal(pmi_g, pmi_o, equal_nan=True)
    assert np.array_equal(idx_g, idx_o)
    m = ~np.isnan(sinr_o)
    assert np.array_equal(m, ~np.isnan(sinr_g))
    assert (np.abs(sinr_g[m] - sinr_o[m]) / np.abs(sinr_o[m])).max() <= 1e-5
    r = PH.pmiSelect(nu, np.zeros_like(hest), 0.05, band)
    assert all(np.isnan(x) for x in r)
    r = PH.pmiSelect(nu, hest, 0.0, band)
    assert all(np.isnan(x) for x in r)


def test_ul_pmi_select_batch(PH):
    """Batched TPMI selection (one launch pair, one sync) equals per-UE pmiSelect calls, including an all-zero estimate."""
    import torch
    rng = np.random.default_rng(77)
    nrb, band, nu, P, R, B = 24, 4, 2, 4, 8, 5
    K = 12 * nrb
    hest = np.zeros((B, K, 1, R, P), dtype=np.complex64)
    sc = np.arange(1, K, 4)
    hest[:, sc, 0] = ((rng.standard_normal((B, sc.size, R, P)) + 1j * rng.standard_normal((B, sc.size, R, P))) / np.sqrt(2)).astype(np.complex64)
    hest[3] = 0                                    # UE without an estimate -> NaN column
    hd = torch.from_numpy(np.ascontiguousarray(hest.transpose(0, 4, 3, 2, 1))).cuda()
    pmi_b, sinr_b, none = PH.pmiSelectBatch(nu, hd, 0.05, band)
    assert list(none) == [0, 0, 0, 1, 0]
    pend = PH.pmiSelectBatchEnqueue(nu, hd, 0.05, band)       # the same report in two halves with unrelated work in between
    junk = torch.randn(1 << 20, device="cuda").sum()
    pmi_e, sinr_e, none_e = pend.finish()
    assert np.array_equal(pmi_e, pmi_b, equal_nan=True) and np.array_equal(sinr_e, sinr_b, equal_nan=True)
    assert list(none_e) == list(none) and bool(torch.isfinite(junk))
    with pytest.raises(Exception):
        pend.finish()
    for b in range(B):
        r = PH.pmiSelect(nu, hest[b], 0.05, band)
        if b == 3:
            assert all(np.isnan(x) for x in r) and np.isnan(pmi_b[:, b]).all()
            continue
        assert np.array_equal(pmi_b[:, b], r[0], equal_nan=True)
        assert np.array_equal(sinr_b[:, :, b], r[1], equal_nan=True)


def test_precoded_sinr(PH):
    """precodedSINR (precodedSINR.m:11-18): single RE and a batch of REs against the float64 oracle."""
    rng = np.random.default_rng(5)
    for nu, P, R in [(1, 2, 4), (2, 4, 8), (4, 4, 16), (3, 8, 8)]:
        H = rng.standard_normal((R, P, 37)) + 1j * rng.standard_normal((R, P, 37))
        W = (rng.standard_normal((P, nu)) + 1j * rng.standard_normal((P, nu))) / np.sqrt(2 * P)
        ref = np.array([C.precoded_sinr_ul(H[:, :, b], 0.3, W) for b in range(37)])
        got = PH.precodedSINR(H, 0.3, W)
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
        assert abs(PH.precodedSINR(H[:, :, 0], 0.3, W) - ref[0]) <= 1e-9 * abs(ref[0])


def test_prg_precode(PH):
    rng = np.random.default_rng(8)
    nrb, L, nu, P, nprg = 24, 14, 2, 8, 6
    K = 12 * nrb
    # PDSCH-like allocation: PRBs 3..20, symbols 2..12, same REs on every layer
    k = np.arange(3 * 12, 21 * 12)
    l = np.arange(2, 13)
    pos = (k[:, None] + K * l[None, :]).reshape(-1, order="F")
    portind = np.stack([pos + 1 + K * L * j for j in range(nu)], axis=1)
    portsym = (rng.standard_normal(portind.shape) + 1j * rng.standard_normal(portind.shape)).astype(np.complex64)
    F = (rng.standard_normal((nu, P, nprg)) + 1j * rng.standard_normal((nu, P, nprg))).astype(np.complex64)
    for nstart in (0, 2):
        s_o, i_o = C.prg_precode((K, L, P), nstart, portsym, portind, F)
        s_g, i_g = PH.prgPrecode((K, L, P), nstart, portsym, portind, F)
        assert np.array_equal(i_g, i_o)
        err = np.abs(s_g - s_o).max() / np.abs(s_o).max()
        print("prgPrecode err", err)
        assert err <= 1e-5
    # layer columns that do not share their RE order (second layer reversed, a few of its REs missing): the grid
    # semantics of prgPrecode.m:131-144 still hold (slow path of the kernel)
    small = pos[:40]
    pi2 = np.stack([small + 1, small[::-1] + 1 + K * L], axis=1)
    pi2[3, 1] = pi2[4, 1]                           # duplicate index: the later symbol wins, RE of row 36 gets no layer-2 symbol
    ps2 = (rng.standard_normal(pi2.shape) + 1j * rng.standard_normal(pi2.shape)).astype(np.complex64)
    s_o, i_o = C.prg_precode((K, L, P), 0, ps2, pi2, F)
    s_g, i_g = PH.prgPrecode((K, L, P), 0, ps2, pi2, F)
    assert np.array_equal(i_g, i_o)
    assert np.abs(s_g - s_o).max() <= 1e-5 * np.abs(s_o).max()
