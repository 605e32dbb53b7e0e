"""dlPMISelect with Type1MultiPanel codebooks (dlPMISelect.m:1351-1772, selection :385-501) on the GPU against the oracle.

The plan keeps the reference's 9-D index set flattened in MATLAB linear order and runs the direct SINR kernel on a beam /
co-phasing table with 2*Ng blocks per column (the table itself is checked on the CPU, tests/test_codebook_cpu.py).
Tolerance: SINR arrays 1e-5 relative (observed ~1e-13), PMISet exact or tie-equivalent at the reference's 4-decimal rounding."""
import importlib

import numpy as np
import pytest

from oracle import comm as C

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("panel,mode,nu,n_rx", [((2, 2, 1), 1, 1, 2), ((2, 2, 1), 1, 2, 4), ((2, 2, 1), 2, 3, 4), ((4, 2, 1), 1, 2, 4),
                                               ((2, 2, 2), 1, 4, 4), ((2, 4, 1), 2, 1, 2)])
def test_dl_pmi_select_multi_panel(gpu, panel, mode, nu, n_rx):
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    Ng, N1, N2 = panel
    O1, O2 = ph._MP_PANELS[panel]
    P = 2 * Ng * N1 * N2
    nrb, sbs = 24, 8
    K = 12 * nrb
    rng = np.random.default_rng(1000 + P + 10 * mode + nu)
    H = ((rng.standard_normal((K, 14, n_rx, P)) + 1j * rng.standard_normal((K, 14, n_rx, P))) / np.sqrt(2)).astype(np.complex64)
    H = (H + np.roll(H, 1, axis=0) + np.roll(H, 2, axis=0)).astype(np.complex64)
    csr = np.ones(N1 * O1 * N2 * O2, dtype=np.uint8)
    csr[3] = 0                                                        # one restricted beam -> NaN entries
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": P, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"CodebookType": "Type1MultiPanel", "PanelDimensions": panel, "CodebookMode": mode, "NSizeBWP": nrb, "NStartBWP": 0,
          "PMIMode": "Subband", "CQIMode": "Wideband", "SubbandSize": sbs, "CodebookSubsetRestriction": csr}
    ocfg = {"N1": N1, "N2": N2, "O1": O1, "O2": O2, "CodebookMode": mode, "NSizeBWP": nrb, "NStartBWP": 0, "PMIMode": "Subband",
            "SubbandSize": sbs, "CodebookSubsetRestriction": csr, "i2Restriction": np.ones(16)}
    re_k, re_l = C.csirs_first_port_res(nrb, 1, 0)
    pm_o, info_o = C.dl_pmi_select_multi_panel(ocfg, Ng, re_k, re_l, nu, H, 0.1)
    pm_g, info_g = ph.dlPMISelect(carrier, csirs, rc, nu, H, 0.1)
    So, Sg = info_o["SINRPerRE"], info_g["SINRPerRE"]
    assert So.shape == Sg.shape and So.ndim == 11
    assert np.array_equal(np.isnan(So), np.isnan(Sg)) and np.isnan(So).any()
    m = ~np.isnan(So)
    err = (np.abs(Sg[m] - So[m]) / np.abs(So[m])).max()
    print(f"{panel} mode {mode} nu {nu}: SINRPerRE {So.shape}, max rel err {err:.2e}; i1 {pm_g['i1']} i2 {pm_g['i2'][:, 0]}")
    assert err <= 1e-5
    Bo, Bg = info_o["SINRPerSubband"], info_g["SINRPerSubband"]
    mb = ~np.isnan(Bo)
    assert np.array_equal(mb, ~np.isnan(Bg)) and (np.abs(Bg[mb] - Bo[mb]) / np.abs(Bo[mb])).max() <= 1e-5
    assert np.abs(info_g["W"] - info_o["W"]).max() <= 1e-14
    assert pm_g["i1"].shape == (6,) and pm_g["i2"].shape == pm_o["i2"].shape
    total = C.matlab_round4(np.nansum(So, axis=(0, 1)))               # [i20 i21 i22 i11 i12 i13 i141 i142 i143]
    g1 = pm_g["i1"].astype(int) - 1
    if not np.array_equal(pm_g["i1"], pm_o["i1"]):                    # tie at the 4-decimal rounding: same rounded maximum
        assert np.isclose(total[:, :, :, g1[0], g1[1], g1[2], g1[3], g1[4], g1[5]].max(), total.max(), rtol=0, atol=1e-4)
    else:
        for sb in range(pm_o["i2"].shape[1]):
            if np.array_equal(pm_g["i2"][:, sb], pm_o["i2"][:, sb], equal_nan=True):
                continue
            t = C.matlab_round4(np.nansum(Bo[sb][:, :, :, :, g1[0], g1[1], g1[2], g1[3], g1[4], g1[5]], axis=0))
            a = pm_g["i2"][:, sb].astype(int) - 1
            assert abs(t[a[0], a[1], a[2]] - t.max()) <= 1e-4, sb


@pytest.mark.parametrize("panel,mode,n_rx", [((2, 2, 1), 1, 4), ((2, 2, 2), 1, 4), ((2, 4, 1), 2, 2)])
def test_ri_cqi_and_report_multi_panel(gpu, panel, mode, n_rx):
    """riSelect / cqiSelect / the fused report over Type1MultiPanel codebooks (riSelect.m:222-231,254-285 with ranks <= 4;
    cqiSelect.m:507,:599-603): RI, the 6-entry i1 / 3-row i2 PMISet and the CQI against the oracle."""
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    Ng, N1, N2 = panel
    O1, O2 = ph._MP_PANELS[panel]
    P = 2 * Ng * N1 * N2
    nrb, sbs = 24, 8
    K = 12 * nrb
    rng = np.random.default_rng(2000 + P + 10 * mode + n_rx)
    H = ((rng.standard_normal((K, 14, n_rx, P)) + 1j * rng.standard_normal((K, 14, n_rx, P))) / np.sqrt(2)).astype(np.complex64)
    H = (H + np.roll(H, 1, axis=0)).astype(np.complex64)
    carrier = {"NSizeGrid": nrb, "NStartGrid": 0, "SymbolsPerSlot": 14}
    csirs = {"NumCSIRSPorts": P, "NumRB": nrb, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0, "Density": "one"}
    rc = {"CodebookType": "Type1MultiPanel", "PanelDimensions": panel, "CodebookMode": mode, "NSizeBWP": nrb, "NStartBWP": 0,
          "PMIMode": "Subband", "CQIMode": "Subband", "SubbandSize": sbs}
    ocfg = {"N1": N1, "N2": N2, "O1": O1, "O2": O2, "CodebookMode": mode, "NSizeBWP": nrb, "NStartBWP": 0, "PMIMode": "Subband",
            "CQIMode": "Subband", "SubbandSize": sbs, "CodebookSubsetRestriction": np.ones(N1 * O1 * N2 * O2), "i2Restriction": np.ones(16),
            "RIRestriction": np.ones(4), "NumCSIRSPorts": P}
    re_k, re_l = C.csirs_first_port_res(nrb, 1, 0)
    table = np.array([-3.46, 1.54, 6.54, 11.05, 13.54, 16.04, 17.54, 20.04, 22.04, 24.43, 26.93, 27.43, 29.43, 32.43, 35.43])
    nvar = 0.05
    ri_o, pm_o = C.ri_select(ocfg, re_k, re_l, H, nvar, n_panels=Ng)
    ri_g, pm_g = ph.riSelect(carrier, csirs, rc, H, nvar)
    assert ri_g == ri_o and 1 <= ri_o <= min(n_rx, 4)
    assert pm_g["i1"].shape == (6,) and pm_g["i2"].shape == pm_o["i2"].shape
    assert np.array_equal(pm_g["i1"], pm_o["i1"]) and np.array_equal(pm_g["i2"], pm_o["i2"], equal_nan=True)
    for nu in range(1, min(n_rx, 4) + 1):
        cq_o, pmc_o, ci_o, _ = C.cqi_select(ocfg, re_k, re_l, nu, H, nvar, table, n_panels=Ng)
        cq_g, pmc_g, ci_g = ph.cqiSelect(carrier, csirs, rc, nu, H, nvar, table)
        assert np.array_equal(pmc_g["i1"], pmc_o["i1"]) and np.array_equal(pmc_g["i2"], pmc_o["i2"], equal_nan=True), nu
        assert np.array_equal(cq_g, cq_o, equal_nan=True), (nu, cq_g, cq_o)
        ref = ci_o["SINRPerSubbandPerCW"]
        assert np.nanmax(np.abs(ci_g["SINRPerSubbandPerCW"] - ref) / np.abs(ref)) <= 1e-5
    rk, pmr, cqr = ph.csiReport(carrier, csirs, rc, H, nvar, table, rankCap=4)
    cq_o, pmc_o, _, _ = C.cqi_select(ocfg, re_k, re_l, int(ri_o), H, nvar, table, n_panels=Ng)
    assert rk == ri_o and np.array_equal(pmr["i1"], pmc_o["i1"]) and np.array_equal(pmr["i2"], pmc_o["i2"], equal_nan=True)
    assert np.array_equal(cqr[:, : cq_o.shape[1]], cq_o, equal_nan=True)
    print(f"{panel} mode {mode}: RI {ri_g}, i1 {pm_g['i1']}, wideband CQI {cqr[0, 0]}")
