"""CPU-only: the C++ Type-I / PUSCH codebook generators (csrc/codebook.cu, pure host code behind the C ABI)
against the oracle, and analytic known answers for the oracle itself (TS 38.214 5.2.2.2.1)."""
import importlib

import numpy as np
import pytest

from oracle import comm as C

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
PANELS = [(2, (1, 1)), (4, (2, 1)), (8, (2, 2)), (8, (4, 1)), (12, (3, 2)), (12, (6, 1)), (16, (4, 2)), (16, (8, 1)),
          (24, (4, 3)), (24, (6, 2)), (24, (12, 1)), (32, (4, 4)), (32, (8, 2)), (32, (16, 1))]


@pytest.mark.parametrize("n_ports,panel", PANELS)
def test_type1_codebook_matches_oracle_and_is_orthonormal(n_ports, panel):
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    for mode in (1, 2):
        ocfg = C.report_config(n_ports, panel if n_ports > 2 else None, 52, 0, mode, subband_size=4)
        for nu in range(1, 9):
            if nu > n_ports or (n_ports == 2 and nu > 2):
                continue
            for var, name in ((0, "ue"), (1, "gnb")):
                Wo = C.type1_single_panel_codebook(ocfg, nu, name)
                rc = {"PanelDimensions": panel, "CodebookMode": mode, "NSizeBWP": 52, "NStartBWP": 0, "NumCSIRSPorts": n_ports}
                Wg = ph._codebook(rc, nu, var)
                assert Wo.shape == Wg.shape
                assert np.abs(Wo - Wg).max() <= 1e-14, (n_ports, panel, mode, nu, name)
            # W'W = I/nu for every unrestricted UE-side precoder
            Wf = Wo.reshape(Wo.shape[0], nu, -1) if name == "gnb" else None
            Wu = C.type1_single_panel_codebook(ocfg, nu, "ue")
            Wf = Wu.reshape(Wu.shape[0], nu, -1)
            G = np.einsum("pic,pjc->ijc", Wf.conj(), Wf)
            assert np.abs(G - np.eye(nu)[:, :, None] / nu).max() <= 1e-14


def test_codebook_restrictions():
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    rng = np.random.default_rng(0)
    csr = (rng.random(64) > 0.3).astype(int)
    i2r = (rng.random(16) > 0.3).astype(int)
    for mode, nu in ((1, 1), (1, 2), (2, 1), (2, 2), (1, 4), (1, 8)):
        ocfg = C.report_config(8, (2, 2), 52, 0, mode, subband_size=4, subset_restriction=csr, i2_restriction=i2r)
        Wo = C.type1_single_panel_codebook(ocfg, nu, "ue")
        rc = {"PanelDimensions": (2, 2), "CodebookMode": mode, "NSizeBWP": 52, "NStartBWP": 0, "NumCSIRSPorts": 8,
              "CodebookSubsetRestriction": csr, "i2Restriction": i2r}
        Wg = ph._codebook(rc, nu, 0)
        assert np.abs(Wo - Wg).max() <= 1e-14
        assert (np.abs(Wo).sum(axis=(0, 1)) == 0).any() and (np.abs(Wo).sum(axis=(0, 1)) > 0).any()


def test_rank1_codeword_closed_form():
    """W = [v_lm ; phi_n v_lm]/sqrt(P), v_lm = u_l (x) u_m with N2 fastest."""
    cfg = C.report_config(16, (4, 2), 52, 0, 1, subband_size=4)
    W = C.type1_single_panel_codebook(cfg, 1, "ue")
    N1, N2, O1, O2 = 4, 2, 4, 4
    for (i2, l, m) in ((0, 0, 0), (1, 5, 3), (3, 15, 7), (2, 9, 1)):
        ul = np.exp(2j * np.pi * l * np.arange(N1) / (O1 * N1))
        um = np.exp(2j * np.pi * m * np.arange(N2) / (O2 * N2))
        v = np.kron(ul, um)
        ref = np.concatenate([v, np.exp(1j * np.pi * i2 / 2) * v]) / 4.0
        assert np.abs(W[:, 0, i2, l, m, 0] - ref).max() < 1e-15


def test_pusch_codebooks_are_consistent():
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    for nu, P in ((1, 1), (1, 2), (2, 2), (1, 4), (2, 4), (3, 4), (4, 4)):
        W = ph.puschCodebook(nu, P)
        assert W.shape[2] == C.max_pusch_tpmi(nu, P) + 1 == ph.maxPUSCHPrecodingMatrixIndicator(nu, P) + 1
        for t in range(W.shape[2]):
            assert np.abs(W[:, :, t] - C.pusch_codebook(nu, P, t).reshape(P, nu)).max() < 1e-15
            # every PUSCH precoder has total power <= 1 and orthogonal columns
            G = W[:, :, t].conj().T @ W[:, :, t]
            assert np.abs(G - np.diag(np.diag(G))).max() < 1e-15
            assert np.trace(G).real <= 1 + 1e-12
        assert len({W[:, :, t].tobytes() for t in range(W.shape[2])}) == W.shape[2]


def test_precoded_sinr_closed_forms():
    rng = np.random.default_rng(4)
    H = rng.standard_normal((4, 8)) + 1j * rng.standard_normal((4, 8))
    cfg = C.report_config(8, (2, 2), 52, 0, 1, subband_size=4)
    W1 = C.type1_single_panel_codebook(cfg, 1, "ue")[:, :, 2, 3, 5, 0]
    assert abs(C.precoded_sinr_dl(H, 0.37, W1)[0] - np.linalg.norm(H @ W1) ** 2 / 0.37) < 1e-9        # rank 1: |Hw|^2/nVar
    W4 = C.type1_single_panel_codebook(cfg, 4, "ue")[:, :, 1, 2, 6, 1]
    A = np.linalg.inv(np.eye(4) + W4.conj().T @ H.conj().T @ H @ W4 / 0.37)
    assert np.abs(C.precoded_sinr_dl(H, 0.37, W4) - (1 / np.real(np.diag(A)) - 1)).max() < 1e-9
    assert abs(C.precoded_sinr_ul(H, np.sqrt(0.37), W4) - np.sum(1 / np.real(np.diag(A)) - 1)) < 1e-9


def test_matlab_round4_and_subbands():
    assert np.array_equal(C.matlab_round4([0.00005, -0.00005, 1.23456, 2.5e-5]), [0.0001, -0.0001, 1.2346, 0.0])
    assert C.subband_info("Subband", 0, 52, 4) == (13, [4] * 13)
    assert C.subband_info("Subband", 3, 50, 8) == (7, [5, 8, 8, 8, 8, 8, 5])
    assert C.subband_info("Wideband", 0, 52, 4) == (1, [52])
    assert C.subband_info("Subband", 0, 20, 4) == (1, [20])


# ---- Type-I multi-panel codebooks (dlPMISelect.m:1351-1772) ---------------------------------------------------------------
_MP_CASES = [((2, 2, 1), 1, 1), ((2, 2, 1), 1, 2), ((2, 2, 1), 2, 1), ((2, 2, 1), 2, 3), ((2, 4, 1), 1, 4), ((2, 2, 2), 2, 2),
             ((4, 2, 1), 1, 1), ((4, 2, 1), 1, 3), ((4, 2, 2), 1, 2), ((2, 4, 2), 2, 4), ((2, 8, 1), 1, 3)]


@pytest.mark.parametrize("panel,mode,nu", _MP_CASES)
def test_multi_panel_codebook_host_builder_equals_oracle(panel, mode, nu):
    import importlib
    com = importlib.import_module(PKG + ".communication")
    Ng, N1, N2 = panel
    O1, O2 = com._MP_PANELS[panel]
    rng = np.random.default_rng(sum(panel) + 10 * mode + nu)
    csr = (rng.random(N1 * O1 * N2 * O2) > 0.15).astype(int)
    csr[rng.integers(csr.size)] = 0                    # at least one restricted beam
    csr[(np.flatnonzero(csr == 0)[0] + 1) % csr.size] = 1   # and at least one allowed
    W = com.pmiType1MultiPanelCodebook({"PanelDimensions": panel, "CodebookMode": mode, "CodebookSubsetRestriction": csr}, nu)
    Wo = C.type1_multi_panel_codebook({"N1": N1, "N2": N2, "O1": O1, "O2": O2, "CodebookMode": mode,
                                        "CodebookSubsetRestriction": csr}, Ng, nu)
    assert W.shape == Wo.shape
    assert np.abs(W - Wo).max() <= 1e-14
    # the beam / co-phasing table the SINR kernels read (2*Ng blocks per column, flattened index set) materialises to the same array
    Wt = com.pmiType1MultiPanelCodebook({"PanelDimensions": panel, "CodebookMode": mode, "CodebookSubsetRestriction": csr}, nu,
                                        from_table=True)
    assert Wt.shape == Wo.shape and np.abs(Wt - Wo).max() <= 1e-14
    # restricted precoders are all zero, every other one has orthogonal columns of power 1/nu
    G = np.einsum("pa...,pb...->ab...", W.conj(), W)
    live = np.abs(W).sum(axis=(0, 1)) > 0
    assert live.any() and (~live).any()
    for a in range(nu):
        for b in range(nu):
            assert np.allclose(G[a, b][live], 1.0 / nu if a == b else 0.0, atol=1e-13)


def test_multi_panel_codebook_known_matrices():
    """The written-out matrices of the reference for single index sets (dlPMISelect.m:1447-1470, :1590-1607, :1680-1700)."""
    phi = lambda x: np.exp(1j * np.pi * x / 2)
    a = lambda x: np.exp(1j * np.pi / 4 + 1j * np.pi * x / 2)
    b = lambda x: np.exp(-1j * np.pi / 4 + 1j * np.pi * x / 2)
    # Ng = 2, mode 1, two layers (Table 5.2.2.2.2-4)
    cfg = {"N1": 2, "N2": 2, "O1": 4, "O2": 4, "CodebookMode": 1}
    W = C.type1_multi_panel_codebook(cfg, 2, 2)
    i20, i11, i12, i13, i141 = 1, 5, 3, 3, 2
    v, vp = C._vlm(2, 2, 4, 4, i11, i12), C._vlm(2, 2, 4, 4, i11 + 4, i12 + 4)      # N1 == N2: k1 = [0 O1 0 O1], k2 = [0 0 O2 O2]
    fn, fp = phi(i20), phi(i141)
    ref = np.block([[v[:, None], vp[:, None]], [fn * v[:, None], -fn * vp[:, None]], [fp * v[:, None], fp * vp[:, None]],
                    [fn * fp * v[:, None], -fn * fp * vp[:, None]]]) / np.sqrt(2 * 16)
    assert np.abs(W[:, :, i20, 0, 0, i11, i12, i13, i141, 0, 0] - ref).max() <= 1e-15
    # Ng = 4, mode 1, one layer (Table 5.2.2.2.2-3)
    cfg = {"N1": 2, "N2": 1, "O1": 4, "O2": 1, "CodebookMode": 1}
    W = C.type1_multi_panel_codebook(cfg, 4, 1)
    i20, i11, p1, p2, p3 = 3, 6, 1, 2, 3
    v = C._vlm(2, 1, 4, 1, i11, 0)
    fn = phi(i20)
    ref = np.concatenate([v, fn * v, phi(p1) * v, fn * phi(p1) * v, phi(p2) * v, fn * phi(p2) * v, phi(p3) * v, fn * phi(p3) * v]) / 4.0
    assert np.abs(W[:, 0, i20, 0, 0, i11, 0, 0, p1, p2, p3] - ref).max() <= 1e-15
    # Ng = 2, mode 2, three layers (Table 5.2.2.2.2-5)
    cfg = {"N1": 4, "N2": 1, "O1": 4, "O2": 1, "CodebookMode": 2}
    W = C.type1_multi_panel_codebook(cfg, 2, 3)
    i20, i21, i22, i11, i13, p1, p2 = 1, 1, 0, 9, 2, 3, 1
    v, vp = C._vlm(4, 1, 4, 1, i11, 0), C._vlm(4, 1, 4, 1, i11 + 3 * 4, 0)          # (N1,N2) = (4,1): k1 = O1*(1:3)
    fn, c1, c2 = phi(i20), a(p1) * b(i21), a(p2) * b(i22)
    col = lambda x, s: np.concatenate([x, s * fn * x, c1 * x, s * c2 * x])
    ref = np.stack([col(v, 1), col(vp, 1), col(v, -1)], axis=1) / np.sqrt(3 * 16)
    assert np.abs(W[:, :, i20, i21, i22, i11, 0, i13, p1, p2, 0] - ref).max() <= 1e-15
    assert W.shape == (16, 3, 2, 2, 2, 16, 1, 3, 4, 4, 1)
    with pytest.raises(ValueError):
        C.type1_multi_panel_codebook(cfg, 4, 1)        # mode 2 exists for two panels only
