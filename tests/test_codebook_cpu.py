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
