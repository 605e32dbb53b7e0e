"""BASELINE config 3 at full size through the report path: 32 CSI-RS ports (4,4) -- the 64-element array virtualised onto the
largest Type-I port count (dlPMISelect.m:623-626) -- 8 receive antennas, 273 PRB, 16-PRB subbands, EVERY rank 1..min(R,P) = 8
scored by riSelect (riSelect.m:254-285), cqiSelect at ranks 3 and 4 (the >= 16-port codebooks with the i13 / theta_p index,
dlPMISelect.m:1179-1210) and the fused report, against the vectorised float64 oracle (itself checked against the loop-faithful
oracle at small sizes in tests/test_oracle_cpu.py).  The reported PMI is then looked up in the gNB-side codebook copy
(pmiType1SinglePanelCodebook.m:348,:358): for ranks 3-4 that copy keeps only slice i13 = 1 -- the reference's scheduler
precodes with a different matrix than the UE selected, reproduced here, not fixed."""
import importlib

import numpy as np
import pytest

from oracle import comm as C
from test_cfg23_gpu import TABLE, _cfg, _channel

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


def _diff_cqi(cq_abs):
    d = cq_abs[1:] - cq_abs[:1]
    off = np.where(np.isnan(d), np.nan, np.where(d == 0, 0, np.where(d == 1, 1, np.where(d >= 2, 2, 3))))
    return np.vstack([cq_abs[:1], off])            # report format of cqiSelect.m:656-677


@pytest.fixture(scope="module")
def case(gpu):
    carrier, csirs, rc, ocfg, re_k, re_l = _cfg(32, (4, 4), 273, 16)
    rng = np.random.default_rng(3300)
    B, R = 2, 8
    Hs = [_channel(rng, 273 * 12, R, 32, taps=6) for _ in range(B)]
    nvar = np.array([0.02, 0.2])
    oracle = [C.csi_report_vectorized(ocfg, re_k, re_l, Hs[b], nvar[b], TABLE, rank_cap=8, return_all=True) for b in range(B)]
    return carrier, csirs, rc, Hs, nvar, oracle


def _same_pmi(got, ref, what):
    assert np.array_equal(got["i1"], ref["i1"]), (what, got["i1"], ref["i1"])
    assert np.array_equal(got["i2"], ref["i2"], equal_nan=True), (what, got["i2"], ref["i2"])


def test_cfg3_ri_select_all_eight_ranks(case):
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    carrier, csirs, rc, Hs, nvar, oracle = case
    RI, pm = ph.riSelect(carrier, csirs, rc, np.stack(Hs, axis=-1), nvar)
    for b, (rank, pmo, cqo, ri_o, keep) in enumerate(oracle):
        assert sorted(keep) == list(range(1, 9))                       # the oracle scored ranks 1..8
        assert RI[b] == ri_o, (b, RI[b], ri_o)
        _same_pmi({"i1": pm["i1"][:, b], "i2": pm["i2"][:, b]}, keep[int(ri_o)][0], f"riSelect UE {b}")
    print("cfg3 RI", RI)


@pytest.mark.parametrize("nu", [3, 4])
def test_cfg3_cqi_select_ranks_with_i13(case, nu):
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    cm = importlib.import_module(PKG + ".communication")
    carrier, csirs, rc, Hs, nvar, oracle = case
    cq, pm, info = ph.cqiSelect(carrier, csirs, rc, nu, np.stack(Hs, axis=-1), nvar, TABLE)
    Wue = ph._codebook(dict(rc, OverSamplingFactors=(4, 4)), nu, variant=0)
    Wgnb = cm.pmiType1SinglePanelCodebook(dict(rc, OverSamplingFactors=(4, 4)), nu)
    assert Wue.shape == Wgnb.shape and Wue.shape[5] == 4                # i13 = theta_p index, 4 values
    assert not np.any(Wgnb[..., 1:]) and np.any(Wue[..., 1:])           # the gNB copy's slices i13 > 1 stay all-zero (:348,:358)
    for b in range(len(Hs)):
        pmo, sel = oracle[b][4][nu]
        _same_pmi({"i1": pm["i1"][:, b], "i2": pm["i2"][:, b]}, pmo, f"cqiSelect nu={nu} UE {b}")
        cqo = _diff_cqi(C.cqi_from_subband_sinr(sel, nu, TABLE))
        assert np.array_equal(cq[:, : cqo.shape[1], b], cqo, equal_nan=True), (b, cq[:, 0, b], cqo[:, 0])
        sb = info["SINRPerSubbandPerCW"][1:, 0, b]
        ref = np.nansum(sel, axis=1)
        assert np.nanmax(np.abs(sb - ref) / np.abs(ref)) <= 1e-5
        # what the reference's scheduler would precode with (schedulerEntity.m:736-777): the gNB copy at the reported indices
        i1 = pm["i1"][:, b].astype(int) - 1
        i2 = int(pm["i2"][0, b]) - 1
        w_ue, w_gnb = Wue[:, :, i2, i1[0], i1[1], i1[2]], Wgnb[:, :, i2, i1[0], i1[1], i1[2]]
        assert np.allclose(np.linalg.norm(w_ue), 1.0)
        if i1[2] > 0:
            assert not np.any(w_gnb)                                     # reported theta_p index > 1: the gNB copy holds zeros there


def test_cfg3_fused_report_rank_cap_8(case):
    ph = importlib.import_module(PKG + ".communication.phyLayer")
    carrier, csirs, rc, Hs, nvar, oracle = case
    RI, pm, cq = ph.csiReport(carrier, csirs, rc, np.stack(Hs, axis=-1), nvar, TABLE, rankCap=8)
    for b, (rank, pmo, cqo, ri_o, keep) in enumerate(oracle):
        assert RI[b] == rank
        _same_pmi({"i1": pm["i1"][:, b], "i2": pm["i2"][:, b]}, pmo, f"report UE {b}")
        cqo = _diff_cqi(cqo)
        assert np.array_equal(cq[:, : cqo.shape[1], b], cqo, equal_nan=True), (b, cq[:, :, b], cqo)
