"""GPU parity of the link-budget kernels (csrc/link.cu) against oracle/link.py: path loss of a batch of links (float64, 1e-12),
path-loss / Rx-gain scaling of device-resident channel matrices (fp32 rounding)."""
import importlib

import numpy as np
import pytest

from oracle import link as L

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


def test_pathloss_batch_matches_oracle(gpu):
    pl = importlib.import_module(PKG + ".communication.pathlossModels")
    rng = np.random.default_rng(4)
    n = 3000
    bs = np.column_stack([rng.uniform(-500, 500, n), rng.uniform(-500, 500, n), rng.uniform(10, 40, n)])
    ue = np.column_stack([rng.uniform(-2000, 2000, n), rng.uniform(-2000, 2000, n), rng.uniform(1.2, 22, n)])
    ue[7] = bs[7]                                                       # identical positions -> 0 dB
    los = rng.integers(0, 2, n)
    for scn in ("UMa", "UMi", "RMa", "InH"):
        got = pl.config5GNRModels(scn, 3.5e9, los, bs, ue)
        ref = np.array([L.path_loss(scn, 3.5e9, int(los[i]), bs[i], ue[i]) for i in range(n)])
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max(), scn
        assert got[7] == 0.0
    got = pl.configFreeSpaceModel(3.5e9, bs, ue)
    ref = np.array([L.path_loss("fspl", 3.5e9, 1, bs[i], ue[i]) for i in range(n)])
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
    one = pl.config5GNRModels("UMa", 3.5e9, 1, [0, 0, 25.0], [100.0, 0, 1.5])     # the reference's one-link call
    assert isinstance(one, float) and abs(one - L.path_loss("UMa", 3.5e9, 1, [0, 0, 25.0], [100.0, 0, 1.5])) < 1e-10
    with pytest.raises(Exception):
        pl.config5GNRModels("InF-SL", 3.5e9, 1, bs[:1], ue[:1])


def test_link_budget_scales_channel_matrices(gpu):
    import torch
    pl = importlib.import_module(PKG + ".communication.pathlossModels")
    g = torch.Generator(device="cuda").manual_seed(3)
    H = torch.view_as_complex(torch.randn(5, 4, 2, 14, 288, 2, device="cuda", generator=g)).contiguous()
    ref = H.cpu().numpy().astype(np.complex128)
    pld = np.array([80.0, 95.5, 110.25, 70.0, 133.0])
    out = pl.applyPathLossAndRxGain(H, pld, 25.5)
    for i in range(5):
        r = L.apply_link_budget(ref[i], pld[i], 25.5)
        assert np.abs(out[i].cpu().numpy() - r).max() <= 2e-7 * np.abs(r).max()
