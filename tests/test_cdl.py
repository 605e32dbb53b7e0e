"""K11 CDL generator: host ray tables vs the NumPy restatement (CPU), known answers of TR 38.901, and the GPU
frequency response vs the oracle synthesis from the same rays."""
import importlib

import numpy as np
import pytest

from oracle import cdl as OC

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"


def test_oracle_pdp_and_k_factor():
    for prof in (0, 1, 2, 3, 4):
        r = OC.build_rays(prof, 300e-9, 5.0, (1, 1, 1), (1, 1, 1), False, False, 7)
        # single vertical isotropic element: |g_m|^2 summed per cluster == normalised cluster power
        pw = np.array([np.sum(np.abs(r["g"][r["cluster"] == n][:20]) ** 2) for n in range(r["tau"].size)])
        assert np.allclose(pw, r["power"], rtol=1e-12)
        assert abs(r["power"].sum() + r["plos"] - 1.0) < 1e-12
        assert np.all(np.abs(r["nu"]) <= 5.0 + 1e-12)
    d = OC.build_rays(3, 300e-9, 5.0, (1, 1, 1), (1, 1, 1), False, False, 1)
    assert abs(10 * np.log10(d["plos"] / d["power"][0]) - 13.3) < 1e-9          # CDL-D K-factor (Table 7.7.1-4)
    assert abs(np.abs(d["g"][-1, 0, 0]) ** 2 - d["plos"]) < 1e-12
    e = OC.build_rays(4, 300e-9, 5.0, (1, 1, 1), (1, 1, 1), False, False, 1)
    assert abs(10 * np.log10(e["plos"] / e["power"][0]) - 22.0) < 1e-9          # CDL-E K-factor (Table 7.7.1-5)
    assert OC.CDL_A.shape == (23, 6) and OC.CDL_C.shape == (24, 6) and OC.CDL_D.shape == (13, 6)
    assert OC.CDL_B.shape == (23, 6) and OC.CDL_E.shape == (14, 6)
    # the strongest cluster of every NLOS table is at 0 dB; mean delays (power-weighted, normalised) of the tables
    assert OC.CDL_A[:, 1].max() == 0 and OC.CDL_B[:, 1].max() == 0 and OC.CDL_C[:, 1].max() == 0
    assert abs(OC.pattern38901(90.0, 0.0) - 10 ** 0.8) < 1e-12 and abs(OC.pattern38901(90.0, 180.0) - 10 ** (-2.2)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("prof,name,tx,rx", [(3, "CDL-D", (1, 4, 2), (1, 1, 2)), (0, "CDL-A", (1, 8, 2), (1, 1, 2)),
                                             (2, "CDL-C", (1, 4, 2), (2, 2, 2)), (1, "CDL-B", (1, 4, 2), (1, 1, 2)),
                                             (4, "CDL-E", (1, 2, 2), (1, 2, 2))])
def test_cdl_rays_and_frequency_response(gpu, prof, name, tx, rx):
    cm = importlib.import_module(PKG + ".communication.channelModels")
    ch = cm.CDLChannel(DelayProfile=name, TransmitAntennaArraySize=tx, ReceiveAntennaArraySize=rx, Seed=1234)
    rays = ch.rays()
    ref = OC.build_rays(prof, 300e-9, 5.0, tx, rx, True, False, 1234)
    assert np.array_equal(rays["cluster"], ref["cluster"])
    assert np.allclose(rays["tau"], ref["tau"], rtol=1e-14, atol=0)
    assert np.allclose(rays["nu"], ref["nu"], rtol=1e-12, atol=1e-12)
    assert np.abs(rays["g"] - ref["g"]).max() <= 1e-13 * np.abs(ref["g"]).max()
    K, scs = 24 * 12, 30e3
    t = 0.0135 + np.arange(14) * 35.7e-6
    H = ch.generate(K, scs, t - 0.0135, t0=0.0135).cpu().numpy().transpose(3, 2, 1, 0)     # -> [K, L, nRx, nTx]
    Href = OC.frequency_response(ref, K, scs, t)
    err = np.abs(H - Href).max() / np.sqrt(np.mean(np.abs(Href) ** 2))
    print(name, "H err / rms", err, "mean |H|^2", np.mean(np.abs(Href) ** 2))
    assert err <= 1e-5
    ch.setKernel(True)   # legacy mma.sync kernel: same 3xTF32 arithmetic as the default tcgen05 kernel
    H2 = ch.generate(K, scs, t - 0.0135, t0=0.0135).cpu().numpy().transpose(3, 2, 1, 0)
    err2 = np.abs(H2 - Href).max() / np.sqrt(np.mean(np.abs(Href) ** 2))
    print(name, "legacy kernel H err / rms", err2, " tcgen05 vs legacy", np.abs(H2 - H).max())
    assert err2 <= 1e-5
    ch.close()


@pytest.mark.gpu
def test_cdl_statistics_over_seeds(gpu):
    """Average channel power over realisations ~ (sum of path powers) * mean element gain / nRx: for a single
    isotropic vertical pair it is 1 (NormalizePathGains, NormalizeChannelOutputs)."""
    cm = importlib.import_module(PKG + ".communication.channelModels")
    acc = []
    for seed in range(24):
        ch = cm.CDLChannel(DelayProfile="CDL-A", TransmitAntennaArraySize=(1, 1, 1), ReceiveAntennaArraySize=(1, 1, 1),
                           TransmitElement="isotropic", Seed=seed)
        H = ch.generate(52 * 12, 15e3, np.arange(14) * 71.4e-6).cpu().numpy()
        acc.append(np.mean(np.abs(H) ** 2))
        ch.close()
    m = float(np.mean(acc))
    print("mean channel power over seeds", m)
    assert 0.8 < m < 1.2


@pytest.mark.gpu
def test_cdl_batch_equals_single(gpu):
    """isac_cdl_generate_batch_dev (blockIdx.z = channel) against one isac_cdl_generate_dev call per channel."""
    import torch
    cm = importlib.import_module(PKG + ".communication.channelModels")
    chans = [cm.CDLChannel("CDL-C", TransmitAntennaArraySize=(1, 4, 2), ReceiveAntennaArraySize=(1, 2, 2), Seed=40 + i)
             for i in range(5)]
    K, scs = 24 * 12, 30e3
    sym = np.arange(14) * 35.7e-6
    t0 = 0.01 * np.arange(5)
    Hb = cm.CDLChannel.generateBatch(chans, K, scs, sym, t0)
    for i, ch in enumerate(chans):
        Hi = ch.generate(K, scs, sym, t0=float(t0[i]))
        assert torch.equal(Hb[i], Hi)
    for ch in chans:
        ch.close()
