"""Known-answer tests of the channel-estimation oracle (oracle/chest.py) and its golden fixture.

nrChannelEstimate is a closed toolbox function (call sites uePhy.m:897, gNBPhy.m:1030) and the reference holds no vectors
for it: PARITY UNPINNED.  These tests pin the estimator this build specifies through closed-form cases."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import chest as OCH

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)


def _flat(K, L, R, P, seed=0):
    rng = np.random.default_rng(seed)
    H0 = rng.standard_normal((R, P)) + 1j * rng.standard_normal((R, P))
    return np.broadcast_to(H0, (K, L, R, P)).copy()


def test_flat_channel_is_recovered_exactly_and_ports_separate():
    nrb, R, P = 24, 3, 4
    ind, sym, cdm = OCH.csirs_row5_layout(nrb, 1, 0, seed=3)
    H = _flat(12 * nrb, 14, R, P)
    He, nv = OCH.channel_estimate(OCH.apply_channel(H, ind, sym), ind, sym, P, cdm)
    assert np.abs(He - H).max() < 1e-13          # the cover codes of the port sharing each CDM group cancel exactly
    assert nv < 1e-25


def test_without_despreading_cdm_ports_leak():
    nrb, P = 12, 4
    ind, sym, _ = OCH.csirs_row5_layout(nrb, 1, 0, seed=4)
    H = _flat(12 * nrb, 14, 2, P, seed=1)
    He, _ = OCH.channel_estimate(OCH.apply_channel(H, ind, sym), ind, sym, P, (1, 1))
    assert np.abs(He - H).max() > 0.1            # CDMLengths matter: LS alone mixes the two ports of a group


def test_linear_channel_no_cdm_interpolates_exactly_between_pilots():
    K, L, R, P = 120, 14, 2, 1
    ks = np.arange(3, K, 6)                       # comb of single-port pilots on symbols 2 and 9
    ind = np.concatenate([1 + ks + K * 2, 1 + ks + K * 9])
    sym = np.exp(1j * np.pi / 4 * (1 + 2 * np.arange(ind.size)))
    k = np.arange(K)[:, None, None, None]
    l = np.arange(L)[None, :, None, None]
    H = (1.0 + 0.01 * k + 0.02j * l) * np.ones((1, 1, R, P))
    He, nv = OCH.channel_estimate(OCH.apply_channel(H, ind, sym), ind, sym, P, (1, 1))
    inner = np.s_[ks[0]:ks[-1] + 1, 2:10]
    assert np.abs(He[inner] - H[inner]).max() < 1e-12                      # linear interpolation in both axes
    assert np.allclose(He[:ks[0], 2], He[ks[0], 2]) and np.allclose(He[ks[-1]:, 5], He[ks[-1], 5])   # constant extrapolation
    assert np.allclose(He[:, 0], He[:, 2]) and np.allclose(He[:, 13], He[:, 9])
    assert nv < 1e-25                                                      # second differences of a linear channel vanish


def test_noise_variance_estimate_is_unbiased():
    nrb, R, P, sig2 = 100, 4, 4, 0.02
    ind, sym, cdm = OCH.csirs_row5_layout(nrb, 1, 0, seed=5)
    K = 12 * nrb
    rng = np.random.default_rng(7)
    noise = np.sqrt(sig2 / 2) * (rng.standard_normal((K, 14, R)) + 1j * rng.standard_normal((K, 14, R)))
    _, nv = OCH.channel_estimate(OCH.apply_channel(_flat(K, 14, R, P), ind, sym, noise), ind, sym, P, cdm)
    assert abs(nv / sig2 - 1.0) < 0.1


def test_averaging_window_reduces_the_error_on_a_flat_channel():
    nrb, R, P = 50, 2, 4
    ind, sym, cdm = OCH.csirs_row5_layout(nrb, 1, 0, seed=6)
    K = 12 * nrb
    rng = np.random.default_rng(8)
    noise = 0.1 * (rng.standard_normal((K, 14, R)) + 1j * rng.standard_normal((K, 14, R)))
    H = _flat(K, 14, R, P)
    rx = OCH.apply_channel(H, ind, sym, noise)
    e0 = np.abs(OCH.channel_estimate(rx, ind, sym, P, cdm)[0] - H).std()
    e5 = np.abs(OCH.channel_estimate(rx, ind, sym, P, cdm, (5, 1))[0] - H).std()
    assert e5 < 0.6 * e0


def test_layout_errors():
    with pytest.raises(ValueError):
        OCH.pilot_layout([1, 2, 3], [1, 1], 12, 14, 1)
    with pytest.raises(ValueError):
        OCH.pilot_layout([1, 12 * 14 * 2 + 1], [1, 1], 12, 14, 2)        # index beyond the grid
    with pytest.raises(ValueError):
        OCH.pilot_layout([1, 2, 13], [1, 1, 1], 12, 14, 1)               # not a product grid


def test_oracle_reproduces_fixture():
    g = np.load(os.path.join(HERE, "golden", "chest_small.npz"))
    K, L, R, P, ind, sym, cdm, H, rx, nv = MG.chest_case()
    He, nve = OCH.channel_estimate(rx, ind, sym, P, cdm)
    assert np.allclose(He, g["Hest"], rtol=0, atol=1e-6)
    assert np.isclose(nve, float(g["nVar"]), rtol=1e-10)
    Ha, nva = OCH.channel_estimate(rx, ind, sym, P, cdm, (3, 1))
    assert np.allclose(Ha, g["Hest_avg"], rtol=0, atol=1e-6)
    assert np.abs(He - H).std() < 0.05 and 0.5 < nve / nv < 2.0           # the estimate tracks the true channel / noise
