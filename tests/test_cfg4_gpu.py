"""BASELINE config 4 at full size: mono-static MUSIC with a 32 x 32 array, 8 targets, 1000 snapshots.

(i)  DoA on the UPA: Ra 1024 x 1024 from X[1024 x 1000] (rank deficient: 1000 snapshots), scan 181 x 361 -> pseudo-spectrum array
     (the reference's UPA peak picker tools.find2DPeaks does not exist, so the array is the output).
(ii) music2D on H[624 x 1000]: Rr 624^2, Rv 1000^2, 1002-point range and 202-point velocity spectra, L = 8.
Tolerance: 1e-4 dB on the spectra (the 1e-5 relative bar of the north star in dB), peak lists exact."""
import importlib
import time

import numpy as np
import pytest

from oracle import sensing as S

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(gpu):
    return importlib.import_module(PKG)


def test_music_doa_32x32_upa_8_targets_1000_snapshots(P):
    nX = nY = 32
    n, N = nX * nY, 1000
    rng = np.random.default_rng(4)
    rp = {"antennaType": {"type": "upa", "nV": nX, "nH": nY, "p": 1, "dV": 0.5, "dH": 0.5}, "azimuthScanScale": 360,
          "azimuthScanGranularity": 1, "elevationScanScale": 180, "elevationScanGranularity": 1}
    mm, nn = np.arange(nX)[None, :], np.arange(nY)[:, None]
    targets = [(20, 30), (-60, 51), (75, 20), (-120, 65), (140, 40), (5, 75), (-30, 12), (100, 58)]
    A = np.stack([np.exp(-2j * np.pi * S.sind(el) * (mm * 0.5 * S.cosd(az) + nn * 0.5 * S.sind(az))).reshape(-1, order="F")
                  for az, el in targets], axis=1)
    sig = rng.standard_normal((8, N)) + 1j * rng.standard_normal((8, N))
    noise = 10 ** (-10 / 20) * (rng.standard_normal((n, N)) + 1j * rng.standard_normal((n, N))) / np.sqrt(2)   # SNR 10 dB
    X = A @ sig + noise
    Ra = X @ X.conj().T / N
    t0 = time.time()
    L, azi, ele, spec = P.sensing.estimation.doaEstimation.music(8, rp, Ra, return_spectrum=True)
    t_gpu = time.time() - t0
    t0 = time.time()
    Lr, _, _, specr = S.music_doa(8, rp, Ra)
    print("32x32 UPA MUSIC: GPU path %.2f s (incl. Jacobi eigen-solver), oracle %.1f s" % (t_gpu, time.time() - t0))
    assert L == Lr == 8 and azi is None and spec.shape == specr.shape == (181, 361)
    err = np.abs(spec - specr).max()
    print("PmusicdB max abs err [dB]", err)
    assert err <= 1e-4
    # rows are elevations -90..90, columns azimuths -180..180 (music.m:50-58); the reference normalises every azimuth column by
    # its own smallest magnitude (music.m:61-63), so a target shows as the maximum of its column, tens of dB above 0
    for az, el in targets:
        col = spec[:, az + 180]
        assert col[el + 90] >= col.max() - 3.0 and col[el + 90] > 30.0, (az, el, col[el + 90], col.max())


def test_music2d_624x1000_8_targets(P):
    nSc, nSym, nAnts = 624, 1000, 4
    rng = np.random.default_rng(41)
    scs, fc = 15.0, 3.5e9
    lam = S.LIGHTSPEED / fc
    Tsri = 1 / (scs * 1e3) + 4.7e-6
    rp = {"fc": fc, "Tsri": Tsri, "cfarEstZone": np.array([[0.0, 500.0], [-50.0, 50.0]]),
          "antennaType": {"type": "ula", "nV": 2, "p": 2, "d": 0.5}, "azimuthScanScale": 360,
          "azimuthScanGranularity": 1, "elevationScanScale": 180, "elevationScanGranularity": 1}
    tx = np.exp(2j * np.pi * rng.random((nSc, nSym, nAnts)))
    k, l = np.arange(nSc)[:, None], np.arange(nSym)[None, :]
    tg = [(60.0, 10.0), (95.5, -22.5), (140.0, 35.0), (188.0, -8.0), (240.5, 18.5), (301.0, -41.0), (366.0, 4.5), (430.5, 27.0)]
    H = np.zeros((nSc, nSym), dtype=complex)
    for i, (r, v) in enumerate(tg):
        H += (1.0 - 0.05 * i) * np.exp(-2j * np.pi * scs * 1e3 * 2 * r * k / S.LIGHTSPEED) * np.exp(2j * np.pi * Tsri * 2 * v * l / lam)
    rx = np.stack([(H * np.exp(-2j * np.pi * a * 0.5 * S.sind(25.0))) * tx[:, :, a] for a in range(nAnts)], axis=2)
    rx += 0.05 * (rng.standard_normal(rx.shape) + 1j * rng.standard_normal(rx.shape))
    rx32, tx32 = rx.astype(np.complex64), tx.astype(np.complex64)
    t0 = time.time()
    got = P.sensing.estimation.music2D(rp, {"scs": scs}, rx32, tx32, numDets=8)
    t_gpu = time.time() - t0
    ref = S.music2d(rp, {"scs": scs}, rx32, tx32, L_override=8)
    print("music2D 624x1000: GPU path %.2f s, sweeps %s, rng %s vel %s" % (t_gpu, got["jacobiSweeps"], got["rngEst"], got["velEst"]))
    assert got["L"] == ref["L"] == 8
    assert got["PrmusicdB"].size == 1002 and got["PvmusicdB"].size == 202
    for key in ("PrmusicdB", "PvmusicdB"):
        err = np.abs(got[key] - ref[key]).max()
        print("   ", key, "max abs err [dB]", err)
        assert err <= 1e-4
    assert np.array_equal(got["rngEst"], ref["rngEst"]) and np.array_equal(got["velEst"], ref["velEst"])
    assert np.array_equal(got["aziEst"], ref["aziEst"])
    for r, v in tg:
        assert np.abs(got["rngEst"] - r).min() <= 0.5 and np.abs(got["velEst"] - v).min() <= 0.5
