"""GPU parity of the sensing chain behind the package API: echo synthesis + OFDM demod (K1/K2),
covariance + MUSIC (K5/K6), the full fft2D estimator and music2D, each against the float64 oracle.

Tolerances: index-valued outputs (rngEst / velEst bins, peak locations, L) exact; complex grids
within 1e-5 of the grid's RMS (fp32 pipeline); float64 MUSIC pseudo-spectra within 1e-5 relative.
"""
import importlib

import numpy as np
import pytest

from oracle import sensing as S

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(gpu):
    return importlib.import_module(PKG)


def _setup(workloads, name, seed=1):
    cell, car, wave = workloads.cell_config(name)
    rp = S.radar_params(cell, car, wave)
    grid, txw = workloads.sensing_tx(name, seed)
    noise = workloads.std_normal_complex(txw.shape, seed + 1)
    return cell, car, wave, rp, grid, txw, noise


def _rel_rms(a, b):
    return np.abs(a - b).max() / np.sqrt(np.mean(np.abs(b) ** 2))


@pytest.mark.parametrize("name", ["tiny", "cfg1", "cfg2"])   # cfg2 = BASELINE config 2 at full size (3276 x 168 x 8 -> 4096 x 256)
def test_mono_static_sensing_matches_oracle(P, workloads, name):
    cell, car, wave, rp, grid, txw, noise = _setup(workloads, name)
    txw32, nz32 = txw.astype(np.complex64), noise.astype(np.complex64)
    ref = S.mono_static_sensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"], nz32)
    got = P.sensing.monoStaticSensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"], noise=nz32)
    assert got.shape == ref.shape
    err = _rel_rms(got, ref)
    print(name, "echoGrid err / rms", err)
    assert err <= 1e-5
    # noiseless: isolates the echo path (no noise floor hiding errors)
    ref0 = S.mono_static_sensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"], np.zeros_like(nz32))
    got0 = P.sensing.monoStaticSensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"])
    err0 = _rel_rms(got0, ref0)
    print(name, "noiseless echoGrid err / rms", err0)
    assert err0 <= 1e-5


def test_basic_radar_channel_matches_oracle_and_nlos(P, workloads):
    cell, car, wave, rp, grid, txw, noise = _setup(workloads, "tiny")
    txw32, nz32 = txw.astype(np.complex64), noise.astype(np.complex64)
    ref = S.basic_radar_channel(txw32, rp, cell["targetLoSConditions"], nz32)
    got = P.sensing.channelModels.basicRadarChannel(txw32, rp, cell["targetLoSConditions"], noise=nz32)
    err = _rel_rms(got, ref)
    print("rxWaveform err / rms", err)
    assert err <= 1e-5
    _lib = importlib.import_module(PKG + "._lib")
    with pytest.raises(_lib.IsacError) as e:  # all targets NLoS (basicRadarChannel.m:57-64)
        P.sensing.channelModels.basicRadarChannel(txw32, rp, np.zeros(1, dtype=int), noise=nz32)
    assert e.value.status == 6


def test_generated_noise_statistics(P, workloads):
    """Philox noise path: echo grid minus the noiseless grid is white with variance Nfft*N0."""
    cell, car, wave, rp, grid, txw, _ = _setup(workloads, "tiny")
    txw32 = txw.astype(np.complex64)
    g0 = P.sensing.monoStaticSensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"])
    g1 = P.sensing.monoStaticSensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"], seed=7)
    g2 = P.sensing.monoStaticSensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"], seed=7)
    assert np.array_equal(g1, g2)
    d = (g1 - g0).ravel()
    var = np.mean(np.abs(d) ** 2)
    expect = wave["Nfft"] * rp["N0"]
    print("noise var ratio", var / expect)
    assert abs(var / expect - 1) < 0.03
    assert abs(np.mean(d)) < 5 * np.sqrt(expect / d.size)
    assert abs(np.mean(d.real * d.imag)) < 0.03 * expect
    # Gaussian shape (Box-Muller on the special-function unit): 4th moment of a complex normal = 2 var^2, tails present
    assert abs(np.mean(np.abs(d) ** 4) / (2 * expect ** 2) - 1) < 0.08
    x = np.concatenate([d.real, d.imag]) / np.sqrt(expect / 2)
    assert abs(np.mean(np.abs(x) > 2.0) - 0.0455) < 0.006 and np.abs(x).max() > 3.5
    # antennas, resource elements and seeds are independent draws
    e = (g1 - g0)
    a0, a1 = e[..., 0].ravel(), e[..., 1].ravel()
    assert abs(np.vdot(a0, a1)) / (np.linalg.norm(a0) * np.linalg.norm(a1)) < 5 / np.sqrt(a0.size)
    assert abs(np.vdot(a0[:-1], a0[1:])) / np.linalg.norm(a0) ** 2 < 5 / np.sqrt(a0.size)
    g3 = P.sensing.monoStaticSensing(txw32, grid.shape, car, rp, cell["targetLoSConditions"], seed=8)
    b0 = (g3 - g0)[..., 0].ravel()
    assert abs(np.vdot(a0, b0)) / (np.linalg.norm(a0) * np.linalg.norm(b0)) < 5 / np.sqrt(a0.size)


@pytest.mark.parametrize("name", ["tiny", "cfg1", "cfg2"])   # cfg2 = BASELINE config 2 at full size (3276 x 168 x 8 -> 4096 x 256)
def test_fft2d_estimator_matches_oracle(P, workloads, name):
    cell, car, wave, rp, grid, txw, noise = _setup(workloads, name)
    rx = S.mono_static_sensing(txw, grid.shape, car, rp, cell["targetLoSConditions"], noise).astype(np.complex64)
    tx = grid.astype(np.complex64)
    cf_o = S.cfar2d_config(rp)
    ref = S.fft2d(rp, cf_o, rx, tx)
    prm = P.sensing.radarParams(cell, car, wave)
    cf = P.sensing.detection.cfar2D(prm)
    got = P.sensing.estimation.fft2D(prm, cf, rx, tx)
    print(name, "rngEst", got["rngEst"], "velEst", got["velEst"], "aziEst", got["aziEst"])
    assert np.array_equal(got["rngEst"], ref["rngEst"])
    assert np.array_equal(got["velEst"], ref["velEst"])
    assert np.array_equal(got["aziEst"], ref["aziEst"])
    assert got["eleEst"].shape == ref["eleEst"].shape and np.all(np.isnan(got["eleEst"]))


def test_fft2d_device_batch_pipeline(P, workloads):
    """Batched device entry: per-map-set results equal the single-map results; spectrum vs oracle."""
    import torch
    cell, car, wave, rp, grid, txw, noise = _setup(workloads, "tiny")
    est = importlib.import_module(PKG + ".sensing.estimation")
    prm = P.sensing.radarParams(cell, car, wave)
    cf = P.sensing.detection.cfar2D(prm)
    B = 3
    rxs, refs = [], []
    for b in range(B):
        nz = workloads.std_normal_complex(txw.shape, 10 + b)
        rx = S.mono_static_sensing(txw, grid.shape, car, rp, cell["targetLoSConditions"], nz).astype(np.complex64)
        rxs.append(rx)
        refs.append(S.fft2d(rp, S.cfar2d_config(rp), rx, grid.astype(np.complex64)))
    plan = est.SensePlan(prm, cf, grid.shape, max_batch=B)
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(3, 2, 1, 0))).cuda()
    rx_d = to_dev(np.stack(rxs, axis=3))
    tx_d = to_dev(np.stack([grid.astype(np.complex64)] * B, axis=3))
    plan.run_dev(rx_d, tx_d, B)
    res = plan.collect(B)
    spec = plan.spectrum(B)
    for b in range(B):
        assert np.array_equal(res[b]["rngEst"], refs[b]["rngEst"])
        assert np.array_equal(res[b]["velEst"], refs[b]["velEst"])
        assert np.array_equal(res[b]["aziEst"], refs[b]["aziEst"])
        assert res[b]["L"] == refs[b]["rngEst"].size
        err = np.abs(spec[b] - refs[b]["PmusicdB"]).max()
        print("PmusicdB max abs err [dB]", err)
        assert err <= 1e-4  # 1e-5 relative on the linear pseudo-spectrum = 8.7e-5 dB
    plan.close()


def test_covariance_and_music_doa(P, workloads):
    import torch
    _lib = importlib.import_module(PKG + "._lib")
    rng = np.random.default_rng(3)
    n, N = 16, 5000
    rp = {"antennaType": {"type": "ula", "nV": 8, "p": 2, "d": 0.5}, "azimuthScanScale": 360,
          "azimuthScanGranularity": 1, "elevationScanScale": 180, "elevationScanGranularity": 1}
    angs = np.array([-35.0, 12.0, 48.0])
    A = np.exp(-2j * np.pi * np.arange(n)[:, None] * 0.5 * S.sind(angs)[None, :])
    sig = (rng.standard_normal((3, N)) + 1j * rng.standard_normal((3, N))) * np.array([[3.0], [2.0], [1.5]])
    X = A @ sig + 0.3 * (rng.standard_normal((n, N)) + 1j * rng.standard_normal((n, N)))
    # grid whose reshape(.,[],nAnts)' equals X: page r holds conj(X[r,:])
    grid = np.conj(X).T.reshape(50, 100, n, order="F").astype(np.complex64)
    Ra_ref = S.antenna_covariance(grid)
    ctx = _lib.get_context(0)
    g_d = torch.from_numpy(np.ascontiguousarray(grid.transpose(2, 1, 0))).cuda()
    Ra = np.zeros((n, n), dtype=np.complex128, order="F")
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_antenna_covariance_dev(ctx.handle, _lib.ptr(g_d), 5000, n, _lib.ptr(Ra)), ctx.handle)
    err = np.abs(Ra - Ra_ref).max() / np.abs(Ra_ref).max()
    print("Ra rel err", err)
    assert err <= 1e-12
    for nd in (3, None, 1, 5):
        L, azi, ele, spec = P.sensing.estimation.doaEstimation.music(nd, rp, Ra_ref, return_spectrum=True)
        Lr, azir, eler, specr = S.music_doa(nd, rp, Ra_ref)
        assert L == Lr
        assert np.array_equal(azi, azir), (nd, azi, azir)
        rel = np.abs(10 ** (spec / 20) - 10 ** (specr / 20)).max()
        print("numDets", nd, "L", L, "azi", azi, "normalised spectrum max abs err", rel)
        assert np.abs(spec - specr).max() <= 1e-4
    # the +-180 deg scan of a ULA is mirror-ambiguous (sin(180-x) = sin(x)): every true angle or its mirror
    # is among the 2*3 strongest peaks
    azi6 = P.sensing.estimation.doaEstimation.music(6, rp, Ra_ref)[1]
    for a in angs:
        assert any(abs(azi6 - x).min() < 1e-9 for x in (a, 180 - a if a > 0 else -180 - a))
    with pytest.raises(_lib.IsacError) as e:
        P.sensing.estimation.doaEstimation.music(0, rp, Ra_ref)
    assert e.value.status == 7


@pytest.mark.parametrize("n,N", [(1, 4096), (2, 30000), (3, 30001), (4, 1234), (5, 60000), (8, 550368), (8, 77), (8, 2), (9, 5000), (16, 9999)])
def test_antenna_covariance_sizes(P, n, N):
    """fft2D.m:106-107 covariance over array sizes on both sides of the single-pass kernel (<= 8 elements, even sample count),
    odd sample counts and sample counts below one CTA; float64 accumulation of exact products, so the bar is 1e-12."""
    import torch
    _lib = importlib.import_module(PKG + "._lib")
    rng = np.random.default_rng(100 * n + N % 97)
    X = (rng.standard_normal((n, N)) + 1j * rng.standard_normal((n, N))).astype(np.complex64)
    X[:, ::7] *= 30.0
    Xd = X.astype(np.complex128)
    Ra_ref = (Xd.conj() @ Xd.T) / N          # Ra(i,j) = mean(conj(x_i) x_j): the orientation isac_antenna_covariance_dev documents
    ctx = _lib.get_context(0)
    g_d = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    Ra = np.zeros((n, n), dtype=np.complex128, order="F")
    ctx.use_torch_stream()
    for rep in range(2):                      # second call: the ticket counters of the single-pass kernel were left at zero
        Ra[:] = 0
        _lib.check(ctx.lib.isac_antenna_covariance_dev(ctx.handle, _lib.ptr(g_d), N, n, _lib.ptr(Ra)), ctx.handle)
        err = np.abs(Ra - Ra_ref).max() / np.abs(Ra_ref).max()
        assert err <= 1e-12, (n, N, rep, err)
        assert np.array_equal(Ra, Ra.conj().T) and np.all(Ra.imag[np.diag_indices(n)] == 0)


@pytest.mark.parametrize("nxy", [(4, 4), (10, 9)])
def test_music_upa_spectrum(P, nxy):
    """UPA branch (music.m:31-63): spectrum array parity (the reference's peak picker does not exist).
    (10,9) = 90 elements exercises the multi-CTA Jacobi path."""
    nX, nY = nxy
    n = nX * nY
    rng = np.random.default_rng(11)
    rp = {"antennaType": {"type": "upa", "nV": nX, "nH": nY, "p": 1, "dV": 0.5, "dH": 0.5}, "azimuthScanScale": 360,
          "azimuthScanGranularity": 4, "elevationScanScale": 180, "elevationScanGranularity": 3}
    N = 400
    mm, nn = np.arange(nX)[None, :], np.arange(nY)[:, None]
    cols = []
    for az, el in ((20.0, 30.0), (-60.0, 51.0)):
        a = np.exp(-2j * np.pi * S.sind(el) * (mm * 0.5 * S.cosd(az) + nn * 0.5 * S.sind(az)))
        cols.append(a.reshape(-1, order="F"))
    A = np.stack(cols, axis=1)
    X = A @ (rng.standard_normal((2, N)) + 1j * rng.standard_normal((2, N))) + 0.2 * (
        rng.standard_normal((n, N)) + 1j * rng.standard_normal((n, N)))
    Ra = X @ X.conj().T / N
    L, azi, ele, spec = P.sensing.estimation.doaEstimation.music(2, rp, Ra, return_spectrum=True)
    Lr, _, _, specr = S.music_doa(2, rp, Ra)
    assert L == Lr == 2 and azi is None
    assert spec.shape == specr.shape
    err = np.abs(spec - specr).max()
    print(nxy, "UPA PmusicdB max abs err [dB]", err)
    assert err <= 1e-4


@pytest.mark.parametrize("nrb_scs_sym", [(24, 15, 17), (52, 15, 30), (106, 30, 29), (273, 30, 28), (11, 15, 5), (6, 60, 57)])
def test_ofdm_modulate_matches_oracle_and_round_trips(P, nrb_scs_sym):
    """sensing.ofdmModulate (gNBPhy.m:599) vs the float64 oracle for every FFT size (128 .. 4096, odd symbol counts that
    cross subframe boundaries), and the size-independent property demodulate(modulate(grid)) == grid."""
    nrb, scs, nsym = nrb_scs_sym
    rng = np.random.default_rng(nrb + nsym)
    nsc, nants = 12 * nrb, 3
    grid = (rng.standard_normal((nsc, nsym, nants)) + 1j * rng.standard_normal((nsc, nsym, nants))).astype(np.complex64)
    car = {"NRBsDL": nrb, "SubcarrierSpacing": scs}
    wave = P.sensing.ofdmModulate(car, grid, scale=37.5)
    ref = S.ofdm_modulate(nrb, scs, grid, 37.5)
    assert wave.shape == ref.shape
    err = np.abs(wave - ref).max() / np.abs(ref).max()
    print(nrb_scs_sym, "Nfft", S.ofdm_info(nrb, scs)["Nfft"], "T", wave.shape[0], "waveform err rel-to-peak", err)
    assert err <= 1e-5   # fp32 IFFT against the float64 oracle
    back = S.ofdm_demodulate(nrb, scs, wave.astype(np.complex128) / 37.5)
    assert back.shape == grid.shape
    assert np.abs(back - grid).max() / np.abs(grid).max() <= 1e-5


def test_ofdm_windowing_and_slot_accumulation(P):
    """nrOFDMModulate with a window (documented W-OLA scheme) vs the oracle, and the device-resident sensing tap
    (gNBPhy.m:604-612): appending DL slots one by one gives the same senTxGrid / senTxWave as modulating them in one piece --
    with the window folded across the slot boundaries."""
    echo = importlib.import_module(PKG + ".sensing._echo")
    nrb, scs = 52, 30
    rng = np.random.default_rng(5)
    nsc, nants, nslots = 12 * nrb, 4, 5
    grid = (rng.standard_normal((nsc, 14 * nslots, nants)) + 1j * rng.standard_normal((nsc, 14 * nslots, nants))).astype(np.complex64)
    car = {"NRBsDL": nrb, "SubcarrierSpacing": scs}
    for N in (0, 18, 72):
        ref = S.ofdm_modulate(nrb, scs, grid, 3.0, windowing=N)
        wave = P.sensing.ofdmModulate(car, grid, scale=3.0, windowing=N)
        assert wave.shape == ref.shape
        assert np.abs(wave - ref).max() / np.abs(ref).max() <= 1e-5, N
        acc = echo.SensingTxAccumulator(car, nants, 14 * nslots, scale=3.0, windowing=N)
        for sl in range(nslots):
            acc.append(grid[:, 14 * sl: 14 * (sl + 1)], slotInSubframe=sl % 2)      # 30 kHz: two slots per subframe
        assert acc.nSym == 14 * nslots and acc.T == ref.shape[0]
        assert np.array_equal(acc.grid().cpu().numpy().transpose(2, 1, 0), grid)
        got = acc.wave().cpu().numpy().T
        assert np.abs(got - ref).max() / np.abs(ref).max() <= 1e-5, N
    if True:   # the window changes only the N samples in front of every symbol but the first
        plain, win = S.ofdm_modulate(nrb, scs, grid, 3.0), S.ofdm_modulate(nrb, scs, grid, 3.0, windowing=18)
        changed = np.flatnonzero(np.abs(plain - win).max(axis=1) > 0)
        assert changed.size <= 18 * (14 * nslots - 1) and changed.size > 0
    with pytest.raises(Exception):
        P.sensing.ofdmModulate(car, grid, scale=1.0, windowing=400)                  # longer than the cyclic prefix


@pytest.mark.parametrize("method", ["mvdrBF", "digitalBF"])
def test_mvdr_and_beamscan_ula(P, method):
    """doaEstimation.mvdrBF / digitalBF, ULA branch (mvdrBF.m:57-89, digitalBF.m:57-90) vs the float64 oracle:
    spectra within 1e-5 relative (8.7e-5 dB), peak lists identical."""
    _lib = importlib.import_module(PKG + "._lib")
    rng = np.random.default_rng(31)
    n, N = 16, 4000
    rp = {"antennaType": {"type": "ula", "nV": 8, "p": 2, "d": 0.5}, "azimuthScanScale": 360,
          "azimuthScanGranularity": 1, "elevationScanScale": 180, "elevationScanGranularity": 1}
    angs = np.array([-41.0, 7.0, 33.0])
    A = np.exp(-2j * np.pi * np.arange(n)[:, None] * 0.5 * S.sind(angs)[None, :])
    sig = (rng.standard_normal((3, N)) + 1j * rng.standard_normal((3, N))) * np.array([[3.0], [2.0], [1.5]])
    X = A @ sig + 0.5 * (rng.standard_normal((n, N)) + 1j * rng.standard_normal((n, N)))
    Ra = X @ X.conj().T / N
    fn = getattr(P.sensing.estimation.doaEstimation, method)
    ref = S.mvdr_bf if method == "mvdrBF" else S.digital_bf
    for nd in (3, 6, 1):
        azi, ele, spec = fn(nd, rp, Ra, return_spectrum=True)
        azir, eler, specr = ref(nd, rp, Ra)
        assert np.array_equal(azi, azir), (nd, azi, azir)
        assert np.all(np.isnan(ele)) and ele.size == azi.size
        err = np.abs(spec - specr).max()
        print(method, "numDets", nd, "azi", azi, "dB spectrum max abs err", err)
        assert err <= 1e-4
    for a in angs:   # mirror-ambiguous +-180 deg scan: the true angle or its mirror is among the 6 strongest peaks
        azi6 = fn(6, rp, Ra)[0]
        assert any(abs(azi6 - x).min() < 1.5 for x in (a, 180 - a if a > 0 else -180 - a))
    with pytest.raises(_lib.IsacError) as e:   # no source-count rule in the beamformers: NPeaks must be >= 1
        fn(None, rp, Ra)
    assert e.value.status == 7


@pytest.mark.parametrize("method", ["mvdrBF", "digitalBF"])
@pytest.mark.parametrize("nxy", [(4, 4), (9, 8)])
def test_mvdr_and_beamscan_upa(P, method, nxy):
    """UPA branch (mvdrBF.m:14-55, digitalBF.m:14-55): spectrum parity incl. MATLAB's column-wise ./max normalisation.
    (9,8) = 72 elements exercises the multi-CTA Jacobi path."""
    nX, nY = nxy
    n = nX * nY
    rng = np.random.default_rng(17)
    rp = {"antennaType": {"type": "upa", "nV": nX, "nH": nY, "p": 1, "dV": 0.5, "dH": 0.5}, "azimuthScanScale": 360,
          "azimuthScanGranularity": 5, "elevationScanScale": 180, "elevationScanGranularity": 4}
    N = 600
    mm, nn = np.arange(nX)[None, :], np.arange(nY)[:, None]
    cols = []
    for az, el in ((25.0, 32.0), (-70.0, 48.0)):
        a = np.exp(-2j * np.pi * S.sind(el) * (mm * 0.5 * S.cosd(az) + nn * 0.5 * S.sind(az)))
        cols.append(a.reshape(-1, order="F"))
    A = np.stack(cols, axis=1)
    X = A @ (rng.standard_normal((2, N)) + 1j * rng.standard_normal((2, N))) + 0.4 * (
        rng.standard_normal((n, N)) + 1j * rng.standard_normal((n, N)))
    Ra = X @ X.conj().T / N
    fn = getattr(P.sensing.estimation.doaEstimation, method)
    ref = S.mvdr_bf if method == "mvdrBF" else S.digital_bf
    azi, ele, spec = fn(2, rp, Ra, return_spectrum=True)
    _, _, specr = ref(2, rp, Ra)
    assert azi is None and ele is None and spec.shape == specr.shape
    err = np.abs(spec - specr).max()
    print(method, nxy, "UPA dB spectrum max abs err", err)
    assert err <= 1e-4


@pytest.mark.parametrize("shape", [(96, 40, 4), (60, 90, 4)])
def test_music2d_matches_oracle(P, shape):
    """music2D (music2D.m:33-123) on tall (nSc>nSym) and wide (nSc<nSym) channel matrices."""
    nSc, nSym, nAnts = shape
    rng = np.random.default_rng(21)
    scs, fc = 30.0, 3.5e9
    lam = S.LIGHTSPEED / fc
    Tsri = 1 / (scs * 1e3) + 5e-6
    rp = {"fc": fc, "Tsri": Tsri, "cfarEstZone": np.array([[50.0, 300.0], [-50.0, 50.0]]),
          "antennaType": {"type": "ula", "nV": 2, "p": 2, "d": 0.5}, "azimuthScanScale": 360,
          "azimuthScanGranularity": 1, "elevationScanScale": 180, "elevationScanGranularity": 1}
    tx = np.exp(2j * np.pi * rng.random((nSc, nSym, nAnts)))
    k, l = np.arange(nSc)[:, None], np.arange(nSym)[None, :]
    H = np.zeros((nSc, nSym), dtype=complex)
    for r, v, amp in ((120.0, 10.0, 1.0), (210.5, -22.5, 0.7)):
        H += amp * np.exp(-2j * np.pi * scs * 1e3 * 2 * r * k / S.LIGHTSPEED) * np.exp(2j * np.pi * Tsri * 2 * v * l / lam)
    ang = np.array([25.0, -40.0, 10.0])
    rx = np.zeros((nSc, nSym, nAnts), dtype=complex)
    for a in range(nAnts):
        rx[:, :, a] = (H * np.exp(-2j * np.pi * a * 0.5 * S.sind(ang[0]))) * tx[:, :, a]
    rx += 0.05 * (rng.standard_normal(rx.shape) + 1j * rng.standard_normal(rx.shape))
    rx32, tx32 = rx.astype(np.complex64), tx.astype(np.complex64)
    for nd in (2, None):
        ref = S.music2d(rp, {"scs": scs}, rx32, tx32, L_override=nd)
        got = P.sensing.estimation.music2D(rp, {"scs": scs}, rx32, tx32, numDets=nd)
        print(shape, "numDets", nd, "L", got["L"], "rng", got["rngEst"], "vel", got["velEst"], "sweeps", got["jacobiSweeps"])
        assert got["L"] == ref["L"]
        assert np.array_equal(got["rngEst"], ref["rngEst"])
        assert np.array_equal(got["velEst"], ref["velEst"])
        assert np.array_equal(got["aziEst"], ref["aziEst"])
        for key in ("PrmusicdB", "PvmusicdB"):
            err = np.abs(got[key] - ref[key]).max()
            print("   ", key, "max abs err [dB]", err)
            assert err <= 1e-4
    assert abs(got["rngEst"][0] - 120.0) <= 0.5 or abs(got["rngEst"][0] - 210.5) <= 0.5
