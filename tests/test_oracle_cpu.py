"""CPU-only known-answer tests that pin the float64 oracle (oracle/sensing.py): the reference ships no golden
vectors, so the restatement is anchored on closed forms and on independent SciPy code."""
import importlib
import math

import numpy as np
import pytest
import scipy.signal

from oracle import sensing as S

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
W = importlib.import_module(PKG + ".workloads")


def test_kaiser_matches_scipy():
    for n in (1, 2, 17, 624, 4096):
        assert np.allclose(S.kaiser(n, 3.0), scipy.signal.windows.kaiser(n, 3.0, sym=True), rtol=1e-13, atol=1e-15)


def test_findpeaks_semantics():
    rng = np.random.default_rng(0)
    y = rng.standard_normal(500)
    pk, loc = S.findpeaks(y, 7)
    ref, _ = scipy.signal.find_peaks(y)
    top = ref[np.argsort(-y[ref], kind="stable")][:7]
    assert np.array_equal(loc - 1, top) and np.array_equal(pk, y[top])
    # plateau -> first sample; end points never peaks; fewer than NPeaks available
    assert np.array_equal(S.findpeaks([0, 1, 3, 3, 3, 1, 5], 4)[1], [3])
    assert np.array_equal(S.findpeaks([5, 1, 2, 1, 4], 3)[1], [3])
    with pytest.raises(ValueError):
        S.findpeaks(y, 0)


def test_sind_exact_and_unique_stable():
    assert S.sind(180.0) == 0 and S.sind(-180.0) == 0 and S.sind(90.0) == 1 and S.cosd(90.0) == 0 and S.sind(360.0) == 0
    assert abs(S.sind(30.0) - 0.5) < 1e-15
    assert np.array_equal(S.unique_stable(np.array([3.0, 1.0, 3.0, 2.0, 1.0])), [3.0, 1.0, 2.0])


def test_cfar_threshold_and_strict_compare():
    assert abs(S.cfar_threshold_factor(24, 1e-9) - 32.91296893587972) < 1e-10
    P = np.ones((40, 30))
    cf = {"CUTIdx": np.array([[10, 11], [12, 12]]), "GuardBandSize": (2, 2), "TrainingBandSize": (1, 1), "Pfa": 1e-9}
    assert S.cfar2d_detect(P, cf).shape[1] == 0
    a = S.cfar_threshold_factor(24, 1e-9)
    P[9, 11] = a * (1 + 1e-9)          # CUT (10,12): just above -> detected
    P[10, 11] = a                       # CUT (11,12): equal to threshold*noise?  its guard cell (10,12) is excluded
    d = S.cfar2d_detect(P, cf)
    assert d.tolist() == [[10], [12]]
    assert np.array_equal(S.cfar2d_detect_exact(P, cf), d)
    with pytest.raises(ValueError):
        S.cfar2d_detect(P, dict(cf, CUTIdx=np.array([[2], [12]])))
    rng = np.random.default_rng(1)
    Q = rng.exponential(size=(64, 48))
    Q[rng.integers(5, 59, 12), rng.integers(5, 43, 12)] *= 200
    rr, cc = np.meshgrid(np.arange(5, 60), np.arange(5, 44), indexing="ij")
    cf2 = dict(cf, CUTIdx=np.stack([rr.reshape(-1, order="F"), cc.reshape(-1, order="F")]), Pfa=1e-4)
    assert np.array_equal(S.cfar2d_detect(Q, cf2), S.cfar2d_detect_exact(Q, cf2))


def test_rdm_closed_form():
    """The kernels' closed form equals the literal restatement of fft2D.m:37-46."""
    rng = np.random.default_rng(0)
    for (nSc, nSym, nA, N, F) in [(300, 77, 3, 512, 64), (288, 42, 4, 512, 64), (200, 10, 2, 256, 16)]:
        rx = rng.standard_normal((nSc, nSym, nA)) + 1j * rng.standard_normal((nSc, nSym, nA))
        tx = rng.standard_normal((nSc, nSym, nA)) + 1j * rng.standard_normal((nSc, nSym, nA))
        ref = S.rdm_2dfft({"nIFFT": N, "nFFT": F}, rx, tx)
        y = np.fft.ifft(rx * np.conj(tx) * S.kaiser(nSc, 3.0)[:, None, None], N, axis=0) * math.sqrt(N)
        M = min(nSym, F)
        sp = np.arange(M)
        ys = y[:, (sp + nSym // 2) % nSym, :] * ((-1.0) ** sp)[None, :, None]
        Z = np.fft.fft(ys, F, axis=1) / math.sqrt(F) * S.kaiser(N, 3.0)[(np.arange(N) - N // 2) % N][:, None, None]
        assert np.abs(Z - ref).max() <= 1e-14 * np.abs(ref).max()


def test_ofdm_roundtrip_and_numerology():
    for nrb, scs, nfft, fs in ((52, 15, 1024, 15.36e6), (273, 30, 4096, 122.88e6), (24, 15, 512, 7.68e6)):
        info = S.ofdm_info(nrb, scs)
        assert info["Nfft"] == nfft and info["SampleRate"] == fs
        assert info["SymbolLengths"].sum() == fs / 1000                # one subframe
        num = W.ofdm_numerology(nrb, scs)
        assert np.array_equal(num["CyclicPrefixLengths"], info["CyclicPrefixLengths"])
    grid = W.qpsk_grid(288, 28, 2, 3)
    wave = W.ofdm_modulate(grid, 24, 15)
    back = S.ofdm_demodulate(24, 15, wave)
    assert back.shape == grid.shape
    assert np.abs(back - grid).max() < 1e-12                              # plain fft(ifft(.)) with the CP-fraction ramp undone


def test_single_target_lands_in_the_predicted_bins():
    """Noiseless target: peak row = s*nIFFT/Nfft + 1; echo passes through every stage."""
    cell, car, wave = W.cell_config("tiny")
    rp = S.radar_params(cell, car, wave)
    grid, txw = W.sensing_tx("tiny", 1)
    rx = S.mono_static_sensing(txw, grid.shape, car, rp, [1], np.zeros(txw.shape, complex))
    P = np.abs(S.rdm_2dfft(rp, rx, grid)[:, :, 0]) ** 2
    r, c = np.unravel_index(np.argmax(P), P.shape)
    s = math.ceil(2 * rp["range"][0] / S.LIGHTSPEED * rp["fs"])
    assert r == s * rp["nIFFT"] // wave["Nfft"]
    # the symbol-axis rotation before the zero-padded Doppler FFT and the mis-stated Tsri (radarParams.m:34-35) bias
    # the velocity axis: only require the right neighbourhood
    assert abs((c - rp["nFFT"] / 2) * rp["vRes"] - rp["velocity"][0]) <= 2 * rp["vRes"]
    assert abs(rp["rRes"] - S.LIGHTSPEED / (2 * 15e3 * 512)) < 1e-9 and rp["nIFFT"] == 512 and rp["nFFT"] == 64


def test_music_on_exact_covariance():
    n = 16
    rp = {"antennaType": {"type": "ula", "nV": 8, "p": 2, "d": 0.5}, "azimuthScanScale": 360, "azimuthScanGranularity": 1}
    angs = np.array([-50.0, 20.0])
    A = np.exp(-2j * np.pi * np.arange(n)[:, None] * 0.5 * S.sind(angs)[None, :])
    Ra = A @ np.diag([2.0, 1.0]) @ A.conj().T + 0.1 * np.eye(n)
    L, azi, ele, PdB = S.music_doa(2, rp, Ra)
    assert L == 2 and np.all(np.isnan(ele))
    # +-180 scan of a ULA is mirror ambiguous: sin(x) = sin(180-x)
    assert set(azi.tolist()) <= {-50.0, -130.0, 20.0, 160.0}
    assert S.determine_num_targets(np.array([0.1, 0.1, 0.1, 0.1, 5.0, 9.0])) >= 1
    # on-grid steering vector is orthogonal to the noise subspace: P = 1/eps at the true angle
    Uann, _ = S.noise_projector(Ra, 2)
    a = np.exp(-2j * np.pi * np.arange(n) * 0.5 * S.sind(20.0))
    assert abs(np.vdot(a, Uann @ a)) < 1e-10


def test_mvdr_and_beamscan_known_answers():
    """mvdrBF / digitalBF on Ra = p a0 a0' + s I (one on-grid source): closed forms by Sherman-Morrison.
    a' Ra^-1 a = (n - p |a'a0|^2 / (s + p n)) / s  and  a' Ra a = p |a'a0|^2 + s n."""
    n, p, s = 16, 3.0, 0.2
    rp = {"antennaType": {"type": "ula", "nV": 8, "p": 2, "d": 0.5}, "azimuthScanScale": 360, "azimuthScanGranularity": 1}
    m = np.arange(n)
    a0 = np.exp(-2j * np.pi * m * 0.5 * S.sind(20.0))
    Ra = p * np.outer(a0, a0.conj()) + s * np.eye(n)
    azi_m, ele_m, PdB_m = S.mvdr_bf(2, rp, Ra)
    azi_d, ele_d, PdB_d = S.digital_bf(2, rp, Ra)
    assert set(azi_m.tolist()) == {20.0, 160.0} and set(azi_d.tolist()) == {20.0, 160.0}   # mirror-ambiguous ULA scan
    assert np.all(np.isnan(ele_m)) and np.all(np.isnan(ele_d))
    ang = np.arange(361) - 180.0
    g = np.array([abs(np.vdot(np.exp(-2j * np.pi * m * 0.5 * S.sind(x)), a0)) ** 2 for x in ang])
    P_m = 1.0 / ((n - p * g / (s + p * n)) / s + S.EPS1)
    P_d = p * g + s * n
    assert np.abs(PdB_m - 20 * np.log10(P_m / P_m.max())).max() < 1e-9
    assert np.abs(PdB_d - 20 * np.log10(P_d / P_d.max())).max() < 1e-9


def test_upa_spectrum_is_normalised_per_azimuth_column():
    """music.m:61-63 on a matrix: Pmusic./max(Pmusic) divides every column by ITS maximum (MATLAB max of a matrix is
    column-wise) -> with Pmusic = -abs(.) every column of the dB map has minimum 0 dB."""
    rp = {"antennaType": {"type": "upa", "nV": 3, "nH": 3, "p": 1}, "azimuthScanScale": 360, "azimuthScanGranularity": 30,
          "elevationScanScale": 180, "elevationScanGranularity": 20}
    rng = np.random.default_rng(2)
    X = rng.standard_normal((9, 40)) + 1j * rng.standard_normal((9, 40))
    Ra = X @ X.conj().T / 40
    for spec in (S.music_doa(1, rp, Ra)[3], S.mvdr_bf(1, rp, Ra)[2], S.digital_bf(1, rp, Ra)[2]):
        assert spec.shape == (9, 12)
        assert np.allclose(spec.min(axis=0), 0.0) and (spec >= 0).all()


def test_ofdm_modulate_is_the_inverse_of_demodulate():
    """oracle nrOFDMModulate restatement (gNBPhy.m:599): demodulate(modulate(grid)) == grid for every CP pattern, and the
    cyclic prefix is the symbol's tail (TS 38.211 5.3.1)."""
    rng = np.random.default_rng(4)
    for nrb, scs, nsym in ((24, 15, 17), (52, 30, 31), (6, 60, 57)):
        nsc = 12 * nrb
        grid = rng.standard_normal((nsc, nsym, 2)) + 1j * rng.standard_normal((nsc, nsym, 2))
        wave = S.ofdm_modulate(nrb, scs, grid, 2.0)
        info = S.ofdm_info(nrb, scs)
        starts = S.ofdm_symbol_starts(info, nsym)
        assert wave.shape[0] == starts[-1] + info["SymbolLengths"][(nsym - 1) % info["SymbolLengths"].size]
        cp0 = int(info["CyclicPrefixLengths"][0])
        assert np.allclose(wave[:cp0], wave[info["Nfft"]: info["Nfft"] + cp0])
        back = S.ofdm_demodulate(nrb, scs, wave / 2.0)
        assert np.abs(back - grid).max() < 1e-12


def test_vectorised_comm_oracle_equals_loop_faithful_oracle():
    """oracle/comm.py holds two restatements of the CSI path: the loop nest of the reference (dl_pmi_select / ri_select /
    cqi_select) and a vectorised one (sinr_per_re_vectorized / csi_report_vectorized) used for the CPU baseline and for the
    full-size GPU parity tests (tests/test_cfg23_gpu.py).  They must agree."""
    from oracle import comm as OC
    rng = np.random.default_rng(77)
    for n_ports, panel, n_rx in ((4, (2, 1), 2), (8, (2, 2), 4)):
        nrb = 24
        cfg = OC.report_config(n_ports, panel, nrb, 0, 1, "Subband", "Subband", 4)
        re_k, re_l = OC.csirs_first_port_res(nrb, 1, 0)
        H = (rng.standard_normal((12 * nrb, 14, n_rx, n_ports)) + 1j * rng.standard_normal((12 * nrb, 14, n_rx, n_ports))) / np.sqrt(2)
        H = H + np.roll(H, 1, axis=0)
        for nu in (1, 2):
            _, info = OC.dl_pmi_select(cfg, re_k, re_l, nu, H, 0.05)
            Sv, Wv = OC.sinr_per_re_vectorized(cfg, re_k, re_l, nu, H, 0.05)
            S = info["SINRPerRE"]
            assert S.shape == Sv.shape and np.array_equal(np.isnan(S), np.isnan(Sv))
            m = ~np.isnan(S)
            assert np.abs(S[m] - Sv[m]).max() <= 1e-10 * np.abs(S[m]).max()
        table = np.array([-3.46, 1.54, 6.54, 11.05, 13.54, 16.04, 17.54, 20.04, 22.04, 24.43, 26.93, 27.43, 29.43, 32.43, 35.43])
        ri, pm = OC.ri_select(cfg, re_k, re_l, H, 0.05)
        rank = int(min(ri, 4))
        cqi, pmc, _, _ = OC.cqi_select(cfg, re_k, re_l, rank, H, 0.05, table)
        rv, pmv, cqv = OC.csi_report_vectorized(cfg, re_k, re_l, H, 0.05, table, rank_cap=4)
        assert rv == rank and np.array_equal(pmv["i1"], pmc["i1"]) and np.array_equal(pmv["i2"], pmc["i2"], equal_nan=True)
        d = cqv[1:] - cqv[:1]       # absolute subband CQIs -> differential report format (cqiSelect.m:656-677)
        off = np.where(np.isnan(d), np.nan, np.where(d == 0, 0, np.where(d == 1, 1, np.where(d >= 2, 2, 3))))
        assert np.array_equal(np.vstack([cqv[:1], off]), cqi[:, : cqv.shape[1]], equal_nan=True)


def test_get_pd_receiver_operating_characteristic():
    """sensing.detection.getPd (getPd.m:1 -> rocpfa, NonfluctuatingCoherent): closed-form known answers."""
    import importlib
    det = importlib.import_module("5g_based_system_level_integrated_sensing_and_communication_simulator_b200.sensing.detection")
    snr = np.linspace(-40.0, 30.0, 71)
    pd = det.getPd([1e-3, 1e-6, 1e-9], snr, 1)
    assert pd.shape == (71, 3)
    assert np.allclose(pd[0], [1e-3, 1e-6, 1e-9], rtol=0.1)          # no signal: Pd -> Pfa
    assert np.all(pd[-1] > 1 - 1e-12)                                # strong signal: Pd -> 1
    assert np.all(np.diff(pd, axis=0) >= 0) and np.all(np.diff(pd, axis=1) <= 0)   # monotone in SNR and in Pfa
    # Pd = 1/2 where sqrt(N SNR) = erfcinv(2 Pfa); N pulses shift the curve by 10 log10 N dB
    from scipy.special import erfcinv
    s_half = 20 * np.log10(erfcinv(2e-6))
    assert abs(float(det.getPd(1e-6, [s_half, s_half], 1)[0, 0]) - 0.5) < 1e-12
    assert abs(float(det.getPd(1e-6, [s_half - 10 * np.log10(16.0)] * 2, 16)[0, 0]) - 0.5) < 1e-12
