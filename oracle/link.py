"""CPU oracle (test infrastructure only): link budget of the channel application step.

Restates +communication/+pathlossModels/config5GNRModels.m:27-36 (-> nrPathLoss, 5G Toolbox: TR 38.901 Table 7.4.1-1 without
shadow fading; toolbox defaults EnvironmentHeight 1 m, BuildingHeight 5 m, StreetWidth 20 m), configFreeSpaceModel.m:1-8
(fspl), and uePhy.m:735-751, :935-950 (DFT fallback matrix, path loss and Rx gain scaling, thermal noise power).
PARITY-UNPINNED: nrPathLoss is closed toolbox code; the formulas are those of TR 38.901 v16 Table 7.4.1-1 as published."""
import math

import numpy as np

LIGHTSPEED = 299792458.0
BOLTZMANN = 1.380649e-23


def _uma_umi(uma, fc, los, d2, d3, h_bs, h_ut):
    f = fc / 1e9
    d_bp = 4.0 * (h_bs - 1.0) * (h_ut - 1.0) * fc / LIGHTSPEED
    a, b1, c2 = (28.0, 22.0, 9.0) if uma else (32.4, 21.0, 9.5)
    if d2 <= d_bp:
        pl_los = a + b1 * math.log10(d3) + 20 * math.log10(f)
    else:
        pl_los = a + 40 * math.log10(d3) + 20 * math.log10(f) - c2 * math.log10(d_bp ** 2 + (h_bs - h_ut) ** 2)
    if los:
        return pl_los
    if uma:
        pl_n = 13.54 + 39.08 * math.log10(d3) + 20 * math.log10(f) - 0.6 * (h_ut - 1.5)
    else:
        pl_n = 35.3 * math.log10(d3) + 22.4 + 21.3 * math.log10(f) - 0.3 * (h_ut - 1.5)
    return max(pl_los, pl_n)


def _rma(fc, los, d2, d3, h_bs, h_ut, h=5.0, w=20.0):
    f = fc / 1e9
    d_bp = 2 * math.pi * h_bs * h_ut * fc / LIGHTSPEED

    def pl1(d):
        return (20 * math.log10(40 * math.pi * d * f / 3) + min(0.03 * h ** 1.72, 10) * math.log10(d)
                - min(0.044 * h ** 1.72, 14.77) + 0.002 * math.log10(h) * d)
    pl_los = pl1(d3) if d2 <= d_bp else pl1(d_bp) + 40 * math.log10(d3 / d_bp)
    if los:
        return pl_los
    pl_n = (161.04 - 7.1 * math.log10(w) + 7.5 * math.log10(h) - (24.37 - 3.7 * (h / h_bs) ** 2) * math.log10(h_bs)
            + (43.42 - 3.1 * math.log10(h_bs)) * (math.log10(d3) - 3) + 20 * math.log10(f)
            - (3.2 * math.log10(11.75 * h_ut) ** 2 - 4.97))
    return max(pl_los, pl_n)


def path_loss(scenario, fc, los, bs, ue):
    """config5GNRModels (scenario 'UMa' | 'UMi' | 'RMa' | 'InH') / configFreeSpaceModel ('fspl') for one link."""
    bs, ue = np.asarray(bs, float), np.asarray(ue, float)
    if np.array_equal(bs, ue):
        return 0.0                                                       # config5GNRModels.m:32-33
    d2 = float(np.hypot(ue[0] - bs[0], ue[1] - bs[1]))
    d3 = float(np.linalg.norm(ue - bs))
    if scenario == "UMa":
        return _uma_umi(True, fc, los, d2, d3, bs[2], ue[2])
    if scenario == "UMi":
        return _uma_umi(False, fc, los, d2, d3, bs[2], ue[2])
    if scenario == "RMa":
        return _rma(fc, los, d2, d3, bs[2], ue[2])
    if scenario == "InH":
        f = fc / 1e9
        pl = 32.4 + 17.3 * math.log10(d3) + 20 * math.log10(f)
        return pl if los else max(pl, 38.3 * math.log10(d3) + 17.30 + 24.9 * math.log10(f))
    if scenario == "fspl":
        return max(20 * math.log10(4 * math.pi * d3 * fc / LIGHTSPEED), 0.0)   # fspl clips at 0 dB
    raise ValueError(scenario)


def thermal_noise_power(noise_figure_db, temperature, sample_rate):
    """Nt of applyThermalNoise (uePhy.m:945-947)."""
    nf = 10 ** (noise_figure_db / 10)
    return BOLTZMANN * (temperature + 290 * (nf - 1)) * sample_rate


def apply_link_budget(H, path_loss_db, rx_gain_db):
    """db2mag(-pathLoss) * waveform, then applyRxGain (uePhy.m:748-751, :935-940), on a channel matrix."""
    return np.asarray(H) * 10 ** (-path_loss_db / 20) * 10 ** (rx_gain_db / 20)


def dft_channel_matrix(n_tx, n_rx):
    """uePhy.m:735-739: H = fft(eye(max(nTx,nRx))); H = H(1:nTx,1:nRx); H = H/norm(H)."""
    m = max(n_tx, n_rx)
    H = np.fft.fft(np.eye(m))[:n_tx, :n_rx]
    return H / np.linalg.norm(H, 2)
