"""``+sensing/+detection`` mirror."""
from __future__ import annotations

import numpy as np


def cfar2D(radaParams):
    """``cfarConfig = sensing.detection.cfar2D(radaParams)`` (reference +sensing/+detection/cfar2D.m:1).

    The reference returns a struct holding CUTIdx and a ``phased.CFARDetector2D`` System object;
    here ``cfarDetector2D`` is a plain dict with the same property names, consumed by
    ``sensing.estimation.fft2D`` (the detector itself runs in csrc/rdm.cu)."""
    nIFFT, nFFT = int(radaParams["nIFFT"]), int(radaParams["nFFT"])
    rngGrid = np.arange(nIFFT, dtype=float) * radaParams["rRes"]
    dopGrid = np.arange(-nFFT // 2, nFFT // 2, dtype=float) * radaParams["vRes"]
    zone = np.asarray(radaParams["cfarEstZone"], float)
    rngIdx = [int(np.argmin(np.abs(rngGrid - zone[0, j]))) + 1 for j in range(2)]
    dopIdx = [int(np.argmin(np.abs(dopGrid - zone[1, j]))) + 1 for j in range(2)]
    rows = np.arange(rngIdx[0], rngIdx[1] + 1)
    cols = np.arange(dopIdx[0], dopIdx[1] + 1)
    cc, rr = np.meshgrid(cols, rows)
    CUTIdx = np.stack([rr.reshape(-1, order="F"), cc.reshape(-1, order="F")], axis=0)
    detector = {"Method": "CA", "ThresholdFactor": "Auto", "ProbabilityFalseAlarm": float(radaParams["Pfa"]),
                "OutputFormat": "Detection index", "GuardBandSize": (2, 2), "TrainingBandSize": (1, 1)}
    return {"CUTIdx": CUTIdx, "cfarDetector2D": detector, "rngIdx": rngIdx, "dopIdx": dopIdx}


def getPd(Pfa, snrdB, nPulses=1):
    """``Pd = sensing.detection.getPd(Pfa, snrdB, nPulses)`` (reference +sensing/+detection/getPd.m:1): detection probability
    over the SNR grid ``snrdB`` for every false-alarm probability in ``Pfa``; returns ``Pd[len(snrdB), len(Pfa)]`` like
    ``rocpfa`` (the reference's figure is not drawn).

    ``rocpfa(Pfa,'MaxSNR',..,'MinSNR',..,'NumPoints',..,'NumPulses',N)`` is a closed Phased Array System Toolbox function;
    with its default signal type ('NonfluctuatingCoherent') it evaluates the textbook receiver operating characteristic of a
    non-fluctuating target in complex white Gaussian noise with coherent detection, N pulses integrated coherently
    (PARITY-UNPINNED against the toolbox):   Pd = 1/2 erfc( erfcinv(2 Pfa) - sqrt(N * SNR) ).
    The SNR points are ``linspace(snrdB(1), snrdB(end), numel(snrdB))`` exactly as the reference passes them (getPd.m:9-12),
    i.e. a non-uniform ``snrdB`` is resampled uniformly, as in the reference.  Host-only, as in the reference."""
    from math import sqrt
    from scipy.special import erfc, erfcinv
    pfa = np.atleast_1d(np.asarray(Pfa, dtype=np.float64)).ravel()
    snr_in = np.atleast_1d(np.asarray(snrdB, dtype=np.float64)).ravel()
    if snr_in.size < 1 or np.any(pfa <= 0) or np.any(pfa >= 1) or nPulses < 1:
        raise ValueError("getPd: Pfa must lie in (0,1), snrdB must be non-empty and nPulses >= 1")
    snr_db = np.linspace(snr_in[0], snr_in[-1], snr_in.size)
    snr = 10.0 ** (snr_db / 10.0)
    return 0.5 * erfc(erfcinv(2.0 * pfa)[None, :] - np.sqrt(float(nPulses) * snr)[:, None])
