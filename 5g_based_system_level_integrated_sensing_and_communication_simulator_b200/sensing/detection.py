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
