"""``+sensing`` package mirror: radarParams, monoStaticSensing, channelModels, detection, estimation."""
from __future__ import annotations

import math

import numpy as np

from . import channelModels, detection, estimation, postProcessing  # noqa: F401
from ._echo import monoStaticSensing, ofdmModulate  # noqa: F401

LIGHTSPEED = 299792458.0   # physconst('Lightspeed')
BOLTZMANN = 1.380649e-23   # physconst('Boltzmann')


def _sind(x):
    """MATLAB sind (exact at multiples of 90 degrees)."""
    x = np.asarray(x, dtype=np.float64)
    r = np.fmod(x, 360.0)
    r = np.where(r > 180.0, r - 360.0, r)
    r = np.where(r < -180.0, r + 360.0, r)
    r = np.where(r > 90.0, 180.0 - r, r)
    r = np.where(r < -90.0, -180.0 - r, r)
    return np.where(np.abs(r) <= 45.0, np.sin(np.deg2rad(r)),
                    np.sign(r) * np.cos(np.deg2rad(90.0 - np.abs(r))))


def _cosd(x):
    return _sind(np.asarray(x, dtype=np.float64) + 90.0)


def radarParams(cellSimuParams, carrierInfo, waveInfo):
    """``radarParams = sensing.radarParams(cellSimuParams, carrierInfo, waveInfo)``
    (reference +sensing/radarParams.m:1).  Host-side scalar arithmetic in float64 (the reference's
    is negligible host code too); every quirk is kept: ``Tsri`` from ``ceil(nSc/8)`` CP samples
    (:34-35), ULA steering dividing the half-wavelength spacing by lambda again (:106-109)."""
    p = cellSimuParams
    nT = int(p["numTargets"])
    coords = (np.asarray(p["targetPosition"], float).reshape(nT, 3) - np.asarray(p["gNBPosition"], float).reshape(1, 3)).T
    x, y, z = coords
    azi = np.rad2deg(np.arctan2(y, x))
    ele = np.rad2deg(np.arctan2(z, np.hypot(x, y)))
    rng = np.sqrt(x * x + y * y + z * z)
    dlRatio = p["numDLSlots"] / len(p["tddPattern"])
    nDLSlots = dlRatio * p["numSlots"]
    nSc = carrierInfo["NRBsDL"] * 12
    nSym = nDLSlots * waveInfo["SymbolsPerSlot"]
    nTxAnts = int(p["gNBTxAnts"])
    c = LIGHTSPEED
    fc = float(p["dlCarrierFreq"])
    scs = carrierInfo["SubcarrierSpacing"] * 1e3
    lam = c / fc
    fs = float(waveInfo["SampleRate"])
    Ts = 1.0 / fs
    Tsri = 1.0 / scs + Ts * math.ceil(nSc / 8)
    NF = 10.0 ** (p["gNBNoiseFigure"] / 10.0)
    Teq = p["gNBTemperature"] + 290.0 * (NF - 1.0)
    N0 = fs * BOLTZMANN * Teq
    Pt = 10.0 ** ((p["gNBTxPower"] - 30.0) / 10.0) * math.sqrt(waveInfo["Nfft"] ** 2 / (nSc * nTxAnts))
    Ar = 10.0 ** (p["gNBRxGain"] / 10.0)
    rcs = np.asarray(p["rcs"], float).reshape(nT)
    v = np.asarray(p["velocity"], float).reshape(nT)
    Pr = Pt * Ar * Ar * (lam ** 2 * rcs) / ((4.0 * np.pi) ** 3 * rng ** 4)
    snrdB = 10.0 * np.log10(Pr / N0)
    out = dict(fc=fc, fs=fs, Tsri=Tsri, N0=N0, nTxAnts=nTxAnts, nTargets=nT, range=rng.copy(), velocity=v,
               largeScaleFading=np.sqrt(Pr / Pt), snrdB=snrdB, txPower=p["gNBTxPower"], Pfa=p["Pfa"])
    nIFFT = 2 ** max(0, math.ceil(math.log2(nSc)))
    out.update(nIFFT=nIFFT, rRes=c / (2.0 * scs * nIFFT), rMax=c / (2.0 * scs))
    nFFT = 2 ** max(0, math.ceil(math.log2(nSym)))
    out.update(nFFT=nFFT, vRes=lam / (2.0 * Tsri * nFFT), vMax=lam / (2.0 * Tsri))
    ant = p["gNBSenAntenna"]
    steer = np.zeros((nTxAnts, nT), dtype=np.complex128)
    if ant["type"] == "upa":
        ax = np.arange(ant["nV"], dtype=float)[None, :] * ant["dV"]
        ay = np.arange(ant["nH"], dtype=float)[:, None] * ant["dH"]
        for t in range(nT):
            a = np.exp(2j * np.pi * _sind(ele[t]) * (ax * _cosd(azi[t]) + ay * _sind(azi[t])) / lam)
            steer[:, t] = a.reshape(-1, order="F")
    else:
        ary = np.arange(nTxAnts, dtype=float) * ant["d"]
        for t in range(nT):
            steer[:, t] = np.exp(2j * np.pi * ary * _sind(azi[t]) / lam)
    out.update(antennaType=dict(ant), azimuthScanScale=360, elevationScanScale=180, azimuthScanGranularity=1,
               elevationScanGranularity=1, RxSteeringVec=steer,
               cfarEstZone=np.asarray(p["detectionArea"], float).reshape(2, 2))
    idx = np.argsort(-snrdB, kind="stable")
    out["targetRealPos"] = [dict(ID=i + 1, Range=rng[j], Velocity=v[j], Elevation=ele[j], Azimuth=azi[j],
                                 snrdB=snrdB[j]) for i, j in enumerate(idx)]
    return out
