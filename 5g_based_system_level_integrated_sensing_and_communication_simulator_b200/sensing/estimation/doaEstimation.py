"""``+sensing/+estimation/+doaEstimation`` mirror."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _lib


def _doa_config(radarEstParams) -> _lib.DoaConfig:
    ant = radarEstParams["antennaType"]
    is_upa = ant["type"] == "upa"
    return _lib.DoaConfig(
        isUpa=1 if is_upa else 0,
        nAnts=0 if is_upa else int(ant["nV"]) * int(ant["p"]),        # ula.numElements (ula.m)
        nX=int(ant["nV"]) if is_upa else 0, nY=int(ant["nH"]) if is_upa else 0, d=0.5,
        aGran=float(radarEstParams["azimuthScanGranularity"]), aMax=float(radarEstParams["azimuthScanScale"]),
        eGran=float(radarEstParams["elevationScanGranularity"]), eMax=float(radarEstParams["elevationScanScale"]))


_MUSIC, _MVDR, _DBF = 0, 1, 2   # ISAC_DOA_* of include/isac_b200.h


def _scan(method, numDets, radarEstParams, Ra):
    ctx = _lib.get_context(None)
    cfg = _doa_config(radarEstParams)
    n = cfg.nX * cfg.nY if cfg.isUpa else cfg.nAnts
    Ra = np.asfortranarray(np.asarray(Ra, dtype=np.complex128))
    if Ra.shape != (n, n):
        raise _lib.IsacError(1, f"Ra must be {n}x{n}")
    aSteps = int(np.floor((cfg.aMax + 1) / cfg.aGran))
    eSteps = int(np.floor((cfg.eMax + 1) / cfg.eGran)) if cfg.isUpa else 1
    L, nA = C.c_int32(), C.c_int32()
    azi = np.zeros(_lib.MAX_PEAKS)
    PdB = np.zeros(aSteps * eSteps)
    ctx.use_own_stream()
    nd = -1 if numDets is None else int(numDets)
    _lib.check(ctx.lib.isac_doa_scan_host(ctx.handle, C.byref(cfg), method, _lib.ptr(Ra), nd, C.byref(L), _lib.ptr(azi),
                                          C.byref(nA), _lib.ptr(PdB), None), ctx.handle)
    if cfg.isUpa:
        return L.value, None, None, PdB.reshape((eSteps, aSteps), order="F")
    a = azi[: nA.value].copy()
    return L.value, a, np.full(a.size, np.nan), PdB


def mvdrBF(numDets, radarEstParams, Ra, return_spectrum=False):
    """``[aziEst, eleEst] = sensing.estimation.doaEstimation.mvdrBF(numDets, radarEstParams, Ra)``
    (reference +sensing/+estimation/+doaEstimation/mvdrBF.m:1): scan of ``1./(aa'*Ra^-1*aa + eps(1))``.
    UPA: (None, None) as the reference's ``tools.find2DPeaks`` does not exist; ``return_spectrum`` appends PmvdrdB."""
    _, azi, ele, spec = _scan(_MVDR, numDets, radarEstParams, Ra)
    return (azi, ele, spec) if return_spectrum else (azi, ele)


def digitalBF(numDets, radarEstParams, Ra, return_spectrum=False):
    """``[aziEst, eleEst] = sensing.estimation.doaEstimation.digitalBF(numDets, radarEstParams, Ra)``
    (reference +sensing/+estimation/+doaEstimation/digitalBF.m:1): beamscan ``aa'*Ra*aa``."""
    _, azi, ele, spec = _scan(_DBF, numDets, radarEstParams, Ra)
    return (azi, ele, spec) if return_spectrum else (azi, ele)


def music(numDets, radarEstParams, Ra, return_spectrum=False):
    """``[L, aziEst, eleEst] = sensing.estimation.doaEstimation.music(numDets, radarEstParams, Ra)``
    (reference +sensing/+estimation/+doaEstimation/music.m:1).  ``numDets=None`` is MATLAB's ``[]``.

    ULA: (L, aziEst, eleEst=NaN...).  UPA: the reference stops at the missing ``tools.find2DPeaks``
    (music.m:69); this returns (L, None, None) and, with ``return_spectrum``, PmusicdB[eSteps x aSteps]."""
    ctx = _lib.get_context(None)
    cfg = _doa_config(radarEstParams)
    n = cfg.nX * cfg.nY if cfg.isUpa else cfg.nAnts
    Ra = np.asfortranarray(np.asarray(Ra, dtype=np.complex128))
    if Ra.shape != (n, n):
        raise _lib.IsacError(1, f"Ra must be {n}x{n}")
    aSteps = int(np.floor((cfg.aMax + 1) / cfg.aGran))
    eSteps = int(np.floor((cfg.eMax + 1) / cfg.eGran)) if cfg.isUpa else 1
    L, nA = C.c_int32(), C.c_int32()
    azi = np.zeros(_lib.MAX_PEAKS)
    PdB = np.zeros(aSteps * eSteps)
    ctx.use_own_stream()
    nd = -1 if numDets is None else int(numDets)
    _lib.check(ctx.lib.isac_music_doa_host(ctx.handle, C.byref(cfg), _lib.ptr(Ra), nd, C.byref(L), _lib.ptr(azi),
                                           C.byref(nA), _lib.ptr(PdB), None), ctx.handle)
    if cfg.isUpa:
        out = (L.value, None, None)
        spec = PdB.reshape((eSteps, aSteps), order="F")
    else:
        a = azi[: nA.value].copy()
        out = (L.value, a, np.full(a.size, np.nan))
        spec = PdB
    return out + (spec,) if return_spectrum else out
