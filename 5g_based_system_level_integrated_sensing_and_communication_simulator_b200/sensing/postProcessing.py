"""``+sensing/+postProcessing`` mirror: host-side bookkeeping on the estimator outputs (no device work, as in the reference)."""
from __future__ import annotations

import numpy as np


def getRMSE(radarEstResults, radarEstParams):
    """``radarEstRMSE = sensing.postProcessing.getRMSE(radarEstResults, radarEstParams)``
    (reference +sensing/+postProcessing/getRMSE.m:1).

    ``radarEstResults``: struct (dict) or struct array (list of dicts) with ``rngEst, velEst, aziEst[, eleEst]``, the
    fields concatenated like ``extractfield`` does (:22-29).  ``radarEstParams``: ``rRes`` and ``tgtRealPos`` (list of
    dicts with ``Range, Velocity, Elevation, Azimuth``, radarParams.m:131-144) and ``antennaType``.
    Every estimate r is matched to the FIRST true target whose range lies within ``rRes`` (:43-52); unmatched estimates
    give NaN (``sqrt(mean(rmmissing(NaN).^2))`` = NaN, :55-58).  Returns NaN when nothing was detected (:31-35).

    Quirks kept: the UPA test is ``isa(antennaType, 'phased.NRRectangularPanelArray')`` (:13), which the simulator's own
    ``parameters.baseStation.antenna.upa`` objects never satisfy, so ``eleRMSE`` is all-NaN unless the caller passes
    ``antennaType = {"type": "phased.NRRectangularPanelArray"}``; the per-detection "RMSE" is the absolute error."""
    res = radarEstResults if isinstance(radarEstResults, (list, tuple)) else [radarEstResults]
    cat = lambda f: np.concatenate([np.atleast_1d(np.asarray(r[f], dtype=np.float64)).ravel() for r in res]) if res else np.zeros(0)
    tgt = radarEstParams["tgtRealPos"]
    real = {f: np.array([float(t[f]) for t in tgt]) for f in ("Range", "Velocity", "Elevation", "Azimuth")}
    is_upa = radarEstParams.get("antennaType", {}).get("type") == "phased.NRRectangularPanelArray"
    rng, vel, azi = cat("rngEst"), cat("velEst"), cat("aziEst")
    ele = cat("eleEst") if is_upa else None
    if rng.size == 0:
        return float("nan")
    n = rng.size
    err = {k: np.full(n, np.nan) for k in ("rng", "vel", "ele", "azi")}
    for r in range(n):
        idx = np.flatnonzero(np.abs(real["Range"] - rng[r]) < radarEstParams["rRes"])
        if idx.size >= 1:
            i = idx[0]
            err["rng"][r] = real["Range"][i] - rng[r]
            err["vel"][r] = real["Velocity"][i] - vel[r]       # MATLAB errors if velEst is shorter than rngEst; so does this
            if is_upa:
                err["ele"][r] = real["Elevation"][i] - ele[r]
            err["azi"][r] = real["Azimuth"][i] - azi[r]
    return {k + "RMSE": np.sqrt(err[k] ** 2) for k in ("rng", "vel", "ele", "azi")}
