"""``+communication/+pathlossModels`` mirror and the link-budget tail of ``uePhy.applyChannelModel`` / ``gNBPhy.applyChannelModel``
(uePhy.m:735-755, :935-950): TR 38.901 path loss for batches of links on the device, receive-gain / path-loss scaling of the
device-resident channel matrices, thermal noise power, DFT fallback channel matrix."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib

_SCENARIOS = {"uma": 0, "umi": 1, "rma": 2, "inh": 3, "fspl": 4}


def _pos(p, n=None):
    a = np.atleast_2d(np.asarray(p, dtype=np.float64))
    if a.shape[-1] != 3 and a.shape[0] == 3:
        a = a.T                                  # MATLAB column vectors [3 x n]
    if n is not None and a.shape[0] == 1 and n > 1:
        a = np.repeat(a, n, axis=0)
    return np.ascontiguousarray(a)


def config5GNRModels(pathLossConfig, carrierFreq, losCondition, bsPosition, uePosition, device=None):
    """``pathLoss = communication.pathlossModels.config5GNRModels(pathLossConfig, carrierFreq, losCondition, bsPosition,
    uePosition)`` (config5GNRModels.m:1; TR 38.901 7.4.1 via nrPathLoss) for one link or a batch: positions [n x 3] (one of
    them may be a single position), ``losCondition`` scalar or [n].  'UMa' | 'UMi' | 'RMa' | 'InH' (the InF-* scenarios raise).
    Returns dB, scalar for one link else [n]."""
    key = str(pathLossConfig).lower()
    if key not in _SCENARIOS or key == "fspl":
        raise _lib.IsacError(1, f"config5GNRModels: scenario {pathLossConfig!r} is not built (UMa, UMi, RMa, InH)")
    return _run(_SCENARIOS[key], carrierFreq, losCondition, bsPosition, uePosition, device)


def configFreeSpaceModel(carrierFreq, bsPosition, uePosition, device=None):
    """``pathLoss = communication.pathlossModels.configFreeSpaceModel(carrierFreq, bsPosition, uePosition)``
    (configFreeSpaceModel.m:1): fspl(distance, lambda)."""
    return _run(_SCENARIOS["fspl"], carrierFreq, 1, bsPosition, uePosition, device)


def _run(scn, fc, los, bs, ue, device):
    ue = _pos(ue)
    bs = _pos(bs, ue.shape[0])
    ue = _pos(ue, bs.shape[0])
    n = ue.shape[0]
    if bs.shape != ue.shape:
        raise _lib.IsacError(1, "bsPosition and uePosition must pair up ([n x 3] each, or one single position)")
    losv = np.ascontiguousarray(np.broadcast_to(np.asarray(los, dtype=np.int32).ravel(), (n,)) if np.size(los) in (1, n)
                                else np.zeros(0, np.int32))
    if losv.size != n:
        raise _lib.IsacError(1, "losCondition must be a scalar or one flag per link")
    out = np.zeros(n)
    ctx = _lib.get_context(device)
    ctx.use_own_stream()
    _lib.check(ctx.lib.isac_pathloss_host(ctx.handle, scn, float(fc), n, _lib.ptr(bs), _lib.ptr(ue), _lib.ptr(losv), _lib.ptr(out)),
               ctx.handle)
    return float(out[0]) if n == 1 else out


def applyPathLossAndRxGain(H, pathLossDb, rxGainDb):
    """``rxWaveform = db2mag(-pathLoss)*rxWaveform; applyRxGain`` (uePhy.m:748-751, :935-940) on device-resident channel
    matrices: ``H`` a torch CUDA complex64 tensor whose leading dimension indexes the links; scaled in place and returned."""
    pl = np.ascontiguousarray(np.atleast_1d(np.asarray(pathLossDb, dtype=np.float64)))
    n = pl.size
    if H.shape[0] != n and n == 1:
        H = H.unsqueeze(0)
    if H.shape[0] != n or not H.is_contiguous():
        raise _lib.IsacError(1, "H must be contiguous with one leading slice per link")
    ctx = _lib.get_context(H.device.index)
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_link_budget_dev(ctx.handle, _lib.ptr(H), H[0].numel(), n, _lib.ptr(pl), float(rxGainDb)), ctx.handle)
    return H


def thermalNoisePower(noiseFigureDb, temperature, sampleRate):
    """Nt of ``applyThermalNoise`` (uePhy.m:942-950): k (T + 290 (NF - 1)) fs."""
    out = C.c_double()
    st = _lib.load().isac_thermal_noise_power(float(noiseFigureDb), float(temperature), float(sampleRate), C.byref(out))
    if st:
        raise _lib.IsacError(st, "thermalNoisePower: invalid argument")
    return out.value


def dftChannelMatrix(numTxAnts, numRxAnts):
    """The channel matrix of the no-CDL branch (uePhy.m:735-739): fft(eye(max(nTx,nRx)))(1:nTx,1:nRx) / norm."""
    H = np.zeros((int(numTxAnts), int(numRxAnts)), dtype=np.complex128, order="F")
    st = _lib.load().isac_dft_channel_matrix(int(numTxAnts), int(numRxAnts), _lib.ptr(H))
    if st:
        raise _lib.IsacError(st, "dftChannelMatrix: invalid argument")
    return H
