"""K1+K2 host mirror: basicRadarChannel / monoStaticSensing (csrc/echo.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ..workloads import ofdm_numerology


class _EchoArgs:
    """Keeps the host arrays referenced by an isac_echo_config alive for the duration of a call."""

    def __init__(self, T, nTx, radarParams, los, carrierInfo=None, nSymTx=0):
        nT = int(radarParams["nTargets"])
        self.range = np.ascontiguousarray(radarParams["range"], dtype=np.float64).reshape(nT)
        self.vel = np.ascontiguousarray(radarParams["velocity"], dtype=np.float64).reshape(nT)
        self.lsf = np.ascontiguousarray(radarParams["largeScaleFading"], dtype=np.float64).reshape(nT)
        self.steer = np.asfortranarray(np.asarray(radarParams["RxSteeringVec"], dtype=np.complex128).reshape(nTx, nT))
        self.los = np.ascontiguousarray(los, dtype=np.int32).reshape(nT)
        if carrierInfo is not None:
            num = ofdm_numerology(int(carrierInfo["NRBsDL"]), float(carrierInfo["SubcarrierSpacing"]))
            self.cp = np.ascontiguousarray(num["CyclicPrefixLengths"], dtype=np.int32)
            nfft, nSc = int(num["Nfft"]), 12 * int(carrierInfo["NRBsDL"])
        else:
            self.cp = np.zeros(1, dtype=np.int32)
            nfft, nSc = 0, 0
        self.cfg = _lib.EchoConfig(
            T=int(T), nTx=int(nTx), nTargets=nT, fc=float(radarParams["fc"]), fs=float(radarParams["fs"]),
            N0=float(radarParams["N0"]), range=self.range.ctypes.data, velocity=self.vel.ctypes.data,
            largeScaleFading=self.lsf.ctypes.data, steeringVec=self.steer.ctypes.data, los=self.los.ctypes.data,
            nfft=nfft, nSc=nSc, nSymTx=int(nSymTx), symbolsPerSubframe=int(self.cp.size), cpLengths=self.cp.ctypes.data)


def _noise_args(noise, seed):
    if noise is None:
        return None, (_lib.NOISE_PHILOX if seed is not None else _lib.NOISE_NONE), int(seed or 0)
    return noise, _lib.NOISE_TENSOR, 0


def _is_dev(a):
    return hasattr(a, "is_cuda") and a.is_cuda


def basicRadarChannel(txWaveform, radarParams, targetLoSConditions, noise=None, seed=None):
    """``rxWaveform = sensing.channelModels.basicRadarChannel(txWaveform, radarParams, targetLoSConditions)``
    (reference +sensing/+channelModels/basicRadarChannel.m:1).

    ``noise``: the ``randn(size)+1j*randn(size)`` draw of :68 as an explicit [T x nTx] complex tensor
    (MATLAB's generator cannot be reproduced); ``noise=None, seed=k`` draws it on the device (Philox);
    ``noise=None, seed=None`` is noiseless.  torch CUDA tensors are laid out [nTx][T]."""
    import torch
    ctx = _lib.get_context(None)
    if _is_dev(txWaveform):
        tx_d = txWaveform
        nTx, T = tx_d.shape
        nz_d = noise
    else:
        tx = np.asarray(txWaveform)
        T, nTx = tx.shape
        tx_d = torch.from_numpy(np.ascontiguousarray(tx.astype(np.complex64).T)).cuda()
        nz_d = None if noise is None else torch.from_numpy(np.ascontiguousarray(np.asarray(noise).astype(np.complex64).T)).cuda()
    nz_d, mode, sd = _noise_args(nz_d, seed)
    args = _EchoArgs(T, nTx, radarParams, targetLoSConditions)
    out = torch.empty((nTx, T), dtype=torch.complex64, device=tx_d.device)
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_radar_channel_dev(ctx.handle, C.byref(args.cfg), _lib.ptr(tx_d), _lib.ptr(nz_d), mode, sd,
                                              _lib.ptr(out)), ctx.handle)
    if _is_dev(txWaveform):
        return out
    return out.cpu().numpy().T.copy()


def monoStaticSensing(txWaveform, txDimension, carrierInfo, radarParams, targetLoSConditions, noise=None, seed=None):
    """``echoGrid = sensing.monoStaticSensing(txWaveform, txDimension, carrierInfo, radarParams, targetLoSConditions)``
    (reference +sensing/monoStaticSensing.m:1): echo synthesis + nrOFDMDemodulate + zero padding, fused in
    one kernel.  Returns [nSc x nSym x nAnts] complex64 (NumPy in, NumPy out; torch CUDA [nTx][T] in,
    torch CUDA [nAnts][nSym][nSc] out).  See ``basicRadarChannel`` for ``noise`` / ``seed``."""
    import torch
    ctx = _lib.get_context(None)
    dev_in = _is_dev(txWaveform)
    if dev_in:
        tx_d = txWaveform
        nTx, T = tx_d.shape
        nz_d = noise
    else:
        tx = np.asarray(txWaveform)
        T, nTx = tx.shape
        tx_d = torch.from_numpy(np.ascontiguousarray(tx.astype(np.complex64).T)).cuda()
        nz_d = None if noise is None else torch.from_numpy(np.ascontiguousarray(np.asarray(noise).astype(np.complex64).T)).cuda()
    nz_d, mode, sd = _noise_args(nz_d, seed)
    args = _EchoArgs(T, nTx, radarParams, targetLoSConditions, carrierInfo, int(txDimension[1]))
    nsym = C.c_int32()
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_mono_static_sensing_dev(ctx.handle, C.byref(args.cfg), None, None, 0, 0, None,
                                                    C.byref(nsym)), ctx.handle)
    out = torch.empty((nTx, nsym.value, args.cfg.nSc), dtype=torch.complex64, device=tx_d.device)
    _lib.check(ctx.lib.isac_mono_static_sensing_dev(ctx.handle, C.byref(args.cfg), _lib.ptr(tx_d), _lib.ptr(nz_d), mode, sd,
                                                    _lib.ptr(out), C.byref(nsym)), ctx.handle)
    if dev_in:
        return out
    return out.cpu().numpy().transpose(2, 1, 0).copy()


class SensingTxAccumulator:
    """Device-resident ``senTxGrid`` / ``senTxWave`` of the gNB PHY's sensing tap (reference +phyLayer/gNBPhy.m:604-612): the
    reference appends every DL slot's grid and waveform with ``cat`` (O(n^2) copying over a run); here the buffers are
    allocated once for ``maxSymbols`` symbols and ``append(slotGrid, slotInSubframe)`` writes the slot's grid and its OFDM
    waveform (``scale * nrOFDMModulate``, optional windowing) straight to their place (isac_ofdm_modulate_ex_dev)."""

    def __init__(self, carrierInfo, nTx, maxSymbols, scale=1.0, windowing=0, device=None):
        import torch
        self.ctx = _lib.get_context(device)
        self.num = ofdm_numerology(int(carrierInfo["NRBsDL"]), float(carrierInfo["SubcarrierSpacing"]))
        self.nSc, self.nTx, self.maxSym = 12 * int(carrierInfo["NRBsDL"]), int(nTx), int(maxSymbols)
        self.scale, self.windowing = float(scale), int(windowing)
        self.cp = np.ascontiguousarray(self.num["CyclicPrefixLengths"], dtype=np.int32)
        # capacity in samples: a run of whole slots never needs more than maxSymbols worst-case symbol lengths
        self.maxT = int(self.maxSym * (self.num["Nfft"] + int(self.cp.max())))
        dev = f"cuda:{self.ctx.device}"
        self.senTxGrid = torch.zeros((self.nTx, self.maxSym, self.nSc), dtype=torch.complex64, device=dev)
        self.senTxWave = torch.zeros((self.nTx, self.maxT), dtype=torch.complex64, device=dev)
        self.nSym, self.T = 0, 0

    def append(self, slotGrid, slotInSubframe=0):
        """``slotGrid``: [nSc x nSymSlot x nTx] NumPy or torch CUDA [nTx][nSymSlot][nSc]; ``slotInSubframe`` selects the slot's
        place in the subframe's cyclic-prefix pattern (the long prefixes sit at the start of every half subframe)."""
        import torch
        if _is_dev(slotGrid):
            g = slotGrid
        else:
            g = torch.from_numpy(np.ascontiguousarray(np.asarray(slotGrid).astype(np.complex64).transpose(2, 1, 0))).to(self.senTxGrid.device)
        nTx, n, nSc = g.shape
        if nTx != self.nTx or nSc != self.nSc or self.nSym + n > self.maxSym:
            raise _lib.IsacError(8, "SensingTxAccumulator: slot grid does not fit (shape or capacity)")
        self.senTxGrid[:, self.nSym: self.nSym + n] = g                     # obj.senTxGrid = cat(2, obj.senTxGrid, txGrid)
        T = C.c_int64()
        src = self.senTxGrid[0, self.nSym]                                  # block start; antenna pages maxSym symbols apart
        self.ctx.use_torch_stream()
        _lib.check(self.ctx.lib.isac_ofdm_modulate_ex_dev(self.ctx.handle, _lib.ptr(src), self.nSc, n, self.nTx, self.maxSym,
                                                          int(self.num["Nfft"]), int(self.cp.size), self.cp.ctypes.data, self.scale,
                                                          self.windowing, int(slotInSubframe) * 14, _lib.ptr(self.senTxWave), self.maxT,
                                                          self.T, C.byref(T)), self.ctx.handle)
        self.nSym += n
        self.T += T.value                                                   # obj.senTxWave = cat(1, obj.senTxWave, txWaveform)

    def grid(self):
        return self.senTxGrid[:, : self.nSym]

    def wave(self):
        return self.senTxWave[:, : self.T]


def ofdmModulate(carrierInfo, txGrid, scale=1.0, windowing=0):
    """``txWaveform = scale * nrOFDMModulate(carrier, txGrid)`` -- the gNB PHY step that produces the waveform handed to
    ``monoStaticSensing`` (reference +phyLayer/gNBPhy.m:599, accumulation for sensing at :604-612), on the device
    (csrc/ofdm.cu).  ``windowing`` = N > 0 samples applies the raised-cosine windowing / overlap of nrOFDMModulate's 'Windowing'
    argument (documented scheme, PARITY-UNPINNED); the default 0 is plain CP-OFDM.

    ``txGrid``: [nSc x nSym x nAnts] NumPy -> returns [T x nAnts] NumPy complex64;
    torch CUDA [nAnts][nSym][nSc] -> torch CUDA [nAnts][T]."""
    import torch
    ctx = _lib.get_context(None)
    num = ofdm_numerology(int(carrierInfo["NRBsDL"]), float(carrierInfo["SubcarrierSpacing"]))
    cp = np.ascontiguousarray(num["CyclicPrefixLengths"], dtype=np.int32)
    dev_in = _is_dev(txGrid)
    if dev_in:
        g_d = txGrid.contiguous()
        nAnts, nSym, nSc = g_d.shape
    else:
        g = np.asarray(txGrid)
        if g.ndim == 2:
            g = g[:, :, None]
        nSc, nSym, nAnts = g.shape
        g_d = torch.from_numpy(np.ascontiguousarray(g.astype(np.complex64).transpose(2, 1, 0))).cuda()
    if nSc != 12 * int(carrierInfo["NRBsDL"]):
        raise _lib.IsacError(1, "txGrid must span 12*NRBsDL subcarriers")
    T = C.c_int64()
    ctx.use_torch_stream()
    if windowing:
        def call(dst, stride):
            return ctx.lib.isac_ofdm_modulate_ex_dev(ctx.handle, _lib.ptr(g_d), nSc, nSym, nAnts, 0, int(num["Nfft"]), int(cp.size),
                                                     cp.ctypes.data, float(scale), int(windowing), 0, dst, stride, 0, C.byref(T))
        _lib.check(call(None, 0), ctx.handle)
        out = torch.empty((nAnts, T.value), dtype=torch.complex64, device=g_d.device)
        _lib.check(call(_lib.ptr(out), T.value), ctx.handle)
    else:
        args = (ctx.handle, _lib.ptr(g_d), nSc, nSym, nAnts, int(num["Nfft"]), int(cp.size), cp.ctypes.data, float(scale))
        _lib.check(ctx.lib.isac_ofdm_modulate_dev(*args, None, C.byref(T)), ctx.handle)
        out = torch.empty((nAnts, T.value), dtype=torch.complex64, device=g_d.device)
        _lib.check(ctx.lib.isac_ofdm_modulate_dev(*args, _lib.ptr(out), C.byref(T)), ctx.handle)
    if dev_in:
        return out
    return out.cpu().numpy().T.copy()
