"""Device plan for the 2D-FFT range-Doppler map + 2D CA-CFAR kernels (csrc/rdm.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib


class RangeDopplerPlan:
    """Owns the windows, the L2-resident range-profile intermediate and the CFAR result buffers.

    Parameters mirror ``radarEstParams`` / ``cfarConfig`` of the reference
    (+sensing/radarParams.m:69-77, +sensing/+detection/cfar2D.m:15-33).
    """

    def __init__(self, nSc, nSym, nAnts, nIFFT, nFFT, cut_rows, cut_cols, pfa,
                 guard=(2, 2), train=(1, 1), kaiser_beta=3.0, max_batch=1, device=None):
        self.ctx = _lib.get_context(device)
        self.lib = self.ctx.lib
        cfg = _lib.RdmConfig(
            nSc=nSc, nSym=nSym, nAnts=nAnts, nIFFT=nIFFT, nFFT=nFFT,
            cutRow0=int(cut_rows[0]), cutRow1=int(cut_rows[1]),
            cutCol0=int(cut_cols[0]), cutCol1=int(cut_cols[1]),
            guardRows=guard[0], guardCols=guard[1], trainRows=train[0], trainCols=train[1],
            maxBatch=max_batch, pfa=float(pfa), kaiserBeta=float(kaiser_beta))
        self.cfg = cfg
        h = C.c_void_p()
        _lib.check(self.lib.isac_rdm_plan_create(self.ctx.handle, C.byref(cfg), C.byref(h)), self.ctx.handle)
        self.handle = h
        a, nt, nc = C.c_double(), C.c_int32(), C.c_int32()
        _lib.check(self.lib.isac_rdm_plan_info(h, C.byref(a), C.byref(nt), C.byref(nc)), self.ctx.handle)
        self.alpha, self.nTrain, self.nCut = a.value, nt.value, nc.value
        self.shape_grid = (nSc, nSym, nAnts)
        self.shape_rdm = (nIFFT, nFFT, nAnts)
        self.max_batch = max_batch

    def set_variant(self, variant):
        """Range-kernel selection for nIFFT = 4096 (isac_rdm_plan_set_variant): 0 lean persistent TMA kernel (default),
        1 first TMA kernel, 2 one CTA per column.  Same results to rounding; used by the parity tests and A/B timing."""
        _lib.check(self.lib.isac_rdm_plan_set_variant(self.handle, int(variant)), self.ctx.handle)

    # -- device path -------------------------------------------------------------------------
    def run_dev(self, rx_dev, tx_dev, batch=1, power_out=None):
        """rx_dev/tx_dev: torch CUDA complex64 tensors laid out [batch][nAnts][nSym][nSc]
        (== MATLAB [nSc x nSym x nAnts x batch]).  Enqueues on torch's current stream."""
        self.ctx.use_torch_stream()
        _lib.check(self.lib.isac_rdm_cfar_dev(self.handle, _lib.ptr(rx_dev), _lib.ptr(tx_dev), batch,
                                              _lib.ptr(power_out)), self.ctx.handle)

    def cfar_dev(self, power_dev, batch=1):
        self.ctx.use_torch_stream()
        _lib.check(self.lib.isac_cfar2d_dev(self.handle, _lib.ptr(power_dev), batch), self.ctx.handle)

    def detections(self, batch=1, max_det=None):
        """-> (counts[batch][nAnts], list over batch of list over antennas of (rowcol[2xN] int32 1-based, peaks[N]))."""
        nA = self.cfg.nAnts
        max_det = int(max_det or self.nCut)
        cnt = np.zeros(nA * batch, dtype=np.int32)
        rc = np.zeros((nA * batch, max_det, 2), dtype=np.int32)
        pk = np.zeros((nA * batch, max_det), dtype=np.float32)
        _lib.check(self.lib.isac_rdm_get_detections(self.handle, batch, max_det, _lib.ptr(cnt), _lib.ptr(rc),
                                                    _lib.ptr(pk)), self.ctx.handle)
        out = []
        for b in range(batch):
            per = []
            for r in range(nA):
                n = int(cnt[b * nA + r])
                per.append((rc[b * nA + r, :n, :].T.copy(), pk[b * nA + r, :n].copy()))
            out.append(per)
        return cnt.reshape(batch, nA), out

    def power(self, batch=1):
        """Host copy of the last power map, MATLAB-shaped [nIFFT x nFFT x nAnts x batch] (Fortran order)."""
        nI, nF, nA = self.shape_rdm
        buf = np.zeros(nI * nF * nA * batch, dtype=np.float32)
        _lib.check(self.lib.isac_rdm_get_power(self.handle, batch, _lib.ptr(buf)), self.ctx.handle)
        return buf.reshape((nI, nF, nA, batch), order="F")

    # -- host path ---------------------------------------------------------------------------
    def run_host(self, rx, tx, batch=1, want_power=False, max_det=None):
        """rx/tx: numpy complex64 Fortran-ordered [nSc x nSym x nAnts (x batch)] (pinned or pageable)."""
        rx = _lib.as_c64(rx)
        tx = _lib.as_c64(tx)
        nA = self.cfg.nAnts
        max_det = int(max_det or self.nCut)
        cnt = np.zeros(nA * batch, dtype=np.int32)
        rc = np.zeros((nA * batch, max_det, 2), dtype=np.int32)
        pk = np.zeros((nA * batch, max_det), dtype=np.float32)
        nI, nF, _ = self.shape_rdm
        pw = np.zeros(nI * nF * nA * batch, dtype=np.float32) if want_power else None
        self.ctx.use_own_stream()
        _lib.check(self.lib.isac_rdm_cfar_host(self.handle, _lib.ptr(rx), _lib.ptr(tx), batch, max_det,
                                               _lib.ptr(cnt), _lib.ptr(rc), _lib.ptr(pk), _lib.ptr(pw)),
                   self.ctx.handle)
        out = []
        for b in range(batch):
            per = []
            for r in range(nA):
                n = int(cnt[b * nA + r])
                per.append((rc[b * nA + r, :n, :].T.copy(), pk[b * nA + r, :n].copy()))
            out.append(per)
        if want_power:
            pw = pw.reshape((nI, nF, nA, batch), order="F")
        return cnt.reshape(batch, nA), out, pw

    def close(self):
        if self.handle:
            self.lib.isac_rdm_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
