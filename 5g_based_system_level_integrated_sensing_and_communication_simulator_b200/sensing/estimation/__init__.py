"""``+sensing/+estimation`` mirror: fft2D, music2D, doaEstimation.music."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _lib
from . import doaEstimation  # noqa: F401
from .doaEstimation import _doa_config

_plans = {}


class SensePlan:
    """Device plan of the fft2D estimator (RDM + CFAR + covariance + MUSIC), see csrc/sense.cu."""

    def __init__(self, radarEstParams, cfar, grid_shape, max_batch=1, device=None, ctx=None):
        # ctx: a dedicated library context (own stream / scratch buffers), e.g. to run the sensing pass on a second
        # CUDA stream concurrently with the COMM kernels of the process-wide context
        self.ctx = ctx if ctx is not None else _lib.get_context(device)
        lib = self.ctx.lib
        nSc, nSym, nAnts = grid_shape
        det = cfar["cfarDetector2D"]
        if det.get("Method", "CA") != "CA":
            raise _lib.IsacError(4, "only the CA detector (cfar2D.m:28) is implemented")
        cut = np.asarray(cfar["CUTIdx"])
        rows = (int(cut[0].min()), int(cut[0].max()))
        cols = (int(cut[1].min()), int(cut[1].max()))
        if cut.shape[1] != (rows[1] - rows[0] + 1) * (cols[1] - cols[0] + 1):
            raise _lib.IsacError(4, "CUTIdx must be the full rectangle built by sensing.detection.cfar2D")
        self.rdm_cfg = _lib.RdmConfig(
            nSc=nSc, nSym=nSym, nAnts=nAnts, nIFFT=int(radarEstParams["nIFFT"]), nFFT=int(radarEstParams["nFFT"]),
            cutRow0=rows[0], cutRow1=rows[1], cutCol0=cols[0], cutCol1=cols[1],
            guardRows=det["GuardBandSize"][0], guardCols=det["GuardBandSize"][1],
            trainRows=det["TrainingBandSize"][0], trainCols=det["TrainingBandSize"][1],
            maxBatch=max_batch, pfa=float(det["ProbabilityFalseAlarm"]), kaiserBeta=3.0)
        self.doa_cfg = _doa_config(radarEstParams)
        h = C.c_void_p()
        _lib.check(lib.isac_sense_plan_create(self.ctx.handle, C.byref(self.rdm_cfg), C.byref(self.doa_cfg),
                                              float(radarEstParams["rRes"]), float(radarEstParams["vRes"]),
                                              C.byref(h)), self.ctx.handle)
        self.handle = h
        self.lib = lib
        self.max_batch = max_batch
        self.rdm_handle = C.c_void_p(lib.isac_sense_plan_rdm(h))
        self.spec_len = int(np.floor((self.doa_cfg.aMax + 1) / self.doa_cfg.aGran)) * (
            int(np.floor((self.doa_cfg.eMax + 1) / self.doa_cfg.eGran)) if self.doa_cfg.isUpa else 1)

    def run_dev(self, rx_dev, tx_dev, batch=1, power_out=None):
        self.ctx.use_torch_stream()
        _lib.check(self.lib.isac_fft2d_dev(self.handle, _lib.ptr(rx_dev), _lib.ptr(tx_dev), batch,
                                           _lib.ptr(power_out)), self.ctx.handle)

    def _alloc_out(self, batch, max_out):
        return dict(rng=np.zeros((batch, max_out)), nr=np.zeros(batch, np.int32), vel=np.zeros((batch, max_out)),
                    nv=np.zeros(batch, np.int32), azi=np.zeros((batch, _lib.MAX_PEAKS)), na=np.zeros(batch, np.int32),
                    L=np.zeros(batch, np.int32), st=np.zeros(batch, np.int32))

    @staticmethod
    def _to_results(o, batch):
        res = []
        for b in range(batch):
            azi = o["azi"][b, : o["na"][b]].copy()
            res.append({"rngEst": o["rng"][b, : o["nr"][b]].copy(), "velEst": o["vel"][b, : o["nv"][b]].copy(),
                        "aziEst": azi, "eleEst": np.full(azi.size, np.nan), "L": int(o["L"][b]),
                        "status": int(o["st"][b])})
        return res

    def collect(self, batch=1, max_out=4096):
        o = self._alloc_out(batch, max_out)
        _lib.check(self.lib.isac_fft2d_collect(self.handle, batch, max_out, _lib.ptr(o["rng"]), _lib.ptr(o["nr"]),
                                               _lib.ptr(o["vel"]), _lib.ptr(o["nv"]), _lib.ptr(o["azi"]),
                                               _lib.ptr(o["na"]), _lib.ptr(o["L"]), _lib.ptr(o["st"])), self.ctx.handle)
        return self._to_results(o, batch)

    def run_host(self, rx, tx, batch=1, max_out=4096):
        rx, tx = _lib.as_c64(rx), _lib.as_c64(tx)
        o = self._alloc_out(batch, max_out)
        self.ctx.use_own_stream()
        _lib.check(self.lib.isac_fft2d_host(self.handle, _lib.ptr(rx), _lib.ptr(tx), batch, max_out, _lib.ptr(o["rng"]),
                                            _lib.ptr(o["nr"]), _lib.ptr(o["vel"]), _lib.ptr(o["nv"]), _lib.ptr(o["azi"]),
                                            _lib.ptr(o["na"]), _lib.ptr(o["L"]), _lib.ptr(o["st"])), self.ctx.handle)
        return self._to_results(o, batch)

    def spectrum(self, batch=1):
        buf = np.zeros((batch, self.spec_len))
        _lib.check(self.lib.isac_fft2d_get_spectrum(self.handle, batch, _lib.ptr(buf)), self.ctx.handle)
        return buf

    def detections(self, batch=1, max_det=None):
        nA = self.rdm_cfg.nAnts
        nCut = (self.rdm_cfg.cutRow1 - self.rdm_cfg.cutRow0 + 1) * (self.rdm_cfg.cutCol1 - self.rdm_cfg.cutCol0 + 1)
        max_det = int(max_det or nCut)
        cnt = np.zeros(nA * batch, np.int32)
        rc = np.zeros((nA * batch, max_det, 2), np.int32)
        pk = np.zeros((nA * batch, max_det), np.float32)
        _lib.check(self.lib.isac_rdm_get_detections(self.rdm_handle, batch, max_det, _lib.ptr(cnt), _lib.ptr(rc),
                                                    _lib.ptr(pk)), self.ctx.handle)
        return cnt.reshape(batch, nA), [[(rc[b * nA + r, :cnt[b * nA + r], :].T.copy(), pk[b * nA + r, :cnt[b * nA + r]].copy())
                                         for r in range(nA)] for b in range(batch)]

    def power(self, batch=1):
        nI, nF, nA = self.rdm_cfg.nIFFT, self.rdm_cfg.nFFT, self.rdm_cfg.nAnts
        buf = np.zeros(nI * nF * nA * batch, np.float32)
        _lib.check(self.lib.isac_rdm_get_power(self.rdm_handle, batch, _lib.ptr(buf)), self.ctx.handle)
        return buf.reshape((nI, nF, nA, batch), order="F")

    def close(self):
        if self.handle:
            self.lib.isac_sense_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _is_device_tensor(a):
    return hasattr(a, "is_cuda") and a.is_cuda


def fft2D(radarEstParams, cfar, rxGrid, txGrid):
    """``estResults = sensing.estimation.fft2D(radarEstParams, cfar, rxGrid, txGrid)``
    (reference +sensing/+estimation/fft2D.m:1).

    rxGrid/txGrid: [nSc x nSym x nAnts] complex (NumPy, any float precision; converted to complex64 like
    the MEX gateway does) or torch CUDA complex64 tensors laid out [nAnts][nSym][nSc].
    Returns dict(rngEst, velEst, aziEst, eleEst).  Errors of the reference (CUT window outside the map,
    zero detections -> findpeaks NPeaks=0) raise ``IsacError`` so a caller's try/except mirrors
    cellSimulation.m:196-202.  The reference's figure (plotRDM, :119) is not produced."""
    if _is_device_tensor(rxGrid):
        nAnts, nSym, nSc = rxGrid.shape
        dev = rxGrid.device.index
    else:
        nSc, nSym, nAnts = np.asarray(rxGrid).shape
        dev = None
    det = cfar["cfarDetector2D"]
    key = (nSc, nSym, nAnts, int(radarEstParams["nIFFT"]), int(radarEstParams["nFFT"]), float(radarEstParams["rRes"]),
           float(radarEstParams["vRes"]), tuple(np.asarray(cfar["CUTIdx"])[:, [0, -1]].ravel().tolist()),
           float(det["ProbabilityFalseAlarm"]), tuple(det["GuardBandSize"]), tuple(det["TrainingBandSize"]),
           repr(sorted(radarEstParams["antennaType"].items())), dev)
    plan = _plans.get(key)
    if plan is None:
        plan = SensePlan(radarEstParams, cfar, (nSc, nSym, nAnts), 1, dev)
        _plans[key] = plan
    if _is_device_tensor(rxGrid):
        plan.run_dev(rxGrid, txGrid, 1)
        res = plan.collect(1)[0]
    else:
        res = plan.run_host(rxGrid, txGrid, 1)[0]
    if res["status"] != 0:
        raise _lib.IsacError(res["status"], "fft2D: no CFAR detection, MUSIC needs numDets >= 1 (music.m:102)")
    return {"rngEst": res["rngEst"], "velEst": res["velEst"], "aziEst": res["aziEst"], "eleEst": res["eleEst"]}


def music2D(rdrEstParams, bsParams, rxGrid, txGrid, numDets=None):
    """``estResults = sensing.estimation.music2D(rdrEstParams, bsParams, rxGrid, txGrid)``
    (reference +sensing/+estimation/music2D.m:1).  ``numDets`` (not in the reference signature)
    overrides the eigen-gap source-count rule when given.  Returns the reference's fields plus the
    dB pseudo-spectra (``PrmusicdB``, ``PvmusicdB``) that the reference only plots."""
    import torch
    ctx = _lib.get_context(None)
    if _is_device_tensor(rxGrid):
        rx_d, tx_d = rxGrid, txGrid
        nAnts, nSym, nSc = rxGrid.shape
    else:
        rx = np.asarray(rxGrid)
        nSc, nSym, nAnts = rx.shape
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.complex64).transpose(2, 1, 0))).cuda()
        rx_d, tx_d = to_dev(rxGrid), to_dev(txGrid)
    zone = np.asarray(rdrEstParams["cfarEstZone"], float)
    cfg = _lib.Music2dConfig(nSc=nSc, nSym=nSym, nAnts=nAnts, scsHz=float(bsParams["scs"]) * 1e3,
                             fc=float(rdrEstParams["fc"]), Tsri=float(rdrEstParams["Tsri"]), rMax=float(zone[0, 1]),
                             vZone=float(zone[1, 1]), doa=_doa_config(rdrEstParams),
                             numDetsOverride=int(numDets or 0))
    rSteps = int(np.floor((zone[0, 1] + 1) / 0.5))
    vSteps = int(np.floor((zone[1, 1] * 2 + 1) / 0.5))
    L, nA, nR, nV, sw = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    azi, rng, vel = (np.zeros(_lib.MAX_PEAKS) for _ in range(3))
    PrdB, PvdB = np.zeros(rSteps), np.zeros(vSteps)
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_music2d_dev(ctx.handle, C.byref(cfg), _lib.ptr(rx_d), _lib.ptr(tx_d), C.byref(L),
                                        _lib.ptr(azi), C.byref(nA), _lib.ptr(rng), C.byref(nR), _lib.ptr(vel),
                                        C.byref(nV), _lib.ptr(PrdB), _lib.ptr(PvdB), C.byref(sw)), ctx.handle)
    a = azi[: nA.value].copy()
    return {"aziEst": a, "eleEst": np.full(a.size, np.nan), "rngEst": rng[: nR.value].copy(),
            "velEst": vel[: nV.value].copy(), "L": L.value, "PrmusicdB": PrdB, "PvmusicdB": PvdB,
            "jacobiSweeps": sw.value}
