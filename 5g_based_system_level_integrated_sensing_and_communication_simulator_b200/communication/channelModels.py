"""``+communication/+channelModels`` mirror."""


def updateCDLModels(simuParams):
    """``delayProfile = communication.channelModels.updateCDLModels(simuParams)`` (updateCDLModels.m:7-15):
    CDL-D for LoS links, CDL-A otherwise (per UE)."""
    return ["CDL-D" if int(los) == 1 else "CDL-A" for los in simuParams["ueLoSConditions"]]


class CDLChannel:
    """Frequency-domain stand-in for the ``nrCDLChannel`` objects the reference builds per UE
    (+parameters/+channelModels/+communication/cdl.m:48-88): ``generate`` returns the channel matrix
    H[K x L x nRx x nTx] that nrChannelEstimate would hand to riSelect/cqiSelect (uePhy.m:897-907).
    Statistical parity only (see csrc/cdl.cu)."""

    PROFILES = {"CDL-A": 0, "CDL-B": 1, "CDL-C": 2, "CDL-D": 3, "CDL-E": 4}

    def __init__(self, DelayProfile="CDL-D", DelaySpread=300e-9, CarrierFrequency=3.5e9, MaximumDopplerShift=5.0,
                 TransmitAntennaArraySize=(1, 8, 2), ReceiveAntennaArraySize=(1, 1, 2), TransmitElement="38.901",
                 ReceiveElement="isotropic", Seed=73, device=None):
        import ctypes as C
        from .. import _lib
        self._lib, self._C = _lib, C
        self.ctx = _lib.get_context(device)
        cfg = _lib.CdlConfig(profile=self.PROFILES[DelayProfile], delaySpread=float(DelaySpread), fc=float(CarrierFrequency),
                             maxDoppler=float(MaximumDopplerShift), txSize=(C.c_int32 * 3)(*TransmitAntennaArraySize[:3]),
                             rxSize=(C.c_int32 * 3)(*ReceiveAntennaArraySize[:3]),
                             txPattern38901=int(TransmitElement == "38.901"), rxPattern38901=int(ReceiveElement == "38.901"),
                             seed=int(Seed))
        h = C.c_void_p()
        _lib.check(self.ctx.lib.isac_cdl_create(self.ctx.handle, C.byref(cfg), C.byref(h)), self.ctx.handle)
        self.handle = h
        self.nTx = int(TransmitAntennaArraySize[0] * TransmitAntennaArraySize[1] * TransmitAntennaArraySize[2])
        self.nRx = int(ReceiveAntennaArraySize[0] * ReceiveAntennaArraySize[1] * ReceiveAntennaArraySize[2])

    def setKernel(self, legacy_mma):
        """False (default): tcgen05/TMEM response kernel; True: legacy mma.sync kernel (isac_cdl_set_kernel)."""
        self._lib.check(self.ctx.lib.isac_cdl_set_kernel(self.handle, int(bool(legacy_mma))), self.ctx.handle)

    def rays(self):
        import numpy as np
        C, lib = self._C, self.ctx.lib
        ncl, nr, nrx, ntx = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        lib.isac_cdl_get_rays(self.handle, C.byref(ncl), C.byref(nr), C.byref(nrx), C.byref(ntx), None, None, None, None)
        tau, nu = np.zeros(ncl.value), np.zeros(nr.value)
        cl = np.zeros(nr.value, np.int32)
        g = np.zeros((nr.value, nrx.value, ntx.value), complex)
        lib.isac_cdl_get_rays(self.handle, C.byref(ncl), C.byref(nr), C.byref(nrx), C.byref(ntx), self._lib.ptr(tau),
                              self._lib.ptr(nu), self._lib.ptr(cl), self._lib.ptr(g))
        return {"tau": tau, "nu": nu, "cluster": cl, "g": g}

    def generate(self, K, scs_hz, sym_times, t0=0.0, out=None):
        """H as a torch CUDA complex64 tensor laid out [nTx][nRx][L][K] (== MATLAB [K x L x nRx x nTx])."""
        import numpy as np
        import torch
        st = np.ascontiguousarray(sym_times, dtype=np.float64)
        if out is None:
            out = torch.empty((self.nTx, self.nRx, st.size, K), dtype=torch.complex64, device=f"cuda:{self.ctx.device}")
        self.ctx.use_torch_stream()
        self._lib.check(self.ctx.lib.isac_cdl_generate_dev(self.handle, int(K), float(scs_hz), int(st.size), self._lib.ptr(st),
                                                           float(t0), self._lib.ptr(out)), self.ctx.handle)
        return out

    @staticmethod
    def generateBatch(channels, K, scs_hz, sym_times, t0, out=None):
        """All ``channels`` (same profile / array sizes) in two launches: H stacked [n][nTx][nRx][L][K]."""
        import numpy as np
        import torch
        c0 = channels[0]
        C, lib_ = c0._C, c0._lib
        st = np.ascontiguousarray(sym_times, dtype=np.float64)
        t0 = np.ascontiguousarray(np.broadcast_to(np.asarray(t0, dtype=np.float64), (len(channels),)))
        if out is None:
            out = torch.empty((len(channels), c0.nTx, c0.nRx, st.size, K), dtype=torch.complex64, device=f"cuda:{c0.ctx.device}")
        hs = (C.c_void_p * len(channels))(*[ch.handle for ch in channels])
        c0.ctx.use_torch_stream()
        lib_.check(c0.ctx.lib.isac_cdl_generate_batch_dev(hs, len(channels), int(K), float(scs_hz), int(st.size), lib_.ptr(st),
                                                          lib_.ptr(t0), lib_.ptr(out)), c0.ctx.handle)
        return out

    def close(self):
        if self.handle:
            self.ctx.lib.isac_cdl_destroy(self.handle)
            self.handle = None
