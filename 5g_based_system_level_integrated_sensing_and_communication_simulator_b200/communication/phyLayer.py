"""``+communication/+phyLayer`` mirror (hot-path functions): dlPMISelect, riSelect, cqiSelect, pmiSelect,
precodedSINR, sinrPerSubband, prgPrecode, maxPUSCHPrecodingMatrixIndicator, and nrChannelEstimate (the toolbox call
right before them, uePhy.m:897 / gNBPhy.m:1030).

MATLAB configuration objects are plain dicts with the same field names:
  carrier      : NSizeGrid, NStartGrid (0), SymbolsPerSlot (14)
  csirs        : NumCSIRSPorts, NumRB, RBOffset (0), SubcarrierLocations (k0), SymbolLocations (l0), Density ('one'|'dot5even'|'dot5odd')
  reportConfig : NSizeBWP, NStartBWP, PanelDimensions, CodebookMode, PMIMode, CQIMode, SubbandSize,
                 CodebookSubsetRestriction, i2Restriction, RIRestriction
Channel matrices: NumPy [K x L x nRx x P] (converted to complex64) or torch CUDA complex64 laid out [P][nRx][L][K].
All arithmetic runs in csrc/comm.cu; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .. import _lib

# TS 38.214 Table 5.2.2.2.2-1: (Ng, N1, N2) -> (O1, O2)
_MP_PANELS = {(2, 2, 1): (4, 1), (2, 4, 1): (4, 1), (4, 2, 1): (4, 1), (2, 2, 2): (4, 4), (2, 8, 1): (4, 1), (4, 4, 1): (4, 1),
              (2, 4, 2): (4, 4), (4, 2, 2): (4, 4)}
_PANELS = {(2, 1): (4, 1), (2, 2): (4, 4), (4, 1): (4, 1), (3, 2): (4, 4), (6, 1): (4, 1), (4, 2): (4, 4), (8, 1): (4, 1),
           (4, 3): (4, 4), (6, 2): (4, 4), (12, 1): (4, 1), (4, 4): (4, 4), (8, 2): (4, 4), (16, 1): (4, 1)}


def _validate_report_config(carrier, n_ports, reportConfig):
    """validateInputs of dlPMISelect.m:511-795 (Type1SinglePanel); raises ValueError with the reference's ids."""
    rc = dict(reportConfig)
    multi = rc.get("CodebookType", "Type1SinglePanel") == "Type1MultiPanel"
    if rc.get("CodebookType", "Type1SinglePanel") not in ("Type1SinglePanel", "Type1MultiPanel"):
        raise ValueError("nr5g:dlPMISelect:InvalidCodebookType")
    nsize = rc.get("NSizeBWP") or carrier["NSizeGrid"]
    nstart = rc.get("NStartBWP")
    nstart = carrier.get("NStartGrid", 0) if nstart is None else nstart
    if nstart < carrier.get("NStartGrid", 0):
        raise ValueError("nr5g:dlPMISelect:InvalidNStartBWP")
    if nsize + nstart > carrier.get("NStartGrid", 0) + carrier["NSizeGrid"]:
        raise ValueError("nr5g:dlPMISelect:InvalidBWPLimits")
    mode = int(rc.get("CodebookMode", 1))
    N1 = N2 = O1 = O2 = 1
    Ng = 0
    if multi:                                           # dlPMISelect.m:629-644 (TS 38.214 Table 5.2.2.2.2-1)
        if "PanelDimensions" not in rc or len(rc["PanelDimensions"]) != 3:
            raise ValueError("nr5g:dlPMISelect:PanelDimensionsMissing")
        Ng, N1, N2 = (int(x) for x in rc["PanelDimensions"])
        if (Ng, N1, N2) not in _MP_PANELS or 2 * Ng * N1 * N2 != n_ports:
            raise ValueError("nr5g:dlPMISelect:InvalidPanelDimensions")
        if mode == 2 and Ng != 2:
            raise ValueError("nr5g:dlPMISelect:InvalidCodebookMode")
        O1, O2 = _MP_PANELS[(Ng, N1, N2)]
    elif n_ports > 2:
        if "PanelDimensions" not in rc:
            raise ValueError("nr5g:dlPMISelect:PanelDimensionsMissing")
        N1, N2 = (int(x) for x in rc["PanelDimensions"])
        if 2 * N1 * N2 != n_ports:
            raise ValueError("nr5g:dlPMISelect:InvalidPanelDimensions")
        if (N1, N2) not in _PANELS:
            raise ValueError("nr5g:dlPMISelect:InvalidPanelConfiguration")
        O1, O2 = _PANELS[(N1, N2)]
    pmi_mode = rc.get("PMIMode", "Wideband")
    cqi_mode = rc.get("CQIMode", "Wideband")
    nsb = 0
    if (pmi_mode.lower() == "subband" or cqi_mode.lower() == "subband") and nsize >= 24:
        if "SubbandSize" not in rc or rc["SubbandSize"] is None:
            raise ValueError("nr5g:dlPMISelect:SubbandSizeMissing")
        valid = [v for (lo, hi), v in {(24, 72): (4, 8), (73, 144): (8, 16), (145, 275): (16, 32)}.items() if lo <= nsize <= hi][0]
        if rc["SubbandSize"] not in valid:
            raise ValueError("nr5g:hDLPMISelect:InvalidSubbandSize")
        nsb = int(rc["SubbandSize"])
    ncsr = N1 * O1 * N2 * O2 if n_ports > 2 else (6 if n_ports == 2 else 1)
    csr = np.ones(ncsr, dtype=np.uint8) if rc.get("CodebookSubsetRestriction") is None else \
        np.ascontiguousarray(rc["CodebookSubsetRestriction"], dtype=np.uint8)
    if csr.size != ncsr:
        raise ValueError("CodebookSubsetRestriction has the wrong length")
    i2r = np.ones(16, dtype=np.uint8) if rc.get("i2Restriction") is None else np.ascontiguousarray(rc["i2Restriction"], dtype=np.uint8)
    rir = np.ones(8, dtype=np.uint8) if rc.get("RIRestriction") is None else np.ascontiguousarray(rc["RIRestriction"], dtype=np.uint8)
    return dict(NSizeBWP=int(nsize), NStartBWP=int(nstart), CodebookMode=mode, N1=N1, N2=N2, O1=O1, O2=O2,
                PMIMode=pmi_mode, CQIMode=cqi_mode, SubbandSize=nsb, csr=csr, i2r=i2r, rir=rir, Ng=Ng)


def _csirs_res(carrier, csirs, v):
    """CSI-RS REs kept by validateInputs (dlPMISelect.m:797-833): port 1, lowest RE of each CDM group, first symbol;
    one RE per occupied PRB (nrCSIRSIndices is toolbox code: restated for density one / dot5)."""
    n_rb = int(csirs.get("NumRB", carrier["NSizeGrid"]))
    rb0 = int(csirs.get("RBOffset", 0))
    k0 = int(np.atleast_1d(csirs.get("SubcarrierLocations", 0))[0])
    l0 = int(np.atleast_1d(csirs.get("SymbolLocations", 0))[0])
    dens = str(csirs.get("Density", "one")).lower()
    prbs = np.arange(rb0, min(rb0 + n_rb, carrier["NSizeGrid"]))
    if dens == "dot5even":
        prbs = prbs[prbs % 2 == 0]
    elif dens == "dot5odd":
        prbs = prbs[prbs % 2 == 1]
    k = 12 * prbs + k0 + 1
    bwp0 = v["NStartBWP"] - carrier.get("NStartGrid", 0)
    keep = (k >= bwp0 * 12 + 1) & (k <= (bwp0 + v["NSizeBWP"]) * 12)          # dlPMISelect.m:352-353
    k = k[keep] - bwp0 * 12                                                   # :356
    return np.ascontiguousarray(k, dtype=np.int32), np.full(k.size, l0 + 1, dtype=np.int32)


class _Csi:
    """isac_csi_config plus the arrays it points to."""

    def __init__(self, carrier, n_ports, n_rx, v, re_k, re_l):
        self.v, self.re_k, self.re_l = v, re_k, re_l
        self.cfg = _lib.CsiConfig(
            nPorts=n_ports, N1=v["N1"], N2=v["N2"], O1=v["O1"], O2=v["O2"], codebookMode=v["CodebookMode"],
            nSizeBWP=v["NSizeBWP"], nStartBWP=v["NStartBWP"], subbandSize=v["SubbandSize"],
            pmiSubband=int(v["PMIMode"].lower() == "subband"), cqiSubband=int(v["CQIMode"].lower() == "subband"),
            K=int(carrier["NSizeGrid"]) * 12, L=int(carrier.get("SymbolsPerSlot", 14)), nRx=int(n_rx),
            subsetRestriction=v["csr"].ctypes.data, i2Restriction=v["i2r"].ctypes.data,
            riRestriction=(C.c_uint8 * 8)(*[int(x) for x in v["rir"]]),
            nRE=int(re_k.size), reK=re_k.ctypes.data, reL=re_l.ctypes.data, nPanels=int(v.get("Ng", 0)))

    def key(self):
        v = self.v
        return (self.cfg.nPorts, v.get("Ng", 0), v["N1"], v["N2"], v["CodebookMode"], v["NSizeBWP"], v["NStartBWP"], v["SubbandSize"],
                v["PMIMode"], v["CQIMode"], self.cfg.K, self.cfg.L, self.cfg.nRx, v["csr"].tobytes(), v["i2r"].tobytes(),
                v["rir"].tobytes(), self.re_k.tobytes(), self.re_l.tobytes())


def _csi_struct(carrier, csirs, reportConfig, n_rx):
    n_ports = int(np.atleast_1d(csirs["NumCSIRSPorts"])[0])
    v = _validate_report_config(carrier, n_ports, reportConfig)
    re_k, re_l = _csirs_res(carrier, csirs, v)
    return _Csi(carrier, n_ports, n_rx, v, re_k, re_l)


def _codebook(reportConfig, nLayers, variant):
    n_ports = 2 * int(np.prod(reportConfig["PanelDimensions"])) if "PanelDimensions" in reportConfig else int(reportConfig["NumCSIRSPorts"])
    if "PanelDimensions" in reportConfig and tuple(reportConfig["PanelDimensions"]) == (1, 1):
        n_ports = 2
    carrier = {"NSizeGrid": reportConfig.get("NSizeBWP") or 24, "NStartGrid": reportConfig.get("NStartBWP") or 0}
    rc = dict(reportConfig)
    rc.setdefault("NSizeBWP", carrier["NSizeGrid"])
    rc.setdefault("NStartBWP", carrier["NStartGrid"])
    rc["PMIMode"], rc["CQIMode"] = "Wideband", "Wideband"
    v = _validate_report_config(carrier, n_ports, rc)
    cs = _Csi(carrier, n_ports, 1, v, np.zeros(0, np.int32), np.zeros(0, np.int32))
    lib = _lib.load()
    dims = (C.c_int32 * 4)()
    st = lib.isac_type1sp_codebook(C.byref(cs.cfg), int(nLayers), int(variant), dims, None)
    if st:
        raise _lib.IsacError(st, "type-1 single-panel codebook: invalid configuration")
    d = [int(x) for x in dims]
    W = np.zeros((n_ports, int(nLayers)) + tuple(d), dtype=np.complex128, order="F")
    st = lib.isac_type1sp_codebook(C.byref(cs.cfg), int(nLayers), int(variant), dims, _lib.ptr(W))
    if st:
        raise _lib.IsacError(st, "type-1 single-panel codebook")
    return W


def _h_to_dev(H):
    """-> (torch CUDA complex64 tensor [batch][P][R][L][K], K, L, R, P, batch)."""
    import torch
    if hasattr(H, "is_cuda"):
        t = H if H.dim() == 5 else H.unsqueeze(0)
        B, P, R, L, K = t.shape
        return t.contiguous(), K, L, R, P, B
    a = np.asarray(H)
    if a.ndim == 4:
        a = a[..., None]
    K, L, R, P, B = a.shape
    dev = _lib.get_context(None).device        # host arrays go to the device of the default context (LOCAL_RANK)
    t = torch.from_numpy(np.ascontiguousarray(a.astype(np.complex64).transpose(4, 3, 2, 1, 0))).to(f"cuda:{dev}")
    return t, K, L, R, P, B


_pmi_plans, _csi_plans = {}, {}
_direct_kernel = False


def setSinrKernel(direct):
    """Choose the SINR kernel of the DL selection functions: False (default) Gram-pair form, True direct H*W form
    (isac_pmi_plan_set_kernel).  Both give the same results to rounding; exposed for the parity tests."""
    global _direct_kernel
    _direct_kernel = bool(direct)
    lib = _lib.get_context(None).lib
    for ent in _pmi_plans.values():
        lib.isac_pmi_plan_set_kernel(ent[0], int(_direct_kernel))
    for ent in _csi_plans.values():
        lib.isac_csi_plan_set_kernel(ent[0], int(_direct_kernel))


def _pmi_plan(cs, nLayers, batch, device=None):
    """Cached isac_pmi_plan of (report configuration, rank) on `device` (the device H lives on); a plan that is too small
    for the batch is destroyed before its replacement takes the cache slot."""
    ctx = _lib.get_context(device)
    key = cs.key() + (nLayers, ctx.device)
    ent = _pmi_plans.get(key)
    if ent is None or ent[1] < batch:
        if ent is not None:
            ctx.lib.isac_pmi_plan_destroy(ent[0])
        h = C.c_void_p()
        _lib.check(ctx.lib.isac_pmi_plan_create(ctx.handle, C.byref(cs.cfg), int(nLayers), int(batch), C.byref(h)), ctx.handle)
        ctx.lib.isac_pmi_plan_set_kernel(h, int(_direct_kernel))
        ent = (h, batch, cs)
        _pmi_plans[key] = ent
    return ctx, ent[0]


def _mp_unflatten(ctx, plan, cs, ranks, i1, i2):
    """Type1MultiPanel report: the plan returns i1[2] / i2 as linear indices into the flattened index sets of the rank that was
    reported (isac_csi_plan_mp_dims); bring PMISet into the reference's form {i1 [6 x B], i2 [3 x nSB x B]}
    (dlPMISelect.m:456-457, :489).  ``ranks``: the rank of every batch entry (NaN -> all-NaN PMISet)."""
    B, nSB = i1.shape[1], i2.shape[0]
    i1o = np.full((6, B), np.nan, order="F")
    i2o = np.full((3, nSB, B), np.nan, order="F")
    cache = {}
    for b in range(B):
        if np.isnan(ranks[b]) or np.any(np.isnan(i1[:, b])):
            continue
        nu = int(ranks[b])
        if nu not in cache:
            mp = (C.c_int32 * 7)()
            _lib.check(ctx.lib.isac_csi_plan_mp_dims(plan, nu, mp), ctx.handle)
            cache[nu] = [int(x) for x in mp]
        mp = cache[nu]
        i1o[:, b] = [i1[0, b], i1[1, b]] + [x + 1 for x in np.unravel_index(int(i1[2, b]) - 1, tuple(mp[3:7]), order="F")]
        for sb in range(nSB):
            if not np.isnan(i2[sb, b]):
                i2o[:, sb, b] = [x + 1 for x in np.unravel_index(int(i2[sb, b]) - 1, tuple(mp[0:3]), order="F")]
    return i1o, i2o


def _csi_plan(cs, batch, device=None):
    ctx = _lib.get_context(device)
    key = cs.key() + (ctx.device,)
    ent = _csi_plans.get(key)
    if ent is None or ent[1] < batch:
        if ent is not None:
            ctx.lib.isac_csi_plan_destroy(ent[0])
        h = C.c_void_p()
        _lib.check(ctx.lib.isac_csi_plan_create(ctx.handle, C.byref(cs.cfg), int(batch), C.byref(h)), ctx.handle)
        ctx.lib.isac_csi_plan_set_kernel(h, int(_direct_kernel))
        ent = (h, batch, cs)
        _csi_plans[key] = ent
    return ctx, ent[0]


def _nvar(nVar, B):
    nv = np.ascontiguousarray(np.broadcast_to(np.asarray(nVar, dtype=np.float64), (B,)))
    if np.any(~np.isfinite(nv)) or np.any(nv < 0):
        raise ValueError("NVAR must be a real, nonnegative, finite scalar")
    return nv


def dlPMISelect(carrier, csirs, reportConfig, nLayers, H, nVar=1e-10, full_grid=False):
    """``[PMISet,info] = communication.phyLayer.dlPMISelect(carrier,csirs,reportConfig,nLayers,H,nVar)``
    (reference +communication/+phyLayer/dlPMISelect.m:1; Type1SinglePanel).

    Returns (PMISet, info): PMISet = {'i1': [i11 i12 i13], 'i2': [per subband]} (1-based, NaN = not reported);
    info = {'SINRPerRE', 'SINRPerSubband', 'W', 'reK', 'reL'}.  info['SINRPerRE'] is
    [nRE x nLayers x i2 x i11 x i12 x i13] at the CSI-RS REs (reK/reL, sorted by subcarrier); with
    ``full_grid=True`` it is scattered into the reference's K x L x ... NaN-filled array (can be GBs).
    A batch of UEs may be passed as H[..., batch] with nVar[batch]: PMISet/info then gain a trailing axis."""
    Hd, K, L, R, P, B = _h_to_dev(H)
    cs = _csi_struct(carrier, csirs, reportConfig, R)
    if P != cs.cfg.nPorts or K != cs.cfg.K:
        raise ValueError("H must be K-by-L-by-nRxAnts-by-NumCSIRSPorts")
    if nLayers > min(R, P):
        raise ValueError("nr5g:hDLPMISelect:InvalidNumLayers")
    nv = _nvar(nVar, B)
    ctx, plan = _pmi_plan(cs, int(nLayers), B, Hd.device.index)
    lib = ctx.lib
    dims = (C.c_int32 * 4)()
    nSB, nC, nRE = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(lib.isac_pmi_plan_info(plan, dims, C.byref(nSB), C.byref(nC), C.byref(nRE), None, None), ctx.handle)
    d = tuple(int(x) for x in dims)
    reK, reL = np.zeros(nRE.value, np.int32), np.zeros(nRE.value, np.int32)
    _lib.check(lib.isac_pmi_plan_info(plan, dims, C.byref(nSB), C.byref(nC), C.byref(nRE), _lib.ptr(reK), _lib.ptr(reL)), ctx.handle)
    ctx.use_torch_stream()
    _lib.check(lib.isac_dl_pmi_select_dev(plan, _lib.ptr(Hd), _lib.ptr(nv), B), ctx.handle)
    i1 = np.zeros((3, B), order="F")
    i2 = np.zeros((nSB.value, B), order="F")
    _lib.check(lib.isac_dl_pmi_collect(plan, B, _lib.ptr(i1), _lib.ptr(i2), None), ctx.handle)
    nCand = int(np.prod(d))
    Ng = int(cs.v.get("Ng", 0))
    if Ng >= 2:
        # Type1MultiPanel (dlPMISelect.m:1351-1772): the plan keeps the 9-D index set [i20 i21 i22 | i11 i12 i13 i141 i142 i143]
        # flattened in MATLAB linear order (i2 = i20,i21,i22; i13' = i13,i141,i142,i143); un-flatten it here
        if nLayers > 4:
            raise ValueError("nr5g:hDLPMISelect:InvalidNumLayers")
        from . import pmiType1MultiPanelCodebook
        mp = (C.c_int32 * 7)()
        _lib.check(lib.isac_pmi_plan_mp_dims(plan, mp), ctx.handle)
        mp = [int(x) for x in mp]
        full = (mp[0], mp[1], mp[2], d[1], d[2], mp[3], mp[4], mp[5], mp[6])
        W = pmiType1MultiPanelCodebook({"PanelDimensions": (Ng, cs.v["N1"], cs.v["N2"]), "CodebookMode": cs.v["CodebookMode"],
                                        "CodebookSubsetRestriction": cs.v["csr"]}, nLayers)
        S = np.zeros((nRE.value, nLayers, nCand, B), order="F")
        Sb = np.zeros((nSB.value, nLayers, nCand, B), order="F")
        _lib.check(lib.isac_dl_pmi_get_info(plan, B, _lib.ptr(S), _lib.ptr(Sb)), ctx.handle)
        S = S.reshape((nRE.value, nLayers) + full + (B,), order="F")
        Sb = Sb.reshape((nSB.value, nLayers) + full + (B,), order="F")
        i1o = np.full((6, B), np.nan, order="F")
        i2o = np.full((3, nSB.value, B), np.nan, order="F")
        for b in range(B):
            if not np.any(np.isnan(i1[:, b])):
                i13f = int(i1[2, b]) - 1
                i1o[:, b] = [i1[0, b], i1[1, b]] + [x + 1 for x in np.unravel_index(i13f, tuple(mp[3:7]), order="F")]
            for sb in range(nSB.value):
                if not np.isnan(i2[sb, b]):
                    i2o[:, sb, b] = [x + 1 for x in np.unravel_index(int(i2[sb, b]) - 1, tuple(mp[0:3]), order="F")]
        if full_grid:
            g = np.full((cs.cfg.nSizeBWP * 12, L, nLayers) + full + (B,), np.nan)
            g[reK - 1, reL - 1] = S
            S = g
        if B == 1:
            i1o, i2o, S, Sb = i1o[:, 0], i2o[:, :, 0], S[..., 0], Sb[..., 0]
        return {"i1": i1o, "i2": i2o}, {"SINRPerRE": S, "SINRPerSubband": Sb, "W": W, "reK": reK, "reL": reL}
    W = _codebook({**reportConfig, "NumCSIRSPorts": P}, nLayers, 0) if P > 1 else np.ones((1, 1, 1, 1, 1, 1), complex)
    if nRE.value and np.any(W):
        S = np.zeros((nRE.value, nLayers, nCand, B), order="F")
        Sb = np.zeros((nSB.value, nLayers, nCand, B), order="F")
        _lib.check(lib.isac_dl_pmi_get_info(plan, B, _lib.ptr(S), _lib.ptr(Sb)), ctx.handle)
        S = S.reshape((nRE.value, nLayers) + d + (B,), order="F")
        Sb = Sb.reshape((nSB.value, nLayers) + d + (B,), order="F")
    else:
        S = np.full((nRE.value, nLayers) + d + (B,), np.nan)
        Sb = np.full((nSB.value, nLayers) + d + (B,), np.nan)
        i1[:] = np.nan
        i2[:] = np.nan
    if full_grid:
        full = np.full((cs.cfg.nSizeBWP * 12, L, nLayers) + d + (B,), np.nan)
        full[reK - 1, reL - 1] = S
        S = full
    if B == 1:
        i1, i2, S, Sb = i1[:, 0], i2[:, 0], S[..., 0], Sb[..., 0]
    return {"i1": i1, "i2": i2}, {"SINRPerRE": S, "SINRPerSubband": Sb, "W": W, "reK": reK, "reL": reL}


def riSelect(carrier, csirs, reportConfig, H, nVar=1e-10):
    """``[RI,PMISet] = communication.phyLayer.riSelect(carrier,csirs,reportConfig,H,nVar)`` (riSelect.m:1)."""
    Hd, K, L, R, P, B = _h_to_dev(H)
    cs = _csi_struct(carrier, csirs, reportConfig, R)
    nv = _nvar(nVar, B)
    ctx, plan = _csi_plan(cs, B, Hd.device.index)
    nSB = len(_subband_sizes(cs.v["PMIMode"], cs.v))
    RI = np.zeros(B)
    i1 = np.zeros((3, B), order="F")
    i2 = np.zeros((nSB, B), order="F")
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_ri_select_dev(plan, _lib.ptr(Hd), _lib.ptr(nv), B, _lib.ptr(RI), _lib.ptr(i1), _lib.ptr(i2)), ctx.handle)
    if int(cs.v.get("Ng", 0)) >= 2:
        # all-NaN totals: the reference returns the PMISet of the last valid rank (riSelect.m:289-292)
        last = max((r + 1 for r in range(min(R, P, 4)) if cs.v["rir"][r]), default=1)
        i1, i2 = _mp_unflatten(ctx, plan, cs, np.where(np.isnan(RI), float(last), RI), i1, i2)
        if B == 1:
            return float(RI[0]), {"i1": i1[:, 0], "i2": i2[:, :, 0]}
        return RI, {"i1": i1, "i2": i2}
    if B == 1:
        return float(RI[0]), {"i1": i1[:, 0], "i2": i2[:, 0]}
    return RI, {"i1": i1, "i2": i2}


def _subband_sizes(mode, v):
    if mode.lower() == "wideband" or v["NSizeBWP"] < 24 or not v["SubbandSize"]:
        return [v["NSizeBWP"]]
    n = v["SubbandSize"]
    first = n - (v["NStartBWP"] % n)
    last = (v["NStartBWP"] + v["NSizeBWP"]) % n or n
    cnt = (v["NSizeBWP"] - (first + last)) // n + 2
    s = [n] * cnt
    s[0], s[-1] = first, last
    return s


def cqiSelect(carrier, csirs, reportConfig, nLayers, H, nVar, SINRTable):
    """``[CQI,PMISet,CQIInfo,PMIInfo] = communication.phyLayer.cqiSelect(carrier,csirs,reportConfig,nLayers,H,nVar,SINRTable)``
    (cqiSelect.m:1, CSI-RS object syntax, PRGSize unset).  Returns (CQI, PMISet, CQIInfo); CQI is
    [rows x nCodewords] with rows = nSubbands+1 in subband CQI mode (wideband CQI first, then the 2-bit
    differential values) else 1.  PMIInfo is available from ``dlPMISelect``."""
    Hd, K, L, R, P, B = _h_to_dev(H)
    cs = _csi_struct(carrier, csirs, reportConfig, R)
    nv = _nvar(nVar, B)
    ctx, plan = _csi_plan(cs, B, Hd.device.index)
    nSB = len(_subband_sizes(cs.v["PMIMode"], cs.v))
    nC = len(_subband_sizes(cs.v["CQIMode"], cs.v))
    rows_full = nC + 1 if nC > 1 else 1
    rows_max = nC + 1
    table = np.ascontiguousarray(SINRTable, dtype=np.float64)
    cqi = np.zeros((rows_max * 2 * B))
    sb = np.zeros((rows_full * 2 * B))
    i1 = np.zeros((3, B), order="F")
    i2 = np.zeros((nSB, B), order="F")
    rows = C.c_int32()
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_cqi_select_dev(plan, int(nLayers), _lib.ptr(Hd), _lib.ptr(nv), B, _lib.ptr(table), table.size,
                                           _lib.ptr(cqi), C.byref(rows), _lib.ptr(i1), _lib.ptr(i2), _lib.ptr(sb)), ctx.handle)
    ncw = int(math.ceil(nLayers / 4))
    cqi = cqi[: rows.value * 2 * B].reshape((rows.value, 2, B), order="F")[:, :ncw]
    sb = sb.reshape((rows_full, 2, B), order="F")[:, :ncw]
    if int(cs.v.get("Ng", 0)) >= 2:
        i1, i2 = _mp_unflatten(ctx, plan, cs, np.full(B, float(nLayers)), i1, i2)
        if B == 1:
            return cqi[..., 0], {"i1": i1[:, 0], "i2": i2[:, :, 0]}, {"SINRPerSubbandPerCW": sb[..., 0]}
        return cqi, {"i1": i1, "i2": i2}, {"SINRPerSubbandPerCW": sb}
    if B == 1:
        return cqi[..., 0], {"i1": i1[:, 0], "i2": i2[:, 0]}, {"SINRPerSubbandPerCW": sb[..., 0]}
    return cqi, {"i1": i1, "i2": i2}, {"SINRPerSubbandPerCW": sb}


class PendingCsiReport:
    """A CSI report whose kernels are enqueued (csiReportEnqueue); ``finish()`` waits for its results only -- work enqueued
    on the stream in between keeps the GPU busy during the host-side RI / CQI tails."""

    def __init__(self, ctx, plan, Hd, B, nSB, nC, table, rankCap, cs=None):
        self.ctx, self.plan, self.Hd, self.B, self.nSB, self.nC = ctx, plan, Hd, B, nSB, nC    # Hd kept alive until finish()
        self.table, self.rankCap, self.cs = table, int(rankCap), cs

    def finish(self):
        B = self.B
        RI = np.zeros(B)
        i1 = np.zeros((3, B), order="F")
        i2 = np.zeros((self.nSB, B), order="F")
        cqi = np.zeros(((self.nC + 1) * 2 * B))
        rows = C.c_int32()
        _lib.check(self.ctx.lib.isac_csi_report_finish(self.plan, _lib.ptr(self.table), self.table.size, self.rankCap, _lib.ptr(RI),
                                                       _lib.ptr(i1), _lib.ptr(i2), _lib.ptr(cqi), C.byref(rows)), self.ctx.handle)
        cqi = cqi[: rows.value * 2 * B].reshape((rows.value, 2, B), order="F")
        self.Hd = None
        if self.cs is not None and int(self.cs.v.get("Ng", 0)) >= 2:
            i1, i2 = _mp_unflatten(self.ctx, self.plan, self.cs, RI, i1, i2)
            if B == 1:
                return float(RI[0]), {"i1": i1[:, 0], "i2": i2[:, :, 0]}, cqi[..., 0]
            return RI, {"i1": i1, "i2": i2}, cqi
        if B == 1:
            return float(RI[0]), {"i1": i1[:, 0], "i2": i2[:, 0]}, cqi[..., 0]
        return RI, {"i1": i1, "i2": i2}, cqi


def csiReportEnqueue(carrier, csirs, reportConfig, H, nVar, SINRTable, rankCap=4):
    """First half of csiReport: launches the SINR / selection kernels of every valid rank and the D2H copy of their results
    on the current torch stream and returns without synchronising (isac_csi_report_enqueue_dev)."""
    Hd, K, L, R, P, B = _h_to_dev(H)
    cs = _csi_struct(carrier, csirs, reportConfig, R)
    nv = _nvar(nVar, B)
    ctx, plan = _csi_plan(cs, B, Hd.device.index)
    nSB = len(_subband_sizes(cs.v["PMIMode"], cs.v))
    nC = len(_subband_sizes(cs.v["CQIMode"], cs.v))
    table = np.ascontiguousarray(SINRTable, dtype=np.float64)
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_csi_report_enqueue_dev(plan, _lib.ptr(Hd), _lib.ptr(nv), B), ctx.handle)
    return PendingCsiReport(ctx, plan, Hd, B, nSB, nC, table, rankCap, cs)


def csiReport(carrier, csirs, reportConfig, H, nVar, SINRTable, rankCap=4):
    """Fused UE CSI report of uePhy.phyRxProcessing (uePhy.m:900-907): ``rank = min(riSelect(...), 4)`` followed
    by ``cqiSelect`` at that rank.  Returns (rank, PMISet, CQI)."""
    return csiReportEnqueue(carrier, csirs, reportConfig, H, nVar, SINRTable, rankCap).finish()


def maxPUSCHPrecodingMatrixIndicator(nlayers, nports):
    """communication.phyLayer.maxPUSCHPrecodingMatrixIndicator (maxPUSCHPrecodingMatrixIndicator.m:14-74)."""
    if nports not in (1, 2, 4):
        raise ValueError("nr5g:hMaxPUSCHPrecodingMatrixIndicator:InvalidNPorts")
    if nlayers > nports:
        raise ValueError("nr5g:hMaxPUSCHPrecodingMatrixIndicator:TooManyLayers")
    return {(1, 1): 0, (1, 2): 5, (1, 4): 27, (2, 2): 2, (2, 4): 21, (3, 4): 6, (4, 4): 4}[(nlayers, nports)]


def puschCodebook(nlayers, nports):
    """nrPUSCHCodebook(nlayers,nports,tpmi).' for every TPMI: W[nports x nlayers x nTPMI] (pmiSelect.m:45)."""
    lib = _lib.load()
    n = C.c_int32()
    st = lib.isac_pusch_codebook(int(nlayers), int(nports), C.byref(n), None)
    if st:
        raise _lib.IsacError(st, "nrPUSCHCodebook: invalid layers/ports")
    W = np.zeros((nports, nlayers, n.value), dtype=np.complex128, order="F")
    lib.isac_pusch_codebook(int(nlayers), int(nports), C.byref(n), _lib.ptr(W))
    return W


def pmiSelect(nlayers, hest, noiseest, bandSize):
    """``[pmi,sinr,subbandIndices] = communication.phyLayer.pmiSelect(nlayers,hest,noiseest,bandSize)``
    (pmiSelect.m:28).  hest: [K x nSym x nRx x nPorts].  Returns NaN scalars when no estimate / zero noise (:60-64)."""
    import torch
    if hasattr(hest, "is_cuda"):
        hd = hest.contiguous()
        P, R, Ls, K = hd.shape
    else:
        a = np.asarray(hest)
        K, Ls, R, P = a.shape
        hd = torch.from_numpy(np.ascontiguousarray(a.astype(np.complex64).transpose(3, 2, 1, 0))).cuda()
    ctx = _lib.get_context(None)
    max_sb = int(math.ceil(K / 12 / bandSize)) + 1
    nT = maxPUSCHPrecodingMatrixIndicator(nlayers, P) + 1
    pmi = np.zeros(max_sb)
    sinr = np.zeros(max_sb * nT)
    idx = np.zeros(max_sb * 2, dtype=np.int32)
    nSB, nTo, none = C.c_int32(), C.c_int32(), C.c_int32()
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_ul_pmi_select_dev(ctx.handle, int(nlayers), _lib.ptr(hd), K, Ls, R, P, float(noiseest), int(bandSize),
                                              max_sb, _lib.ptr(pmi), _lib.ptr(sinr), _lib.ptr(idx), C.byref(nSB), C.byref(nTo),
                                              C.byref(none)), ctx.handle)
    if none.value:
        return np.nan, np.nan, np.nan
    n = nSB.value
    return pmi[:n].copy(), sinr[: n * nT].reshape((n, nT), order="F"), idx[: 2 * n].reshape((n, 2), order="F")


class PendingPmiSelect:
    """A batched pmiSelect whose kernels are enqueued (pmiSelectBatchEnqueue); ``finish()`` waits for its results only."""

    def __init__(self, ctx, hd, B, P, K, nlayers, bandSize):
        self.ctx, self.hd, self.B = ctx, hd, B      # hd kept alive until finish()
        self.max_sb = int(math.ceil(K / 12 / bandSize)) + 1
        self.nT = maxPUSCHPrecodingMatrixIndicator(nlayers, P) + 1

    def finish(self):
        B, max_sb, nT = self.B, self.max_sb, self.nT
        pmi = np.zeros(max_sb * B)
        sinr = np.zeros(max_sb * nT * B)
        none = np.zeros(B, dtype=np.int32)
        nSB, nTo = C.c_int32(), C.c_int32()
        _lib.check(self.ctx.lib.isac_ul_pmi_select_batch_finish(self.ctx.handle, max_sb, _lib.ptr(pmi), _lib.ptr(sinr), C.byref(nSB),
                                                                C.byref(nTo), _lib.ptr(none)), self.ctx.handle)
        self.hd = None
        n = nSB.value
        pmi = pmi.reshape((max_sb, B), order="F")[:n].copy()
        sinr = np.stack([sinr.reshape((max_sb * nT, B), order="F")[: n * nT, b].reshape((n, nT), order="F") for b in range(B)], axis=2)
        pmi[:, none != 0] = np.nan
        sinr[:, :, none != 0] = np.nan
        return pmi, sinr, none


def pmiSelectBatchEnqueue(nlayers, hest, noiseest, bandSize):
    """First half of pmiSelectBatch: kernels + asynchronous result copy on the current torch stream, no synchronisation
    (isac_ul_pmi_select_batch_enqueue_dev).  One report may be pending per context."""
    hd = hest.contiguous()
    B, P, R, Ls, K = hd.shape
    ctx = _lib.get_context(None)
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_ul_pmi_select_batch_enqueue_dev(ctx.handle, int(nlayers), _lib.ptr(hd), K, Ls, R, P, float(noiseest),
                                                            int(bandSize), B), ctx.handle)
    return PendingPmiSelect(ctx, hd, B, P, K, nlayers, bandSize)


def pmiSelectBatch(nlayers, hest, noiseest, bandSize):
    """pmiSelect for a batch of estimates resident on the device: hest torch complex64 [batch][nPorts][nRx][nSym][K].
    Returns (pmi [nSB x batch], sinr [nSB x nTPMI x batch], none [batch]); one synchronisation for the batch."""
    return pmiSelectBatchEnqueue(nlayers, hest, noiseest, bandSize).finish()


def precodedSINR(H, sigma, W):
    """``sinr = communication.phyLayer.precodedSINR(H,sigma,W)`` (precodedSINR.m:11-18): LMMSE SINR of the precoded
    channel summed over the layers.  H [nRx x nPorts] (or [nRx x nPorts x nRE] for a batch of REs sharing W),
    W [nPorts x nLayers]; float64 throughout as in the reference."""
    H = np.asarray(H, dtype=np.complex128)
    W = np.ascontiguousarray(np.asarray(W, dtype=np.complex128).T)      # column-major [P x nu]
    single = H.ndim == 2
    if single:
        H = H[:, :, None]
    R, P, B = H.shape
    nu = W.shape[0]
    if W.shape[1] != P:
        raise ValueError("W must be nPorts-by-nLayers")
    Hc = np.ascontiguousarray(H.transpose(2, 1, 0))                     # [B][P][R] == column-major [R x P x B]
    out = np.zeros(B)
    ctx = _lib.get_context(None)
    _lib.check(ctx.lib.isac_precoded_sinr_host(ctx.handle, _lib.ptr(Hc), R, P, float(sigma), _lib.ptr(W), nu, B, _lib.ptr(out)),
               ctx.handle)
    return float(out[0]) if single else out


def sinrPerSubband(sinr, bandSize):
    """``[sinrSubband, subbandIndices] = communication.phyLayer.sinrPerSubband(sinr, bandSize)`` (sinrPerSubband.m:12-35).
    Host helper (a few hundred flops): the fused device version lives inside pmiSelect."""
    sinr = np.asarray(sinr, dtype=np.float64)
    nrb = sinr.shape[0] / 12
    r = nrb / bandSize
    n = int(math.ceil(r))
    out = np.zeros((n, sinr.shape[2]))
    idx = np.zeros((n, 2), dtype=int)
    for s in range(n):
        lo = 12 * bandSize * s + 1
        hi = 12 * bandSize * (s + 1) if s < int(math.floor(r)) else int(12 * bandSize * r)
        idx[s] = (lo, hi)
        blk = sinr[lo - 1: hi]
        with np.errstate(invalid="ignore", divide="ignore"):
            out[s] = blk.sum(axis=(0, 1)) / np.count_nonzero(blk.sum(axis=2))
    return out, idx


def prgPrecode(siz, nstartgrid, portsym, portind, F):
    """``[antsym,antind] = communication.phyLayer.prgPrecode(siz,nstartgrid,portsym,portind,F)`` (prgPrecode.m:53).
    portsym / portind: [NRE x nLayers] (1-based linear indices into a K x L x nLayers grid); F: [nLayers x P x NPRG].
    Returns antsym complex64 [NRE x P], antind int [NRE x P]."""
    import torch
    F = np.asarray(F)
    if F.ndim == 2:
        F = F[:, :, None]
    nu, P, nprg = F.shape
    ps = np.asarray(portsym).reshape(-1, nu, order="F")
    pi = np.asarray(portind).reshape(-1, nu, order="F")
    nre = ps.shape[0]
    ctx = _lib.get_context(None)
    d = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asfortranarray(a.astype(dt)).ravel(order="F"))).cuda()
    ps_d, pi_d, F_d = d(ps, np.complex64), d(pi, np.int32), d(F, np.complex64)
    out_s = torch.empty(nre * P, dtype=torch.complex64, device="cuda")
    out_i = torch.empty(nre * P, dtype=torch.int32, device="cuda")
    ctx.use_torch_stream()
    _lib.check(ctx.lib.isac_prg_precode_dev(ctx.handle, int(siz[0]), int(siz[1]), int(nstartgrid), _lib.ptr(ps_d), _lib.ptr(pi_d),
                                            nre, nu, _lib.ptr(F_d), P, nprg, _lib.ptr(out_s), _lib.ptr(out_i)), ctx.handle)
    return (out_s.cpu().numpy().reshape((nre, P), order="F"), out_i.cpu().numpy().astype(np.int64).reshape((nre, P), order="F"))


# ---- channel estimation ----------------------------------------------------------------------------
class ChannelEstimator:
    """Device plan of ``nrChannelEstimate`` for one reference-signal layout (csrc/chest.cu).

    refInd / refSym follow the toolbox convention (1-based column-major linear indices into the K x L x nPorts grid, any
    array shape).  ``run_dev`` keeps everything on the device: rxGrid [batch][nRx][L][K] -> Hest [batch][P][nRx][L][K]
    (== MATLAB K x L x nRx x P x batch), the layout dlPMISelect / riSelect / cqiSelect / csiReport / pmiSelectBatch consume."""

    def __init__(self, K, L, nRx, nPorts, refInd, refSym, CDMLengths=(1, 1), AveragingWindow=(0, 0), max_batch=1,
                 device=None, ctx=None):
        self.ctx = ctx if ctx is not None else _lib.get_context(device)
        self.K, self.L, self.nRx, self.P, self.max_batch = int(K), int(L), int(nRx), int(nPorts), int(max_batch)
        ind = np.ascontiguousarray(np.asarray(refInd).reshape(-1, order="F"), dtype=np.int32)
        sym = np.ascontiguousarray(np.asarray(refSym).reshape(-1, order="F"), dtype=np.complex64)
        if ind.size != sym.size:
            raise ValueError("refInd and refSym disagree in size")
        h = C.c_void_p()
        _lib.check(self.ctx.lib.isac_chest_plan_create(self.ctx.handle, self.K, self.L, self.nRx, self.P, ind.size, _lib.ptr(ind),
                                                       _lib.ptr(sym), int(CDMLengths[0]), int(CDMLengths[1]),
                                                       int(AveragingWindow[0]), int(AveragingWindow[1]), self.max_batch,
                                                       C.byref(h)), self.ctx.handle)
        self.handle = h

    def run_dev(self, rx_dev, batch=1, H_out=None, sync=True):
        """rx_dev: torch CUDA complex64 [batch][nRx][L][K].  -> (Hest torch [batch][P][nRx][L][K], nVar [batch] or None)."""
        import torch
        if H_out is None:
            H_out = torch.empty((batch, self.P, self.nRx, self.L, self.K), dtype=torch.complex64, device=rx_dev.device)
        nvar = np.zeros(batch) if sync else None
        self.ctx.use_torch_stream()
        _lib.check(self.ctx.lib.isac_channel_estimate_dev(self.handle, _lib.ptr(rx_dev), int(batch), _lib.ptr(H_out),
                                                          _lib.ptr(nvar)), self.ctx.handle)
        return H_out, nvar

    def nvar(self, batch=1):
        out = np.zeros(batch)
        _lib.check(self.ctx.lib.isac_chest_get_nvar(self.handle, int(batch), _lib.ptr(out)), self.ctx.handle)
        return out

    def close(self):
        if self.handle:
            self.ctx.lib.isac_chest_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nrChannelEstimate(rxGrid, refInd, refSym, nPorts=None, CDMLengths=(1, 1), AveragingWindow=(0, 0)):
    """``[Hest, nVar] = nrChannelEstimate(rxGrid, refInd, refSym, 'CDMLengths', cdmLen, 'AveragingWindow', win)`` as the
    reference calls it (uePhy.m:897, gNBPhy.m:1030).  rxGrid: NumPy [K x L x nRx]; returns Hest [K x L x nRx x P] complex64
    and the scalar noise-variance estimate.  nPorts defaults to the port count implied by max(refInd)."""
    import torch
    rx = np.asarray(rxGrid)
    if rx.ndim == 2:
        rx = rx[:, :, None]
    K, L, R = rx.shape
    ind = np.asarray(refInd)
    P = int(nPorts) if nPorts else int((int(ind.max()) - 1) // (K * L) + 1)
    est = ChannelEstimator(K, L, R, P, refInd, refSym, CDMLengths, AveragingWindow, 1)
    rx_d = torch.from_numpy(np.ascontiguousarray(rx.astype(np.complex64).transpose(2, 1, 0))).cuda(est.ctx.device)[None]
    H, nvar = est.run_dev(rx_d, 1)
    out = H[0].permute(3, 2, 1, 0).cpu().numpy()       # [K, L, R, P]
    est.close()
    return out, float(nvar[0])
