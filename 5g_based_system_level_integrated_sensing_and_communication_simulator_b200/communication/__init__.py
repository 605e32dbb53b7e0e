"""``+communication`` package mirror (hot-path functions only): phyLayer.*, pmiType1SinglePanelCodebook,
the CSI-RS / SRS / CQI-table setup helpers that feed them, and channelModels."""
from __future__ import annotations

import numpy as np

from . import channelModels, pathlossModels, phyLayer  # noqa: F401
from .phyLayer import _csi_struct, _validate_report_config
from .. import _lib


def pmiType1SinglePanelCodebook(reportConfig, nLayers):
    """``W = communication.pmiType1SinglePanelCodebook(reportConfig,nLayers)``
    (reference +communication/pmiType1SinglePanelCodebook.m:1) — the gNB-side copy, including its two
    deviations from the UE-side copy.  ``reportConfig`` needs PanelDimensions,
    OverSamplingFactors (or a valid panel), CodebookMode, CodebookSubsetRestriction, i2Restriction.
    Returns W[P, nLayers, i2, i11, i12, i13] complex128."""
    return phyLayer._codebook(reportConfig, nLayers, variant=1)


_MP_PANELS = phyLayer._MP_PANELS   # TS 38.214 Table 5.2.2.2.2-1: (Ng, N1, N2) -> (O1, O2)   (dlPMISelect.m:629-644)


def pmiType1MultiPanelCodebook(reportConfig, nLayers, from_table=False):
    """``Wmp = getPMIType1MultiPanelCodebook(reportConfig,nLayers)`` (reference +communication/+phyLayer/dlPMISelect.m:1351-1772,
    TS 38.214 Tables 5.2.2.2.2-1..-6).  ``reportConfig``: PanelDimensions = (Ng, N1, N2), CodebookMode (2 only with Ng = 2),
    optional OverSamplingFactors and CodebookSubsetRestriction (N1*O1*N2*O2 bits).
    Returns Wmp[P, nLayers, i20, i21, i22, i11, i12, i13, i141, i142, i143] complex128 (host code, isac_type1mp_codebook).
    ``from_table`` materialises the array from the beam / co-phasing table the SINR kernels read instead (consistency check).
    Selection over it: ``communication.phyLayer.dlPMISelect`` with a three-element PanelDimensions."""
    import ctypes as C
    Ng, N1, N2 = (int(x) for x in reportConfig["PanelDimensions"])
    if (Ng, N1, N2) not in _MP_PANELS:
        raise ValueError("nr5g:dlPMISelect:InvalidPanelDimensions")
    O1, O2 = reportConfig.get("OverSamplingFactors") or _MP_PANELS[(Ng, N1, N2)]
    csr = reportConfig.get("CodebookSubsetRestriction")
    csr = None if csr is None else np.ascontiguousarray(np.asarray(csr).ravel() != 0, dtype=np.uint8)
    if csr is not None and csr.size != N1 * O1 * N2 * O2:
        raise ValueError("nr5g:dlPMISelect:InvalidCodebookSubsetRestriction")
    cfg = _lib.CsiConfig(nPorts=2 * Ng * N1 * N2, N1=N1, N2=N2, O1=int(O1), O2=int(O2),
                         codebookMode=int(reportConfig.get("CodebookMode", 1)),
                         subsetRestriction=csr.ctypes.data if csr is not None else None)
    lib = _lib.load()
    fn = lib.isac_type1mp_codebook_from_table if from_table else lib.isac_type1mp_codebook
    dims = (C.c_int32 * 9)()
    st = fn(C.byref(cfg), Ng, int(nLayers), dims, None)
    if st:
        raise _lib.IsacError(st, "type-1 multi-panel codebook: invalid configuration")
    W = np.zeros((2 * Ng * N1 * N2, int(nLayers)) + tuple(int(x) for x in dims), dtype=np.complex128, order="F")
    st = fn(C.byref(cfg), Ng, int(nLayers), dims, _lib.ptr(W))
    if st:
        raise _lib.IsacError(st, "type-1 multi-panel codebook")
    return W


def csirsPanelDimensions(antennaPorts, choice=0):
    """communication.csirsPanelDimensions (csirsPanelDimensions.m:4-18).  The reference picks one of the
    valid panels at random (``randperm``); here the pick is the explicit ``choice`` index."""
    table = {4: [(2, 1)], 8: [(2, 2), (4, 1)], 12: [(3, 2), (6, 1)], 16: [(4, 2), (8, 1)],
             24: [(4, 3), (6, 2), (12, 1)], 32: [(4, 4), (8, 2), (16, 1)]}
    return table[antennaPorts][choice]


def subbandSize(prb, choice=0):
    """communication.subbandSize (subbandSize.m:5-14); explicit ``choice`` instead of ``randperm``."""
    if 24 <= prb <= 72:
        opts = (4, 8)
    elif 73 <= prb <= 144:
        opts = (8, 16)
    elif 145 <= prb <= 275:
        opts = (16, 32)
    else:
        raise ValueError("NumRBs is out of limit")
    return opts[choice]


def setupCSIRS(numRBs, panel_choice=0, subband_choice=0):
    """communication.setupCSIRS (setupCSIRS.m:5-24): row 5 (4 ports), period [5 2], subband CQI/PMI, mode 1."""
    csirs = {"NumCSIRSPorts": 4, "RowNumber": 5, "NumRB": numRBs, "RBOffset": 0, "SubcarrierLocations": 1,
             "SymbolLocations": 0, "Density": "one", "CSIRSPeriod": (5, 2), "CSIRSType": "nzp", "CDMType": "FD-CDM2"}
    rep = {"PanelDimensions": csirsPanelDimensions(4, panel_choice), "CQIMode": "Subband", "PMIMode": "Subband",
           "SubbandSize": subbandSize(numRBs, subband_choice), "CodebookMode": 1}
    return [csirs], [rep]


def setupSINRtoCQIMappingTable():
    """communication.setupSINRtoCQIMappingTable (setupSINRtoCQIMappingTable.m:7-11)."""
    dl = np.array([-3.46, 1.54, 6.54, 11.05, 13.54, 16.04, 17.54, 20.04, 22.04, 24.43, 26.93, 27.43, 29.43, 32.43, 35.43])
    ul = np.array([-5.46, -0.46, 4.54, 9.05, 11.54, 14.04, 15.54, 18.04, 20.04, 22.43, 24.93, 25.43, 27.43, 30.43, 33.43])
    return {"downlinkSINR90pc": dl, "uplinkSINR90pc": ul}
