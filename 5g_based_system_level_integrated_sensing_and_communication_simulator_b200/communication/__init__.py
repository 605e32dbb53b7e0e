"""``+communication`` package mirror (hot-path functions only): phyLayer.*, pmiType1SinglePanelCodebook,
the CSI-RS / SRS / CQI-table setup helpers that feed them, and channelModels."""
from __future__ import annotations

import numpy as np

from . import channelModels, phyLayer  # noqa: F401
from .phyLayer import _csi_struct, _validate_report_config
from .. import _lib


def pmiType1SinglePanelCodebook(reportConfig, nLayers):
    """``W = communication.pmiType1SinglePanelCodebook(reportConfig,nLayers)``
    (reference +communication/pmiType1SinglePanelCodebook.m:1) — the gNB-side copy, including its two
    deviations from the UE-side copy.  ``reportConfig`` needs PanelDimensions,
    OverSamplingFactors (or a valid panel), CodebookMode, CodebookSubsetRestriction, i2Restriction.
    Returns W[P, nLayers, i2, i11, i12, i13] complex128."""
    return phyLayer._codebook(reportConfig, nLayers, variant=1)


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
