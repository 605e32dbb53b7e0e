"""Float64 oracle for the COMM half of the hot path (test infrastructure only; PARITY UNPINNED).

Restates ``+communication/+phyLayer/{dlPMISelect,riSelect,cqiSelect,pmiSelect,precodedSINR,
sinrPerSubband,prgPrecode}.m`` and ``+communication/pmiType1SinglePanelCodebook.m`` of the reference.
Loop nests follow the reference (this is also the "loop-faithful" CPU baseline of BASELINE.md).
Toolbox pieces that are not in the repository (nrCSIRSIndices, nrPUSCHCodebook, nrLayerDemap,
nrExtractResources) are restated from TS 38.211 / 38.214 and tagged PARITY-UNPINNED.

Indices returned to the caller are 1-based like MATLAB's; NaN marks "not reported".
"""
from __future__ import annotations

import math
import warnings

import numpy as np

# TS 38.214 Table 5.2.2.2.1-2 (dlPMISelect.m:625-628)
_PANEL_CONFIGS = {(2, 1): (4, 1), (2, 2): (4, 4), (4, 1): (4, 1), (3, 2): (4, 4), (6, 1): (4, 1), (4, 2): (4, 4),
                  (8, 1): (4, 1), (4, 3): (4, 4), (6, 2): (4, 4), (12, 1): (4, 1), (4, 4): (4, 4), (8, 2): (4, 4),
                  (16, 1): (4, 1)}


def matlab_round4(x):
    """round(x,4,'decimal'): half away from zero (dlPMISelect.m:449,486,492)."""
    x = np.asarray(x, dtype=np.float64)
    return np.sign(x) * np.floor(np.abs(x) * 1e4 + 0.5) / 1e4


def report_config(n_ports, panel=None, n_size_bwp=None, n_start_bwp=0, codebook_mode=1, pmi_mode="Subband",
                  cqi_mode="Subband", subband_size=None, subset_restriction=None, i2_restriction=None,
                  ri_restriction=None):
    """Validated report configuration (dlPMISelect.m:511-851 validateInputs, Type1SinglePanel only)."""
    cfg = {"NSizeBWP": n_size_bwp, "NStartBWP": n_start_bwp, "CodebookType": "Type1SinglePanel",
           "CodebookMode": codebook_mode, "PMIMode": pmi_mode, "CQIMode": cqi_mode, "PRGSize": None,
           "NumCSIRSPorts": n_ports}
    N1 = N2 = O1 = O2 = 1
    if n_ports > 2:
        N1, N2 = int(panel[0]), int(panel[1])
        if 2 * N1 * N2 != n_ports:
            raise ValueError("nr5g:dlPMISelect:InvalidPanelDimensions")
        if (N1, N2) not in _PANEL_CONFIGS:
            raise ValueError("nr5g:dlPMISelect:InvalidPanelConfiguration")
        O1, O2 = _PANEL_CONFIGS[(N1, N2)]
    cfg["PanelDimensions"] = (N1, N2)
    cfg["OverSamplingFactors"] = (O1, O2)
    nsb = None
    if pmi_mode == "Subband" or cqi_mode == "Subband":
        if n_size_bwp >= 24:
            valid = {(24, 72): (4, 8), (73, 144): (8, 16), (145, 275): (16, 32)}
            ok = [v for (lo, hi), v in valid.items() if lo <= n_size_bwp <= hi][0]
            if subband_size not in ok:
                raise ValueError("nr5g:hDLPMISelect:InvalidSubbandSize")
            nsb = subband_size
    cfg["SubbandSize"] = nsb
    if n_ports > 2:
        L = N1 * O1 * N2 * O2
        cfg["CodebookSubsetRestriction"] = np.ones(L, dtype=int) if subset_restriction is None else np.asarray(subset_restriction)
    elif n_ports == 2:
        cfg["CodebookSubsetRestriction"] = np.ones(6, dtype=int) if subset_restriction is None else np.asarray(subset_restriction)
    else:
        cfg["CodebookSubsetRestriction"] = np.ones(1, dtype=int)
    cfg["i2Restriction"] = np.ones(16, dtype=int) if i2_restriction is None else np.asarray(i2_restriction)
    cfg["RIRestriction"] = np.ones(8, dtype=int) if ri_restriction is None else np.asarray(ri_restriction)
    return cfg


def subband_info(mode, n_start_bwp, n_size_bwp, nsbprb):
    """getDownlinkPMISubbandInfo / getSubbandInfo (dlPMISelect.m:1836-1887, cqiSelect.m:1208-1244)."""
    if mode.lower() == "wideband" or n_size_bwp < 24:
        return 1, [n_size_bwp]
    first = nsbprb - (n_start_bwp % nsbprb)
    last = (n_start_bwp + n_size_bwp) % nsbprb or nsbprb
    n = (n_size_bwp - (first + last)) // nsbprb + 2
    sizes = [nsbprb] * n
    sizes[0] = first
    sizes[-1] = last
    return n, sizes


def csirs_first_port_res(n_rb, k0=1, l0=0, density=1.0, rb_offset=0):
    """CSI-RS RE subscripts (1-based k, l) that dlPMISelect keeps (dlPMISelect.m:797-833): port 1, lowest
    RE of each CDM group, first symbol.  PARITY-UNPINNED (nrCSIRSIndices): one RE per occupied PRB at
    subcarrier k0 and symbol l0 for density 1; every other PRB for density 0.5 (setupCSIRS.m: row 5,
    SubcarrierLocations 1, SymbolLocations 0)."""
    step = 1 if density >= 1 else 2
    prbs = np.arange(rb_offset, n_rb, step)
    return 12 * prbs + k0 + 1, np.full(prbs.size, l0 + 1)


# ----------------------------------------------------------------------------------------------
# Type-I single-panel codebook (TS 38.214 Tables 5.2.2.2.1-1 ... -12)
# ----------------------------------------------------------------------------------------------
def _vlm(N1, N2, O1, O2, l, m):
    """getVlm (dlPMISelect.m:1774-1782)."""
    um = np.exp(2j * np.pi * m * np.arange(N2) / (O2 * N2))
    ul = np.exp(2j * np.pi * l * np.arange(N1) / (O1 * N1))
    return (ul[:, None] * um[None, :]).reshape(-1)          # reshape((ul.*um).',[],1): N2 fastest


def _vbarlm(N1, N2, O1, O2, l, m):
    """getVbarlm (dlPMISelect.m:1784-1793)."""
    um = np.exp(2j * np.pi * m * np.arange(N2) / (O2 * N2))
    ul = np.exp(2j * np.pi * l * np.arange(N1 // 2) / (O1 * N1 / 2))
    return (ul[:, None] * um[None, :]).reshape(-1)


def _restricted(csr, bits, n, i2r):
    """isRestricted (dlPMISelect.m:1795-1823)."""
    ridx = np.flatnonzero(np.asarray(csr) == 0)
    lm = any(int(b) in ridx for b in np.atleast_1d(bits))
    i2 = n in np.flatnonzero(np.asarray(i2r) == 0)
    return lm, i2


def type1_single_panel_codebook(cfg, n_layers, variant="ue"):
    """W[P, nLayers, i2, i11, i12, i13] (restricted entries all-zero).

    variant 'ue'  : getPMIType1SinglePanelCodebook (dlPMISelect.m:853-1349)
    variant 'gnb' : communication.pmiType1SinglePanelCodebook (pmiType1SinglePanelCodebook.m:46-554) with its
                    two deviations: rank 3-4 / >=16 ports written without the i13 index (:348,:358) and
                    rank 2 / mode 2 / N2>1 using floor(i2/4) for l', m' (:225,:227)."""
    N1, N2 = cfg["PanelDimensions"]
    O1, O2 = cfg["OverSamplingFactors"]
    mode = cfg["CodebookMode"]
    csr, i2r = cfg["CodebookSubsetRestriction"], cfg["i2Restriction"]
    phi = lambda x: np.exp(1j * np.pi * x / 2)
    P = 2 * N1 * N2
    nu = n_layers
    if P == 2:
        if nu == 1:
            W = np.zeros((2, 1, 4), complex)
            for i, v in enumerate(([1, 1], [1, 1j], [1, -1], [1, -1j])):
                if csr[i]:
                    W[:, 0, i] = np.array(v) / math.sqrt(2)
        else:
            W = np.zeros((2, 2, 2), complex)
            for i, v in enumerate(([[1, 1], [1, -1]], [[1, 1], [1j, -1j]])):
                if csr[4 + i]:
                    W[:, :, i] = np.array(v) / 2
        return W[:, :, :, None, None, None]
    lm_add = [(0, 0), (1, 0), (0, 1), (1, 1)]
    V = lambda l, m: _vlm(N1, N2, O1, O2, l, m)
    if nu == 1:
        if mode == 1:
            W = np.zeros((P, 1, 4, N1 * O1, N2 * O2, 1), complex)
            for i11 in range(N1 * O1):
                for i12 in range(N2 * O2):
                    for i2 in range(4):
                        a, b = _restricted(csr, N2 * O2 * i11 + i12, i2, i2r)
                        if not (a or b):
                            v = V(i11, i12)
                            W[:, 0, i2, i11, i12, 0] = np.concatenate([v, phi(i2) * v]) / math.sqrt(P)
        else:
            n12 = 1 if N2 == 1 else N2 * O2 // 2
            W = np.zeros((P, 1, 16, N1 * O1 // 2, n12, 1), complex)
            for i11 in range(N1 * O1 // 2):
                for i12 in range(n12):
                    for i2 in range(16):
                        f = i2 // 4
                        if N2 == 1:
                            l, m = 2 * i11 + f, 0
                        else:
                            l, m = 2 * i11 + lm_add[f][0], 2 * i12 + lm_add[f][1]
                        n = i2 % 4
                        a, b = _restricted(csr, N2 * O2 * l + m, i2, i2r)
                        if not (a or b):
                            v = V(l, m)
                            W[:, 0, i2, i11, i12, 0] = np.concatenate([v, phi(n) * v]) / math.sqrt(P)
        return W
    if nu == 2:
        if N1 > N2 and N2 > 1:
            k1, k2 = [0, O1, 0, 2 * O1], [0, 0, O2, 0]
        elif N1 == N2:
            k1, k2 = [0, O1, 0, O1], [0, 0, O2, O2]
        elif N1 == 2 and N2 == 1:
            k1, k2 = [0, O1], [0, 0]
        else:
            k1, k2 = [0, O1, 2 * O1, 3 * O1], [0, 0, 0, 0]
        n13 = len(k1)
        if mode == 1:
            W = np.zeros((P, 2, 2, N1 * O1, N2 * O2, n13), complex)
            for i11 in range(N1 * O1):
                for i12 in range(N2 * O2):
                    for i13 in range(n13):
                        for i2 in range(2):
                            a, b = _restricted(csr, N2 * O2 * i11 + i12, i2, i2r)
                            if not (a or b):
                                v, vp = V(i11, i12), V(i11 + k1[i13], i12 + k2[i13])
                                ph = phi(i2)
                                W[:, :, i2, i11, i12, i13] = np.block([[v[:, None], vp[:, None]],
                                                                        [ph * v[:, None], -ph * vp[:, None]]]) / math.sqrt(2 * P)
        else:
            n12 = 1 if N2 == 1 else N2 * O2 // 2
            W = np.zeros((P, 2, 8, N1 * O1 // 2, n12, n13), complex)
            for i11 in range(N1 * O1 // 2):
                for i12 in range(n12):
                    for i13 in range(n13):
                        for i2 in range(8):
                            f = i2 // 2
                            fp = i2 // 4 if variant == "gnb" else f       # pmiType1SinglePanelCodebook.m:225,227
                            if N2 == 1:
                                l, lp, m, mp = 2 * i11 + f, 2 * i11 + f + k1[i13], 0, 0
                            else:
                                l = 2 * i11 + lm_add[f][0]
                                lp = 2 * i11 + k1[i13] + lm_add[fp][0]
                                m = 2 * i12 + lm_add[f][1]
                                mp = 2 * i12 + k2[i13] + lm_add[fp][1]
                            n = i2 % 2
                            a, b = _restricted(csr, N2 * O2 * l + m, i2, i2r)
                            if not (a or b):
                                v, vp = V(l, m), V(lp, mp)
                                ph = phi(n)
                                W[:, :, i2, i11, i12, i13] = np.block([[v[:, None], vp[:, None]],
                                                                        [ph * v[:, None], -ph * vp[:, None]]]) / math.sqrt(2 * P)
        return W
    if nu in (3, 4):
        if P < 16:
            tab = {(2, 1): ([O1], [0]), (4, 1): ([O1, 2 * O1, 3 * O1], [0, 0, 0]),
                   (6, 1): ([O1, 2 * O1, 3 * O1, 4 * O1], [0, 0, 0, 0]), (2, 2): ([O1, 0, O1], [0, O2, O2]),
                   (3, 2): ([O1, 0, O1, 2 * O1], [0, O2, O2, 0])}
            k1, k2 = tab[(N1, N2)]
            n13 = len(k1)
            W = np.zeros((P, nu, 2, N1 * O1, N2 * O2, n13), complex)
            for i11 in range(N1 * O1):
                for i12 in range(N2 * O2):
                    for i13 in range(n13):
                        for i2 in range(2):
                            a, b = _restricted(csr, N2 * O2 * i11 + i12, i2, i2r)
                            if a or b:
                                continue
                            v, vp = V(i11, i12)[:, None], V(i11 + k1[i13], i12 + k2[i13])[:, None]
                            ph = phi(i2)
                            if nu == 3:
                                M = np.block([[v, vp, v], [ph * v, ph * vp, -ph * v]]) / math.sqrt(3 * P)
                            else:
                                M = np.block([[v, vp, v, vp], [ph * v, ph * vp, -ph * v, -ph * vp]]) / math.sqrt(4 * P)
                            W[:, :, i2, i11, i12, i13] = M
            return W
        W = np.zeros((P, nu, 2, N1 * O1 // 2, N2 * O2, 4), complex)
        for i11 in range(N1 * O1 // 2):
            for i12 in range(N2 * O2):
                for i13 in range(4):
                    for i2 in range(2):
                        th = np.exp(1j * np.pi * i13 / 4)
                        l, m = i11, i12
                        ph = phi(i2)
                        bits = [(N2 * O2 * (2 * l - 1) + m) % (N1 * O1 * N2 * O2), N2 * O2 * (2 * l) + m, N2 * O2 * (2 * l + 1) + m]
                        a, b = _restricted(csr, bits, i2, i2r)
                        if a or b:
                            continue
                        vb = _vbarlm(N1, N2, O1, O2, l, m)[:, None]
                        if nu == 3:
                            M = np.block([[vb, vb, vb], [th * vb, -th * vb, th * vb], [ph * vb, ph * vb, -ph * vb],
                                          [ph * th * vb, -ph * th * vb, -ph * th * vb]]) / math.sqrt(3 * P)
                        else:
                            M = np.block([[vb, vb, vb, vb], [th * vb, -th * vb, th * vb, -th * vb],
                                          [ph * vb, ph * vb, -ph * vb, -ph * vb],
                                          [ph * th * vb, -ph * th * vb, -ph * th * vb, ph * th * vb]]) / math.sqrt(4 * P)
                        if variant == "gnb":
                            W[:, :, i2, i11, i12, 0] = M      # pmiType1SinglePanelCodebook.m:348,358 (no i13 index)
                        else:
                            W[:, :, i2, i11, i12, i13] = M
        return W
    if nu in (5, 6):
        n12 = 1 if N2 == 1 else N2 * O2
        W = np.zeros((P, nu, 2, N1 * O1, n12, 1), complex)
        for i11 in range(N1 * O1):
            for i12 in range(n12):
                for i2 in range(2):
                    if N2 == 1:
                        l, lp, ld, m, mp, md = i11, i11 + O1, i11 + 2 * O1, 0, 0, 0
                    else:
                        l, lp, ld, m, mp, md = i11, i11 + O1, i11 + O1, i12, i12, i12 + O2
                    a, b = _restricted(csr, N2 * O2 * l + m, i2, i2r)
                    if a or b:
                        continue
                    v, vp, vd = V(l, m)[:, None], V(lp, mp)[:, None], V(ld, md)[:, None]
                    ph = phi(i2)
                    if nu == 5:
                        M = np.block([[v, v, vp, vp, vd], [ph * v, -ph * v, vp, -vp, vd]]) / math.sqrt(5 * P)
                    else:
                        M = np.block([[v, v, vp, vp, vd, vd], [ph * v, -ph * v, ph * vp, -ph * vp, vd, -vd]]) / math.sqrt(6 * P)
                    W[:, :, i2, i11, i12, 0] = M
        return W
    # 7, 8 layers
    if N2 == 1:
        n12 = 1
        n11 = N1 * O1 // 2 if N1 == 4 else N1 * O1
    else:
        n11 = N1 * O1
        n12 = N2 * O2 if ((N1 == 2 and N2 == 2) or (N1 > 2 and N2 > 2)) else N2 * O2 // 2
    W = np.zeros((P, nu, 2, n11, n12, 1), complex)
    for i11 in range(n11):
        for i12 in range(n12):
            for i2 in range(2):
                if N2 == 1:
                    ls, ms = [i11, i11 + O1, i11 + 2 * O1, i11 + 3 * O1], [0, 0, 0, 0]
                else:
                    ls, ms = [i11, i11 + O1, i11, i11 + O1], [i12, i12, i12 + O2, i12 + O2]
                a, b = _restricted(csr, N2 * O2 * ls[0] + ms[0], i2, i2r)
                if a or b:
                    continue
                v, vp, vd, vt = (V(ls[j], ms[j])[:, None] for j in range(4))
                ph = phi(i2)
                if nu == 7:
                    M = np.block([[v, v, vp, vd, vd, vt, vt], [ph * v, -ph * v, ph * vp, vd, -vd, vt, -vt]]) / math.sqrt(7 * P)
                else:
                    M = np.block([[v, v, vp, vp, vd, vd, vt, vt],
                                  [ph * v, -ph * v, ph * vp, -ph * vp, vd, -vd, vt, -vt]]) / math.sqrt(8 * P)
                W[:, :, i2, i11, i12, 0] = M
    return W


# ----------------------------------------------------------------------------------------------
# SINR
# ----------------------------------------------------------------------------------------------
def type1_multi_panel_codebook(cfg, n_panels, n_layers):
    """getPMIType1MultiPanelCodebook (dlPMISelect.m:1351-1772): TS 38.214 Tables 5.2.2.2.2-1..-6, written out table by
    table like the reference.  ``cfg``: N1, N2, O1, O2, CodebookMode, CodebookSubsetRestriction, i2Restriction;
    ``n_panels`` = Ng in {2, 4} (mode 2 only with Ng = 2).
    -> W [P, nLayers, i20, i21, i22, i11, i12, i13, i141, i142, i143] complex128, restricted precoders all zero."""
    Ng, N1, N2, O1, O2 = int(n_panels), cfg["N1"], cfg["N2"], cfg["O1"], cfg["O2"]
    mode = cfg["CodebookMode"]
    if Ng not in (2, 4) or (mode == 2 and Ng != 2) or not 1 <= n_layers <= 4:
        raise ValueError("nr5g:dlPMISelect:InvalidPanelDimensions")
    P = 2 * Ng * N1 * N2                                                       # :1385
    phi = lambda x: np.exp(1j * np.pi * x / 2)                                 # :1390-1392
    a = lambda x: np.exp(1j * np.pi / 4 + 1j * np.pi * x / 2)
    b = lambda x: np.exp(-1j * np.pi / 4 + 1j * np.pi * x / 2)
    n11, n12, n141 = N1 * O1, N2 * O2, 4                                       # :1396-1400
    if mode == 1:                                                              # :1405-1419
        n142, n143 = (1, 1) if Ng == 2 else (4, 4)
        n21 = n22 = 1
    else:
        n142, n143, n21, n22 = 4, 1, 2, 2
    if n_layers == 1:                                                          # :1424-1426
        n13, n20, k1, k2 = 1, 4, [0], [0]
    elif n_layers == 2:                                                        # :1518-1536 (Table 5.2.2.2.1-3)
        n20 = 2
        if N1 > N2 and N2 > 1:
            k1, k2 = [0, O1, 0, 2 * O1], [0, 0, O2, 0]
        elif N1 == N2:
            k1, k2 = [0, O1, 0, O1], [0, 0, O2, O2]
        elif N1 == 2 and N2 == 1:
            k1, k2 = [0, O1], [0, 0]
        else:
            k1, k2 = [0, O1, 2 * O1, 3 * O1], [0, 0, 0, 0]
        n13 = len(k1)
    else:                                                                      # :1615-1636 (Table 5.2.2.2.2-2)
        n20 = 2
        tab = {(2, 1): ([O1], [0]), (4, 1): ([O1, 2 * O1, 3 * O1], [0, 0, 0]), (8, 1): ([O1, 2 * O1, 3 * O1, 4 * O1], [0] * 4),
               (2, 2): ([O1, 0, O1], [0, O2, O2]), (4, 2): ([O1, 0, O1, 2 * O1], [0, O2, O2, 0])}
        if (N1, N2) not in tab:
            raise ValueError("nr5g:dlPMISelect:InvalidPanelDimensions")
        k1, k2 = tab[(N1, N2)]
        n13 = len(k1)
    W = np.zeros((P, n_layers, n20, n21, n22, n11, n12, n13, n141, n142, n143), dtype=np.complex128)
    csr = cfg.get("CodebookSubsetRestriction", np.ones(n11 * n12))
    i2r = cfg.get("i2Restriction", np.ones(16))
    for i11 in range(n11):
        for i12 in range(n12):
            for i13 in range(n13):
                if _restricted(csr, N2 * O2 * i11 + i12, None, i2r)[0]:      # only v_lm restriction applies (:1434, :1546)
                    continue
                v = _vlm(N1, N2, O1, O2, i11, i12)
                vp = _vlm(N1, N2, O1, O2, i11 + k1[i13], i12 + k2[i13])
                for i141 in range(n141):
                    for i142 in range(n142):
                        for i143 in range(n143):
                            for i20 in range(n20):
                                for i21 in range(n21):
                                    for i22 in range(n22):
                                        if mode == 1:
                                            fn = phi(i20)
                                            cp = [1.0, phi(i141)] if Ng == 2 else [1.0, phi(i141), phi(i142), phi(i143)]
                                            plus = lambda x: np.concatenate([np.concatenate([c * x, c * fn * x]) for c in cp])
                                            minus = lambda x: np.concatenate([np.concatenate([c * x, -c * fn * x]) for c in cp])
                                        else:                                   # Ng = 2, mode 2 (:1489-1508 etc.)
                                            fn = phi(i20)
                                            c1, c2 = a(i141) * b(i21), a(i142) * b(i22)
                                            plus = lambda x: np.concatenate([x, fn * x, c1 * x, c2 * x])
                                            minus = lambda x: np.concatenate([x, -fn * x, c1 * x, -c2 * x])
                                        cols = {1: [plus(v)], 2: [plus(v), minus(vp)], 3: [plus(v), plus(vp), minus(v)],
                                                4: [plus(v), plus(vp), minus(v), minus(vp)]}[n_layers]
                                        W[:, :, i20, i21, i22, i11, i12, i13, i141, i142, i143] = (
                                            np.stack(cols, axis=1) / np.sqrt(n_layers * P))
    return W


def precoded_sinr_dl(H, n_var, W):
    """getPrecodedSINR (dlPMISelect.m:1825-1834): per-layer LMMSE SINR."""
    nu = W.shape[1]
    noise = n_var * np.eye(nu)
    den = noise @ np.linalg.inv((W.conj().T @ H.conj().T) @ H @ W + noise)
    return np.real(1.0 / np.diag(den) - 1.0)


def precoded_sinr_ul(H, sigma, W):
    """communication.phyLayer.precodedSINR (precodedSINR.m:11-18): summed over layers."""
    nu = W.shape[1]
    noise = sigma ** 2 * np.eye(nu)
    den = noise @ np.linalg.inv((W.conj().T @ H.conj().T) @ H @ W + noise)
    return float(np.real(np.sum(1.0 / np.diag(den) - 1.0)))


# ----------------------------------------------------------------------------------------------
# a9: dlPMISelect
# ----------------------------------------------------------------------------------------------
def dl_pmi_select_multi_panel(cfg, n_panels, re_k, re_l, n_layers, H, n_var=1e-10):
    """dlPMISelect with CodebookType = 'Type1MultiPanel' (dlPMISelect.m:385-501, multi-panel branches).  The reference walks its
    9-D index set [i20 i21 i22 | i11 i12 i13 i141 i142 i143] in MATLAB linear order (find(...,1) at :455 and :489), so the
    selection equals the single-panel selection on the array flattened to [i20*i21*i22, i11, i12, i13*i141*i142*i143].
    -> PMISet {i1 [6], i2 [3 x nSB]} (:456-457, :489), info with the 9-D SINR arrays."""
    Wmp = type1_multi_panel_codebook(cfg, n_panels, n_layers)
    P, nu = Wmp.shape[:2]
    d = Wmp.shape[2:]
    flat = (d[0] * d[1] * d[2], d[3], d[4], d[5] * d[6] * d[7] * d[8])
    pm, info = dl_pmi_select(dict(cfg, NumCSIRSPorts=P), re_k, re_l, n_layers, H, n_var,
                             W_override=np.asfortranarray(Wmp).reshape((P, nu) + flat, order="F"))
    n_sb = pm["i2"].size
    i1 = np.full(6, np.nan)
    if not np.any(np.isnan(pm["i1"])):
        i1[:2] = pm["i1"][:2]
        i1[2:] = [x + 1 for x in np.unravel_index(int(pm["i1"][2]) - 1, d[5:9], order="F")]
    i2 = np.full((3, n_sb), np.nan)
    for sb in range(n_sb):
        if not np.isnan(pm["i2"][sb]):
            i2[:, sb] = [x + 1 for x in np.unravel_index(int(pm["i2"][sb]) - 1, d[0:3], order="F")]
    out = {"W": Wmp}
    for key in ("SINRPerRE", "SINRPerSubband"):
        a = info[key]
        out[key] = None if a is None else np.asfortranarray(a).reshape(a.shape[:2] + tuple(d), order="F")
    return {"i1": i1, "i2": i2}, out


def dl_pmi_select(cfg, re_k, re_l, n_layers, H, n_var=1e-10, K=None, L=14, compact=True, W_override=None):
    """``[PMISet,info] = dlPMISelect(carrier,csirs,reportConfig,nLayers,H,nVar)`` (dlPMISelect.m:307-509),
    Type1SinglePanel.  ``re_k/re_l``: 1-based CSI-RS RE subscripts relative to the BWP (validateInputs :797-833).
    ``compact``: info['SINRPerRE'] is [nRE, nLayers, i2, i11, i12, i13] at the CSI-RS REs instead of the
    reference's K x L x ... NaN-filled array (same numbers, without the NaN padding)."""
    n_var = max(float(n_var), 1e-10)                                                    # :846-848
    n_sb, sb_sizes = subband_info(cfg["PMIMode"], cfg["NStartBWP"], cfg["NSizeBWP"], cfg["SubbandSize"])
    P = cfg["NumCSIRSPorts"]
    if W_override is not None:
        W = W_override
    else:
        W = np.ones((1, 1, 1, 1, 1, 1), complex) if P == 1 else type1_single_panel_codebook(cfg, n_layers, "ue")
    _, _, n2, n11, n12, n13 = W.shape
    sizes = (n2, n11, n12, n13)
    re_k = np.asarray(re_k, dtype=int)
    re_l = np.asarray(re_l, dtype=int)
    n_re = re_k.size
    if n_re == 0 or not np.any(W):
        return ({"i1": np.full(3, np.nan), "i2": np.full(n_sb, np.nan)},
                {"SINRPerRE": None, "SINRPerSubband": np.full((n_sb, n_layers) + sizes, np.nan), "W": W})
    Hp = np.transpose(np.asarray(H), (2, 3, 0, 1))                                       # :384
    S = np.full((n_re, n_layers) + sizes, np.nan)
    for e in range(n_re):                                                                 # :385-428
        Ht = Hp[:, :, re_k[e] - 1, re_l[e] - 1]
        for i11 in range(n11):
            for i12 in range(n12):
                for i13 in range(n13):
                    for i2 in range(n2):
                        cw = W[:, :, i2, i11, i12, i13]
                        if np.any(cw):
                            S[e, :, i2, i11, i12, i13] = precoded_sinr_dl(Ht, n_var, cw)
    pm = {}
    if np.all(np.isnan(S)):
        pm["i1"], pm["i2"] = np.full(3, np.nan), np.array([np.nan])
    else:
        total = matlab_round4(np.nansum(S, axis=(0, 1)))                                  # :444-449
        lin = int(np.flatnonzero(total.reshape(-1, order="F") == total.max())[0])         # :453
        i2, i11, i12, i13 = np.unravel_index(lin, sizes, order="F")
        pm["i1"], pm["i2"] = np.array([i11 + 1, i12 + 1, i13 + 1], float), np.array([i2 + 1.0])
    warnings.filterwarnings("ignore", message="Mean of empty slice")
    sub = np.full((n_sb, n_layers) + sizes, np.nan)
    i2_out = np.full(n_sb, np.nan)
    i1 = np.ones(3, dtype=int) if np.any(np.isnan(pm["i1"])) else pm["i1"].astype(int)
    start = 0
    for sb in range(n_sb):                                                                # :471-501
        lo, hi = start * 12 + 1, (start + sb_sizes[sb]) * 12
        sel = (re_k >= lo) & (re_k <= hi)
        if sel.any():
            # mean over k per symbol, then over symbols (mean(mean(.,'omitnan'),'omitnan'))
            acc = []
            for l in np.unique(re_l[sel]):
                with np.errstate(invalid="ignore"):
                    acc.append(np.nanmean(S[sel & (re_l == l)], axis=0))
            with np.errstate(invalid="ignore"):
                sub[sb] = np.nanmean(np.stack(acc), axis=0)
            t = matlab_round4(np.nansum(sub[sb][:, :, i1[0] - 1, i1[1] - 1, i1[2] - 1], axis=0))   # :492
            i2_out[sb] = int(np.argmax(t)) + 1                                            # :496
        start += sb_sizes[sb]
    pm["i2"] = i2_out                                                                     # :476,:496 overwrite per subband
    info = {"SINRPerRE": S if compact else _scatter(S, re_k, re_l, K or cfg["NSizeBWP"] * 12, L),
            "SINRPerSubband": sub, "W": W}
    return pm, info


def _scatter(S, re_k, re_l, K, L):
    full = np.full((K, L) + S.shape[1:], np.nan)
    full[re_k - 1, re_l - 1] = S
    return full


# ----------------------------------------------------------------------------------------------
# a12: riSelect
# ----------------------------------------------------------------------------------------------
def _mp_flat_codebook(cfg, n_panels, n_layers):
    """Type1MultiPanel codebook with its 9-D index set flattened in MATLAB linear order to [i2', i11, i12, i13'] (see
    dl_pmi_select_multi_panel), and the 9 index-set lengths."""
    Wmp = type1_multi_panel_codebook(cfg, n_panels, n_layers)
    P, nu = Wmp.shape[:2]
    d = Wmp.shape[2:]
    flat = (d[0] * d[1] * d[2], d[3], d[4], d[5] * d[6] * d[7] * d[8])
    return np.asfortranarray(Wmp).reshape((P, nu) + flat, order="F"), d


def mp_unflatten_pmi(pm, d):
    """Flattened PMISet {i1 [3], i2 [nSB]} -> the multi-panel form {i1 [6], i2 [3 x nSB]} (dlPMISelect.m:456-457, :489)."""
    n_sb = np.asarray(pm["i2"]).size
    i1 = np.full(6, np.nan)
    if not np.any(np.isnan(pm["i1"])):
        i1[:2] = pm["i1"][:2]
        i1[2:] = [x + 1 for x in np.unravel_index(int(pm["i1"][2]) - 1, d[5:9], order="F")]
    i2 = np.full((3, n_sb), np.nan)
    for sb in range(n_sb):
        if not np.isnan(pm["i2"][sb]):
            i2[:, sb] = [x + 1 for x in np.unravel_index(int(pm["i2"][sb]) - 1, d[0:3], order="F")]
    return {"i1": i1, "i2": i2}


def _pmi_any_panel(cfg, n_panels, re_k, re_l, r, H, n_var):
    """dlPMISelect on the (flattened) index set of either codebook type -> (flattened PMISet, info, index-set lengths or None)."""
    if n_panels >= 2:
        Wf, d = _mp_flat_codebook(cfg, n_panels, r)
        pm, info = dl_pmi_select(dict(cfg, NumCSIRSPorts=Wf.shape[0]), re_k, re_l, r, H, n_var, W_override=Wf)
        return pm, info, d
    pm, info = dl_pmi_select(cfg, re_k, re_l, r, H, n_var)
    return pm, info, None


def ri_select(cfg, re_k, re_l, H, n_var=1e-10, n_panels=0):
    """``[RI,PMISet] = riSelect(carrier,csirs,reportConfig,H,nVar)`` (riSelect.m:198-294).  ``n_panels`` >= 2: CodebookType
    'Type1MultiPanel' (ranks 1-4, RIRestriction of 4 bits, riSelect.m:222-231, :449-456; PMISet in the multi-panel form)."""
    n_sb, _ = subband_info(cfg["PMIMode"], cfg["NStartBWP"], cfg["NSizeBWP"], cfg["SubbandSize"])
    P, R = H.shape[3], H.shape[2]
    max_rank = min(R, P, 4) if n_panels >= 2 else min(R, P)
    valid = [r for r in range(1, max_rank + 1) if cfg["RIRestriction"][r - 1]]
    if not valid or len(re_k) == 0:
        if n_panels >= 2:
            return np.nan, {"i1": np.full(6, np.nan), "i2": np.full((3, n_sb), np.nan)}
        return np.nan, {"i1": np.full(3, np.nan), "i2": np.full(n_sb, np.nan)}
    best, total = -np.inf, np.full(max_rank, np.nan)
    RI, pm_best, pm = np.nan, None, None
    dims = {}
    for r in valid:                                                                       # :254-285
        pm, info, dims[r] = _pmi_any_panel(cfg, n_panels, re_k, re_l, r, H, n_var)
        sb_sinr = np.full((n_sb, r), np.nan)
        if not np.any(np.isnan(pm["i1"])):
            i1 = pm["i1"].astype(int)
            for s in range(n_sb):
                if not np.isnan(pm["i2"][s]):
                    sb_sinr[s] = info["SINRPerSubband"][s, :, int(pm["i2"][s]) - 1, i1[0] - 1, i1[1] - 1, i1[2] - 1] * r
            with np.errstate(invalid="ignore"):
                layer = np.nanmean(sb_sinr, axis=0)                                       # :278
            total[r - 1] = np.sum(layer[layer >= 1])                                      # :282
        if total[r - 1] > best + 0.1:                                                     # :284
            best, RI, pm_best = total[r - 1], r, pm
    if np.all(np.isnan(total)):
        return np.nan, (mp_unflatten_pmi(pm, dims[valid[-1]]) if n_panels >= 2 else pm)
    return RI, (mp_unflatten_pmi(pm_best, dims[int(RI)]) if n_panels >= 2 else pm_best)


# ----------------------------------------------------------------------------------------------
# a13: cqiSelect
# ----------------------------------------------------------------------------------------------
def layer_demap_sums(layer_sinr):
    """cellfun(@sum, nrLayerDemap(.)) (cqiSelect.m:617): PARITY-UNPINNED TS 38.211 Table 7.3.1.3-1:
    <=4 layers -> one codeword; otherwise floor(n/2) layers to CW0 and the rest to CW1."""
    v = np.asarray(layer_sinr, float)
    n = v.size
    if n <= 4:
        return np.array([v.sum()])
    h = n // 2
    return np.array([v[:h].sum(), v[h:].sum()])


def get_cqi(lin_sinr, table):
    """getCQI (cqiSelect.m:697-722)."""
    if np.isnan(lin_sinr):
        return np.nan
    with np.errstate(divide="ignore"):
        db = 10.0 * np.log10(lin_sinr)
    idx = np.flatnonzero(np.asarray(table) <= db)
    return 0.0 if idx.size == 0 else float(idx[-1] + 1)


def cqi_select(cfg, re_k, re_l, n_layers, H, n_var, sinr_table, L=14, n_panels=0):
    """``[CQI,PMISet,CQIInfo,PMIInfo] = cqiSelect(carrier,csirs,reportConfig,nLayers,H,nVar,SINRTable)``
    (cqiSelect.m:411-695, CSI-RS object syntax, no PRGSize).  ``n_panels`` >= 2: Type1MultiPanel (the SINR look-ups of :589-603
    on the flattened index set; PMISet returned in the multi-panel form, PMIInfo arrays stay flattened)."""
    n_cq, cq_sizes = subband_info(cfg["CQIMode"], cfg["NStartBWP"], cfg["NSizeBWP"], cfg["SubbandSize"])
    n_cw = int(math.ceil(n_layers / 4))
    pm, info, mp_dims = _pmi_any_panel(cfg, n_panels, re_k, re_l, n_layers, H, n_var)     # :507
    S = info["SINRPerRE"]
    re_k = np.asarray(re_k, int)
    re_l = np.asarray(re_l, int)
    sinr_sb = np.full((n_cq, n_layers), np.nan)
    nan_pm = np.all(np.isnan(pm["i1"])) and np.all(np.isnan(pm["i2"]))
    if not nan_pm:
        i1 = pm["i1"].astype(int)
        if cfg["PMIMode"].lower() == "wideband":                                          # :586-596 getSubbandSINR
            start = 0
            for s in range(n_cq):
                lo, hi = start * 12 + 1, (start + cq_sizes[s]) * 12
                sel = (re_k >= lo) & (re_k <= hi)
                if sel.any() and not np.isnan(pm["i2"][0]):
                    acc = [np.nanmean(S[sel & (re_l == l)][:, :, int(pm["i2"][0]) - 1, i1[0] - 1, i1[1] - 1, i1[2] - 1], axis=0)
                           for l in np.unique(re_l[sel])]
                    sinr_sb[s] = np.nanmean(np.stack(acc), axis=0)
                start += cq_sizes[s]
        else:                                                                             # :604-614
            for s in range(len(pm["i2"])):
                if not np.isnan(pm["i2"][s]):
                    sinr_sb[s] = info["SINRPerSubband"][s, :, int(pm["i2"][s]) - 1, i1[0] - 1, i1[1] - 1, i1[2] - 1]
    sb_cw = np.zeros((n_cq, n_cw))
    for s in range(n_cq):                                                                 # :610-627
        sb_cw[s] = layer_demap_sums(sinr_sb[s]) if not np.any(np.isnan(sinr_sb[s])) else np.nan
    if sb_cw.shape[0] > 1:
        with np.errstate(invalid="ignore"):
            sb_cw = np.vstack([np.nanmean(sb_cw, axis=0), sb_cw])                         # :631-633
    pm_out = mp_unflatten_pmi(pm, mp_dims) if n_panels >= 2 else pm
    if nan_pm:
        ns = 0 if n_cq == 1 else n_cq
        return np.full((ns + 1, n_cw), np.nan), pm_out, {"SINRPerSubbandPerCW": np.full((ns + 1, n_cw), np.nan)}, info
    cq_all = np.vectorize(lambda x: get_cqi(x, sinr_table))(sb_cw)                        # :653
    if cfg["CQIMode"].lower() == "subband":
        diff = cq_all[1:] - cq_all[0]                                                     # :661
        off = np.full(diff.shape, np.nan)
        off[diff == 0] = 0
        off[diff == 1] = 1
        off[diff >= 2] = 2
        off[diff <= -1] = 3
        cqi = np.vstack([cq_all[0:1], off])                                               # :677
    else:
        cqi = cq_all[0:1]
    return cqi, pm_out, {"SINRPerSubbandPerCW": sb_cw, "SubbandCQI": cq_all}, info


# ----------------------------------------------------------------------------------------------
# a14: UL pmiSelect (+ nrPUSCHCodebook tables)
# ----------------------------------------------------------------------------------------------
def max_pusch_tpmi(n_layers, n_ports):
    """maxPUSCHPrecodingMatrixIndicator (maxPUSCHPrecodingMatrixIndicator.m:14-74)."""
    return {(1, 1): 0, (1, 2): 5, (1, 4): 27, (2, 2): 2, (2, 4): 21, (3, 4): 6, (4, 4): 4}[(n_layers, n_ports)]


def _m(rows):
    return np.array(rows, dtype=complex)


def pusch_codebook(n_layers, n_ports, tpmi):
    """``nrPUSCHCodebook(nlayers,nports,tpmi).'`` -> W [nports x nlayers] (pmiSelect.m:45).
    PARITY-UNPINNED: TS 38.211 Tables 6.3.1.5-1 ... -7 (transform precoding disabled), recalled from the spec."""
    j = 1j
    if n_ports == 1:
        return _m([[1]])
    if n_ports == 2 and n_layers == 1:
        return _m([[1, 0], [0, 1], [1, 1], [1, -1], [1, j], [1, -j]][tpmi])[:, None] / math.sqrt(2)
    if n_ports == 2 and n_layers == 2:
        return [_m([[1, 0], [0, 1]]) / math.sqrt(2), _m([[1, 1], [1, -1]]) / 2, _m([[1, 1], [j, -j]]) / 2][tpmi]
    if n_layers == 1:
        t = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 1, 0], [1, 0, -1, 0], [1, 0, j, 0], [1, 0, -j, 0],
             [0, 1, 0, 1], [0, 1, 0, -1], [0, 1, 0, j], [0, 1, 0, -j], [1, 1, 1, 1], [1, 1, j, j], [1, 1, -1, -1], [1, 1, -j, -j],
             [1, j, 1, j], [1, j, j, -1], [1, j, -1, -j], [1, j, -j, 1], [1, -1, 1, -1], [1, -1, j, -j], [1, -1, -1, 1],
             [1, -1, -j, j], [1, -j, 1, -j], [1, -j, j, 1], [1, -j, -1, j], [1, -j, -j, -1]]
        return _m(t[tpmi])[:, None] / 2
    if n_layers == 2:
        sel = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
        if tpmi < 6:
            W = np.zeros((4, 2), complex)
            W[sel[tpmi][0], 0] = 1
            W[sel[tpmi][1], 1] = 1
            return W / 2
        if tpmi < 14:
            ab = [(1, -j), (1, j), (-j, 1), (-j, -1), (-1, -j), (-1, j), (j, 1), (j, -1)][tpmi - 6]
            return _m([[1, 0], [0, 1], [ab[0], 0], [0, ab[1]]]) / 2
        t = [[[1, 1], [1, 1], [1, -1], [1, -1]], [[1, 1], [1, 1], [j, -j], [j, -j]], [[1, 1], [j, j], [1, -1], [j, -j]],
             [[1, 1], [j, j], [j, -j], [-1, 1]], [[1, 1], [-1, -1], [1, -1], [-1, 1]], [[1, 1], [-1, -1], [j, -j], [-j, j]],
             [[1, 1], [-j, -j], [1, -1], [-j, j]], [[1, 1], [-j, -j], [j, -j], [1, -1]]]
        return _m(t[tpmi - 14]) / (2 * math.sqrt(2))
    if n_layers == 3:
        t = [(_m([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0]]), 2), (_m([[1, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1]]), 2),
             (_m([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, 0, 1]]), 2),
             (_m([[1, 1, 1], [1, -1, 1], [1, 1, -1], [1, -1, -1]]), 2 * math.sqrt(3)),
             (_m([[1, 1, 1], [1, -1, 1], [j, j, -j], [j, -j, -j]]), 2 * math.sqrt(3)),
             (_m([[1, 1, 1], [-1, 1, -1], [1, 1, -1], [-1, 1, 1]]), 2 * math.sqrt(3)),
             (_m([[1, 1, 1], [-1, 1, -1], [j, j, -j], [-j, j, j]]), 2 * math.sqrt(3))]
        return t[tpmi][0] / t[tpmi][1]
    t = [(np.eye(4, dtype=complex), 2), (_m([[1, 1, 0, 0], [0, 0, 1, 1], [1, -1, 0, 0], [0, 0, 1, -1]]), 2 * math.sqrt(2)),
         (_m([[1, 1, 0, 0], [0, 0, 1, 1], [j, -j, 0, 0], [0, 0, j, -j]]), 2 * math.sqrt(2)),
         (_m([[1, 1, 1, 1], [1, -1, 1, -1], [1, 1, -1, -1], [1, -1, -1, 1]]), 4),
         (_m([[1, 1, 1, 1], [1, -1, 1, -1], [j, j, -j, -j], [j, -j, -j, j]]), 4)]
    return t[tpmi][0] / t[tpmi][1]


def sinr_per_subband(sinr, band_size):
    """sinrPerSubband (sinrPerSubband.m:12-35)."""
    nrb = sinr.shape[0] / 12
    r = nrb / band_size
    n_sb = int(math.ceil(r))
    out = np.zeros((n_sb, sinr.shape[2]))
    idx = np.zeros((n_sb, 2), dtype=int)
    for s in range(n_sb):
        lo = 12 * band_size * s + 1
        hi = 12 * band_size * (s + 1) if s < int(math.floor(r)) else int(12 * band_size * r)
        idx[s] = (lo, hi)
        blk = sinr[lo - 1: hi]
        with np.errstate(invalid="ignore", divide="ignore"):
            out[s] = blk.sum(axis=(0, 1)) / np.count_nonzero(blk.sum(axis=2))
    return out, idx


def pmi_select(n_layers, hest, noise_est, band_size):
    """``[pmi,sinr,subbandIndices] = pmiSelect(nlayers,hest,noiseest,bandSize)`` (pmiSelect.m:28-66)."""
    hest = np.asarray(hest)
    K, Ls, R, P = hest.shape
    mx = max_pusch_tpmi(n_layers, P)
    sinr = np.zeros((K, Ls, mx + 1))
    mask = hest.sum(axis=(2, 3)) != 0                                                     # :35
    if not mask.any() or noise_est == 0:
        return np.nan, np.nan, np.nan
    sigma = math.sqrt(noise_est)
    ks, ls = np.nonzero(mask)
    for t in range(mx + 1):
        W = pusch_codebook(n_layers, P, t)
        for k, l in zip(ks, ls):
            sinr[k, l, t] = precoded_sinr_ul(hest[k, l], sigma, W)                        # :51
    bands, idx = sinr_per_subband(sinr, band_size)                                        # :55
    pmi = np.argmax(bands, axis=1).astype(float)                                          # :56 (first max), 0-based (:58)
    pmi[np.isnan(bands[:, 0])] = np.nan
    return pmi, bands, idx


# ----------------------------------------------------------------------------------------------
# a15: prgPrecode
# ----------------------------------------------------------------------------------------------
def prg_precode(siz, nstartgrid, portsym, portind, F):
    """``[antsym,antind] = prgPrecode(siz,nstartgrid,portsym,portind,F)`` (prgPrecode.m:53-144).
    portsym/portind: [NRE x nLayers] (1-based linear indices into a K x L x nLayers grid); F: [nLayers x P x NPRG].
    PARITY-UNPINNED (nrExtractResources): the output keeps the RE positions of portind's first plane and
    projects them onto all P antenna planes -> antsym/antind [NRE x P]."""
    F = np.asarray(F)
    if F.ndim == 2:
        F = F[:, :, None]
    nu, P, nprg = F.shape
    K, Lsym = int(siz[0]), int(siz[1])
    nrb = K // 12
    pd = int(math.ceil((nrb + nstartgrid) / nprg))                                        # getPRGSet :94-100
    prgset = np.repeat(np.arange(1, nprg + 1), pd)[nstartgrid: nstartgrid + nrb]
    portind = np.asarray(portind, dtype=np.int64)
    portsym = np.asarray(portsym)
    re = (portind - 1) % (K * Lsym)                                                       # position within a plane
    k = re % K
    prg = prgset[k // 12]                                                                 # :80-84
    ant = np.zeros((K * Lsym, P), complex)
    for g in range(1, nprg + 1):                                                          # :118-138
        sel = prg == g
        if not sel.any():
            continue
        port = np.zeros((K * Lsym, nu), complex)
        lin = portind[sel] - 1                                                            # portgrid(indin(thisprg)) = symin(thisprg)
        port[lin % (K * Lsym), lin // (K * Lsym)] = portsym[sel]
        ant += port @ F[:, :, g - 1]
    pos = re[:, 0] if re.ndim == 2 else re
    antind = pos[:, None] + 1 + (K * Lsym) * np.arange(P)[None, :]
    return ant[pos, :], antind


# ----------------------------------------------------------------------------------------------
# vectorised variants (same arithmetic, batched linear algebra) — used as the best-effort CPU baseline
# ----------------------------------------------------------------------------------------------
def sinr_per_re_vectorized(cfg, re_k, re_l, n_layers, H, n_var):
    """SINRPerRE [nRE, nLayers, i2, i11, i12, i13] with stacked np.linalg.inv instead of the loop nest."""
    n_var = max(float(n_var), 1e-10)
    P = cfg["NumCSIRSPorts"]
    W = np.ones((1, 1, 1, 1, 1, 1), complex) if P == 1 else type1_single_panel_codebook(cfg, n_layers, "ue")
    sizes = W.shape[2:]
    Wf = W.reshape(W.shape[0], n_layers, -1, order="F")                      # [P, nu, nCand] (MATLAB linear order)
    valid = np.abs(Wf).sum(axis=(0, 1)) > 0
    Hs = np.asarray(H)[np.asarray(re_k) - 1, np.asarray(re_l) - 1]           # [nRE, R, P]
    G = np.einsum("erp,pvc->ecrv", Hs.astype(complex), Wf[:, :, valid])
    A = np.einsum("ecrv,ecrw->ecvw", G.conj(), G) + n_var * np.eye(n_layers)
    d = np.real(np.diagonal(np.linalg.inv(A), axis1=2, axis2=3))             # [nRE, nValid, nu]
    S = np.full((Hs.shape[0], n_layers, Wf.shape[2]), np.nan)
    S[:, :, valid] = np.transpose(1.0 / (n_var * d) - 1.0, (0, 2, 1))
    return S.reshape((Hs.shape[0], n_layers) + sizes, order="F"), W


def pmi_at_rank_vectorized(cfg, re_k, re_l, r, H, n_var):
    """dlPMISelect at rank r (dlPMISelect.m:385-501) from the vectorised SINR array: (PMISet, SINRPerSubband at the reported
    PMI [nSB x r]).  Same selection logic as dl_pmi_select above."""
    re_k = np.asarray(re_k, int)
    n_sb, sb_sizes = subband_info(cfg["PMIMode"], cfg["NStartBWP"], cfg["NSizeBWP"], cfg["SubbandSize"])
    S, _ = sinr_per_re_vectorized(cfg, re_k, re_l, r, H, n_var)
    sizes = S.shape[2:]
    total = matlab_round4(np.nansum(S, axis=(0, 1)))
    lin = int(np.flatnonzero(total.reshape(-1, order="F") == total.max())[0])
    i2, i11, i12, i13 = np.unravel_index(lin, sizes, order="F")
    i2s, sel = np.full(n_sb, np.nan), np.full((n_sb, r), np.nan)
    start = 0
    for sb in range(n_sb):
        m = (re_k >= start * 12 + 1) & (re_k <= (start + sb_sizes[sb]) * 12)
        if m.any():
            sub = np.nanmean(S[m][:, :, :, i11, i12, i13], axis=0)           # [nu, n2] (single CSI-RS symbol)
            t = matlab_round4(np.nansum(sub, axis=0))
            i2s[sb] = int(np.argmax(t)) + 1
            sel[sb] = sub[:, int(i2s[sb]) - 1]
        start += sb_sizes[sb]
    return {"i1": np.array([i11 + 1, i12 + 1, i13 + 1.0]), "i2": i2s}, sel


def cqi_from_subband_sinr(sel, rank, sinr_table):
    """cqiSelect.m:610-653 from SINRPerSubband at the reported PMI: per-codeword sums (nrLayerDemap), wideband row = mean over
    the subbands when there is more than one, CQI look-up.  Returns absolute CQIs [rows x nCW]."""
    n_sb = sel.shape[0]
    n_cw = int(math.ceil(rank / 4))
    cw = np.stack([layer_demap_sums(sel[s]) if not np.any(np.isnan(sel[s])) else np.full(n_cw, np.nan) for s in range(n_sb)])
    full = np.vstack([np.nanmean(cw, axis=0), cw]) if n_sb > 1 else cw
    return np.vectorize(lambda x: get_cqi(x, sinr_table))(full)


def csi_report_vectorized(cfg, re_k, re_l, H, n_var, sinr_table, rank_cap=4, return_all=False):
    """riSelect + cqiSelect (uePhy.m:900-907) using the vectorised SINR kernel; selection logic identical to
    dl_pmi_select / ri_select / cqi_select above.  return_all: also the unclipped RI and every rank's (PMISet, SINR)."""
    R, P = H.shape[2], H.shape[3]
    best, RI, keep = -np.inf, np.nan, {}
    for r in range(1, min(R, P) + 1):
        if not cfg["RIRestriction"][r - 1]:
            continue
        pm, sel = pmi_at_rank_vectorized(cfg, re_k, re_l, r, H, n_var)
        with np.errstate(invalid="ignore"):
            layer = np.nanmean(sel * r, axis=0)
        tot_r = np.sum(layer[layer >= 1])
        keep[r] = (pm, sel)
        if tot_r > best + 0.1:
            best, RI = tot_r, r
    rank = int(min(RI, rank_cap))
    pm, sel = keep[rank]
    cq = cqi_from_subband_sinr(sel, rank, sinr_table)
    if return_all:
        return rank, pm, cq, RI, keep
    return rank, pm, cq
