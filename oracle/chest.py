"""Float64 oracle of the pilot-based channel estimator (test infrastructure only; PARITY UNPINNED).

SURVEY 8(f) row 1: the step immediately before the COMM hot path.  The reference obtains the channel matrix it hands to
riSelect / cqiSelect / pmiSelect from the closed 5G Toolbox function

    [Hest, nVar] = nrChannelEstimate(rxGrid, refInd, refSym, 'CDMLengths', cdmLen [, 'AveragingWindow', [F T]])

(call sites +communication/+phyLayer/uePhy.m:897 (CSI-RS, cdmLen from CDMType :889-895), gNBPhy.m:1030 (SRS,
'AveragingWindow',[0 7], cdmLen from srsCDMLengths.m), uePhy.m:836 / gNBPhy.m:935 (PDSCH / PUSCH DM-RS)).  The toolbox source
is not in the repository and no golden vectors exist, so this file states the estimator this build implements, following
the processing steps the toolbox documents (least-squares estimates at the reference REs, CDM despreading, interpolation
over the slot, noise estimation) with every free choice spelled out:

  1. LS estimate at every reference RE of port p:      z = rx / s                                  (per receive antenna)
  2. CDM despreading: the reference REs of a port are cut into blocks of FD consecutive reference subcarriers x TD
     consecutive reference symbols (cdmLen = [FD TD]); each block is replaced by the mean of its LS estimates, located at
     the block's mean subcarrier / symbol position.  (The orthogonal cover codes of the other ports sharing the block
     cancel in the mean when the channel is flat over the block.)
  3. optional moving average over F x T neighbouring blocks ('AveragingWindow'; odd sizes, window truncated at the edges;
     0 or 1 = none -- the toolbox's SNR-dependent automatic choice for 0 is NOT reproduced).
  4. frequency: linear interpolation between block centres, constant (nearest-block) extrapolation outside their span.
  5. time: linear interpolation between block rows, constant extrapolation outside (one reference symbol -> the estimate
     is held over the slot).
  6. nVar: the despread estimates of neighbouring blocks differ by channel variation plus noise of variance
     sigma^2 / (FD*TD) each; the second difference d_b = h_{b+1} - 2 h_b + h_{b-1} removes a locally linear channel and has
     variance 6 sigma^2 / (FD*TD), so   nVar = FD*TD/6 * mean |d_b|^2   over blocks, block rows, antennas and ports
     (0 when a port has fewer than three blocks along frequency).

refInd / refSym use the toolbox convention: 1-based column-major linear indices into a K x L x P grid and the reference
symbols transmitted there (any layout; port = floor((ind-1) / (K*L))).  Every port must use the same number of reference
subcarriers on each of its reference symbols (true for CSI-RS, SRS and DM-RS patterns).
"""
from __future__ import annotations

import numpy as np


def pilot_layout(ref_ind, ref_sym, K, L, P):
    """-> per port p: dict(k=[nK] sorted 0-based subcarriers, l=[nL] sorted 0-based symbols, s=[nL, nK] complex symbols)."""
    ind = np.asarray(ref_ind).reshape(-1, order="F").astype(np.int64) - 1
    sym = np.asarray(ref_sym).reshape(-1, order="F").astype(np.complex128)
    if ind.size != sym.size:
        raise ValueError("refInd and refSym disagree in size")
    if ind.size and (ind.min() < 0 or ind.max() >= K * L * P):
        raise ValueError("refInd outside the K x L x P grid")
    port = ind // (K * L)
    rem = ind % (K * L)
    kk, ll = rem % K, rem // K
    out = []
    for p in range(P):
        sel = port == p
        if not sel.any():
            raise ValueError(f"port {p} has no reference REs")
        ks, ls = np.unique(kk[sel]), np.unique(ll[sel])
        S = np.zeros((ls.size, ks.size), dtype=np.complex128)
        seen = np.zeros(S.shape, dtype=bool)
        ki = np.searchsorted(ks, kk[sel])
        li = np.searchsorted(ls, ll[sel])
        S[li, ki] = sym[sel]
        seen[li, ki] = True
        if not seen.all():
            raise ValueError("reference REs of a port must form a (symbols x subcarriers) product grid")
        out.append({"k": ks, "l": ls, "s": S})
    return out


def _blocks(pos, n):
    """Cut sorted positions into consecutive blocks of n (last one may be shorter) -> list of index arrays."""
    return [np.arange(i, min(i + n, pos.size)) for i in range(0, pos.size, n)]


def _moving_average(D, F, T):
    """Truncated moving average over F blocks in frequency (axis 1) and T block rows (axis 0); D [nBt, nBf, ...]."""
    out = D
    for axis, w in ((1, F), (0, T)):
        if w <= 1:
            continue
        h = w // 2
        n = out.shape[axis]
        acc = np.zeros_like(out)
        for i in range(n):
            lo, hi = max(0, i - h), min(n, i + h + 1)
            sl = [slice(None)] * out.ndim
            sl[axis] = slice(lo, hi)
            dst = [slice(None)] * out.ndim
            dst[axis] = i
            acc[tuple(dst)] = out[tuple(sl)].mean(axis=axis)
        out = acc
    return out


def _interp_table(centres, n):
    """For positions 0..n-1: lower block index and weight of linear interpolation with constant extrapolation."""
    c = np.asarray(centres, dtype=np.float64)
    x = np.arange(n, dtype=np.float64)
    if c.size == 1:
        return np.zeros(n, dtype=np.int64), np.zeros(n)
    lo = np.clip(np.searchsorted(c, x, side="right") - 1, 0, c.size - 2)
    w = (x - c[lo]) / (c[lo + 1] - c[lo])
    w = np.clip(w, 0.0, 1.0)     # constant extrapolation outside the span of the block centres
    return lo, w


def channel_estimate(rx_grid, ref_ind, ref_sym, n_ports, cdm_lengths=(1, 1), averaging_window=(0, 0)):
    """-> (H [K, L, R, P] complex128, nVar float).  See the module docstring for the algorithm."""
    rx = np.asarray(rx_grid, dtype=np.complex128)
    if rx.ndim == 2:
        rx = rx[:, :, None]
    K, L, R = rx.shape
    P = int(n_ports)
    FD, TD = int(cdm_lengths[0]), int(cdm_lengths[1])
    F, T = int(averaging_window[0]), int(averaging_window[1])
    lay = pilot_layout(ref_ind, ref_sym, K, L, P)
    H = np.zeros((K, L, R, P), dtype=np.complex128)
    nsum, ncnt = 0.0, 0
    for p, pl in enumerate(lay):
        ks, ls, S = pl["k"], pl["l"], pl["s"]
        Z = rx[ks[None, :], ls[:, None], :] * (np.conj(S) / np.abs(S) ** 2)[:, :, None]     # [nL, nK, R]  step 1
        bf, bt = _blocks(ks, FD), _blocks(ls, TD)
        D = np.zeros((len(bt), len(bf), R), dtype=np.complex128)                            # step 2
        for it, tt in enumerate(bt):
            for jf, ff in enumerate(bf):
                D[it, jf] = Z[np.ix_(tt, ff)].reshape(-1, R).mean(axis=0)
        kc = np.array([ks[ff].mean() for ff in bf])
        lc = np.array([ls[tt].mean() for tt in bt])
        if len(bf) >= 3:                                                                    # step 6 (before averaging)
            d = D[:, 2:] - 2.0 * D[:, 1:-1] + D[:, :-2]
            nsum += float((np.abs(d) ** 2).sum())
            ncnt += d.size
        D = _moving_average(D, F, T)                                                        # step 3
        flo, fw = _interp_table(kc, K)                                                      # step 4
        if len(bf) == 1:
            Fq = np.repeat(D[:, :1], K, axis=1)
        else:
            Fq = D[:, flo] * (1.0 - fw)[None, :, None] + D[:, flo + 1] * fw[None, :, None]  # [nBt, K, R]
        tlo, tw = _interp_table(lc, L)                                                      # step 5
        if len(bt) == 1:
            Hp = np.repeat(Fq[:1], L, axis=0)
        else:
            Hp = Fq[tlo] * (1.0 - tw)[:, None, None] + Fq[tlo + 1] * tw[:, None, None]      # [L, K, R]
        H[:, :, :, p] = Hp.transpose(1, 0, 2)
    n_var = (FD * TD / 6.0) * nsum / ncnt if ncnt else 0.0
    return H, n_var


# ---- synthetic reference-signal layouts for the tests / benchmarks ----------------------------------------------------
def csirs_row5_layout(n_rb, k0=1, l0=0, seed=0, L=14):
    """4-port CSI-RS row 5 of TS 38.211 Table 7.4.1.5.3-1 (the reference's configuration, setupCSIRS.m:8-10): density 1,
    FD-CDM2, CDM group j at (k0 + {0,1}, l0 + j), j = 0,1; ports 2j, 2j+1 share group j with cover codes [+1 +1], [+1 -1].
    QPSK base sequence r(m) drawn from `seed` (nrCSIRS' Gold sequence is not reproduced).  -> (refInd [nRE x 4] 1-based,
    refSym [nRE x 4], cdm_lengths)."""
    K = 12 * n_rb
    rng = np.random.default_rng(seed)
    ind = np.zeros((2 * n_rb, 4), dtype=np.int64)
    sym = np.zeros((2 * n_rb, 4), dtype=np.complex128)
    for j in range(2):
        r = (rng.integers(0, 2, (n_rb, 2)) * 2 - 1 + 1j * (rng.integers(0, 2, (n_rb, 2)) * 2 - 1)) / np.sqrt(2)
        for s in range(2):
            p = 2 * j + s
            wf = np.array([1.0, 1.0 if s == 0 else -1.0])
            k = (12 * np.arange(n_rb)[:, None] + k0 + np.arange(2)[None, :]).reshape(-1)
            ind[:, p] = 1 + k + K * (l0 + j) + K * L * p
            sym[:, p] = (r * wf[None, :]).reshape(-1)
    return ind, sym, (2, 1)


def apply_channel(H, ref_ind, ref_sym, noise=None):
    """rxGrid [K, L, R] = sum_p H[:, :, :, p] * x_p (+ noise) for the reference grid described by refInd / refSym."""
    K, L, R, P = H.shape
    ind = np.asarray(ref_ind).reshape(-1, order="F").astype(np.int64) - 1
    sym = np.asarray(ref_sym).reshape(-1, order="F").astype(np.complex128)
    X = np.zeros(K * L * P, dtype=np.complex128)
    X[ind] = sym
    X = X.reshape((K, L, P), order="F")
    rx = np.einsum("klrp,klp->klr", H, X)
    if noise is not None:
        rx = rx + noise
    return rx
