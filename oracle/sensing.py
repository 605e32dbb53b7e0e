"""Float64 oracle for the SENSING half of the hot path (test infrastructure only).

Restates, line by line, the reference MATLAB under ``+sensing`` (paths relative to
the reference root).  PARITY UNPINNED: the reference has no tests/golden vectors
and cannot run here (see ``oracle/__init__.py``).  Toolbox behaviour that is not
in the repository is restated from public documentation and tagged
``PARITY-UNPINNED`` per function.

Array conventions: NumPy arrays indexed exactly like the MATLAB arrays
(``grid[k, l, r]``); complex128 / float64 throughout; every *returned index* is
1-based like MATLAB's.
"""
from __future__ import annotations

import math
import numpy as np

LIGHTSPEED = 299792458.0          # physconst('Lightspeed')
BOLTZMANN = 1.380649e-23          # physconst('Boltzmann')
EPS1 = 2.220446049250313e-16      # eps(1)


# ----------------------------------------------------------------------------
# small MATLAB built-ins
# ----------------------------------------------------------------------------
def sind(x):
    """MATLAB ``sind``: exact at multiples of 90 deg (radarParams.m:95,109; music.m:44,82)."""
    x = np.asarray(x, dtype=np.float64)
    r = np.fmod(x, 360.0)
    r = np.where(r > 180.0, r - 360.0, r)
    r = np.where(r < -180.0, r + 360.0, r)
    # reflect into [-90, 90]
    r = np.where(r > 90.0, 180.0 - r, r)
    r = np.where(r < -90.0, -180.0 - r, r)
    out = np.where(np.abs(r) <= 45.0,
                   np.sin(np.deg2rad(r)),
                   np.sign(r) * np.cos(np.deg2rad(90.0 - np.abs(r))))
    return out


def cosd(x):
    """MATLAB ``cosd`` (exact at multiples of 90 deg)."""
    return sind(np.asarray(x, dtype=np.float64) + 90.0)


def nextpow2(x):
    """MATLAB ``nextpow2`` for positive x."""
    return int(math.ceil(math.log2(x))) if x > 1 else 0


def kaiser(n, beta):
    """``kaiser(n,beta)`` (Signal Processing Toolbox, called fft2D.m:135).

    PARITY-UNPINNED: w[k] = I0(beta*sqrt(1-((k-a)/a)^2))/I0(beta), a=(n-1)/2
    (MathWorks doc).  Cross-checked against scipy.signal.windows.kaiser in tests.
    """
    if n == 1:
        return np.ones(1)
    k = np.arange(n, dtype=np.float64)
    a = (n - 1) / 2.0
    return np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - ((k - a) / a) ** 2))) / np.i0(beta)


def findpeaks(y, npeaks):
    """``findpeaks(y,'NPeaks',L,'SortStr','descend')`` (music.m:102, music2D.m:120-121).

    PARITY-UNPINNED: strict local maxima, first sample of a plateau, end points
    excluded, sorted by height descending (stable), truncated to L.
    Returns (peaks, 1-based locations).
    """
    y = np.asarray(y, dtype=np.float64).ravel()
    if npeaks is None or npeaks < 1 or int(npeaks) != npeaks:
        raise ValueError("findpeaks: NPeaks must be a positive integer")
    n = y.size
    locs = []
    i = 1
    while i < n - 1:
        if y[i] > y[i - 1]:
            j = i
            while j < n - 1 and y[j + 1] == y[i]:
                j += 1
            if j < n - 1 and y[j + 1] < y[i]:
                locs.append(i)
            i = j + 1
        else:
            i += 1
    locs = np.asarray(locs, dtype=np.int64)
    if locs.size == 0:
        return np.zeros(0), np.zeros(0, dtype=np.int64)
    order = np.argsort(-y[locs], kind="stable")
    locs = locs[order][: int(npeaks)]
    return y[locs], locs + 1


def unique_stable(x):
    """``unique(x,'stable')`` (fft2D.m:99): exact-equality dedupe keeping first occurrence."""
    x = np.asarray(x).ravel()
    seen = set()
    out = []
    for v in x.tolist():
        if v not in seen:
            seen.add(v)
            out.append(v)
    return np.asarray(out, dtype=x.dtype)


# ----------------------------------------------------------------------------
# OFDM numerology (toolbox: nrOFDMInfo / nrOFDMDemodulate)
# ----------------------------------------------------------------------------
def ofdm_info(nrb, scs_khz):
    """``nrOFDMInfo(NRB,scs)`` as used at cdl.m:54, gNBPhy.m:772.

    PARITY-UNPINNED (5G Toolbox): Nfft = smallest power of two >= 128 with
    12*NRB/Nfft <= 0.85; SampleRate = Nfft*scs; normal-CP lengths from
    TS 38.211 5.3.1 scaled by Nfft/2048.
    """
    nfft = 128
    while 12.0 * nrb / nfft > 0.85:
        nfft *= 2
    mu = int(round(math.log2(scs_khz / 15.0)))
    slots_per_subframe = 2 ** mu
    sym_per_slot = 14
    nsym_sf = sym_per_slot * slots_per_subframe
    # TS 38.211 5.3.1: N_cp = 144*kappa*2^-mu (+16*kappa for l=0 and l=7*2^mu);
    # at the sample rate Nfft*scs, kappa*2^-mu time units == Nfft/2048 samples.
    base = 144 * nfft // 2048
    extra = (16 * nfft // 2048) * (2 ** mu)
    cp = np.full(nsym_sf, base, dtype=np.int64)
    cp[0] += extra
    cp[nsym_sf // 2] += extra
    return {
        "Nfft": nfft,
        "SampleRate": float(nfft * scs_khz * 1e3),
        "CyclicPrefixLengths": cp,
        "SymbolLengths": cp + nfft,
        "SymbolsPerSlot": sym_per_slot,
        "SlotsPerSubframe": slots_per_subframe,
        "SlotsPerFrame": 10 * slots_per_subframe,
        "SymbolsPerSubframe": nsym_sf,
    }


def ofdm_symbol_starts(info, nsym):
    """Sample index of the first CP sample of each of ``nsym`` symbols (periodic per subframe)."""
    lens = info["SymbolLengths"]
    per = lens.size
    starts = np.zeros(nsym, dtype=np.int64)
    acc = 0
    for s in range(nsym):
        starts[s] = acc
        acc += int(lens[s % per])
    return starts


def ofdm_demodulate(nrb, scs_khz, wave, cp_fraction=0.5):
    """``nrOFDMDemodulate(carrier, wave)`` as called at monoStaticSensing.m:16.

    PARITY-UNPINNED (5G Toolbox): whole symbols only; FFT window starts
    ``fix(cp*0.5)`` samples into the CP; the early start is undone by a
    per-subcarrier phase ramp; plain (unnormalised) FFT; central 12*NRB bins.
    Returns grid [nSc x nSym x nAnts].
    """
    info = ofdm_info(nrb, scs_khz)
    nfft = info["Nfft"]
    nsc = 12 * nrb
    wave = np.asarray(wave, dtype=np.complex128)
    if wave.ndim == 1:
        wave = wave[:, None]
    T, nants = wave.shape
    lens = info["SymbolLengths"]
    per = lens.size
    # number of whole symbols
    nsym, acc = 0, 0
    while acc + int(lens[nsym % per]) <= T:
        acc += int(lens[nsym % per])
        nsym += 1
    starts = ofdm_symbol_starts(info, nsym)
    kk = np.arange(nsc) - nsc // 2                     # subcarrier frequency index
    bins = np.mod(kk, nfft)
    grid = np.zeros((nsc, nsym, nants), dtype=np.complex128)
    for s in range(nsym):
        cp = int(info["CyclicPrefixLengths"][s % per])
        off = int(math.floor(cp * cp_fraction))      # fix()
        seg = wave[starts[s] + off: starts[s] + off + nfft, :]
        X = np.fft.fft(seg, axis=0)
        ramp = np.exp(2j * np.pi * kk * (cp - off) / nfft)
        grid[:, s, :] = X[bins, :] * ramp[:, None]
    return grid


def ofdm_modulate(nrb, scs_khz, grid, scale=1.0, windowing=0):
    """``scale * nrOFDMModulate(carrier, grid)`` as called at gNBPhy.m:599 (the waveform the gNB PHY accumulates for
    sensing, gNBPhy.m:604-612).

    PARITY-UNPINNED (5G Toolbox): TS 38.211 5.3.1 CP-OFDM -- subcarrier k of the grid on FFT bin (k - nSc/2) mod Nfft,
    ``ifft`` (1/Nfft), cyclic prefix = copy of the symbol's tail, symbols laid end to end from a subframe boundary.
    ``windowing`` = N > 0 samples: the W-OLA scheme documented for the 'Windowing' argument -- the symbol's cyclic extension
    grows by N samples in front of its prefix, that head is shaped by p[i] = 0.5 (1 - sin(pi (N + 1 - 2 i) / (2 N))), i = 1..N,
    the last N samples of the symbol before it by 1 - p, and the two are added (waveform length unchanged; the first symbol
    has nothing in front of it).  The toolbox's DEFAULT window length (a table over SCS and NRB) is not public, so the default
    here is N = 0, the exact inverse of ``ofdm_demodulate`` above.  Returns [T x nAnts]."""
    info = ofdm_info(nrb, scs_khz)
    nfft = info["Nfft"]
    grid = np.asarray(grid, dtype=np.complex128)
    if grid.ndim == 2:
        grid = grid[:, :, None]
    nsc, nsym, nants = grid.shape
    starts = ofdm_symbol_starts(info, nsym)
    per = info["SymbolLengths"].size
    T = int(starts[-1] + info["SymbolLengths"][(nsym - 1) % per])
    wave = np.zeros((T, nants), dtype=np.complex128)
    bins = np.mod(np.arange(nsc) - nsc // 2, nfft)
    for s in range(nsym):
        spec = np.zeros((nfft, nants), dtype=np.complex128)
        spec[bins, :] = grid[:, s, :]
        x = np.fft.ifft(spec, axis=0) * scale
        cp = int(info["CyclicPrefixLengths"][s % per])
        if windowing > 0 and s > 0:
            N = int(windowing)
            rise = 0.5 * (1.0 - np.sin(np.pi * (N + 1 - 2 * np.arange(1, N + 1)) / (2 * N)))[:, None]
            st = int(starts[s])
            wave[st - N: st, :] = (1.0 - rise) * wave[st - N: st, :] + rise * x[nfft - cp - N: nfft - cp, :]
        wave[starts[s]: starts[s] + cp, :] = x[nfft - cp:, :]
        wave[starts[s] + cp: starts[s] + cp + nfft, :] = x
    return wave


# ----------------------------------------------------------------------------
# a1: sensing.radarParams  (+sensing/radarParams.m:1-146)
# ----------------------------------------------------------------------------
def radar_params(cell, carrier_info, wave_info):
    """``sensing.radarParams`` (+sensing/radarParams.m:12-144), quirks kept:
    Tsri uses ceil(nSc/8) samples as CP (:34-35); ULA steering divides the
    half-wavelength spacing by lambda again (:106-109)."""
    rp = {}
    n_t = int(cell["numTargets"])
    tpos = np.asarray(cell["targetPosition"], dtype=np.float64).reshape(n_t, 3)
    gpos = np.asarray(cell["gNBPosition"], dtype=np.float64).reshape(3)
    coords = tpos.T - gpos[:, None]                                   # :12
    x, y, z = coords
    azi_rad = np.arctan2(y, x)                                         # cart2sph :13
    ele_rad = np.arctan2(z, np.hypot(x, y))
    rng = np.sqrt(x * x + y * y + z * z)
    azi = np.rad2deg(azi_rad)
    ele = np.rad2deg(ele_rad)

    dl_ratio = cell["numDLSlots"] / len(cell["tddPattern"])            # :18
    n_dl_slots = dl_ratio * cell["numSlots"]                            # :19
    n_sc = carrier_info["NRBsDL"] * 12                                  # :20
    n_sym = n_dl_slots * wave_info["SymbolsPerSlot"]                    # :21
    uf = 1
    ut = 1
    n_tx = int(cell["gNBTxAnts"])

    c = LIGHTSPEED
    fc = float(cell["dlCarrierFreq"])
    scs = carrier_info["SubcarrierSpacing"] * 1e3
    lam = c / fc
    fs = float(wave_info["SampleRate"])
    Ts = 1.0 / fs
    Tofdm = 1.0 / scs
    Tcp = Ts * math.ceil(n_sc / 8)                                      # :34
    Tsri = Tofdm + Tcp                                                  # :35

    NF = 10.0 ** (cell["gNBNoiseFigure"] / 10.0)
    Teq = cell["gNBTemperature"] + 290.0 * (NF - 1.0)
    N0 = fs * BOLTZMANN * Teq                                           # :40
    Pt = 10.0 ** ((cell["gNBTxPower"] - 30.0) / 10.0) * math.sqrt(
        wave_info["Nfft"] ** 2 / (carrier_info["NRBsDL"] * 12 * n_tx))  # :41
    Ar = 10.0 ** (cell["gNBRxGain"] / 10.0)
    At = Ar

    rcs = np.asarray(cell["rcs"], dtype=np.float64).reshape(n_t)
    r = rng.copy()
    v = np.asarray(cell["velocity"], dtype=np.float64).reshape(n_t)
    Pr = Pt * At * Ar * (lam ** 2 * rcs) / ((4.0 * np.pi) ** 3 * r ** 4)  # :49
    snr = Pr / N0
    snr_db = 10.0 * np.log10(snr)

    rp.update(fc=fc, fs=fs, Tsri=Tsri, N0=N0, nTxAnts=n_tx, nTargets=n_t,
              range=r, velocity=v, largeScaleFading=np.sqrt(Pr / Pt),
              snrdB=snr_db, txPower=cell["gNBTxPower"], Pfa=cell["Pfa"])

    n_ifft = 2 ** nextpow2(n_sc / uf)                                   # :69
    rp["nIFFT"] = n_ifft
    rp["rRes"] = c / (2.0 * (scs * uf) * n_ifft)                        # :71
    rp["rMax"] = c / (2.0 * (scs * uf))
    n_fft = 2 ** nextpow2(n_sym / ut)                                   # :75
    rp["nFFT"] = n_fft
    rp["vRes"] = lam / (2.0 * (Tsri * ut) * n_fft)                      # :77
    rp["vMax"] = lam / (2.0 * (Tsri * ut))

    ant = cell["gNBSenAntenna"]
    steer = np.zeros((n_tx, n_t), dtype=np.complex128)
    if ant["type"] == "upa":                                            # :84-101
        ax = np.arange(ant["nV"], dtype=np.float64) * ant["dV"]         # 1 x nX
        ay = (np.arange(ant["nH"], dtype=np.float64) * ant["dH"])[:, None]  # nY x 1
        for t in range(n_t):
            a = np.exp(2j * np.pi * sind(ele[t]) *
                       (ax[None, :] * cosd(azi[t]) + ay * sind(azi[t])) / lam)
            steer[:, t] = a.reshape(-1, order="F")
    else:                                                               # :103-116
        ary = np.arange(n_tx, dtype=np.float64) * ant["d"]
        for t in range(n_t):
            steer[:, t] = np.exp(2j * np.pi * ary * sind(azi[t]) / lam)
    rp["antennaType"] = dict(ant)
    rp["azimuthScanScale"] = 360
    rp["elevationScanScale"] = 180
    rp["azimuthScanGranularity"] = 1
    rp["elevationScanGranularity"] = 1
    rp["RxSteeringVec"] = steer
    rp["cfarEstZone"] = np.asarray(cell["detectionArea"], dtype=np.float64).reshape(2, 2)

    idx = np.argsort(-snr_db, kind="stable")                            # :131
    rp["targetRealPos"] = [
        dict(ID=i + 1, Range=r[j], Velocity=v[j], Elevation=ele[j], Azimuth=azi[j], snrdB=snr_db[j])
        for i, j in enumerate(idx)
    ]
    return rp


# ----------------------------------------------------------------------------
# a2: sensing.channelModels.basicRadarChannel
# ----------------------------------------------------------------------------
def basic_radar_channel(tx_wave, rp, los, noise_std_normal):
    """``basicRadarChannel`` (+sensing/+channelModels/basicRadarChannel.m:8-74), literal.

    ``noise_std_normal`` replaces ``randn(size)+1j*randn(size)`` (:68): a complex
    array [T x nAnts] whose real/imag parts are unit-variance normals (MATLAB's
    RNG stream cannot be reproduced, so noise is always an explicit input).
    """
    tx = np.asarray(tx_wave, dtype=np.complex128)
    T, n_tx = tx.shape
    c = LIGHTSPEED
    fc = rp["fc"]
    lam = c / fc
    Ts = 1.0 / rp["fs"]
    n_t = rp["nTargets"]
    delay = 2.0 * np.asarray(rp["range"]) / c                          # :21
    shift = np.ceil(delay / Ts).astype(np.int64)                        # :22
    fd = 2.0 * np.asarray(rp["velocity"]) / lam                         # :25
    t_tx = np.arange(T, dtype=np.float64) * Ts                          # :29
    txm = tx * np.exp(2j * np.pi * fc * t_tx)[:, None]                  # :30-31
    lsf = np.asarray(rp["largeScaleFading"])
    A = np.asarray(rp["RxSteeringVec"], dtype=np.complex128)
    echoes = []
    for i in range(n_t):
        if los[i] == 1:
            s = int(shift[i])
            e = np.concatenate([np.zeros((s, n_tx), dtype=np.complex128), txm[: T - s, :]], axis=0)  # :42
            t_ch = np.arange(e.shape[0], dtype=np.float64) * Ts
            e = e * np.exp(2j * np.pi * fd[i] * t_ch)[:, None]          # :44-45
            e = e * lsf[i]                                              # :48
            e = (e @ A[:, i:i + 1]) @ A[:, i:i + 1].T                   # :51 (plain transpose)
            echoes.append(e)
    if not echoes:
        raise ValueError("basicRadarChannel: no LoS target (reference yields an empty waveform, :59,:64)")
    rx = np.sum(np.stack(echoes, axis=2), axis=2)                        # :64
    n0 = math.sqrt(rp["N0"] / 2.0)                                       # :67
    rx = rx + n0 * np.asarray(noise_std_normal, dtype=np.complex128)     # :68-69
    t_rx = np.arange(rx.shape[0], dtype=np.float64) * Ts
    rx = rx * np.exp(-2j * np.pi * fc * t_rx)[:, None]                   # :73-74
    return rx


# ----------------------------------------------------------------------------
# a3: sensing.monoStaticSensing
# ----------------------------------------------------------------------------
def mono_static_sensing(tx_wave, tx_dimension, carrier_info, rp, los, noise_std_normal):
    """``monoStaticSensing`` (+sensing/monoStaticSensing.m:8-21)."""
    echo = basic_radar_channel(tx_wave, rp, los, noise_std_normal)                    # :13
    grid = ofdm_demodulate(carrier_info["NRBsDL"], carrier_info["SubcarrierSpacing"], echo)  # :16
    if grid.shape[1] < tx_dimension[1]:                                               # :19-21
        pad = np.zeros((grid.shape[0], tx_dimension[1] - grid.shape[1], grid.shape[2]), dtype=grid.dtype)
        grid = np.concatenate([grid, pad], axis=1)
    return grid


# ----------------------------------------------------------------------------
# a4: sensing.detection.cfar2D  /  a6: phased.CFARDetector2D step
# ----------------------------------------------------------------------------
def cfar2d_config(rp, guard=(2, 2), train=(1, 1)):
    """``sensing.detection.cfar2D`` (+sensing/+detection/cfar2D.m:15-33).

    Returns CUTIdx [2 x nCUT] (1-based [row; col], rows fastest) and the CA-CFAR
    settings of the phased.CFARDetector2D object built at :27-33.
    """
    n_ifft, n_fft = rp["nIFFT"], rp["nFFT"]
    rng_grid = np.arange(n_ifft, dtype=np.float64) * rp["rRes"]                   # :17
    dop_grid = np.arange(-n_fft // 2, n_fft // 2, dtype=np.float64) * rp["vRes"]  # :18
    zone = np.asarray(rp["cfarEstZone"], dtype=np.float64)
    rng_idx = [int(np.argmin(np.abs(rng_grid - zone[0, j]))) + 1 for j in range(2)]  # :21
    dop_idx = [int(np.argmin(np.abs(dop_grid - zone[1, j]))) + 1 for j in range(2)]  # :22
    cols = np.arange(dop_idx[0], dop_idx[1] + 1)
    rows = np.arange(rng_idx[0], rng_idx[1] + 1)
    cc, rr = np.meshgrid(cols, rows)                                              # :23
    cut = np.stack([rr.reshape(-1, order="F"), cc.reshape(-1, order="F")], axis=0)  # :24
    return {"CUTIdx": cut.astype(np.int64), "rngIdx": rng_idx, "dopIdx": dop_idx,
            "GuardBandSize": tuple(guard), "TrainingBandSize": tuple(train),
            "Pfa": float(rp["Pfa"]), "Method": "CA"}


def cfar_threshold_factor(n_train, pfa):
    """CA-CFAR 'Auto' threshold factor alpha = N (Pfa^(-1/N) - 1) (PARITY-UNPINNED, toolbox)."""
    return n_train * (pfa ** (-1.0 / n_train) - 1.0)


def cfar2d_detect(rd_response, cfar):
    """``phased.CFARDetector2D`` step, CA, 'Detection index' (fft2D.m:62).

    PARITY-UNPINNED (Phased Array System Toolbox): training band of width T
    around a guard band G around the CUT; N = (2(G+T)+1)^2 - (2G+1)^2 cells;
    detect iff x > alpha * mean(training) (strict); a CUT whose window leaves
    the matrix is an error; output columns follow the supplied CUT order.
    Returns [2 x nDet] 1-based.
    """
    P = np.asarray(rd_response, dtype=np.float64)
    gr, gc = cfar["GuardBandSize"]
    tr, tc = cfar["TrainingBandSize"]
    hr, hc = gr + tr, gc + tc
    n_train = (2 * hr + 1) * (2 * hc + 1) - (2 * gr + 1) * (2 * gc + 1)
    alpha = cfar_threshold_factor(n_train, cfar["Pfa"])
    cut = cfar["CUTIdx"]
    rows, cols = cut[0] - 1, cut[1] - 1
    if (rows.min() - hr < 0 or rows.max() + hr >= P.shape[0]
            or cols.min() - hc < 0 or cols.max() + hc >= P.shape[1]):
        raise ValueError("CFARDetector2D: CUT training window exceeds the input matrix")
    # box sums via summed-area table would reorder additions; sum explicitly (small CUT sets)
    dets = []
    for r0, c0 in zip(rows.tolist(), cols.tolist()):
        outer = P[r0 - hr: r0 + hr + 1, c0 - hc: c0 + hc + 1].sum()
        inner = P[r0 - gr: r0 + gr + 1, c0 - gc: c0 + gc + 1].sum()
        noise = (outer - inner) / n_train
        if P[r0, c0] > alpha * noise:
            dets.append((r0 + 1, c0 + 1))
    if not dets:
        return np.zeros((2, 0), dtype=np.int64)
    return np.asarray(dets, dtype=np.int64).T


def cfar2d_detect_exact(rd_response, cfar):
    """Same detector, but the 24 training cells are summed one by one in float64
    (no outer-inner cancellation).  This is the summation the CUDA kernel
    performs, so it is the bit-exact comparator for a given float32 RDM."""
    P = np.asarray(rd_response, dtype=np.float64)
    gr, gc = cfar["GuardBandSize"]
    tr, tc = cfar["TrainingBandSize"]
    hr, hc = gr + tr, gc + tc
    offs = [(dr, dc) for dc in range(-hc, hc + 1) for dr in range(-hr, hr + 1)
            if abs(dr) > gr or abs(dc) > gc]
    n_train = len(offs)
    alpha = cfar_threshold_factor(n_train, cfar["Pfa"])
    cut = cfar["CUTIdx"]
    rows, cols = cut[0] - 1, cut[1] - 1
    if (rows.min() - hr < 0 or rows.max() + hr >= P.shape[0]
            or cols.min() - hc < 0 or cols.max() + hc >= P.shape[1]):
        raise ValueError("CFARDetector2D: CUT training window exceeds the input matrix")
    acc = np.zeros(rows.size, dtype=np.float64)
    for dr, dc in offs:
        acc = acc + P[rows + dr, cols + dc]
    thr = alpha * (acc / n_train)
    hit = P[rows, cols] > thr
    return np.stack([rows[hit] + 1, cols[hit] + 1], axis=0).astype(np.int64)


# ----------------------------------------------------------------------------
# a5: sensing.estimation.fft2D
# ----------------------------------------------------------------------------
def rdm_2dfft(rp, rx_grid, tx_grid, beta=3.0):
    """The windowed 2D-FFT range-Doppler map of fft2D.m:37-46, literal
    (including the dimension-less ifftshift/fftshift over all three axes and the
    'Doppler' window that is applied along the range axis)."""
    rx = np.asarray(rx_grid, dtype=np.complex128)
    tx = np.asarray(tx_grid, dtype=np.complex128)
    n_sc, n_sym, n_ants = rx.shape
    n_ifft, n_fft = int(rp["nIFFT"]), int(rp["nFFT"])
    chan = rx * np.conj(tx)                                                # :37
    rng_win = kaiser(n_sc, beta)[:, None, None]                            # :40,:146
    dop_win = kaiser(n_ifft, beta)[:, None, None]                          # :147 (length nIFFT!)
    chl = chan * rng_win                                                   # :43
    rng_ifft = np.fft.ifftshift(np.fft.ifft(chl, n_ifft, axis=0) * math.sqrt(n_ifft))   # :44 (all dims)
    rng_ifft = rng_ifft * dop_win                                          # :45
    rdm = np.fft.fftshift(np.fft.fft(rng_ifft, n_fft, axis=1) / math.sqrt(n_fft))       # :46 (all dims)
    return rdm


def antenna_covariance(rx_grid):
    """Ra = X X^H / (nSc nSym), X = reshape(rxGrid,[],nAnts)' (fft2D.m:106-107, music2D.m:57-58)."""
    rx = np.asarray(rx_grid, dtype=np.complex128)
    n_sc, n_sym, n_ants = rx.shape
    X = rx.reshape(n_sc * n_sym, n_ants, order="F").conj().T
    return X @ X.conj().T / (n_sc * n_sym)


def fft2d(rp, cfar, rx_grid, tx_grid, rd_power_override=None, detector=cfar2d_detect):
    """``sensing.estimation.fft2D`` (+sensing/+estimation/fft2D.m:31-115).

    ``rd_power_override`` ([nIFFT x nFFT x nAnts] real) lets a test run the
    detection/estimation tail on a given (e.g. GPU float32) power map.
    Returns dict(rngEst, velEst, aziEst, eleEst, detections=[per-antenna 2xN], rdm).
    """
    n_sc, n_sym, n_ants = np.asarray(rx_grid).shape
    n_fft = int(rp["nFFT"])
    rdm = None
    if rd_power_override is None:
        rdm = rdm_2dfft(rp, rx_grid, tx_grid)
    all_rng, all_vel, det_list, peak_list = [], [], [], []
    for r in range(n_ants):                                                 # :59
        rd = np.abs(rdm[:, :, r]) ** 2 if rd_power_override is None else \
            np.asarray(rd_power_override[:, :, r], dtype=np.float64)        # :61
        det = detector(rd, cfar)                                            # :62
        n_det = det.shape[1]
        peaks = rd[det[0] - 1, det[1] - 1] if n_det else np.zeros(0)        # :74
        rng_idx = det[0] - 1                                                # :77
        vel_idx = det[1] - n_fft / 2 - 1                                    # :78
        rng_est = rng_idx * rp["rRes"]                                      # :81
        vel_est = vel_idx * rp["vRes"]                                      # :82
        idx = np.argsort(-peaks, kind="stable")                             # :89
        all_rng.append(rng_est[idx])
        all_vel.append(vel_est[idx])
        det_list.append(det)
        peak_list.append(peaks)
    all_rng = np.concatenate(all_rng) if all_rng else np.zeros(0)
    all_vel = np.concatenate(all_vel) if all_vel else np.zeros(0)
    u_rng = unique_stable(all_rng)                                          # :99
    u_vel = unique_stable(all_vel)
    Ra = antenna_covariance(rx_grid)                                        # :106-107
    num_dets = u_rng.size                                                   # :110
    L, azi, ele, pm = music_doa(num_dets, rp, Ra)                           # :111
    return {"rngEst": u_rng, "velEst": u_vel, "aziEst": azi, "eleEst": ele,
            "detections": det_list, "peaks": peak_list, "rdm": rdm, "Ra": Ra, "PmusicdB": pm}


# ----------------------------------------------------------------------------
# a7: sensing.estimation.doaEstimation.music
# ----------------------------------------------------------------------------
def determine_num_targets(V):
    """Eigen-gap rule (doaEstimation/music.m:109-125), literal on eig()'s ascending order."""
    V = np.asarray(V, dtype=np.float64)
    delta = -np.diff(V)
    n = delta.size
    half_mean = np.mean(delta[int(math.ceil((n + 1) / 2.0)) - 1:])
    eps_ = 1.0
    return int(np.argmax(delta - (1.0 + eps_) * half_mean)) + 1


def noise_projector(R, L):
    """eig -> sort descending -> Un Un^H (music.m:19-29, music2D.m:77-89).
    Hermitian solver (the projector is basis-invariant)."""
    R = np.asarray(R, dtype=np.complex128)
    w, U = np.linalg.eigh((R + R.conj().T) / 2.0)
    order = np.argsort(-w, kind="stable")
    U = U[:, order]
    Un = U[:, L:]
    return Un @ Un.conj().T, w


def music_doa(num_dets, rp, Ra):
    """``doaEstimation.music`` (+sensing/+estimation/+doaEstimation/music.m:11-104).

    ULA: returns (L, aziEst[deg], eleEst=NaN, PmusicdB[1 x 361]).
    UPA: the reference calls the non-existent tools.find2DPeaks (:69) so peak
    lists are undefined; returns (L, None, None, PmusicdB[eSteps x aSteps]) where
    PmusicdB follows :61-63 literally (Pmusic = -abs(.), ``./max(Pmusic)`` = MATLAB's
    column-wise maximum of a matrix, i.e. every azimuth column is normalised by its own
    smallest magnitude over the elevation scan).
    """
    ant = rp["antennaType"]
    d = 0.5
    Ra = np.asarray(Ra, dtype=np.complex128)
    w_asc = np.linalg.eigvalsh((Ra + Ra.conj().T) / 2.0)                  # :19 (ascending)
    if num_dets is None:
        L = determine_num_targets(w_asc)                                  # :23
    else:
        L = int(num_dets)
    Uann, _ = noise_projector(Ra, L)                                       # :27-29
    if ant["type"] == "upa":                                               # :31-71
        nx, ny = int(ant["nV"]), int(ant["nH"])
        a_gran, e_gran = rp["azimuthScanGranularity"], rp["elevationScanGranularity"]
        a_max, e_max = rp["azimuthScanScale"], rp["elevationScanScale"]
        a_steps = int(math.floor((a_max + 1) / a_gran))
        e_steps = int(math.floor((e_max + 1) / e_gran))
        mm = np.arange(nx, dtype=np.float64)[None, :]
        nn = np.arange(ny, dtype=np.float64)[:, None]
        P = np.zeros((e_steps, a_steps))
        for e in range(e_steps):
            el = e * e_gran - e_max / 2.0
            for a in range(a_steps):
                az = a * a_gran - a_max / 2.0
                aa = np.exp(-2j * np.pi * sind(el) * (mm * d * cosd(az) + nn * d * sind(az)))
                aa = aa.reshape(-1, order="F")
                q = np.vdot(aa, Uann @ aa)
                P[e, a] = np.abs(1.0 / (q + EPS1))        # :56 then abs of :61
        return L, None, None, _upa_normalise_db(P)

    n_ants = int(ant["nV"]) * int(ant["p"])                                # numElements (ula.m)
    gran = rp["azimuthScanGranularity"]
    a_max = rp["azimuthScanScale"]
    a_steps = int(math.floor((a_max + 1) / gran))
    nn = np.arange(n_ants, dtype=np.float64)
    P = np.zeros(a_steps)
    for a in range(a_steps):                                               # :87-91
        ang = a * gran - a_max / 2.0
        aa = np.exp(-2j * np.pi * nn * d * sind(ang))
        q = np.vdot(aa, Uann @ aa)
        P[a] = np.abs(1.0 / (q + EPS1))
    PdB = 20.0 * np.log10(P / P.max())                                     # :94-96
    _, locs = findpeaks(PdB, L)                                            # :102
    azi = (locs - 1) * gran - a_max / 2.0                                  # :103
    ele = np.full(azi.size, np.nan)
    return L, azi, ele, PdB


def _upa_normalise_db(P_abs):
    """music.m:61-63 / mvdrBF.m:43-45 / digitalBF.m:43-45 on a [eSteps x aSteps] matrix:
    ``P = -abs(P); Pn = P./max(P); PdB = mag2db(Pn)``.  ``max`` of a matrix runs along the
    first dimension (column-wise; along the only row when eSteps == 1), implicit expansion
    divides every column by its own maximum; the maximum of the negated magnitudes is minus
    the column's smallest magnitude."""
    P = -np.abs(np.asarray(P_abs, dtype=np.float64))
    mx = P.max(axis=0, keepdims=True) if P.shape[0] > 1 else P.max(axis=1, keepdims=True)
    return 20.0 * np.log10(P / mx)


def _beamformer_doa(num_dets, rp, Ra, kind):
    """Shared body of ``mvdrBF`` (+sensing/+estimation/+doaEstimation/mvdrBF.m:11-92) and ``digitalBF``
    (+sensing/+estimation/+doaEstimation/digitalBF.m:11-93): they differ only in the scanned quantity,
    ``1./(aa'*Ra^-1*aa + eps(1))`` (mvdrBF.m:40,73) vs ``aa'*Ra*aa`` (digitalBF.m:40,73)."""
    ant = rp["antennaType"]
    d = 0.5                                                                 # :12
    Ra = np.asarray(Ra, dtype=np.complex128)
    Q = np.linalg.inv(Ra) if kind == "mvdr" else Ra                         # Ra^-1

    def power(aa):
        q = np.vdot(aa, Q @ aa)
        return 1.0 / (q + EPS1) if kind == "mvdr" else q

    if ant["type"] == "upa":                                                # :14-55
        nx, ny = int(ant["nV"]), int(ant["nH"])
        a_gran, e_gran = rp["azimuthScanGranularity"], rp["elevationScanGranularity"]
        a_max, e_max = rp["azimuthScanScale"], rp["elevationScanScale"]
        a_steps = int(math.floor((a_max + 1) / a_gran))
        e_steps = int(math.floor((e_max + 1) / e_gran))
        mm = np.arange(nx, dtype=np.float64)[None, :]
        nn = np.arange(ny, dtype=np.float64)[:, None]
        P = np.zeros((e_steps, a_steps))
        for e in range(e_steps):
            el = e * e_gran - e_max / 2.0
            for a in range(a_steps):
                az = a * a_gran - a_max / 2.0
                aa = np.exp(-2j * np.pi * sind(el) * (mm * d * cosd(az) + nn * d * sind(az)))
                P[e, a] = np.abs(power(aa.reshape(-1, order="F")))
        # tools.find2DPeaks (:51) does not exist in the reference: peak lists undefined
        return None, None, _upa_normalise_db(P)
    n_ants = int(ant["nV"]) * int(ant["p"])                                 # numElements (ula.m)
    gran = rp["azimuthScanGranularity"]
    a_max = rp["azimuthScanScale"]
    a_steps = int(math.floor((a_max + 1) / gran))
    nn = np.arange(n_ants, dtype=np.float64)
    P = np.zeros(a_steps)
    for a in range(a_steps):                                                # :70-74
        ang = a * gran - a_max / 2.0
        aa = np.exp(-2j * np.pi * nn * d * sind(ang))
        P[a] = np.abs(power(aa))                                            # :77
    PdB = 20.0 * np.log10(P / P.max())                                      # :78-79
    _, locs = findpeaks(PdB, int(num_dets))                                 # :85
    azi = (locs - 1) * gran - a_max / 2.0                                   # :86
    return azi, np.full(azi.size, np.nan), PdB                              # :87


def mvdr_bf(num_dets, rp, Ra):
    """``[aziEst, eleEst] = doaEstimation.mvdrBF(numDets, radarEstParams, Ra)`` (mvdrBF.m:1).
    Returns (aziEst, eleEst, PmvdrdB); UPA: (None, None, PmvdrdB[eSteps x aSteps])."""
    return _beamformer_doa(num_dets, rp, Ra, "mvdr")


def digital_bf(num_dets, rp, Ra):
    """``[aziEst, eleEst] = doaEstimation.digitalBF(numDets, radarEstParams, Ra)`` (digitalBF.m:1).
    Returns (aziEst, eleEst, PdbfdB); UPA: (None, None, PdbfdB[eSteps x aSteps])."""
    return _beamformer_doa(num_dets, rp, Ra, "dbf")


# ----------------------------------------------------------------------------
# a8: sensing.estimation.music2D
# ----------------------------------------------------------------------------
def music2d(rp, bs_params, rx_grid, tx_grid, L_override=None):
    """``sensing.estimation.music2D`` (+sensing/+estimation/music2D.m:33-123)."""
    rx = np.asarray(rx_grid, dtype=np.complex128)
    tx = np.asarray(tx_grid, dtype=np.complex128)
    n_sc, n_sym, n_ants = rx.shape
    scs = bs_params["scs"] * 1e3
    c = LIGHTSPEED
    lam = c / rp["fc"]
    T = rp["Tsri"]
    zone = np.asarray(rp["cfarEstZone"], dtype=np.float64)
    r_max = zone[0, 1]
    v_max = zone[1, 1] * 2
    r_gran = v_gran = 0.5
    r_steps = int(math.floor((r_max + 1) / r_gran))
    v_steps = int(math.floor((v_max + 1) / v_gran))
    Ra = antenna_covariance(rx)                                            # :57-58
    L, azi, ele, pm = music_doa(L_override, rp, Ra)                        # :61
    H = (rx * np.conj(tx))[:, :, 0]                                        # :67-68
    Rr = H @ H.conj().T / n_sym                                            # :71
    Rv = H.T @ np.conj(H) / n_sc                                           # :72
    Urnn, _ = noise_projector(Rr, L)                                       # :77-82
    Uvnn, _ = noise_projector(Rv, L)                                       # :84-89
    nn = np.arange(n_sc, dtype=np.float64)
    mm = np.arange(n_sym, dtype=np.float64)
    Pr = np.zeros(r_steps)
    Pv = np.zeros(v_steps)
    for i in range(r_steps):                                               # :98-102
        ar = np.exp(-2j * np.pi * scs * 2 * (i * r_gran) * nn / c)
        Pr[i] = np.abs(1.0 / np.vdot(ar, Urnn @ ar))
    for i in range(v_steps):                                               # :104-108
        av = np.exp(2j * np.pi * T * 2 * (i * v_gran - v_max / 2.0) * mm / lam)
        Pv[i] = np.abs(1.0 / np.vdot(av, Uvnn @ av))
    PrdB = 20.0 * np.log10(Pr / Pr.max())                                  # :111-113
    PvdB = 20.0 * np.log10(Pv / Pv.max())                                  # :115-117
    _, rl = findpeaks(PrdB, L)                                             # :120
    _, vl = findpeaks(PvdB, L)                                             # :121
    return {"L": L, "aziEst": azi, "eleEst": ele,
            "rngEst": (rl - 1) * r_gran, "velEst": (vl - 1) * v_gran - v_max / 2.0,
            "PrmusicdB": PrdB, "PvmusicdB": PvdB, "Prmusic": Pr, "Pvmusic": Pv,
            "Ra": Ra, "Rr": Rr, "Rv": Rv}
