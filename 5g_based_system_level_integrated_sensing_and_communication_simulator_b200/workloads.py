"""Synthetic workloads of BASELINE.json's configs (input generation only; NumPy on the host).

Nothing here is on the hot path: these helpers build the ``cellSimuParams``-shaped dictionaries the
reference's scenario/launcher layer would produce (+scenarios/openStreetMapCity.m:11-117,
+simulation/assignCellSimulationParameters.m:27-101) and the transmit grid / waveform that the
gNB PHY accumulates for sensing (gNBPhy.m:599-612: ``txWaveform = signalAmp*nrOFDMModulate(txGrid)``).
"""
from __future__ import annotations

import math

import numpy as np


def ofdm_numerology(nrb: int, scs_khz: float) -> dict:
    """nrOFDMInfo-equivalent numerology (cdl.m:54, gNBPhy.m:772): Nfft, sample rate, CP lengths."""
    nfft = 128
    while 12.0 * nrb / nfft > 0.85:
        nfft *= 2
    mu = int(round(math.log2(scs_khz / 15.0)))
    nsym_sf = 14 * (2 ** mu)
    base = 144 * nfft // 2048
    extra = (16 * nfft // 2048) * (2 ** mu)
    cp = np.full(nsym_sf, base, dtype=np.int64)
    cp[0] += extra
    cp[nsym_sf // 2] += extra
    return {"Nfft": nfft, "SampleRate": float(nfft * scs_khz * 1e3), "CyclicPrefixLengths": cp,
            "SymbolLengths": cp + nfft, "SymbolsPerSlot": 14, "SlotsPerSubframe": 2 ** mu,
            "SlotsPerFrame": 10 * 2 ** mu, "SymbolsPerSubframe": nsym_sf}


def symbol_starts(num: dict, nsym: int) -> np.ndarray:
    lens = num["SymbolLengths"]
    idx = np.arange(nsym) % lens.size
    return np.concatenate([[0], np.cumsum(lens[idx])[:-1]]).astype(np.int64)


def waveform_length(num: dict, nsym: int) -> int:
    lens = num["SymbolLengths"]
    return int(np.sum(lens[np.arange(nsym) % lens.size]))


def qpsk_grid(nsc: int, nsym: int, nants: int, seed: int) -> np.ndarray:
    """Unit-power QPSK on every RE, [nSc x nSym x nAnts] complex128."""
    rng = np.random.default_rng(seed)
    b = rng.integers(0, 4, size=(nsc, nsym, nants))
    return np.exp(1j * (np.pi / 4 + np.pi / 2 * b))


def ofdm_modulate(grid: np.ndarray, nrb: int, scs_khz: float, windowing: int = 0) -> np.ndarray:
    """CP-OFDM modulation (IFFT + cyclic prefix) of [nSc x nSym x nAnts] -> [T x nAnts]; symbol timing starts at a subframe
    boundary like nrOFDMDemodulate assumes.  ``windowing`` = N > 0: the documented W-OLA scheme of nrOFDMModulate's 'Windowing'
    argument (PARITY-UNPINNED): each symbol's cyclic extension grows by N samples in front of its prefix, that head is shaped
    by the rising raised cosine p[i] = 0.5 (1 - sin(pi (N + 1 - 2 i) / (2 N))), i = 1..N, the last N samples of the symbol
    before it by the falling one (1 - p), and the two overlap-add; the first symbol has nothing in front of it."""
    num = ofdm_numerology(nrb, scs_khz)
    nfft = num["Nfft"]
    nsc, nsym, nants = grid.shape
    starts = symbol_starts(num, nsym)
    T = waveform_length(num, nsym)
    wave = np.zeros((T, nants), dtype=np.complex128)
    bins = np.mod(np.arange(nsc) - nsc // 2, nfft)
    spec = np.zeros((nfft, nsym, nants), dtype=np.complex128)
    spec[bins, :, :] = grid
    td = np.fft.ifft(spec, axis=0)
    for s in range(nsym):
        cp = int(num["CyclicPrefixLengths"][s % num["CyclicPrefixLengths"].size])
        st = int(starts[s])
        wave[st: st + cp, :] = td[nfft - cp:, s, :]
        wave[st + cp: st + cp + nfft, :] = td[:, s, :]
    if windowing > 0:
        N = int(windowing)
        rise = 0.5 * (1.0 - np.sin(np.pi * (N + 1 - 2 * np.arange(1, N + 1)) / (2 * N)))[:, None]
        for s in range(1, nsym):
            cp = int(num["CyclicPrefixLengths"][s % num["CyclicPrefixLengths"].size])
            st = int(starts[s])
            head = td[nfft - cp - N: nfft - cp, s, :]
            wave[st - N: st, :] = (1.0 - rise) * wave[st - N: st, :] + rise * head
    return wave


def _target_positions(ranges, azimuths_deg, gnb_pos, height=1.5):
    """Place targets at the given slant ranges / azimuths at a fixed height (target.height, scenario :33)."""
    out = []
    for r, az in zip(ranges, azimuths_deg):
        dz = height - gnb_pos[2]
        ground = math.sqrt(max(r * r - dz * dz, 0.0))
        out.append([gnb_pos[0] + ground * math.cos(math.radians(az)),
                    gnb_pos[1] + ground * math.sin(math.radians(az)), height])
    return np.asarray(out, dtype=np.float64)


CONFIGS = {
    # name: nrb, scs, num_slots, n_ula_v, pol, ranges, velocities, azimuths
    "tiny": dict(nrb=24, scs=15, num_slots=5, nV=2, p=2, ranges=[90.0], vel=[8.0], azi=[20.0]),
    "cfg1": dict(nrb=52, scs=15, num_slots=100, nV=2, p=2, ranges=[120.0], vel=[7.0], azi=[25.0]),
    "cfg2": dict(nrb=273, scs=30, num_slots=20, nV=4, p=2,
                 ranges=[80.0, 150.0, 260.0, 400.0], vel=[-12.0, -3.0, 5.0, 20.0],
                 azi=[-40.0, -10.0, 15.0, 50.0]),
}


def cell_config(name: str) -> tuple[dict, dict, dict]:
    """(cellSimuParams, carrierInfo, waveInfo) for one of BASELINE.json's sensing configs.

    Radio constants follow the shipped scenario (+scenarios/openStreetMapCity.m:44-62):
    3.5 GHz, 46 dBm, 25.5 dB Rx gain, NF 6 dB, 290 K, TDD DDDSU, Pfa 1e-9, zone [50 500; -50 50].
    """
    c = CONFIGS[name]
    num = ofdm_numerology(c["nrb"], c["scs"])
    gnb_pos = np.array([0.0, 0.0, 30.0])
    n_ants = c["nV"] * c["p"]
    cell = {
        "numTargets": len(c["ranges"]),
        "targetPosition": _target_positions(c["ranges"], c["azi"], gnb_pos),
        "gNBPosition": gnb_pos,
        "numDLSlots": 3, "tddPattern": list("DDDSU"), "numSlots": c["num_slots"],
        "gNBTxAnts": n_ants, "dlCarrierFreq": 3.5e9,
        "gNBNoiseFigure": 6.0, "gNBTemperature": 290.0, "gNBTxPower": 46.0, "gNBRxGain": 25.5,
        "rcs": np.ones(len(c["ranges"])), "velocity": np.asarray(c["vel"], dtype=np.float64),
        "gNBSenAntenna": {"type": "ula", "nV": c["nV"], "p": c["p"], "d": 0.5},
        "detectionArea": np.array([[50.0, 500.0], [-50.0, 50.0]]), "Pfa": 1e-9,
        "targetLoSConditions": np.ones(len(c["ranges"]), dtype=np.int64),
    }
    carrier = {"SubcarrierSpacing": c["scs"], "NRBsDL": c["nrb"]}
    wave = {"SampleRate": num["SampleRate"], "SymbolsPerSlot": 14, "Nfft": num["Nfft"],
            "SlotsPerSubframe": num["SlotsPerSubframe"], "SymbolLengths": num["SymbolLengths"],
            "CyclicPrefixLengths": num["CyclicPrefixLengths"]}
    return cell, carrier, wave


def sensing_tx(name: str, seed: int = 1):
    """senTxGrid [nSc x nSym x nTx] (unit QPSK) and senTxWave [T x nTx] (= signalAmp * OFDM-mod)
    for a config, the way gNBPhy accumulates them over the DL slots (gNBPhy.m:591-612)."""
    cell, carrier, wave = cell_config(name)
    c = CONFIGS[name]
    n_dl_slots = int(round(3 / 5 * c["num_slots"]))
    nsym = 14 * n_dl_slots
    nsc = 12 * c["nrb"]
    n_tx = cell["gNBTxAnts"]
    grid = qpsk_grid(nsc, nsym, n_tx, seed)
    amp = 10.0 ** ((cell["gNBTxPower"] - 30.0) / 20.0) * math.sqrt(wave["Nfft"] ** 2 / (nsc * n_tx))
    tx_wave = amp * ofdm_modulate(grid, c["nrb"], c["scs"])
    return grid, tx_wave


def std_normal_complex(shape, seed: int) -> np.ndarray:
    """randn(size)+1j*randn(size) stand-in (basicRadarChannel.m:68) with a fixed NumPy seed."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 5: the openStreetMapCity scenario scaled to 19 gNBs / 100 UEs / 20 moving targets
# ---------------------------------------------------------------------------------------------------------------------
RADIO = {
    # the shipped scenario's radio (+scenarios/openStreetMapCity.m:44-62): 100 MHz @ 30 kHz -> 273 PRB, 8-element dual-polarised
    # ULA (16 Tx), 2-antenna UEs, row-5 CSI-RS (4 ports, setupCSIRS.m:8), 2-port SRS (setupSRS.m:18), one frame = 20 slots
    "shipped": dict(nrb=273, scs=30, nV=8, p=2, ue_ants=2, csirs_ports=4, panel=(2, 1), subband=16, num_slots=20),
    # same structure at 24 PRB / 15 kHz / 4 Tx for tests whose oracle has to finish in seconds
    "small": dict(nrb=24, scs=15, nV=2, p=2, ue_ants=2, csirs_ports=4, panel=(2, 1), subband=4, num_slots=10),
    # BASELINE config 3: 64-element gNB array virtualised onto 32 CSI-RS ports (4,4) -- the largest Type-I port count
    # (dlPMISelect.m:623-626) --, 4-antenna UEs, 273 PRB; the 16-element Rx sub-array serves the 2-port SRS
    "cfg3": dict(nrb=273, scs=30, nV=8, p=2, ue_ants=4, csirs_ports=32, panel=(4, 4), subband=16, num_slots=20),
}


def hex_gnb_positions(num_cells: int, radius: float, roi, height: float) -> np.ndarray:
    """gNB sites of the reference's hexagonal layout (+networkTopology/+wraparound/generateWrapAround.m:94-150,
    getgNBPositions): lattice points (i*1.5r, j*r*sqrt(3)/2) with i + j even inside the region of interest, in the
    reference's loop order (i outer, j inner), truncated to ``num_cells``.  Returns [n x 3]."""
    xw, yw = roi[0] / 2.0, roi[1] / 2.0
    dx, dy = 1.5 * radius, radius * math.sqrt(3.0) / 2.0
    mx, my = math.ceil(xw / dx), math.ceil(yw / dy)
    out = []
    for i in range(-mx, mx + 1):
        for j in range(-my, my + 1):
            if (i + j) % 2 == 0 and abs(i * dx) <= xw and abs(j * dy) <= yw:
                out.append([i * dx, j * dy, height])
    if len(out) < num_cells:
        raise ValueError(f"the region of interest holds only {len(out)} sites of radius {radius}")
    return np.asarray(out[:num_cells], dtype=np.float64)


def scenario_cfg5(seed: int = 5, n_gnb: int = 19, n_ue: int = 100, n_targets: int = 20, radius: float = 80.0,
                  radio: str = "shipped", roi=(500.0, 760.0)) -> dict:
    """BASELINE config 5: ``n_gnb`` hexagonally placed gNBs inside the cached OSM city's extent (the shipped scenario places
    its single gNB inside the same city, +scenarios/openStreetMapCity.m:20-43), ``n_ue`` UEs attached to the nearest site,
    ``n_targets`` targets moving on straight lines, each attached to the gNB it was dropped around (attachedTargets is static
    in the reference, openStreetMapCity.m:41).  Deterministic in ``seed``; nothing depends on how cells are later sharded."""
    rng = np.random.default_rng(seed)
    gnb = hex_gnb_positions(n_gnb, radius, roi, 25.0)
    ue = np.column_stack([rng.uniform(-roi[0] / 2, roi[0] / 2, n_ue), rng.uniform(-roi[1] / 2, roi[1] / 2, n_ue), np.full(n_ue, 1.5)])
    ue_cell = np.argmin(((ue[:, None, :2] - gnb[None, :, :2]) ** 2).sum(axis=2), axis=1)
    tgt_cell = rng.integers(0, n_gnb, n_targets)
    rr, az = rng.uniform(60.0, 180.0, n_targets), np.deg2rad(rng.uniform(-60.0, 60.0, n_targets))
    tgt = gnb[tgt_cell] + np.column_stack([rr * np.cos(az), rr * np.sin(az), np.zeros(n_targets)])
    tgt[:, 2] = 1.5
    speed, head = rng.uniform(2.0, 15.0, n_targets), rng.uniform(0, 2 * np.pi, n_targets)
    vel = np.column_stack([speed * np.cos(head), speed * np.sin(head), np.zeros(n_targets)])
    load = rng.uniform(0.5, 1.0, (n_gnb, 64))                 # DL load of cell c in frame f (fraction of the Tx power radiated)
    return {"gnb": gnb, "ue": ue, "ue_cell": ue_cell, "target0": tgt, "target_vel": vel, "target_cell": tgt_cell,
            "rcs": np.ones(n_targets), "radio": radio, "seed": seed, "load": load, "frame_time": 10e-3,
            "fc": 3.5e9, "txPower": 46.0, "rxGainUE": 0.0, "noiseFigureUE": 9.0, "scenario": "UMa"}


def cfg5_target_state(scn: dict, frame: int):
    """Positions of every target at the start of CPI ``frame`` and their radial velocities towards the attached gNB
    (positive = receding, the sign convention of radarParams.velocity, basicRadarChannel.m:25)."""
    pos = scn["target0"] + scn["target_vel"] * (frame * scn["frame_time"])
    los_vec = pos - scn["gnb"][scn["target_cell"]]
    unit = los_vec / np.linalg.norm(los_vec, axis=1, keepdims=True)
    return pos, (scn["target_vel"] * unit).sum(axis=1)


def cfg5_cell_params(scn: dict, cell: int, frame: int, ue_los=None, tgt_los=None) -> tuple[dict, dict, dict]:
    """(cellSimuParams, carrierInfo, waveInfo) of one cell for CPI ``frame``: the fields
    +simulation/assignCellSimulationParameters.m:27-101 flattens for cellSimulation, plus the UE list of the cell.
    ``ue_los`` / ``tgt_los``: LoS flags of ALL UEs / targets of the scenario towards their own gNB (checkLoS,
    networkSimulation.m:138,154); default all LoS."""
    r = RADIO[scn["radio"]]
    num = ofdm_numerology(r["nrb"], r["scs"])
    pos, radial = cfg5_target_state(scn, frame)
    ti = np.flatnonzero(scn["target_cell"] == cell)
    ui = np.flatnonzero(scn["ue_cell"] == cell)
    n_ants = r["nV"] * r["p"]
    cellp = {
        "cellID": int(cell), "frame": int(frame),
        "numTargets": int(ti.size), "targetIDs": ti, "targetPosition": pos[ti].reshape(-1, 3),
        "gNBPosition": scn["gnb"][cell],
        "numDLSlots": 3, "tddPattern": list("DDDSU"), "numSlots": r["num_slots"],
        "gNBTxAnts": n_ants, "dlCarrierFreq": scn["fc"],
        "gNBNoiseFigure": 6.0, "gNBTemperature": 290.0, "gNBTxPower": scn["txPower"], "gNBRxGain": 25.5,
        "rcs": scn["rcs"][ti], "velocity": radial[ti],
        "gNBSenAntenna": {"type": "ula", "nV": r["nV"], "p": r["p"], "d": 0.5},
        "detectionArea": np.array([[50.0, 500.0], [-50.0, 50.0]]), "Pfa": 1e-9,
        "targetLoSConditions": (np.ones(ti.size, dtype=np.int64) if tgt_los is None else np.asarray(tgt_los)[ti].astype(np.int64)),
        "ueIDs": ui, "uePosition": scn["ue"][ui].reshape(-1, 3),
        "ueLoSConditions": (np.ones(ui.size, dtype=np.int64) if ue_los is None else np.asarray(ue_los)[ui].astype(np.int64)),
        "ueTxAnts": r["ue_ants"], "radio": scn["radio"], "txLoad": float(scn["load"][cell, frame % scn["load"].shape[1]]),
    }
    carrier = {"SubcarrierSpacing": r["scs"], "NRBsDL": r["nrb"]}
    wave = {"SampleRate": num["SampleRate"], "SymbolsPerSlot": 14, "Nfft": num["Nfft"],
            "SlotsPerSubframe": num["SlotsPerSubframe"], "SymbolLengths": num["SymbolLengths"],
            "CyclicPrefixLengths": num["CyclicPrefixLengths"]}
    cellp["carrierInfo"], cellp["waveInfo"] = carrier, wave
    return cellp, carrier, wave


def cfg5_sensing_grid(scn: dict, cell: int) -> np.ndarray:
    """senTxGrid of a cell's CPI: unit QPSK on every RE of the DL symbols of a frame, seeded by (scenario, cell); the same
    payload is transmitted every frame (what changes from CPI to CPI are the targets, the noise and the channels)."""
    r = RADIO[scn["radio"]]
    nsym = 14 * int(round(3 / 5 * r["num_slots"]))
    return qpsk_grid(12 * r["nrb"], nsym, r["nV"] * r["p"], 1_000_003 * scn["seed"] + 1009 * cell)


def scenario_cfg3(seed: int = 3, ue_per_cell: int = 4, radius: float = 500.0) -> dict:
    """BASELINE config 3: 7 hexagonal cells (centre + first ring of the reference's lattice, generateWrapAround.m:94-150,
    radius 500 m), ``ue_per_cell`` UEs dropped uniformly in each cell, 32-port Type-I reports + the per-PRB SINR grid; no radar
    targets (the configuration is about the COMM half).  Same structure as scenario_cfg5."""
    rng = np.random.default_rng(seed)
    dx, dy = 1.5 * radius, radius * math.sqrt(3.0) / 2.0
    sites = [(0, 0)] + [(i, j) for i, j in ((0, 2), (1, 1), (1, -1), (0, -2), (-1, -1), (-1, 1))]
    gnb = np.array([[i * dx, j * dy, 25.0] for i, j in sites])
    n_ue = 7 * ue_per_cell
    ue_cell = np.repeat(np.arange(7), ue_per_cell)
    rr, az = radius * 0.9 * np.sqrt(rng.uniform(0.01, 1.0, n_ue)), rng.uniform(0, 2 * np.pi, n_ue)
    ue = gnb[ue_cell] + np.column_stack([rr * np.cos(az), rr * np.sin(az), np.zeros(n_ue)])
    ue[:, 2] = 1.5
    return {"gnb": gnb, "ue": ue, "ue_cell": ue_cell, "target0": np.zeros((0, 3)), "target_vel": np.zeros((0, 3)),
            "target_cell": np.zeros(0, dtype=int), "rcs": np.zeros(0), "radio": "cfg3", "seed": seed,
            "load": rng.uniform(0.5, 1.0, (7, 64)), "frame_time": 10e-3, "fc": 3.5e9, "txPower": 46.0, "rxGainUE": 0.0,
            "noiseFigureUE": 9.0, "scenario": "UMa", "sinr_grid": True}
