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


def ofdm_modulate(grid: np.ndarray, nrb: int, scs_khz: float) -> np.ndarray:
    """Plain CP-OFDM modulation (IFFT + cyclic prefix, no windowing) of [nSc x nSym x nAnts]
    -> [T x nAnts]; symbol timing starts at a subframe boundary like nrOFDMDemodulate assumes."""
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
