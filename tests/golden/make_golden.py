"""Regenerate the fixtures under tests/golden/ (run from the repo root: ``python tests/golden/make_golden.py``).

The reference is MATLAB + 5G/Phased-Array toolboxes and cannot be executed in this image, so these vectors are NOT
reference outputs: they freeze the float64 oracle (oracle/, itself PARITY-UNPINNED) on small seeded cases.  They
serve two purposes: (1) `tests/test_golden_cpu.py` detects drift of the oracle itself (NumPy/SciPy upgrades, edits);
(2) `tests/test_golden_gpu.py` checks the CUDA path against committed numbers, independent of a live oracle run.
Inputs are regenerated from seeds by `workloads.py` (QPSK grids, OFDM waveforms) and NumPy `default_rng`.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import geometry as OG  # noqa: E402
from oracle import cdl as OCDL  # noqa: E402
from oracle import chest as OCH  # noqa: E402
from oracle import comm as OC  # noqa: E402
from oracle import sensing as OS  # noqa: E402

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
HERE = os.path.dirname(os.path.abspath(__file__))


def sensing_case(name="tiny", seed=1):
    W = importlib.import_module(PKG + ".workloads")
    cell, car, wave = W.cell_config(name)
    rp = OS.radar_params(cell, car, wave)
    grid, txw = W.sensing_tx(name, seed)
    noise = W.std_normal_complex(txw.shape, seed + 1)
    rx = OS.mono_static_sensing(txw, grid.shape, car, rp, cell["targetLoSConditions"], noise)
    cf = OS.cfar2d_config(rp)
    return cell, car, wave, rp, cf, grid, txw, noise, rx


def make_sensing():
    cell, car, wave, rp, cf, grid, txw, noise, rx = sensing_case()
    rx32, tx32 = rx.astype(np.complex64), grid.astype(np.complex64)
    ref = OS.fft2d(rp, cf, rx32, tx32)
    P = np.abs(ref["rdm"]) ** 2
    det = ref["detections"]
    out = {
        "echo_grid_sample": rx[::7, ::5, :],                      # float64 echo grid, decimated (monoStaticSensing)
        "echo_grid_rms": np.sqrt(np.mean(np.abs(rx) ** 2)),
        "power_peak": P.max(), "power_sum": P.sum(axis=(0, 1)),   # per-antenna checksums of |RDM|^2
        "power_sample": P[::9, ::3, :],
        "n_det": np.array([d.shape[1] for d in det]),
        "det_rows": np.concatenate([d[0] for d in det]), "det_cols": np.concatenate([d[1] for d in det]),
        "rngEst": ref["rngEst"], "velEst": ref["velEst"], "aziEst": ref["aziEst"],
        "PmusicdB": ref["PmusicdB"],
        "cfar_alpha": OS.cfar_threshold_factor(24, rp["Pfa"]),
        "nIFFT": rp["nIFFT"], "nFFT": rp["nFFT"],
    }
    np.savez_compressed(os.path.join(HERE, "sensing_tiny.npz"), **out)


def comm_case(n_ports, panel, nrb, n_rx, seed, sb=4):
    rng = np.random.default_rng(seed)
    K = 12 * nrb
    H = ((rng.standard_normal((K, 14, n_rx, n_ports)) + 1j * rng.standard_normal((K, 14, n_rx, n_ports))) / np.sqrt(2)).astype(np.complex64)
    H = (H + np.roll(H, 1, axis=0) + np.roll(H, 2, axis=0)).astype(np.complex64)
    ocfg = OC.report_config(n_ports, panel, nrb, 0, 1, "Subband", "Subband", sb)
    re_k, re_l = OC.csirs_first_port_res(nrb, 1, 0)
    return ocfg, re_k, re_l, H, 0.1


COMM_CASES = {"p4": (4, (2, 1), 24, 4, 11), "p8": (8, (2, 2), 24, 8, 12)}


def make_comm():
    out = {}
    table = np.array([-5.8, -4.0, -2.0, 0.1, 2.1, 4.0, 6.0, 7.9, 9.8, 11.7, 13.7, 15.6, 17.5, 19.5, 21.4])  # any monotone table
    out["cqi_table"] = table
    for tag, (P, panel, nrb, R, seed) in COMM_CASES.items():
        ocfg, re_k, re_l, H, nv = comm_case(P, panel, nrb, R, seed)
        for nu in (1, 2, min(R, P)):
            pm, info = OC.dl_pmi_select(ocfg, re_k, re_l, nu, H, nv)
            out[f"{tag}_nu{nu}_i1"] = pm["i1"]
            out[f"{tag}_nu{nu}_i2"] = pm["i2"]
            out[f"{tag}_nu{nu}_sinr_sb_sum"] = np.nansum(info["SINRPerSubband"], axis=(0, 1))
            out[f"{tag}_nu{nu}_sinr_re_sample"] = info["SINRPerRE"][::5, :, ...].reshape(-1)[::97]
        ri, pm = OC.ri_select(ocfg, re_k, re_l, H, nv)
        out[f"{tag}_ri"] = ri
        cqi, pmc, _, _ = OC.cqi_select(ocfg, re_k, re_l, int(ri), H, nv, table)
        out[f"{tag}_cqi"] = cqi
    # UL TPMI selection
    rng = np.random.default_rng(21)
    K = 12 * 24
    hest = np.zeros((K, 14, 8, 4), dtype=np.complex64)
    sc = np.arange(1, K, 4)
    hest[sc, 13] = ((rng.standard_normal((sc.size, 8, 4)) + 1j * rng.standard_normal((sc.size, 8, 4))) / np.sqrt(2)).astype(np.complex64)
    pmi, sinr, idx = OC.pmi_select(2, hest, 0.05, 4)
    out["ul_pmi"], out["ul_sinr"], out["ul_idx"] = pmi, sinr, idx
    # codebook checksums (Type-I single panel, PUSCH)
    for tag, (P, panel, nrb, R, seed) in COMM_CASES.items():
        ocfg = OC.report_config(P, panel, nrb, 0, 1, "Subband", "Subband", 4)
        for nu in range(1, min(R, P) + 1):
            Wc = OC.type1_single_panel_codebook(ocfg, nu, "ue")
            w = np.arange(1, Wc.size + 1).reshape(Wc.shape, order="F")
            out[f"{tag}_cb{nu}_shape"] = np.array(Wc.shape)
            out[f"{tag}_cb{nu}_checksum"] = np.array([np.sum(Wc * w), np.sum(np.abs(Wc) ** 2)])
    np.savez_compressed(os.path.join(HERE, "comm_small.npz"), **out)


def make_cdl():
    rays = OCDL.build_rays(2, 300e-9, 5.0, (1, 4, 2), (1, 2, 2), True, False, 73)   # 2 = CDL-C
    H = OCDL.frequency_response(rays, 24 * 12, 30e3, np.arange(14) * 35.7e-6)
    out = {"tau": rays["tau"], "power": rays["power"], "nu": rays["nu"], "g_abs2_sum": np.sum(np.abs(rays["g"]) ** 2, axis=(1, 2)),
           "H_sample": H[::17, ::3], "H_power": np.mean(np.abs(H) ** 2)}
    np.savez_compressed(os.path.join(HERE, "cdl_c.npz"), **out)


def chest_case(seed=21, nrb=24, R=2, snr_db=25.0):
    """4-port CSI-RS row 5 (setupCSIRS.m:8-10) through a smooth frequency-selective channel + AWGN -> rxGrid."""
    K, L, P = 12 * nrb, 14, 4
    ind, sym, cdm = OCH.csirs_row5_layout(nrb, 1, 0, seed=seed)
    rng = np.random.default_rng(seed)
    taps = (rng.standard_normal((3, R, P)) + 1j * rng.standard_normal((3, R, P))) * np.array([1.0, 0.5, 0.25])[:, None, None]
    k = np.arange(K)[:, None, None, None]
    H = sum(taps[t][None, None] * np.exp(-2j * np.pi * k * t * 3 / 1024.0) for t in range(3)) * np.ones((1, L, 1, 1))
    sig = 10 ** (-snr_db / 20)
    noise = sig / np.sqrt(2) * (rng.standard_normal((K, L, R)) + 1j * rng.standard_normal((K, L, R)))
    rx = OCH.apply_channel(H, ind, sym, noise)
    return K, L, R, P, ind, sym, cdm, H, rx, sig ** 2


def make_chest():
    K, L, R, P, ind, sym, cdm, H, rx, nv = chest_case()
    He, nve = OCH.channel_estimate(rx, ind, sym, P, cdm)
    Ha, nva = OCH.channel_estimate(rx, ind, sym, P, cdm, (3, 1))
    np.savez_compressed(os.path.join(HERE, "chest_small.npz"), Hest=He.astype(np.complex64), nVar=nve,
                        Hest_avg=Ha.astype(np.complex64), nVar_avg=nva, nVar_true=nv)


def city_links(fp_flat, fp_off, heights, seed=5, n=1500):
    """Seeded link set over the city's bounding box: users / targets at street level and on roofs, gNB-like antennas."""
    rng = np.random.default_rng(seed)
    lo, hi = fp_flat.min(axis=1), fp_flat.max(axis=1)
    ue = np.column_stack([rng.uniform(lo[0] - 20, hi[0] + 20, n), rng.uniform(lo[1] - 20, hi[1] + 20, n),
                          rng.choice([1.5, 1.5, 1.5, 12.0, 40.0], n)])
    ant = np.column_stack([rng.uniform(lo[0], hi[0], n), rng.uniform(lo[1], hi[1], n), rng.uniform(20.0, 45.0, n)])
    # degenerate / boundary cases: a user exactly on a building corner (winding "invalid" rule), a link parallel to the
    # ceilings (division by zero -> NaN -> not blocked by that wall), coincident user and antenna
    ue[0] = [fp_flat[0, 0], fp_flat[1, 0], 0.0]
    ue[1, 2] = ant[1, 2] = heights[0]
    ue[2] = ant[2]
    return ue, ant


def make_city():
    """Condense the reference's cached city (its only data fixture: dataFiles/blockages/OSM_city.json, the file
    city.loadCityFromFile reads, city.m:116-143) and freeze the oracle's LoS decisions for a seeded link set."""
    import json
    src = "/root/reference/dataFiles/blockages/OSM_city.json"
    if not os.path.exists(src):
        print("skipping osm_city.npz: the reference tree is not mounted")
        return
    b = json.load(open(src))["buildings"]
    fps = [np.asarray(x["floorPlan"], dtype=np.float64) for x in b]
    heights = np.array([float(x["height"]) for x in b])
    fp_off = np.zeros(len(fps) + 1, dtype=np.int64)
    fp_off[1:] = np.cumsum([f.shape[1] for f in fps])
    fp_flat = np.concatenate(fps, axis=1)
    ue, ant = city_links(fp_flat, fp_off, heights)
    los = OG.check_los(list(zip(fps, heights)), ue, ant)
    los_one = OG.check_los(list(zip(fps, heights)), ue, ant[7:8])
    np.savez_compressed(os.path.join(HERE, "osm_city.npz"), fp_flat=fp_flat, fp_off=fp_off, heights=heights, ue=ue, ant=ant,
                        los=los, los_one_antenna=los_one)
    print("osm_city: %d buildings, %d links, %.1f %% LoS" % (len(fps), los.size, 100.0 * los.mean()))


if __name__ == "__main__":
    make_city()
    make_chest()
    make_sensing()
    make_comm()
    make_cdl()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
