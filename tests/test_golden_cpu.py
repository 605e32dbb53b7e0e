"""The float64 oracle against the committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py).

The fixtures freeze the oracle on small seeded cases; they are not reference (MATLAB) outputs - see the generator's
header.  This test catches drift of the oracle itself; tests/test_golden_gpu.py checks the CUDA path against the
same numbers."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MG)


def _load(name):
    return np.load(os.path.join(HERE, "golden", name))


def test_sensing_oracle_reproduces_fixture():
    g = _load("sensing_tiny.npz")
    cell, car, wave, rp, cf, grid, txw, noise, rx = MG.sensing_case()
    assert np.allclose(rx[::7, ::5, :], g["echo_grid_sample"], rtol=1e-12, atol=1e-12 * float(g["echo_grid_rms"]))
    ref = MG.OS.fft2d(rp, cf, rx.astype(np.complex64), grid.astype(np.complex64))
    P = np.abs(ref["rdm"]) ** 2
    assert np.allclose(P[::9, ::3, :], g["power_sample"], rtol=1e-10, atol=1e-12 * float(g["power_peak"]))
    assert np.array_equal(np.concatenate([d[0] for d in ref["detections"]]), g["det_rows"])
    assert np.array_equal(np.concatenate([d[1] for d in ref["detections"]]), g["det_cols"])
    assert np.array_equal(ref["rngEst"], g["rngEst"]) and np.array_equal(ref["velEst"], g["velEst"])
    assert np.array_equal(ref["aziEst"], g["aziEst"])
    assert np.allclose(ref["PmusicdB"], g["PmusicdB"], rtol=0, atol=1e-6)
    assert int(g["n_det"].sum()) > 0 and g["rngEst"].size > 0


def test_comm_oracle_reproduces_fixture():
    g = _load("comm_small.npz")
    for tag, (P, panel, nrb, R, seed) in MG.COMM_CASES.items():
        ocfg, re_k, re_l, H, nv = MG.comm_case(P, panel, nrb, R, seed)
        nu = 2
        pm, info = MG.OC.dl_pmi_select(ocfg, re_k, re_l, nu, H, nv)
        assert np.array_equal(pm["i1"], g[f"{tag}_nu{nu}_i1"])
        assert np.array_equal(pm["i2"], g[f"{tag}_nu{nu}_i2"], equal_nan=True)
        assert np.allclose(np.nansum(info["SINRPerSubband"], axis=(0, 1)), g[f"{tag}_nu{nu}_sinr_sb_sum"], rtol=1e-10)
        for q in range(1, min(R, P) + 1):
            Wc = MG.OC.type1_single_panel_codebook(ocfg, q, "ue")
            w = np.arange(1, Wc.size + 1).reshape(Wc.shape, order="F")
            assert np.array_equal(np.array(Wc.shape), g[f"{tag}_cb{q}_shape"])
            assert np.allclose(np.array([np.sum(Wc * w), np.sum(np.abs(Wc) ** 2)]), g[f"{tag}_cb{q}_checksum"], rtol=1e-12)


def test_cdl_oracle_reproduces_fixture():
    g = _load("cdl_c.npz")
    rays = MG.OCDL.build_rays(2, 300e-9, 5.0, (1, 4, 2), (1, 2, 2), True, False, 73)
    assert np.allclose(rays["tau"], g["tau"], rtol=1e-14) and np.allclose(rays["nu"], g["nu"], rtol=1e-12, atol=1e-12)
    assert np.allclose(np.sum(np.abs(rays["g"]) ** 2, axis=(1, 2)), g["g_abs2_sum"], rtol=1e-12)
    assert 0.05 < float(g["H_power"]) < 5.0            # one realisation; element pattern gain included
