"""GPU parity: batched LoS / blockage test (csrc/los.cu) vs the float64 oracle -- decisions bit-exact."""
import importlib
import os

import numpy as np
import pytest

from oracle import geometry as G

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def blk(gpu):
    return importlib.import_module(PKG + ".networkTopology.blockages")


def _fixture():
    z = np.load(os.path.join(HERE, "golden", "osm_city.npz"))
    off = z["fp_off"]
    return z, [(z["fp_flat"][:, off[i]:off[i + 1]], float(z["heights"][i])) for i in range(off.size - 1)]


def test_osm_city_los_matches_golden_and_oracle(blk):
    """The reference's cached city (81 buildings, 666 walls): 1500 seeded links incl. a user on a building corner, a link
    parallel to the ceilings and a zero-length link; element-wise pairs and the one-antenna form."""
    z, buildings = _fixture()
    c = blk.city(buildings)
    assert c.nWalls == sum(b[0].shape[1] for b in buildings)   # nCorner-1 side walls + 1 ceiling per building
    los = c.checkLoS(z["ue"], z["ant"])
    assert np.array_equal(los, z["los"])
    assert np.array_equal(c.checkLoS(z["ue"], z["ant"][7]), z["los_one_antenna"])
    sel = np.r_[0:40]
    assert np.array_equal(los[sel], G.check_los(buildings, z["ue"][sel], z["ant"][sel]))   # live oracle on a slice
    c.close()


def test_random_city_and_box_known_answers(blk):
    rng = np.random.default_rng(8)
    buildings = []
    for _ in range(12):   # convex and L-shaped floor plans
        x0, y0 = rng.uniform(-100, 100, 2)
        w, d = rng.uniform(8, 30, 2)
        if rng.random() < 0.5:
            fp = np.array([[x0, x0 + w, x0 + w, x0, x0], [y0, y0, y0 + d, y0 + d, y0]])
        else:
            fp = np.array([[x0, x0 + w, x0 + w, x0 + w / 2, x0 + w / 2, x0, x0],
                           [y0, y0, y0 + d / 2, y0 + d / 2, y0 + d, y0 + d, y0]])
        buildings.append((fp, float(rng.uniform(5, 40))))
    c = blk.city(buildings)
    n = 4000
    ue = np.column_stack([rng.uniform(-150, 150, n), rng.uniform(-150, 150, n), rng.uniform(0, 50, n)])
    ant = np.column_stack([rng.uniform(-150, 150, n), rng.uniform(-150, 150, n), rng.uniform(10, 60, n)])
    got = c.checkLoS(ue, ant)
    ref = G.check_los(buildings, ue, ant)
    assert np.array_equal(got, ref)
    assert 0.02 < got.mean() < 0.98
    c.close()
    box = blk.city([(np.array([[0, 10, 10, 0, 0], [0, 0, 10, 10, 0]], float), 20.0)])
    ue = np.array([[-5, 5, 1.5], [-5, 5, 30.0], [-5, -5, 1.5], [5, 5, 1.5], [60, 5, 10.0]])
    assert box.checkLoS(ue, [30.0, 5.0, 10.0]).tolist() == [False, True, True, False, False]
    box.close()


def test_city_argument_errors(blk):
    _lib = importlib.import_module(PKG + "._lib")
    with pytest.raises(_lib.IsacError):
        blk.city([])
    with pytest.raises(_lib.IsacError):   # collinear floor plan -> degenerate ceiling
        blk.city([(np.array([[0.0, 1.0, 2.0, 0.0], [0.0, 0.0, 0.0, 0.0]]), 5.0)])
    c = blk.city([(np.array([[0, 10, 10, 0, 0], [0, 0, 10, 10, 0]], float), 20.0)])
    with pytest.raises(_lib.IsacError):
        c.checkLoS(np.zeros((4, 3)), np.zeros((3, 3)))
    c.close()
