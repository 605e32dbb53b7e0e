"""CPU: the LoS / blockage oracle (oracle/geometry.py) -- analytic known answers and the frozen OSM-city fixture."""
import os

import numpy as np

from oracle import geometry as G

HERE = os.path.dirname(os.path.abspath(__file__))


def _fixture():
    z = np.load(os.path.join(HERE, "golden", "osm_city.npz"))
    off = z["fp_off"]
    buildings = [(z["fp_flat"][:, off[i]:off[i + 1]], float(z["heights"][i])) for i in range(off.size - 1)]
    return z, buildings


def test_box_building_known_answers():
    """One 10 x 10 x 20 m box (building.m:61-73 walls + ceiling).  Links through the box are blocked, links over or beside
    it are not; a user inside is blocked; the reference projects onto the INFINITE line through user and antenna
    (wallBlockage.m:121-123), so a box behind the antenna on that line blocks as well."""
    box = (np.array([[0, 10, 10, 0, 0], [0, 0, 10, 10, 0]], float), 20.0)
    ant = np.array([[30.0, 5.0, 10.0]])
    ue = np.array([[-5, 5, 1.5],       # straight through the box
                   [-5, 5, 30.0],      # passes over the roof (z = 27 m above the near wall)
                   [-5, -5, 1.5],      # passes beside it
                   [5, 5, 1.5],        # inside the building
                   [60, 5, 10.0]])     # box BEHIND the antenna on the same line: still "blocked" in the reference
    assert G.check_los([box], ue, ant).tolist() == [False, True, True, False, False]
    # plane of a side wall: normal +-x or +-y, distance = its coordinate
    w = G.building_walls(*box)
    assert len(w) == 5 and w[-1].shape == (3, 5)
    n, d = G.wall_plane(w[1])          # edge (10,0)-(10,10): the plane x = 10
    assert np.allclose(np.abs(n), [1, 0, 0]) and abs(abs(d) - 10.0) < 1e-12
    # winding number: 2*pi inside, 0 outside, 1 (flag) exactly on a corner
    c = w[-1]
    nz = np.array([0.0, 0.0, 1.0])
    wn = G.winding_number(c, nz, np.array([[5.0, 5.0, 20.0], [20.0, 20.0, 20.0], [0.0, 0.0, 20.0]]).T)
    assert abs(wn[0] - 2 * np.pi) < 1e-12 and wn[1] < 1e-12 and wn[2] == 1.0


def test_link_parallel_to_a_wall_is_not_blocked_by_it():
    """n'(ue-ant) = 0 -> division by zero -> NaN winding number -> `NaN > 0.1` is false (wallBlockage.m:123-127)."""
    c = np.array([[0, 10, 10, 0], [0, 0, 0, 0], [0, 0, 5, 5]], float)   # wall in the plane y = 0
    n, d = G.wall_plane(c)
    ue, ant = np.array([[2.0, 0.0, 1.0]]).T, np.array([[8.0, 0.0, 2.0]]).T     # link inside that plane
    assert not G.wall_check_blockage(c, n, d, ue, ant)[0]


def test_osm_city_fixture_matches_the_oracle():
    """tests/golden/osm_city.npz (condensed from the reference's dataFiles/blockages/OSM_city.json by make_golden.py)
    freezes the oracle's decisions: guards against drift of the restatement."""
    z, buildings = _fixture()
    assert len(buildings) == 81
    sel = np.r_[0:60, 700:760]
    los = G.check_los(buildings, z["ue"][sel], z["ant"][sel])
    assert np.array_equal(los, z["los"][sel])
    los1 = G.check_los(buildings, z["ue"][sel], z["ant"][7:8])
    assert np.array_equal(los1, z["los_one_antenna"][sel])
    assert 0.05 < z["los"].mean() < 0.95
