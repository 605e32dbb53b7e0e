"""CPU restatement (NumPy float64) of the reference's LoS / blockage geometry -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package; the
product path (5g_..._b200/) never does.  PARITY UNPINNED: the reference ships no tests or recorded outputs; MATLAB cannot
run here.  The only reference artefact usable as a fixture is its cached city, dataFiles/blockages/OSM_city.json
(81 buildings), which tests/golden/make_golden.py condenses into tests/golden/osm_city.npz.

Follows (file:line relative to /root/reference):
  +networkTopology/+blockages/wallBlockage.m:31-74   constructor (plane normal / distance)
  +networkTopology/+blockages/wallBlockage.m:96-148  checkBlockage (projection along the link + winding number)
  +networkTopology/+blockages/wallBlockage.m:178-222 getWindingNumber
  +networkTopology/+blockages/building.m:37-99       walls + ceiling of a building
  +networkTopology/+blockages/building.m:113-137     building.checkBlockage (any wall)
  +networkTopology/+blockages/openStreetMapCity.m:67-95  checkLoS (any building), call sites networkSimulation.m:138,154
"""
from __future__ import annotations

import numpy as np


def wall_plane(corner_list):
    """wallBlockage constructor (wallBlockage.m:62-71): ``vectors = c1 - c(2:end)``, ``basis = orth(vectors)``,
    ``normVec = cross(b1,b2)/norm``, ``normDist = normVec'*c1``.  ``orth`` = left singular vectors of the range
    (MATLAB: SVD based); the sign of the normal is arbitrary and irrelevant (abs() of the winding sum, :216)."""
    c = np.asarray(corner_list, dtype=np.float64)
    vectors = c[:, [0]] - c[:, 1:]
    u, s, _ = np.linalg.svd(vectors, full_matrices=False)
    tol = max(vectors.shape) * np.spacing(s.max())
    basis = u[:, s > tol]
    n = np.cross(basis[:, 0], basis[:, 1])
    n = n / np.linalg.norm(n)
    return n, float(n @ c[:, 0])


def winding_number(corner_list, norm_vec, point):
    """getWindingNumber (wallBlockage.m:178-222) for points [3 x n] -> [n]."""
    poly = np.asarray(corner_list, dtype=np.float64)
    point = np.asarray(point, dtype=np.float64)
    n_pts, n_c = point.shape[1], poly.shape[1]
    with np.errstate(all="ignore"):
        vec = poly[:, None, :] - point[:, :, None]                       # [3 x nPoints x nCorners]   :199-200
        len_vec = np.sqrt((vec ** 2).sum(axis=0, keepdims=True))        # vecnorm(vec,2,1)           :201
        invalid = len_vec < 1e-10                                        # :203
        vec = vec / len_vec                                              # :207
        shift = np.roll(vec, 1, axis=2)                                  # circshift(vec,1,3)         :209
        dotv = (shift * vec).sum(axis=0)                                 # :210
        crossv = np.cross(shift, vec, axis=0)                            # :211
        ang = np.arctan2((norm_vec[:, None, None] * crossv).sum(axis=0), dotv)   # :215
        wn = np.zeros(n_pts)
        for k in range(n_c):                                             # sum(diffAngle,3): left to right
            wn = wn + ang[:, k]
        wn = np.abs(wn)                                                  # :220
    wn[invalid[0].sum(axis=1) > 0] = 1.0                                 # :222
    return wn


def wall_check_blockage(corner_list, norm_vec, norm_dist, ue, ant):
    """wallBlockage.checkBlockage (wallBlockage.m:121-127): project the user onto the wall plane ALONG THE LINK
    (the infinite line, not the segment -- a wall behind the antenna on that line also blocks: reference behaviour)
    and test the winding number against 0.1.  ue, ant: [3 x n] (ant may be [3 x 1])."""
    ue = np.asarray(ue, dtype=np.float64)
    ant = np.asarray(ant, dtype=np.float64)
    with np.errstate(all="ignore"):
        vec = ue - ant
        t = (norm_dist - norm_vec @ ue) / (norm_vec @ vec)
        proj = ue + vec * t[None, :]
    return winding_number(corner_list, norm_vec, proj) > 0.1


def building_walls(floor_plan, height):
    """building constructor (building.m:61-73): one 4-corner wall per floor-plan edge, then the ceiling polygon."""
    fp = np.asarray(floor_plan, dtype=np.float64)
    walls = []
    for i in range(fp.shape[1] - 1):
        ll = np.array([fp[0, i], fp[1, i], 0.0])
        lr = np.array([fp[0, i + 1], fp[1, i + 1], 0.0])
        ul = np.array([fp[0, i], fp[1, i], height])
        ur = np.array([fp[0, i + 1], fp[1, i + 1], height])
        walls.append(np.stack([ll, lr, ur, ul], axis=1))
    walls.append(np.vstack([fp, np.full((1, fp.shape[1]), float(height))]))
    return walls


def check_los(buildings, ue_pos, ant_pos):
    """openStreetMapCity.checkLoS (openStreetMapCity.m:67-95) for element-wise link pairs.

    buildings: list of (floorPlan [2 x nCorner], height); ue_pos [n x 3]; ant_pos [n x 3] or [1 x 3]
    (the reference is called with one link at a time, networkSimulation.m:138,154).  Returns bool [n]: True = LoS."""
    ue = np.atleast_2d(np.asarray(ue_pos, dtype=np.float64)).T
    ant = np.atleast_2d(np.asarray(ant_pos, dtype=np.float64)).T
    blocked = np.zeros(ue.shape[1])
    for fp, h in buildings:
        b = np.zeros(ue.shape[1])
        for w in building_walls(fp, h):
            n, d = wall_plane(w)
            b = b + wall_check_blockage(w, n, d, ue, ant)                # building.m:129-131
        blocked = blocked + (b > 0)                                      # building.m:135, openStreetMapCity.m:84-86
    return ~(blocked > 0)                                                # :90-93
