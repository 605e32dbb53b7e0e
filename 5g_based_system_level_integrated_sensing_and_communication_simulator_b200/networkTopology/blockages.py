"""``+networkTopology/+blockages`` mirror: the city layout and its batched LoS test (csrc/los.cu)."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from .. import _lib


def building_walls(floorPlan, height):
    """Wall polygons of a building as its constructor builds them (reference +networkTopology/+blockages/building.m:61-73):
    one 4-corner wall [lowerLeft, lowerRight, upperRight, upperLeft] per floor-plan edge, then the ceiling polygon at
    z = height.  ``floorPlan``: [2 x nCorner] closed polygon (first corner repeated last).  Returns a list of [3 x n]."""
    fp = np.asarray(floorPlan, dtype=np.float64)
    if fp.ndim != 2 or fp.shape[0] != 2 or fp.shape[1] < 3:
        raise _lib.IsacError(1, "floorPlan must be [2 x nCorner] with at least 3 corners")
    h = float(height)
    walls = []
    for i in range(fp.shape[1] - 1):
        x0, y0, x1, y1 = fp[0, i], fp[1, i], fp[0, i + 1], fp[1, i + 1]
        walls.append(np.array([[x0, x1, x1, x0], [y0, y1, y1, y0], [0.0, 0.0, h, h]]))
    walls.append(np.vstack([fp, np.full((1, fp.shape[1]), h)]))
    return walls


class city:
    """``networkTopology.blockages.city`` / ``openStreetMapCity`` restricted to what the simulation loop uses: the building
    list and ``checkLoS`` (reference +networkTopology/+blockages/openStreetMapCity.m:67-95, city.m:116-143)."""

    def __init__(self, buildings, device=None):
        """``buildings``: iterable of dicts with ``floorPlan`` [2 x nCorner] and ``height`` (the fields city.m:128-133
        reads from the JSON file), or (floorPlan, height) tuples."""
        self.buildings = []
        walls = []
        for b in buildings:
            fp, h = (b["floorPlan"], b["height"]) if isinstance(b, dict) else b
            fp = np.asarray(fp, dtype=np.float64)
            self.buildings.append((fp, float(h)))
            walls += building_walls(fp, h)
        if not walls:
            raise _lib.IsacError(1, "a city needs at least one building")
        off = np.zeros(len(walls) + 1, dtype=np.int32)
        off[1:] = np.cumsum([w.shape[1] for w in walls])
        corners = np.asfortranarray(np.concatenate(walls, axis=1))     # [3 x nCorners]
        self.nWalls = len(walls)
        self.ctx = _lib.get_context(device)
        h_ = C.c_void_p()
        _lib.check(self.ctx.lib.isac_city_create(self.ctx.handle, self.nWalls, _lib.ptr(off), _lib.ptr(corners), C.byref(h_)),
                   self.ctx.handle)
        self.handle = h_

    @classmethod
    def loadCityFromFile(cls, loadFile, device=None):
        """city.loadCityFromFile (city.m:116-143): the JSON written by saveCityToFile, e.g. the reference's cached
        ``dataFiles/blockages/OSM_city.json``."""
        with open(loadFile) as fh:
            data = json.load(fh)
        return cls(data["buildings"], device=device)

    def checkLoS(self, uePos, antPos):
        """``losDecision = simuLayout.checkLoS(uePos, antPos)`` for a batch of links (the reference is called once per UE
        and per target, networkSimulation.m:138,154).  ``uePos`` [n x 3] (or [3]); ``antPos`` [n x 3] for element-wise
        pairs or [3] / [1 x 3] for one antenna.  Returns a bool array [n] (True = line of sight)."""
        ue = np.ascontiguousarray(np.atleast_2d(np.asarray(uePos, dtype=np.float64)))
        ant = np.ascontiguousarray(np.atleast_2d(np.asarray(antPos, dtype=np.float64)))
        if ue.shape[1] != 3 or ant.shape[1] != 3 or ant.shape[0] not in (1, ue.shape[0]):
            raise _lib.IsacError(1, "uePos must be [n x 3] and antPos [n x 3] or [1 x 3]")
        los = np.zeros(ue.shape[0], dtype=np.int32)
        self.ctx.use_own_stream()
        _lib.check(self.ctx.lib.isac_city_check_los_host(self.handle, ue.shape[0], _lib.ptr(ue), _lib.ptr(ant), ant.shape[0],
                                                         _lib.ptr(los)), self.ctx.handle)
        return los.astype(bool)

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.isac_city_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


openStreetMapCity = city   # the shipped scenario's layout class (+networkTopology/+blockages/openStreetMapCity.m)
