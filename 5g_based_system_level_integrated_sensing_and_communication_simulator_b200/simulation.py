"""``+simulation`` mirror (hot-path driver only).

``cellSimulation`` runs the hot path of one cell the way ``simulation.cellSimulation`` does after its slot loop
(reference +simulation/cellSimulation.m:141-145, :189-202): radar parameters, CFAR configuration, mono-static
sensing of the accumulated Tx waveform, fft2D estimation.  The MAC/RLC/APP slot loop stays in the reference.
``networkSimulation`` shards independent cells over ranks (reference +simulation/networkSimulation.m:57-60 loops
over them serially; cells never interact) and gathers the per-cell result records.
"""
from __future__ import annotations

import numpy as np


def shard_cells(n_cells: int, world_size: int, rank: int):
    """Block-cyclic assignment of cells to ranks (19 cells over 8 ranks -> 3,3,3,2,2,2,2,2)."""
    return list(range(rank, n_cells, world_size))


def cellSimulation(cellSimuParams, senTxWave=None, senTxGrid=None, noise=None, seed=0):
    """Sensing pass of ``[comResults, senResults] = simulation.cellSimulation(cellSimuParams)``
    (cellSimulation.m:141-145,189-202).  ``senTxWave`` / ``senTxGrid`` are what gNBPhy accumulates
    (gNBPhy.m:604-612).  A failing estimator yields ``senResults = nan`` like the reference's try/catch (:196-202)."""
    from . import _lib, sensing
    p = cellSimuParams
    carrierInfo, waveInfo = p["carrierInfo"], p["waveInfo"]
    radarParams = sensing.radarParams(p, carrierInfo, waveInfo)                      # :144
    cfarConfig = sensing.detection.cfar2D(radarParams)                               # :145
    senRxGrid = sensing.monoStaticSensing(senTxWave, np.asarray(senTxGrid).shape, carrierInfo, radarParams,
                                          p["targetLoSConditions"], noise=noise, seed=seed)   # :194
    try:
        senResults = sensing.estimation.fft2D(radarParams, cfarConfig, senRxGrid, senTxGrid)   # :197
    except _lib.IsacError:
        senResults = float("nan")                                                    # :198-202
    return {"cellID": p.get("cellID", 0)}, senResults


def networkSimulation(cell_params_list, cell_fn=cellSimulation, cell_args=None, group=None):
    """Cells -> ranks, no data-path collective; results gathered to every rank with ``all_gather_object``
    (KB-sized records).  Works on any ``torch.distributed`` backend (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    n = len(cell_params_list)
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    mine = shard_cells(n, world, rank)
    local = {}
    for i in mine:
        args = cell_args[i] if cell_args is not None else ()
        local[i] = cell_fn(cell_params_list[i], *args)
    if world == 1:
        return [local[i] for i in range(n)]
    gathered = [None] * world
    dist.all_gather_object(gathered, local, group=group)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(n)]
