"""CPU-only (gloo, world size 2): cells shard across ranks without any data-path collective and the gathered
per-cell records equal the single-process result (reference: networkSimulation.m:57-60 loops cells serially)."""
import importlib
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"


def _fake_cell(params):
    # stands in for the GPU hot path of one cell: deterministic record derived from the cell's parameters
    return {"cellID": params["cellID"], "rngEst": [params["cellID"] * 10.0 + k for k in range(params["numTargets"])]}


def _worker(rank, world, port, n_cells, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = importlib.import_module(PKG + ".simulation")
    cells = [{"cellID": i, "numTargets": 1 + i % 3} for i in range(n_cells)]
    res = sim.networkSimulation(cells, cell_fn=_fake_cell)
    mine = sim.shard_cells(n_cells, world, rank)
    out.put((rank, res, mine))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_cells_block_cyclic():
    sim = importlib.import_module(PKG + ".simulation")
    sizes = [len(sim.shard_cells(19, 8, r)) for r in range(8)]
    assert sizes == [3, 3, 3, 2, 2, 2, 2, 2]                     # 19 cells over 8 GPUs
    assert sorted(sum((sim.shard_cells(7, 8, r) for r in range(8)), [])) == list(range(7))
    assert sim.shard_cells(7, 8, 7) == []                        # 7-cell hex layout: one rank idle


def test_network_simulation_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n_cells = _free_port(), 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cells, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [_fake_cell({"cellID": i, "numTargets": 1 + i % 3}) for i in range(n_cells)]
    for rank, res, mine in got:
        assert res == expect                                      # every rank holds every cell's record, in cell order
        assert mine == list(range(rank, n_cells, 2))


def test_network_simulation_single_process():
    sim = importlib.import_module(PKG + ".simulation")
    cells = [{"cellID": i, "numTargets": 2} for i in range(3)]
    assert sim.networkSimulation(cells, cell_fn=_fake_cell) == [_fake_cell(c) for c in cells]


def test_get_rmse_host_logic():
    """sensing.postProcessing.getRMSE (getRMSE.m:1): first-match within rRes, NaN for unmatched estimates, NaN when empty."""
    import importlib
    import numpy as np
    pp = importlib.import_module("5g_based_system_level_integrated_sensing_and_communication_simulator_b200.sensing.postProcessing")
    params = {"rRes": 1.2, "antennaType": {"type": "ula"},
              "tgtRealPos": [{"Range": 80.0, "Velocity": -12.0, "Elevation": 1.0, "Azimuth": 30.0},
                             {"Range": 150.0, "Velocity": 5.0, "Elevation": 2.0, "Azimuth": -40.0}]}
    res = [{"rngEst": [80.5, 149.4], "velEst": [-11.0, 5.5], "aziEst": [31.0, -40.0]},
           {"rngEst": [300.0], "velEst": [0.0], "aziEst": [0.0]}]
    out = pp.getRMSE(res, params)
    assert np.allclose(out["rngRMSE"][:2], [0.5, 0.6]) and np.isnan(out["rngRMSE"][2])
    assert np.allclose(out["velRMSE"][:2], [1.0, 0.5]) and np.allclose(out["aziRMSE"][:2], [1.0, 0.0])
    assert np.all(np.isnan(out["eleRMSE"]))
    empty = pp.getRMSE({"rngEst": [], "velEst": [], "aziEst": []}, params)
    assert isinstance(empty, float) and np.isnan(empty)
