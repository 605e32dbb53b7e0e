"""BASELINE config 5 (openStreetMapCity scenario: 19 gNBs, 100 UEs, 20 moving targets, cells sharded 3,3,3,2,2,2,2,2 over 8
ranks) through the package's per-frame driver (simulation.networkFrames), at the test radio size (24 PRB, same structure as
the shipped 273-PRB radio; the bench runs the shipped size):

* every cell's records are identical whether the 19 cells run in one process or sharded over 8 ranks (last
  row) -- emulated in-process and with two real processes that exchange records through torch.distributed;
* one cell is checked against the float64 oracle end to end: CSI report of a CSI-RS occasion from the device-resident channel
  matrices (after path loss / Rx gain) and the sensing estimates of the CPI with the targets where they are in that frame;
* LoS flags come from the reference's cached OSM city through the device LoS kernel and equal the oracle's decisions."""
import importlib
import os
import socket

import numpy as np
import pytest

from oracle import comm as OC
from oracle import geometry as OG
from oracle import sensing as OS

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu
N_FRAMES = 2


def _buildings():
    z = np.load(os.path.join(HERE, "golden", "osm_city.npz"))
    off = z["fp_off"]
    return [(z["fp_flat"][:, off[i]:off[i + 1]], float(z["heights"][i])) for i in range(off.size - 1)]


@pytest.fixture(scope="module")
def world1(gpu):
    sim = importlib.import_module(PKG + ".simulation")
    W = importlib.import_module(PKG + ".workloads")
    scn = W.scenario_cfg5(radio="small")
    hp = sim.HotPath(scn, city_buildings=_buildings())
    recs = sim.networkFrames(scn, N_FRAMES, interference=True, hp=hp)
    return sim, W, scn, hp, recs


def test_cfg5_layout_and_records(world1):
    sim, W, scn, hp, recs = world1
    assert scn["gnb"].shape == (19, 3) and scn["ue"].shape == (100, 3) and scn["target0"].shape == (20, 3)
    assert [len(sim.shard_cells(19, 8, r)) for r in range(8)] == [3, 3, 3, 2, 2, 2, 2, 2]
    assert len(recs) == N_FRAMES and recs[0].shape[0] == 19
    n_ue = np.bincount(scn["ue_cell"], minlength=19)
    n_tg = np.bincount(scn["target_cell"], minlength=19)
    for f in range(N_FRAMES):
        assert np.array_equal(recs[f][:, 0], n_ue.astype(float))
        assert not np.any(np.isnan(recs[f][n_ue > 0, 1]))               # the scheduled UE's PDSCH was precoded
        sen_ok = recs[f][:, -(4 + 2 * sim.REC_RNG + sim.REC_AZI)]
        assert np.all(sen_ok[n_tg == 0] == 0.0)                           # no target attached -> senResults = NaN
    assert sum(recs[0][c, -(4 + 2 * sim.REC_RNG + sim.REC_AZI)] for c in range(19)) >= 3   # cells whose targets are in line of sight detect them
    assert not np.array_equal(recs[0], recs[1], equal_nan=True)           # the channels and the targets moved


def test_cfg5_los_flags_match_oracle(world1):
    sim, W, scn, hp, recs = world1
    ue_los, tgt_los = hp.los_flags(1)
    tpos, _ = W.cfg5_target_state(scn, 1)
    b = _buildings()
    assert np.array_equal(ue_los, OG.check_los(b, scn["ue"], scn["gnb"][scn["ue_cell"]]))
    assert np.array_equal(tgt_los, OG.check_los(b, tpos, scn["gnb"][scn["target_cell"]]))
    assert 0 < ue_los.sum() < ue_los.size                                  # the city blocks some links, not all


def test_cfg5_rank_invariance_emulated_8_ranks(world1):
    """The cells of each of 8 ranks computed by a fresh driver instance: same records as the single-process run."""
    sim, W, scn, hp, recs = world1
    for f in range(N_FRAMES):
        full = np.full_like(recs[f], np.nan)
        summary = hp.tx_summary(list(range(19)), f)
        for r in range(8):
            hp_r = sim.HotPath(scn, city_buildings=_buildings())
            ue_los, tgt_los = hp_r.los_flags(f)
            for c in sim.shard_cells(19, 8, r):
                cellp, _, _ = W.cfg5_cell_params(scn, c, f, ue_los, tgt_los)
                com, sen = sim.cellFrame(hp_r, cellp, summary)
                full[c] = sim.pack_record(hp_r, com, sen)
        assert np.array_equal(full, recs[f], equal_nan=True), f


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK="0")
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)        # both ranks share cuda:0; records travel through gloo
    sim = importlib.import_module(PKG + ".simulation")
    W = importlib.import_module(PKG + ".workloads")
    scn = W.scenario_cfg5(radio="small")
    recs = sim.networkFrames(scn, N_FRAMES, city_buildings=_buildings(), interference=True)
    out.put((rank, [r.copy() for r in recs]))
    dist.destroy_process_group()


def test_cfg5_rank_invariance_two_processes(world1):
    import torch.multiprocessing as mp
    sim, W, scn, hp, recs = world1
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, rr in got:
        for f in range(N_FRAMES):
            assert np.array_equal(rr[f], recs[f], equal_nan=True), (rank, f)


def test_cfg5_one_cell_against_the_oracle(world1):
    sim, W, scn, hp, _ = world1
    n_tg = np.bincount(scn["target_cell"], minlength=19)
    n_ue = np.bincount(scn["ue_cell"], minlength=19)
    cell = int(np.flatnonzero((n_tg > 0) & (n_ue > 0))[0])
    frame = 1
    ue_los, tgt_los = hp.los_flags(frame)
    tgt_los[:] = True                                                      # keep every target of the checked cell visible
    cellp, car, wave = W.cfg5_cell_params(scn, cell, frame, ue_los, tgt_los)
    # COMM: the report of every occasion from the very channel matrices the device used
    hp.keep_H = []
    com = hp.comm_frame(cellp, None)
    kept, hp.keep_H = hp.keep_H, None
    r = W.RADIO["small"]
    ocfg = OC.report_config(r["csirs_ports"], r["panel"], r["nrb"], 0, 1, "Subband", "Subband", r["subband"])
    re_k, re_l = OC.csirs_first_port_res(r["nrb"], 1, 0)
    for cells_of, o, H, nvar in kept:
        assert set(cells_of) == {cell}
        for u in range(H.shape[0]):
            Hm = H[u].transpose(3, 2, 1, 0)                                # [P][R][L][K] -> [K x L x R x P]
            rank, pmo, cqo = OC.csi_report_vectorized(ocfg, re_k, re_l, Hm, nvar[u], hp.table, rank_cap=4)
            assert com["RI"][o, u] == rank, (o, u)
            assert np.array_equal(com["i1"][o, :, u], pmo["i1"]) and np.array_equal(com["i2"][o, :, u], pmo["i2"], equal_nan=True)
            assert com["CQI"][o, u] == cqo[0, 0]
    # SENSING: explicit noise tensor (MATLAB's randn cannot be reproduced), oracle on the same grid / waveform
    grid = W.cfg5_sensing_grid(scn, cell)
    amp = 10.0 ** ((cellp["gNBTxPower"] - 30.0) / 20.0) * np.sqrt(wave["Nfft"] ** 2 / (grid.shape[0] * grid.shape[2]))
    txw = (amp * W.ofdm_modulate(grid, r["nrb"], r["scs"])).astype(np.complex64)
    noise = W.std_normal_complex(txw.shape, 77).astype(np.complex64)
    sen = hp.sensing_cpi(cellp, noise=noise)
    rp = OS.radar_params(cellp, car, wave)
    rx = OS.mono_static_sensing(txw, grid.shape, car, rp, cellp["targetLoSConditions"], noise).astype(np.complex64)
    ref = OS.fft2d(rp, OS.cfar2d_config(rp), rx, grid.astype(np.complex64))
    assert isinstance(sen, dict)
    assert np.array_equal(sen["rngEst"], ref["rngEst"]) and np.array_equal(sen["velEst"], ref["velEst"])
    assert np.array_equal(sen["aziEst"], ref["aziEst"])
    # the moving target shows up where it is in THIS frame
    tpos, radial = W.cfg5_target_state(scn, frame)
    d = np.linalg.norm(tpos[cellp["targetIDs"]] - scn["gnb"][cell], axis=1)
    assert min(abs(sen["rngEst"][0] - x) for x in d) <= 2 * rp["rRes"]
