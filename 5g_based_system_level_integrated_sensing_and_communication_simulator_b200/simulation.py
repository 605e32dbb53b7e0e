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


# ---------------------------------------------------------------------------------------------------------------------
# Per-frame hot path of a multi-cell scenario (BASELINE config 5): what uePhy / gNBPhy / cellSimulation call per cell and
# frame, driven from Python.  COMM: CSI-RS occasions -> CDL channel matrix of every UE (profile by LoS, updateCDLModels.m:7-15)
# -> path loss / Rx gain / noise (uePhy.m:743-755) -> fused RI / PMI / CQI report (uePhy.m:886-932); SRS occasions -> UL channel
# -> TPMI selection (gNBPhy.m:983-1062); the scheduled UE's PDSCH precoded with the gNB-side codebook at the reported PMI
# (schedulerEntity.m:736-777 -> gNBPhy.m:775-831).  SENSING: cellSimulation.m:141-145,189-202 with the targets where they
# are in that CPI.  Cells never interact in the reference; the optional inter-cell interference term (north_star) enters only
# through nVar and is exchanged once per frame.
# ---------------------------------------------------------------------------------------------------------------------
REC_RNG, REC_AZI = 16, 8


class HotPath:
    """Everything the cells of one process share for a scenario: radio constants, plans (cached in the package modules), CDL
    channel objects per UE, the city for LoS tests."""

    def __init__(self, scn, device=None, city_buildings=None):
        import importlib
        from . import _lib, communication, workloads
        self.scn, self.W = scn, workloads
        self.ph = communication.phyLayer
        self.cm = communication.channelModels
        self.pl = communication.pathlossModels
        self.comm = communication
        self.sensing = importlib.import_module(__package__ + ".sensing")
        self.ctx = _lib.get_context(device)
        self.device = self.ctx.device
        r = workloads.RADIO[scn["radio"]]
        self.r = r
        self.K = 12 * r["nrb"]
        self.num = workloads.ofdm_numerology(r["nrb"], r["scs"])
        self.sym_t = np.ascontiguousarray(workloads.symbol_starts(self.num, 14) / self.num["SampleRate"], dtype=np.float64)
        self.slot_dur = 1e-3 / self.num["SlotsPerSubframe"]
        self.carrier = {"NSizeGrid": r["nrb"], "NStartGrid": 0, "SymbolsPerSlot": 14}
        self.csirs = {"NumCSIRSPorts": r["csirs_ports"], "NumRB": r["nrb"], "RBOffset": 0, "SubcarrierLocations": 1,
                      "SymbolLocations": 0, "Density": "one"}
        self.rc = {"NSizeBWP": r["nrb"], "NStartBWP": 0, "PanelDimensions": r["panel"], "CodebookMode": 1, "PMIMode": "Subband",
                   "CQIMode": "Subband", "SubbandSize": r["subband"], "OverSamplingFactors": (4, 1) if r["panel"][1] == 1 else (4, 4)}
        self.table = np.ascontiguousarray(communication.setupSINRtoCQIMappingTable()["downlinkSINR90pc"], dtype=np.float64)
        self.nSB = -(-r["nrb"] // r["subband"])
        self.csi_slots = [s for s in range(r["num_slots"]) if s % 5 == 2]          # CSIRSPeriod [5 2] (setupCSIRS.m:11)
        self.srs_slots = [s for s in range(r["num_slots"]) if s % 5 == 4]          # the UL slot of DDDSU
        self.noise_re = self.pl.thermalNoisePower(scn["noiseFigureUE"], 290.0, self.num["SampleRate"]) / self.num["Nfft"]
        self.noise_re_gnb = self.pl.thermalNoisePower(6.0, 290.0, self.num["SampleRate"]) / self.num["Nfft"]
        self.city = None
        if city_buildings is not None:
            from .networkTopology import blockages
            self.city = blockages.city(city_buildings, device=self.device)
        self._channels = {}
        self._pl_cache = {}
        self._xgain = {}
        self._grids = {}            # per cell: (pinned host senTxGrid [nAnts][nSym][nSc], NumPy view of it)
        self.keep_H = None          # tests: a list that receives (cellID, occasion, H after the link budget, nVar)
        self.maxUE = int(np.bincount(scn["ue_cell"], minlength=scn["gnb"].shape[0]).max())

    # -- geometry ------------------------------------------------------------------------------------------------------
    def los_flags(self, frame):
        """LoS of every UE and every target towards its own gNB (networkSimulation.m:138,154), one batched device call."""
        scn = self.scn
        tpos, _ = self.W.cfg5_target_state(scn, frame)
        if self.city is None:
            return np.ones(scn["ue"].shape[0], bool), np.ones(tpos.shape[0], bool)
        pts = np.vstack([scn["ue"], tpos])
        ant = np.vstack([scn["gnb"][scn["ue_cell"]], scn["gnb"][scn["target_cell"]]])
        los = self.city.checkLoS(pts, ant)
        return los[: scn["ue"].shape[0]], los[scn["ue"].shape[0]:]

    def tx_summary(self, cells, frame):
        """Per-cell transmit summary of a frame, the payload of the inter-cell exchange: [cell, x, y, z, W per RE]."""
        scn = self.scn
        p_re = 10 ** ((scn["txPower"] - 30.0) / 10.0) / self.K
        out = np.zeros((len(cells), 5))
        for q, c in enumerate(cells):
            out[q] = [c, *scn["gnb"][c], p_re * scn["load"][c, frame % scn["load"].shape[1]]]
        return out

    def interference(self, cellp, summary):
        """Inter-cell interference power per RE at every UE of the cell: sum over the OTHER cells of their per-RE Tx power
        through the UMa path loss of that link (LoS from the city) and the UE's Rx gain.  No reference counterpart (cells are
        independent there, networkSimulation.m:57-60): an extension named by BASELINE.json, off by default in the parity tests."""
        ue = np.asarray(cellp["uePosition"]).reshape(-1, 3)
        if ue.shape[0] == 0:
            return np.zeros(0)
        c = int(cellp["cellID"])
        g = self._xgain.get(c)
        if g is None:   # linear gains UE <- every gNB site: the sites, the UEs and the city are static, so one device pass per cell
            gnb = self.scn["gnb"]
            uu = np.repeat(ue, gnb.shape[0], axis=0)
            bb = np.tile(gnb, (ue.shape[0], 1))
            los = self.city.checkLoS(uu, bb) if self.city is not None else np.ones(uu.shape[0], bool)
            pl = self.pl.config5GNRModels(self.scn["scenario"], self.scn["fc"], los.astype(np.int32), bb, uu, device=self.device)
            g = 10 ** ((self.scn["rxGainUE"] - np.atleast_1d(pl)) / 10.0)
            g = g.reshape(ue.shape[0], gnb.shape[0])
            self._xgain[c] = g
        p_re = np.zeros(self.scn["gnb"].shape[0])
        p_re[summary[:, 0].astype(int)] = summary[:, 4]
        p_re[c] = 0.0                                   # the own cell is the signal, not interference
        return g @ p_re

    # -- COMM ----------------------------------------------------------------------------------------------------------
    def _channel(self, ue_id, los, uplink):
        key = (int(ue_id), bool(los), bool(uplink))
        ch = self._channels.get(key)
        if ch is None:
            prof = "CDL-D" if los else "CDL-A"                                     # updateCDLModels.m:11-13
            P = self.r["csirs_ports"]
            if uplink:   # SRS: 2 UE ports -> the gNB's Rx array (cdl.m:69-88)
                ch = self.cm.CDLChannel(prof, TransmitAntennaArraySize=(1, 1, 2), ReceiveAntennaArraySize=(1, self.r["nV"], self.r["p"]),
                                        Seed=50_000 + int(ue_id), device=self.device)
            else:        # CSI-RS ports -> the UE's antennas (cdl.m:48-67)
                ch = self.cm.CDLChannel(prof, TransmitAntennaArraySize=(1, P // 2, 2), ReceiveAntennaArraySize=(1, self.r["ue_ants"] // 2, 2),
                                        Seed=73 + int(ue_id), device=self.device)
            self._channels[key] = ch
        return ch

    def _batch(self, ue_ids, los, uplink, t0, sym_t):
        """Channel matrices of the UEs in list order: one generateBatch per delay profile present."""
        import torch
        chans = [self._channel(u, l, uplink) for u, l in zip(ue_ids, los)]
        c0 = chans[0]
        H = torch.empty((len(chans), c0.nTx, c0.nRx, len(sym_t), self.K), dtype=torch.complex64, device=f"cuda:{self.device}")
        for flag in (True, False):
            idx = [i for i, l in enumerate(los) if bool(l) == flag]
            if not idx:
                continue
            part = self.cm.CDLChannel.generateBatch([chans[i] for i in idx], self.K, self.r["scs"] * 1e3, sym_t, t0)
            H[idx] = part
        return H

    def comm_frame(self, cellp, interf=None):
        """COMM share of one frame of one cell.  Returns the comResults record (arrays indexed by the cell's UE order)."""
        return self.comm_frames([cellp], [interf])[0]

    def _path_loss(self, cellp):
        """UMa path loss of the cell's UE links (uePhy.m:743-747); UEs and the city are static, so one device call per cell."""
        key = (int(cellp["cellID"]), tuple(np.asarray(cellp["ueLoSConditions"]).astype(int).tolist()))
        pl = self._pl_cache.get(key)
        if pl is None:
            ue = np.asarray(cellp["uePosition"]).reshape(-1, 3)
            pl = np.atleast_1d(self.pl.config5GNRModels(self.scn["scenario"], self.scn["fc"], np.asarray(key[1], np.int32),
                                                       cellp["gNBPosition"], ue, device=self.device))
            self._pl_cache[key] = pl
        return pl

    def comm_frames(self, cellps, interfs=None):
        """COMM share of one frame for several cells at once: the UEs of all cells that report in the same slot go through the
        CDL generator, the link budget and the fused report together (batches of up to 32 UEs, the library's limit) -- the
        per-UE results do not depend on who shares the batch, so the records are the same as cell by cell."""
        interfs = interfs if interfs is not None else [None] * len(cellps)
        scn, nSB, nOcc, nSrs = self.scn, self.nSB, len(self.csi_slots), len(self.srs_slots)
        recs, ues = [], []          # ues: (cell index, position in the cell, ue id, los, path loss, nVar, txLoad)
        for ci, (cellp, interf) in enumerate(zip(cellps, interfs)):
            ue_ids = np.asarray(cellp["ueIDs"], int)
            n = ue_ids.size
            recs.append({"ueIDs": ue_ids, "RI": np.full((nOcc, n), np.nan), "i1": np.full((nOcc, 3, n), np.nan),
                         "i2": np.full((nOcc, nSB, n), np.nan), "CQI": np.full((nOcc, n), np.nan),
                         "ulPMI": np.full((nSrs, nSB, n), np.nan), "precodeEnergy": np.nan})
            if n == 0:
                continue
            los = np.asarray(cellp["ueLoSConditions"]).astype(bool)
            pl_db = self._path_loss(cellp)
            nvar = self.noise_re + (np.zeros(n) if interf is None else np.asarray(interf))
            for q in range(n):
                ues.append((ci, q, int(ue_ids[q]), bool(los[q]), float(pl_db[q]), float(nvar[q]), float(cellp["txLoad"]),
                            float(cellp["gNBRxGain"])))
        if not ues:
            return recs
        frame_t0 = cellps[0]["frame"] * scn["frame_time"]
        p_tx = 10 ** ((scn["txPower"] - 30.0) / 10.0) / (self.K * self.r["csirs_ports"])                # W per RE and port at full load
        p_ul = 10 ** ((23.0 - 30.0) / 10.0) / (self.K * 2)                                               # 23 dBm UE, 2 SRS ports
        for c0 in range(0, len(ues), 32):
            grp = ues[c0: c0 + 32]
            ids, los = [g[2] for g in grp], [g[3] for g in grp]
            pl = np.array([g[4] for g in grp])
            nvar = np.array([g[5] for g in grp])
            load = np.array([g[6] for g in grp])
            for o, slot in enumerate(self.csi_slots):
                H = self._batch(ids, los, False, frame_t0 + slot * self.slot_dur, self.sym_t)
                self.pl.applyPathLossAndRxGain(H, pl - 10 * np.log10(p_tx * load), scn["rxGainUE"])    # uePhy.m:743-751
                if self.keep_H is not None:
                    self.keep_H.append(([cellps[g[0]]["cellID"] for g in grp], o, H.cpu().numpy(), nvar.copy()))
                RI, pm, cq = self.ph.csiReport(self.carrier, self.csirs, self.rc, H, nvar, self.table, rankCap=4)   # uePhy.m:900-907
                m = len(grp)
                if scn.get("sinr_grid"):   # BASELINE config 3: the per-PRB SINR grid at the reported rank (info.SINRPerRE, dlPMISelect.m:505)
                    ranks = np.atleast_1d(RI)
                    for nu in sorted({int(x) for x in ranks if not np.isnan(x)}):
                        sel = [j for j in range(m) if ranks[j] == nu]
                        _, info = self.ph.dlPMISelect(self.carrier, self.csirs, self.rc, nu, H[sel], nvar[sel])
                        S = np.asarray(info["SINRPerRE"])
                        S = S.reshape(S.shape + (1,)) if len(sel) == 1 else S
                        for q, j in enumerate(sel):
                            recs[grp[j][0]].setdefault("sinrGridSum", np.zeros(len(recs[grp[j][0]]["ueIDs"])))[grp[j][1]] = np.nansum(S[..., q])
                RI, i1, i2 = np.atleast_1d(RI), np.asarray(pm["i1"]).reshape(3, m), np.asarray(pm["i2"]).reshape(nSB, m)
                cq = np.asarray(cq).reshape(np.asarray(cq).shape[0], -1, m)
                for j, g in enumerate(grp):
                    r = recs[g[0]]
                    r["RI"][o, g[1]], r["i1"][o, :, g[1]], r["i2"][o, :, g[1]], r["CQI"][o, g[1]] = RI[j], i1[:, j], i2[:, j], cq[0, 0, j]
            rxg = grp[0][7]
            for o, slot in enumerate(self.srs_slots):
                Hul = self._batch(ids, los, True, frame_t0 + slot * self.slot_dur, self.sym_t[13:14])
                self.pl.applyPathLossAndRxGain(Hul, pl - 10 * np.log10(p_ul), rxg)                      # gNBPhy.m:852-860
                pmi, _, none = self.ph.pmiSelectBatch(2, Hul, self.noise_re_gnb, self.r["subband"])     # gNBPhy.m:1035, rank 2 (cellSimulation.m:16)
                pmi = np.asarray(pmi).reshape(-1, len(grp))
                pmi = np.where(np.asarray(none).reshape(1, -1) != 0, np.nan, pmi)
                for j, g in enumerate(grp):
                    recs[g[0]]["ulPMI"][o, : pmi.shape[0], g[1]] = pmi[:, j]
        for ci, cellp in enumerate(cellps):
            r = recs[ci]
            if r["ueIDs"].size:
                r["precodeEnergy"] = self._precode_scheduled(cellp, (r["RI"][-1], r["i1"][-1], r["i2"][-1]))
        return recs

    def _precode_scheduled(self, cellp, last):
        """PDSCH of the first UE of the cell precoded per PRG with the gNB-side codebook at its reported PMI
        (schedulerEntity.m:736-777 -> prgPrecode, gNBPhy.m:822).  Returns the energy of the antenna symbols (a checksum)."""
        RI, i1, i2 = last
        if np.isnan(RI[0]):
            return np.nan
        nu = int(RI[0])
        W = self.comm.pmiType1SinglePanelCodebook(self.rc, nu)                                       # [P, nu, i2, i11, i12, i13]
        a, b, c = (int(v) - 1 for v in i1[:, 0])
        P, nSB, K = self.r["csirs_ports"], self.nSB, self.K
        F = np.zeros((nu, P, nSB), np.complex64)
        for sb in range(nSB):
            q = i2[sb, 0]
            F[:, :, sb] = W[:, :, (0 if np.isnan(q) else int(q) - 1), a, b, c].T
        rng = np.random.default_rng(7_000_003 * self.scn["seed"] + 1013 * cellp["cellID"] + cellp["frame"])
        pos = (np.arange(K)[:, None] + K * np.arange(2, 14)[None, :]).T.reshape(-1)
        portind = np.stack([pos + 1 + K * 14 * j for j in range(nu)], axis=1).astype(np.int32)
        portsym = np.exp(1j * (np.pi / 4 + np.pi / 2 * rng.integers(0, 4, (pos.size, nu)))).astype(np.complex64)
        sym, _ = self.ph.prgPrecode((K, 14, P), 0, portsym, portind, F)
        return float(np.sum(np.abs(sym.astype(np.complex128)) ** 2))

    # -- SENSING -------------------------------------------------------------------------------------------------------
    def sensing_cpi(self, cellp, noise=None, tx_grid=None):
        """cellSimulation.m:141-145,189-202 for one CPI with the Tx grid generated and modulated on the device."""
        import torch
        from . import _lib
        if int(cellp["numTargets"]) == 0:
            return float("nan")
        carrier, wave = cellp["carrierInfo"], cellp["waveInfo"]
        rp = self.sensing.radarParams(cellp, carrier, wave)
        cf = self.sensing.detection.cfar2D(rp)
        if tx_grid is None:        # the cell's Tx payload: generated once, kept in pinned host memory, uploaded every CPI
            ent = self._grids.get(int(cellp["cellID"]))
            if ent is None:
                g = self.W.cfg5_sensing_grid(self.scn, cellp["cellID"])
                ent = (torch.from_numpy(np.ascontiguousarray(g.astype(np.complex64).transpose(2, 1, 0))).pin_memory(), g.shape)
                self._grids[int(cellp["cellID"])] = ent
            g_d = ent[0].to(f"cuda:{self.device}", non_blocking=True)
            shape = ent[1]
        else:
            g_d = torch.from_numpy(np.ascontiguousarray(tx_grid.astype(np.complex64).transpose(2, 1, 0))).to(f"cuda:{self.device}")
            shape = tx_grid.shape
        nsc, ntx = shape[0], shape[2]
        amp = 10.0 ** ((cellp["gNBTxPower"] - 30.0) / 20.0) * np.sqrt(wave["Nfft"] ** 2 / (nsc * ntx))   # signalAmp (gNBPhy.m:599)
        w_d = self.sensing.ofdmModulate(carrier, g_d, amp)
        seed = 9_000_011 * self.scn["seed"] + 1019 * cellp["cellID"] + cellp["frame"]
        try:
            nz = None if noise is None else torch.from_numpy(np.ascontiguousarray(noise.astype(np.complex64).T)).to(w_d.device)
            rx = self.sensing.monoStaticSensing(w_d, shape, carrier, rp, cellp["targetLoSConditions"], noise=nz,
                                                seed=None if noise is not None else seed)
            return self.sensing.estimation.fft2D(rp, cf, rx, g_d)
        except _lib.IsacError:
            return float("nan")                                                       # cellSimulation.m:198-202


def cellFrame(hp, cellp, summary=None, noise=None):
    """One frame of one cell: (comResults, senResults)."""
    interf = hp.interference(cellp, summary) if summary is not None else None
    return hp.comm_frame(cellp, interf), hp.sensing_cpi(cellp, noise)


def pack_record(hp, com, sen):
    """Fixed-size float64 record of a cell-frame (NaN padded) -- what the ranks exchange."""
    n, U = com["ueIDs"].size, hp.maxUE
    nOcc, nSB, nSrs = com["RI"].shape[0], hp.nSB, com["ulPMI"].shape[0]
    ue = np.full((U, 2 + nOcc * (5 + nSB) + nSrs * nSB), np.nan)
    grid = com.get("sinrGridSum")
    for q in range(n):
        row = [float(com["ueIDs"][q]), np.nan if grid is None else float(grid[q])]   # checksum of the per-PRB SINR grid (cfg3)
        for o in range(nOcc):
            row += [com["RI"][o, q], *com["i1"][o, :, q], com["CQI"][o, q], *com["i2"][o, :, q]]
        for o in range(nSrs):
            row += list(com["ulPMI"][o, :, q])
        ue[q] = row
    s = np.full(4 + 2 * REC_RNG + REC_AZI, np.nan)
    if isinstance(sen, dict):
        r, v, a = (np.atleast_1d(np.asarray(sen[k], float)) for k in ("rngEst", "velEst", "aziEst"))
        s[0], s[1], s[2], s[3] = 1.0, r.size, v.size, a.size
        s[4: 4 + min(r.size, REC_RNG)] = r[:REC_RNG]
        s[4 + REC_RNG: 4 + REC_RNG + min(v.size, REC_RNG)] = v[:REC_RNG]
        s[4 + 2 * REC_RNG: 4 + 2 * REC_RNG + min(a.size, REC_AZI)] = a[:REC_AZI]
    else:
        s[0] = 0.0
    return np.concatenate([[float(n), com["precodeEnergy"]], ue.ravel(), s])


def networkFrames(scn, n_frames, city_buildings=None, device=None, interference=False, group=None, hp=None, on_frame=None, frame0=0):
    """``n_frames`` frames of every cell of the scenario, cells block-cyclic over the ranks of ``group`` (reference:
    the serial cell loop of networkSimulation.m:57-60).  Per frame: (optional) all-gather of the per-cell transmit summaries for
    the interference term, the cells of this rank, all-gather of their fixed-size records.  Returns records[frame] =
    float64 array [nCells x recordLength] (identical on every rank, and for every world size)."""
    import torch
    import torch.distributed as dist
    hp = hp or HotPath(scn, device=device, city_buildings=city_buildings)
    n = scn["gnb"].shape[0]
    dist_on = dist.is_available() and dist.is_initialized()
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist_on else (1, 0)
    mine = shard_cells(n, world, rank)
    per_rank = -(-n // world)
    use_cuda = dist_on and dist.get_backend(group) == "nccl"
    dev = f"cuda:{hp.device}" if use_cuda else "cpu"

    def gather(local_rows, width):
        """rows of this rank's cells (in `mine` order) -> rows of all cells in cell order"""
        if world == 1:
            return local_rows
        buf = torch.full((per_rank, width), float("nan"), dtype=torch.float64, device=dev)
        if len(mine):
            buf[: len(mine)] = torch.from_numpy(local_rows).to(dev)
        out = torch.empty((world * per_rank, width), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, buf, group=group)
        out = out.cpu().numpy().reshape(world, per_rank, width)
        full = np.empty((n, width))
        for r in range(world):
            cells_r = shard_cells(n, world, r)
            full[cells_r] = out[r, : len(cells_r)]
        return full

    records = []
    for f in range(frame0, frame0 + n_frames):
        ue_los, tgt_los = hp.los_flags(f)
        summary = None
        if interference:
            summary = gather(hp.tx_summary(mine, f), 5)
        cellps = [hp.W.cfg5_cell_params(scn, c, f, ue_los, tgt_los)[0] for c in mine]
        interfs = [hp.interference(cp, summary) if summary is not None else None for cp in cellps]
        coms = hp.comm_frames(cellps, interfs) if cellps else []
        rows = [pack_record(hp, com, hp.sensing_cpi(cp)) for cp, com in zip(cellps, coms)]
        width = len(rows[0]) if rows else len(pack_record(hp, hp.comm_frame({"ueIDs": np.zeros(0, int)}), float("nan")))
        local = np.stack(rows) if rows else np.zeros((0, width))
        records.append(gather(local, width))
        if on_frame is not None:
            on_frame(f, records[-1])
    return records
