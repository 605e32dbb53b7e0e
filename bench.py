#!/usr/bin/env python
"""bench.py — throughput of the B200-native ISAC hot path (contract: see the task statement).

Metric (BASELINE.json): cell-subframes/sec.  One *step* = ``--frames-per-step`` (default 10) consecutive 10 ms frames
(10 subframes each) of hot-path work for each of ``cells_per_gpu`` cells of BASELINE config 2 (1 gNB, 8 UE, 4 targets, 8x8,
273 PRB @ 30 kHz, TDD DDDSU -> 168 DL symbols per frame), so the default K = 20 steps give a >= 1 s timed region:
    sensing : mono-static echo synthesis + OFDM demod of the frame's DL waveform (K1+K2),
              2D-FFT range-Doppler map + 2D CA-CFAR (K3+K4), antenna covariance + MUSIC DoA (K5+K6);
    comm    : per UE and CSI-RS occasion RI/PMI/CQI selection over the Type-I codebook (K7-K9), UL TPMI
              selection per SRS occasion (K12), PRG precoding of the DL slots (K10), with the channel
              matrices H produced on the device by the CDL generator (K11)          [stages listed in
              config.stages are the ones inside the timed region].
`value`  : cell-subframes/s with every input resident in HBM (device-timed, CUDA events, max over ranks).
`e2e`    : same metric through the package's host API (pinned host buffers in, host results out).
`roofline`: dominant kernel group, algorithmic bytes / CUDA-event time measured inside the timed region.
`cpu_baseline` / ``--impl reference``: the float64 NumPy restatement of the reference (oracle/) on the
host cores — MATLAB itself cannot run here (no MATLAB/Octave, closed toolboxes); labelled "port".
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
SUBFRAMES_PER_FRAME = 10
WORKLOAD = "cfg2: 1 gNB, 8 UE, 4 targets, 8x8, 273 PRB @30 kHz, frames of 168 DL symbols per cell"
STAGES = ["echo_demod", "rdm_2dfft", "cfar2d", "covariance", "music_doa", "cdl_dl_generate", "csi_report(ri+pmi+cqi)",
          "cdl_ul_generate", "ul_tpmi_select", "prg_precode"]
# measured sustained DFMA rate of one B200 (tools/micro/fp64_peak.cu, profiles/r2_fp64_peak.txt): the roofline of the PMI / SINR
# group, which runs its Cholesky factorisations on the FP64 pipe
FP64_PEAK_TFLOPS = 32.8
# FP64 flops of one 32-UE cfg2 CSI report (8 ports (2,2), ranks 1-8, 273 REs): DFMA = 2, DMUL / DADD = 1, thread-level counts
# of the SASS instructions executed (ncu source page of profiles/r2_pmi_fused_v3_summary.txt, scaled by UEs / 32)
PMI_FLOPS_PER_UE_REPORT = 2.39e8


def bench_config(cells, frames):
    """The `config` object of the JSON line -- identical for the B200 arm and the reference arm."""
    return {"workload": WORKLOAD, "cells_per_gpu": cells, "frames_per_step": frames, "stages": STAGES,
            "l2_policy": "inputs (waveforms + grids of all cells, 330 MB per frame at 4 cells) larger than L2; no flush"}


# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch group from the committed ncu capture
# profiles/r1_rdm_warm_traffic_v5.txt (--cache-control none: 357.9 MB per chain of 4 cfg2 map-sets = 89.5 MB per map-set
# against 104.0 MB algorithmic: with the L2 eviction hints the range profiles stay in L2, and part of the power-map
# write-back (33.5 MB per map-set, dirty in the 126 MB L2) drains after the chain's last kernel, outside the capture).
TRAFFIC = {"rdm_2dfft+cfar": lambda cells: int(89.5e6 * cells)}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.time(), line.strip()))
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.samples:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def build_cell_inputs(W, seed):
    """Synthetic cfg2 inputs of one cell-frame (host, complex64): senTxGrid, senTxWave."""
    grid, txw = W.sensing_tx("cfg2", seed)
    return grid.astype(np.complex64), txw.astype(np.complex64)


class CommWorkload:
    """COMM share of one cfg2 cell-frame (20 slots @30 kHz, 8 UEs, 8x8, 273 PRB, CDL-C):
    4 CSI-RS occasions (period 5 slots, setupCSIRS.m:11) x 8 UEs: CDL channel matrix + fused RI/PMI/CQI report;
    20 SRS occasions (period 8 slots, setupSRS.m:13; 2.5 per UE per frame): UL CDL + TPMI selection over the full band
    (the reference hands pmiSelect the interpolated estimate of the SRS symbol, gNBPhy.m:1030-1035);
    12 DL slots: PRG precoding of a full-band 2-layer PDSCH (12 symbols) and its DM-RS."""

    N_UE, NRB, SCS = 8, 273, 30e3

    def __init__(self, P, cells, device):
        import ctypes as C
        import torch
        self.C, self.torch, self.P, self.cells = C, torch, P, cells
        self._lib = P._lib
        ph = importlib.import_module(PKG + ".communication.phyLayer")
        cm = importlib.import_module(PKG + ".communication.channelModels")
        self.ctx = self._lib.get_context(device)
        K = self.NRB * 12
        self.K = K
        self.carrier = {"NSizeGrid": self.NRB, "NStartGrid": 0, "SymbolsPerSlot": 14}
        self.csirs = {"NumCSIRSPorts": 8, "NumRB": self.NRB, "RBOffset": 0, "SubcarrierLocations": 1, "SymbolLocations": 0,
                      "Density": "one"}
        self.rc = {"NSizeBWP": self.NRB, "NStartBWP": 0, "PanelDimensions": (2, 2), "CodebookMode": 1, "PMIMode": "Subband",
                   "CQIMode": "Subband", "SubbandSize": 16}
        self.cs = ph._csi_struct(self.carrier, self.csirs, self.rc, 8)
        self.table = np.ascontiguousarray(P.communication.setupSINRtoCQIMappingTable()["downlinkSINR90pc"], dtype=np.float64)
        num = P.workloads.ofdm_numerology(self.NRB, 30)
        starts = P.workloads.symbol_starts(num, 14) / num["SampleRate"]
        self.sym_t = np.ascontiguousarray(starts, dtype=np.float64)
        self.sym13 = np.ascontiguousarray(starts[13:14], dtype=np.float64)
        self.dl = [[cm.CDLChannel("CDL-C", TransmitAntennaArraySize=(1, 4, 2), ReceiveAntennaArraySize=(2, 2, 2),
                                  Seed=1000 * c + u, device=device) for u in range(self.N_UE)] for c in range(cells)]
        self.ul = [[cm.CDLChannel("CDL-C", TransmitAntennaArraySize=(1, 1, 2), ReceiveAntennaArraySize=(1, 4, 2),
                                  Seed=5000 + 1000 * c + u, device=device) for u in range(self.N_UE)] for c in range(cells)]
        dev = f"cuda:{device}"
        self.nb = cells * self.N_UE                                        # UEs of all cells report together (same slot)
        assert self.nb <= 32, "cells_per_gpu <= 4 (PMI batch limit 32)"
        _, self.csi_plan = ph._csi_plan(self.cs, self.nb)
        self.H = torch.empty((self.nb, 8, 8, 14, K), dtype=torch.complex64, device=dev)       # [cell*ue][P][R][L][K]
        self.nul = cells * 4                                                # 4 UEs share an SRS slot (setupSRS.m:13)
        self.hest = torch.empty((self.nul, 2, 8, 1, K), dtype=torch.complex64, device=dev)    # [cell*ue][P][R][1][K]
        self.nvar = np.full(self.nb, 10 ** (-15 / 10))
        nSB = (self.NRB + 15) // 16
        self.RI = np.zeros(self.nb)
        self.i1 = np.zeros((3, self.nb), order="F")
        self.i2 = np.zeros((nSB, self.nb), order="F")
        self.cqi = np.zeros((nSB + 1) * 2 * self.nb)
        self.rows = C.c_int32()
        self.ul_pmi = np.zeros((nSB + 1) * self.nul)
        self.ul_sinr = np.zeros((nSB + 1) * 3 * self.nul)
        self.ul_none = np.zeros(self.nul, dtype=np.int32)
        self.ul_n = (C.c_int32(), C.c_int32())
        self.dl_handles = (C.c_void_p * self.nb)(*[self.dl[c][u].handle for c in range(cells) for u in range(self.N_UE)])
        self.ul_handles = [(C.c_void_p * self.nul)(*[self.ul[c][4 * grp + q].handle for c in range(cells) for q in range(4)])
                           for grp in range(2)]
        self.t0_dl = np.zeros(self.nb)
        self.t0_ul = np.zeros(self.nul)
        # PDSCH: full band, symbols 2..13 (12 symbols), 2 layers, PRG size 2 -> 137 PRGs; DM-RS: symbol 2, 6 REs/PRB
        L, nu, Pp = 14, 2, 8
        self.nprg = (self.NRB + 1) // 2
        g = torch.Generator(device=dev).manual_seed(5)
        k = torch.arange(K, device=dev)
        def alloc(syms, ksel):
            pos = (ksel[:, None] + K * torch.tensor(syms, device=dev)[None, :]).T.reshape(-1)
            ind = torch.stack([pos + 1 + K * L * j for j in range(nu)], dim=0).to(torch.int32).contiguous()   # [nu][NRE] == MATLAB [NRE x nu]
            ind = ind[None].repeat(cells, 1, 1).contiguous()                                                 # every cell: same allocation
            sym = torch.view_as_complex(torch.randn(cells, nu, pos.numel(), 2, device=dev, generator=g)).contiguous()
            return sym, ind, pos.numel()
        self.pdsch = alloc(list(range(2, 14)), k)
        self.dmrs = alloc([2], k[::2])
        self.F = torch.view_as_complex(torch.randn(cells, self.nprg, Pp, nu, 2, device=dev, generator=g)).contiguous()  # == MATLAB [nu x P x NPRG x cells]
        self.out_sym = torch.empty(cells * self.pdsch[2] * Pp, dtype=torch.complex64, device=dev)
        self.out_ind = torch.empty(cells * self.pdsch[2] * Pp, dtype=torch.int32, device=dev)
        self.slot_t = 0.5e-3
        # argument tuples of the per-slot calls, marshalled once (every buffer is allocated above and reused each frame):
        # the precoding kernels take ~13 us, so per-call pointer marshalling in Python would leave the GPU waiting
        ptr = self._lib.ptr
        h = self.ctx.handle
        self.prg_args = [(h, self.K, 14, 0, ptr(sym), ptr(ind), nre, 2, ptr(self.F), 8, self.nprg, self.cells, ptr(self.out_sym),
                          ptr(self.out_ind)) for sym, ind, nre in (self.pdsch, self.dmrs)]
        self.cdl_dl_args = (self.dl_handles, self.nb, self.K, self.SCS, 14, ptr(self.sym_t), ptr(self.t0_dl), ptr(self.H))
        self.cdl_ul_args = [(self.ul_handles[g], self.nul, self.K, self.SCS, 1, ptr(self.sym13), ptr(self.t0_ul), ptr(self.hest))
                            for g in range(2)]
        self.csi_enq_args = (self.csi_plan, ptr(self.H), ptr(self.nvar), self.nb)
        self.csi_fin_args = (self.csi_plan, ptr(self.table), self.table.size, 4, ptr(self.RI), ptr(self.i1), ptr(self.i2),
                             ptr(self.cqi), C.byref(self.rows))
        self.ul_enq_args = (h, 2, ptr(self.hest), self.K, 1, 8, 2, 0.05, 16, self.nul)
        self.ul_fin_args = (h, self.ul_pmi.size // self.nul, ptr(self.ul_pmi), ptr(self.ul_sinr), C.byref(self.ul_n[0]),
                            C.byref(self.ul_n[1]), ptr(self.ul_none))

    def step(self, step, fillers=()):
        """One frame of COMM work.  `fillers`: up to four callables, one per CSI-RS occasion, that enqueue independent work
        (the bench passes pieces of the cell's sensing pass) behind the report's kernels before the host blocks on the
        report's results -- the GPU then stays busy during the host-side RI / CQI tails."""
        lib, ctx, C = self.ctx.lib, self.ctx, self.C
        fillers = list(fillers)
        ptr, check = self._lib.ptr, self._lib.check
        ctx.use_torch_stream()
        frame_t0 = 0.010 * step
        srs = {3: 0, 4: 1, 11: 0, 12: 1, 19: 0}   # UEs 0-3 at slots 3, 11, 19, UEs 4-7 at slots 4, 12 (period 8, offset 3 + floor(ue/4))
        ul_pending = csi_pending = False
        for slot in range(20):                    # slot order of the frame (TDD DDDSU @30 kHz); the GPU's cells advance together
            if slot % 5 < 3:                      # DL slot: PDSCH + DM-RS precoding of every cell (gNBPhy.m:822,826)
                if csi_pending:
                    csi_pending = False
                    check(lib.isac_csi_report_finish(*self.csi_fin_args), ctx.handle)
                for a in self.prg_args:
                    check(lib.isac_prg_precode_batch_dev(*a), ctx.handle)
            if slot % 5 == 2:                     # CSI-RS occasion (period 5 slots): channel of all UEs + fused RI/PMI/CQI report
                self.t0_dl[:] = frame_t0 + slot * self.slot_t
                check(lib.isac_cdl_generate_batch_dev(*self.cdl_dl_args), ctx.handle)
                check(lib.isac_csi_report_enqueue_dev(*self.csi_enq_args), ctx.handle)
                if fillers:
                    fillers.pop(0)()
                    ctx.use_torch_stream()
                    check(lib.isac_csi_report_finish(*self.csi_fin_args), ctx.handle)
                else:                             # nothing to put behind the report: collect it where it is first needed (the
                    csi_pending = True            #   next DL slot's precoding), behind the S / U slots' SRS work
            if ul_pending:                        # TPMI results of the previous slot's SRS occasion: their kernels ran behind
                ul_pending = False                #   this slot's precoding launches, so the host only collects here
                check(lib.isac_ul_pmi_select_batch_finish(*self.ul_fin_args), ctx.handle)
            if slot in srs:                       # SRS occasion: UL channel of the 4 UEs of the group + TPMI selection
                self.t0_ul[:] = frame_t0 + slot * self.slot_t
                check(lib.isac_cdl_generate_batch_dev(*self.cdl_ul_args[srs[slot]]), ctx.handle)
                # pmiSelect sees the INTERPOLATED estimate of the SRS symbol, i.e. every subcarrier of the band
                # (Hest(:,srsSymbols,:,:) out of nrChannelEstimate, gNBPhy.m:1030-1035), not only the comb-4 REs
                check(lib.isac_ul_pmi_select_batch_enqueue_dev(*self.ul_enq_args), ctx.handle)
                ul_pending = True
        if csi_pending:                           # report of the frame's last occasion (its host tail runs behind slot 19's SRS kernels)
            check(lib.isac_csi_report_finish(*self.csi_fin_args), ctx.handle)
        if ul_pending:                            # SRS occasion in the frame's last slot
            check(lib.isac_ul_pmi_select_batch_finish(*self.ul_fin_args), ctx.handle)

    def d2h_bytes_per_step(self):
        return 4 * (self.RI.nbytes + self.i1.nbytes + self.i2.nbytes + self.cqi.nbytes) + 5 * (self.ul_pmi.nbytes + self.ul_sinr.nbytes)

    def algorithmic_bytes(self):
        K = self.K
        # bytes of H written per launch pair, averaged over the 4 DL (nb channels x 14 symbols x 8x8) and 5 UL
        # (nul channels x 1 symbol x 8x2) batched generations of a step (8 bytes per channel coefficient)
        total = 4 * self.nb * 8 * K * 14 * 8 * 8 + 5 * self.nul * 8 * K * 1 * 8 * 2
        return {"cdl": total / 9.0}


def write_timeline(path, base_event, contexts, ms_total, frames):
    """Poor man's timeline of the timed region (no nsys in the image): the CUDA-event stamps of every kernel group, merged over
    the contexts, with the idle gaps between consecutive groups and what ran on either side of them."""
    rec = sorted((r for cx in contexts for r in cx.profile_timeline(base_event)), key=lambda r: r[1])
    busy, gaps, by_pair, last_end, last_name = 0.0, [], {}, 0.0, "start"
    for name, a, b in rec:
        busy += b - a
        g = a - last_end
        if g > 0.002:
            gaps.append((g, last_name, name, a))
            k = (last_name, name)
            by_pair[k] = (by_pair.get(k, (0.0, 0))[0] + g, by_pair.get(k, (0.0, 0))[1] + 1)
        if b > last_end:
            last_end, last_name = b, name
    with open(path, "w") as f:
        f.write("# bench.py --timeline: %d kernel groups over %d frames, %.3f ms timed, %.3f ms inside groups (%.1f %%), "
                "%.3f ms in %d gaps > 2 us\n" % (len(rec), frames, ms_total, busy, 100 * busy / ms_total,
                                                 sum(g[0] for g in gaps), len(gaps)))
        f.write("# idle time by (group before, group after), per frame:\n")
        for (a, b), (t, n) in sorted(by_pair.items(), key=lambda kv: -kv[1][0]):
            f.write("#   %-14s -> %-14s %8.1f us/frame in %5.1f gaps/frame (mean %.1f us)\n" % (a, b, 1e3 * t / frames, n / frames, 1e3 * t / n))
        f.write("# one frame in the middle of the region (begin_us, duration_us, gap_before_us, group):\n")
        lo = ms_total * (frames // 2) / frames
        hi = ms_total * (frames // 2 + 1) / frames
        prev = None
        for name, a, b in rec:
            if a >= lo and a < hi:
                f.write("%10.1f %8.1f %8.1f  %s\n" % (1e3 * (a - lo), 1e3 * (b - a), 1e3 * (a - prev) if prev is not None else 0.0, name))
            prev = max(prev, b) if prev is not None else b


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = importlib.import_module(PKG)
    _lib = P._lib
    W = P.workloads
    est = importlib.import_module(PKG + ".sensing.estimation")
    echo = importlib.import_module(PKG + ".sensing._echo")
    ctx = _lib.get_context(local)
    cells = args.cells_per_gpu

    cell, car, wave = W.cell_config("cfg2")
    rp = P.sensing.radarParams(cell, car, wave)
    cf = P.sensing.detection.cfar2D(rp)
    los = cell["targetLoSConditions"]
    # per-cell inputs: distinct QPSK payloads (cells are independent, networkSimulation.m:57-60)
    host_grid, host_wave = [], []
    uniq = min(cells, 2)  # generating 47 MB waveforms on the host is slow; reuse payloads round-robin
    for c in range(uniq):
        g, w = build_cell_inputs(W, 1000 * rank + c + 1)
        host_grid.append(g)
        host_wave.append(w)
    nSc, nSym, nTx = host_grid[0].shape
    T = host_wave[0].shape[0]
    to_dev_grid = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda()
    to_dev_wave = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    tx_grid_d = torch.stack([to_dev_grid(host_grid[c % uniq]) for c in range(cells)])   # [cells][nAnts][nSym][nSc]
    tx_wave_d = [to_dev_wave(host_wave[c % uniq]) for c in range(cells)]                # each [nTx][T]
    rx_grid_d = torch.empty_like(tx_grid_d)
    # The sensing pass and the COMM slots of a cell are independent (cellSimulation.m runs the sensing pass after the slot
    # loop on the accumulated Tx grid only).  The sensing chain has its own library context (own scratch buffers and
    # profiling slots); its kernels are enqueued behind the CSI reports' kernels so that they fill the GPU while the host is
    # busy with the (synchronising) report tails.
    ctx_s = _lib.Context(local) if args.sense_ctx == "own" else ctx
    plan = est.SensePlan(rp, cf, (nSc, nSym, nTx), max_batch=cells, device=local, ctx=ctx_s)
    eargs = echo._EchoArgs(T, nTx, rp, los, car, nSym)
    import ctypes as C
    nsym_out = C.c_int32()

    def sensing_pieces(step):
        # device-resident leg: ONE stream, so that the CUDA-event intervals around each kernel group (the live roofline
        # figures) are not stretched by kernels of another stream competing for the SMs.  The sensing pass of the frame is
        # enqueued in three pieces behind the kernels of the first three CSI reports (CommWorkload.step `fillers`): it is
        # independent of the COMM slots (cellSimulation.m:191-197 runs it on the accumulated Tx grid only).
        def echo(c0, c1):
            def run():
                ctx_s.use_torch_stream()
                for c in range(c0, c1):
                    _lib.check(ctx_s.lib.isac_mono_static_sensing_dev(ctx_s.handle, C.byref(eargs.cfg), _lib.ptr(tx_wave_d[c]),
                                                                      None, _lib.NOISE_PHILOX, 7919 * step + c,
                                                                      _lib.ptr(rx_grid_d[c]), C.byref(nsym_out)), ctx_s.handle)
            return run
        half = (cells + 1) // 2
        return [echo(0, half), echo(half, cells), lambda: plan.run_dev(rx_grid_d, tx_grid_d, cells)]

    comm = CommWorkload(P, cells, local)

    F = args.frames_per_step

    def step_dev(step):
        for f in range(F):                 # one step = F consecutive frames of every cell (a >= 1 s timed region at the default K)
            comm.step(step * F + f, sensing_pieces(step * F + f))

    # ---- device-resident timing ----------------------------------------------------------------
    for i in range(args.warmup):
        step_dev(i)
    torch.cuda.synchronize()
    res = plan.collect(cells)
    n_est = [len(r["rngEst"]) for r in res]
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    for cx in (ctx, ctx_s):
        cx.profile_collect()
        cx.profile_enable(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step_dev(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    if args.timeline:
        write_timeline(args.timeline, e0, [ctx] + ([ctx_s] if ctx_s is not ctx else []), ms_total, args.steps * F)
    prof, launches = ctx.profile_collect()
    if ctx_s is not ctx:
        prof_s, launches_s = ctx_s.profile_collect()
        prof.update(prof_s)                                     # disjoint kernel groups (sensing vs COMM)
        launches += launches_s
    if os.environ.get("ISAC_BENCH_DEBUG"):
        print("prof", {k: (round(v[0], 3), v[1]) for k, v in prof.items()}, "ms_total", round(ms_total, 3), file=sys.stderr)
    for cx in (ctx, ctx_s):
        cx.profile_enable(False)
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = cells * SUBFRAMES_PER_FRAME * F * world / (ms_step * 1e-3)

    # ---- end-to-end through the host API (pinned host in, host results out) -------------------------
    # Per step and cell the transmit GRID crosses PCIe (pinned -> device); the waveform the gNB PHY derives from it
    # (txWaveform = signalAmp * nrOFDMModulate(txGrid), gNBPhy.m:599) is produced on the device by
    # isac_ofdm_modulate_dev, then the simulation-level sensing pass (cellSimulation.m:191-197) runs and estResults come
    # back to the host.  COMM results (PMI/RI/CQI/TPMI) land on the host inside comm.step().
    pin_grid = [torch.from_numpy(np.ascontiguousarray(host_grid[c % uniq].transpose(2, 1, 0))).pin_memory() for c in range(uniq)]
    # Two sets of device staging buffers: the copy stream fills set (s+1)%2 with the next step's inputs while the compute
    # stream works on set s%2, so PCIe traffic overlaps the COMM/sensing kernels of the step before.
    stage_grid = [tx_grid_d, torch.empty_like(tx_grid_d)]
    wave_d = [torch.empty((nTx, T), dtype=torch.complex64, device="cuda") for _ in range(cells)]
    num = W.ofdm_numerology(int(car["NRBsDL"]), float(car["SubcarrierSpacing"]))
    cp_len = np.ascontiguousarray(num["CyclicPrefixLengths"], dtype=np.int32)
    amp = 10.0 ** ((cell["gNBTxPower"] - 30.0) / 20.0) * np.sqrt(num["Nfft"] ** 2 / (nSc * nTx))   # signalAmp (gNBPhy.m:596-599)
    T_out = C.c_int64()
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event(), torch.cuda.Event()]     # set s holds the inputs of its step
    consumed = [torch.cuda.Event(), torch.cuda.Event()]   # the sensing pass that read set s has been enqueued and finished

    def upload(step):
        s = step % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            for c in range(cells):
                stage_grid[s][c].copy_(pin_grid[c % uniq], non_blocking=True)
            copied[s].record(copy_stream)

    def step_e2e(step, last):
        # One stream, as in the device-resident leg: the step's sensing pass (device OFDM modulation of the uploaded grids,
        # echo synthesis + demodulation, fft2D chain) is enqueued in three pieces behind the kernels of the first three CSI
        # reports, so it runs while the host is busy with the reports' RI / CQI tails.  The copy stream uploads the next
        # step's grids meanwhile.
        s = step % 2
        if not last:
            upload(step + 1)

        def echo(c0, c1, first):
            def run():
                ctx_s.use_torch_stream()
                if first:
                    torch.cuda.current_stream().wait_event(copied[s])
                for c in range(c0, c1):
                    _lib.check(ctx_s.lib.isac_ofdm_modulate_dev(ctx_s.handle, _lib.ptr(stage_grid[s][c]), nSc, nSym, nTx,
                                                                int(num["Nfft"]), int(cp_len.size), cp_len.ctypes.data, float(amp),
                                                                _lib.ptr(wave_d[c]), C.byref(T_out)), ctx_s.handle)
                    _lib.check(ctx_s.lib.isac_mono_static_sensing_dev(ctx_s.handle, C.byref(eargs.cfg), _lib.ptr(wave_d[c]), None,
                                                                      _lib.NOISE_PHILOX, 7919 * step + c, _lib.ptr(rx_grid_d[c]),
                                                                      C.byref(nsym_out)), ctx_s.handle)
            return run

        def chain():
            plan.run_dev(rx_grid_d, stage_grid[s], cells)
            consumed[s].record(torch.cuda.current_stream())

        half = (cells + 1) // 2
        comm.step(step, [echo(0, half, True), echo(half, cells, False), chain])   # CSI / TPMI reports land on the host
        return plan.collect(cells)   # D2H of detections / estimates (synchronises)

    e2e_steps = max(1, args.steps) * F   # frames: the same K steps of F frames as the device-resident leg
    torch.cuda.synchronize()
    for ev in consumed:
        ev.record(torch.cuda.current_stream())
    upload(0)
    step_e2e(0, True)
    torch.cuda.synchronize()
    assert T_out.value == T, (T_out.value, T)
    if world > 1:
        dist.barrier()
    t0 = time.time()
    upload(1)                      # every timed step's inputs cross PCIe inside the timed region
    for i in range(e2e_steps):
        out = step_e2e(i + 1, i == e2e_steps - 1)
    torch.cuda.synchronize()
    t_e2e = time.time() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = cells * SUBFRAMES_PER_FRAME * world * e2e_steps / t_e2e
    n_est_e2e = [len(r["rngEst"]) for r in out]
    h2d = F * cells * (nSc * nSym * nTx * 8)
    d2h = F * (cells * (nTx * 4 + 64 * 8 + 16) + comm.d2h_bytes_per_step())  # detections/estimates + CSI / TPMI reports

    # ---- the drop-in cost of the CSI functions: H of one CSI-RS occasion from PINNED HOST memory ---------------------------
    # dlPMISelect / riSelect / cqiSelect take H as a host array (uePhy.m:897-907: Hest comes out of nrChannelEstimate on the host);
    # through the MEX / _host path every occasion's H (8 bytes x K x 14 x nRx x P per UE) crosses PCIe before the report runs.
    host_h = None
    if not args.skip_host_h:
        Hpin = torch.empty(comm.H.shape, dtype=torch.complex64).pin_memory()
        Hpin.copy_(comm.H)
        Hdev = torch.empty_like(comm.H)
        ph = importlib.import_module(PKG + ".communication.phyLayer")
        torch.cuda.synchronize()
        reps = 3
        t0h = time.time()
        for _ in range(reps):
            Hdev.copy_(Hpin, non_blocking=True)
            ph.csiReport(comm.carrier, comm.csirs, comm.rc, Hdev, comm.nvar, comm.table, rankCap=4)
        torch.cuda.synchronize()
        t_occ = (time.time() - t0h) / reps
        frame_s = t_e2e / e2e_steps
        # a frame has 4 occasions; their H transfers replace the device CDL generation of the e2e leg
        host_h = {"ms_per_occasion": round(t_occ * 1e3, 2), "h2d_bytes_per_occasion": int(Hpin.numel() * 8),
                  "pcie_GBps": round(Hpin.numel() * 8 / t_occ / 1e9, 1),
                  "value_if_every_occasion_crossed_pcie": round(cells * SUBFRAMES_PER_FRAME * world / (frame_s + 4 * t_occ), 1),
                  "note": "one CSI-RS occasion (%d UEs): pinned host H -> device -> fused RI/PMI/CQI report -> host; the MATLAB drop-in "
                          "pays this per occasion, the device-resident pipeline (CDL or channel estimation on the GPU) does not" % comm.nb}

    sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- roofline of the dominant kernel group ---------------------------------------------------
    peak, peak_src = _peaks()
    nIFFT, nFFT = rp["nIFFT"], rp["nFFT"]
    alg = {
        "rdm_2dfft+cfar": (16 * nSc * nSym * nTx + 4 * nIFFT * nFFT * nTx) * cells,        # 104.0 MB / map-set
        "echo_demod": (8 * T * nTx + 8 * nSc * nSym * nTx),                                # per launch (one cell)
    }
    groups = {}
    rdm_ms = sum(prof.get(k, (0.0, 0))[0] for k in ("rdm_range", "rdm_doppler", "cfar"))
    rdm_n = prof.get("rdm_range", (0.0, 1))[1]
    groups["rdm_2dfft+cfar"] = (rdm_ms, rdm_n)
    groups["echo_demod"] = prof.get("echo_demod", (0.0, 1))
    for name in ("pmi_sinr", "cdl", "ul_tpmi", "prg_precode", "covariance", "music"):
        if name in prof:
            groups[name] = prof[name]
    alg.update(comm.algorithmic_bytes())
    dom = max(groups, key=lambda k: groups[k][0])
    roof = {}
    for name, (ms, n) in groups.items():
        if name == "pmi_sinr" and n and ms > 0:   # FP64-pipe bound: Cholesky / inverse-diagonal of every candidate in float64
            ach = PMI_FLOPS_PER_UE_REPORT * comm.nb / (ms / n * 1e-3) / 1e12
            roof[name] = {"bound": "fp64", "achieved": round(ach, 2), "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                          "frac": round(ach / FP64_PEAK_TFLOPS, 4), "avg_launch_us": round(ms / n * 1e3, 2),
                          "share_of_step": round(ms / ms_total, 4),
                          "peak_source": "measured DFMA rate (tools/micro/fp64_peak.cu, profiles/r2_fp64_peak.txt)",
                          "fp64_pipe_pct_ncu": 36.0, "ncu": "profiles/r2_pmi_fused_v3_summary.txt"}
            continue
        if n and ms > 0 and name not in alg:
            roof[name] = {"bound": "latency/alu", "avg_launch_us": round(ms / n * 1e3, 2), "share_of_step": round(ms / ms_total, 4)}
            continue
        if n and ms > 0:
            ach = alg[name] / (ms / n * 1e-3) / 1e9
            roof[name] = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(ach / peak, 4), "traffic": TRAFFIC.get(name, lambda c: None)(cells),
                          "avg_launch_us": round(ms / n * 1e3, 2),
                          "share_of_step": round(ms / ms_total, 4), "peak_source": peak_src}
    if rank == 0:
        line = {
            "metric": "cell_subframes_per_sec", "value": round(value, 2), "unit": "cell-subframes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (MUSIC/CFAR compare in f64)",
            "data": "synthetic",
            "config": bench_config(cells, F),
            "details": {"rd_map_sets_per_sec": round(cells * F * world / (ms_step * 1e-3), 1),
                        "rd_map_sets_per_sec_rdm_kernels": round(cells * world / (rdm_ms / max(rdm_n, 1) * 1e-3), 1) if rdm_ms > 0 else None,
                        "detections_sanity": n_est[:4], "timed_region_s": round(ms_total * 1e-3, 3),
                        "input_bytes_per_frame": int(cells * (T * nTx * 8 + nSc * nSym * nTx * 8))},
            "e2e": {"value": round(e2e_value, 2), "unit": "cell-subframes/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "api": "pinned txGrid -> device OFDM modulation (gNBPhy.m:599) -> simulation-level sensing pass "
                           "(cellSimulation.m:191-197) -> estResults on the host; CSI/TPMI reports on the host",
                    "detections_sanity": n_est_e2e[:4], "host_H_occasion": host_h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": dict(roof.get("rdm_2dfft+cfar", {}),
                             kernel="rdm_2dfft+cfar: one chain of range IFFT + Doppler FFT launches (one pair per cell) + 2 CFAR "
                                    "launches, linked by programmatic dependent launch; bytes and time are per chain"),
            "roofline_all": roof, "dominant_group": dom,
        }
        if args.cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def run_cfg5(args):
    """--workload cfg5: BASELINE config 5 (openStreetMapCity scenario: 19 gNBs, 100 UEs, 20 moving targets, shipped 273-PRB radio)
    through simulation.networkFrames.  STRONG scaling: the 19 cells are sharded block-cyclic over the N ranks (3,3,3,2,2,2,2,2 at
    N = 8); per frame the ranks all-gather the per-cell transmit summaries (inter-cell interference term) and the fixed-size
    per-cell records over NCCL -- both collectives are INSIDE the timed region.  One step = one frame of all 19 cells."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = importlib.import_module(PKG)
    sim = importlib.import_module(PKG + ".simulation")
    W = P.workloads
    if args.workload == "cfg3":      # BASELINE config 3: 7-cell hex, 4 UE/cell, 32-port Type-I reports + per-PRB SINR grid, no city
        scn, buildings = W.scenario_cfg3(), None
        wl = ("cfg3: 7-cell hex, 4 UE/cell, 64-element array on 32 CSI-RS ports (4,4), Type-I reports of every rank + per-PRB SINR "
              "grid, 273 PRB, one frame of all cells per step, cells sharded over the GPUs")
    else:
        scn = W.scenario_cfg5(radio=args.cfg5_radio)
        z = np.load(os.path.join(ROOT, "tests", "golden", "osm_city.npz"))
        off = z["fp_off"]
        buildings = [(z["fp_flat"][:, off[i]:off[i + 1]], float(z["heights"][i])) for i in range(off.size - 1)]
        wl = ("cfg5: openStreetMapCity scenario, 19 gNB / 100 UE / 20 moving targets, %d PRB, one frame of all cells per step, "
              "cells sharded over the GPUs" % W.RADIO[scn["radio"]]["nrb"])
    hp = sim.HotPath(scn, device=local, city_buildings=buildings)
    ctx = hp.ctx
    n_cells = scn["gnb"].shape[0]
    frame = [0]

    def run(n):
        recs = sim.networkFrames(scn, n, interference=True, hp=hp, frame0=frame[0])
        frame[0] += n
        return recs

    run(max(1, args.warmup))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ctx.profile_collect()
    ctx.profile_enable(True)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    recs = run(args.steps)                       # host results of every cell on every rank: this IS the end-to-end path
    torch.cuda.synchronize()
    t_wall1 = time.time()
    _, launches = ctx.profile_collect()
    ctx.profile_enable(False)
    t = t_wall1 - t_wall0
    if world > 1:
        dist.barrier()
        tt = torch.tensor([t, float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(tt[1:], op=dist.ReduceOp.SUM)
        t, launches = float(tt[0].item()), int(tt[1].item())
    sampler.stop()
    value = n_cells * SUBFRAMES_PER_FRAME * args.steps / t
    r = W.RADIO[scn["radio"]]
    nsym = 14 * int(round(3 / 5 * r["num_slots"]))
    h2d = sum(12 * r["nrb"] * nsym * r["nV"] * r["p"] * 8 for c in range(n_cells) if (scn["target_cell"] == c).any())
    if rank == 0:
        det = int(sum(x[:, -(4 + 2 * sim.REC_RNG + sim.REC_AZI)].sum() for x in recs))
        line = {"metric": "cell_subframes_per_sec", "value": round(value, 2), "unit": "cell-subframes/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 (PMI / MUSIC / CFAR compare in f64)", "data": "synthetic",
                "config": {"workload": wl, "cells": n_cells, "shard_sizes": [len(sim.shard_cells(n_cells, world, q)) for q in range(world)],
                           "collectives_in_timed_region": ["all_gather(tx summaries [cells x 5] f64)", "all_gather(records [cells x %d] f64)" % recs[0].shape[1]],
                           "l2_policy": "every cell-frame streams fresh grids / waveforms (> L2 per rank at the shipped radio); no flush"},
                "e2e": {"value": round(value, 2), "unit": "cell-subframes/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(recs[0].size * 8),
                        "api": "simulation.networkFrames: host scenario in, per-cell host records of every cell out (timed with the host clock, "
                               "max over ranks); `value` is this same end-to-end figure -- the driver has no device-only leg"},
                "gpu_launches": int(launches), "clocks": sampler.summary(t_wall0, t_wall1),
                "details": {"cell_frames_with_detections": det, "frames": args.steps}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# CPU arms: the float64 NumPy restatement of the reference (oracle/), WHOLE cell-frames, no extrapolation.
# One cfg2 cell-frame = the task list below in slot order (TDD DDDSU @30 kHz): per DL slot the PDSCH and DM-RS precoding, per
# CSI-RS occasion the channel + fused RI/PMI/CQI report of each of the 8 UEs, per SRS occasion the UL channel + TPMI selection of
# each of its 4 UEs, and at the end of the frame the sensing pass (echo synthesis + demodulation, fft2D).  Every task is
# stateless (its inputs derive from the seed), so tasks can be spread over processes and over the steps of a run.
# ----------------------------------------------------------------------------------------------
_TABLE = np.array([-3.46, 1.54, 6.54, 11.05, 13.54, 16.04, 17.54, 20.04, 22.04, 24.43, 26.93, 27.43, 29.43, 32.43, 35.43])


def cell_frame_tasks():
    tasks = []
    srs = {3, 4, 11, 12, 19}
    for slot in range(20):
        if slot % 5 < 3:
            tasks += [("prg", slot, 0), ("prg", slot, 1)]
        if slot % 5 == 2:
            tasks += [("csi", slot, u) for u in range(8)]
        if slot in srs:
            tasks += [("srs", slot, u) for u in range(4)]
    tasks.append(("sense", 20, 0))
    return tasks


def run_task(task, seed):
    """Execute one task of a cell-frame on the CPU.  Returns (seconds, fft2D seconds or 0)."""
    from oracle import cdl as OCd
    from oracle import comm as OCm
    kind, slot, u = task
    nrb, K = 273, 273 * 12
    t_sym = np.arange(14) * 35.7e-6
    t0 = time.time()
    t_fft = 0.0
    if kind == "csi":
        cfg = OCm.report_config(8, (2, 2), nrb, 0, 1, "Subband", "Subband", 16)
        re_k, re_l = OCm.csirs_first_port_res(nrb, 1, 0)
        rays = OCd.build_rays(2, 300e-9, 5.0, (1, 4, 2), (2, 2, 2), True, False, seed * 1000 + slot * 10 + u)
        H = OCd.frequency_response(rays, K, 30e3, t_sym)
        OCm.csi_report_vectorized(cfg, re_k, re_l, H, 10 ** -1.5, _TABLE)
    elif kind == "srs":
        rays = OCd.build_rays(2, 300e-9, 5.0, (1, 1, 2), (1, 4, 2), True, False, seed * 1000 + 500 + slot * 10 + u)
        h = OCd.frequency_response(rays, K, 30e3, t_sym[13:14])
        OCm.pmi_select(2, h, 0.05, 16)                      # full band, as on the GPU leg
    elif kind == "prg":
        rng = np.random.default_rng(seed * 1000 + slot * 2 + u)
        L, nu, P, nprg = 14, 2, 8, 137
        ks, syms = (np.arange(K), np.arange(2, 14)) if u == 0 else (np.arange(0, K, 2), np.array([2]))   # PDSCH / DM-RS
        pos = (ks[:, None] + K * syms[None, :]).T.reshape(-1)
        ind = np.stack([pos + 1 + K * L * j for j in range(nu)], axis=1)
        sym = rng.standard_normal((pos.size, nu)) + 1j * rng.standard_normal((pos.size, nu))
        Fm = rng.standard_normal((nu, P, nprg)) + 1j * rng.standard_normal((nu, P, nprg))
        OCm.prg_precode((K, L, P), 0, sym, ind, Fm)
    else:
        from oracle import sensing as S
        W = importlib.import_module(PKG + ".workloads")
        cell, car, wave = W.cell_config("cfg2")
        rp = S.radar_params(cell, car, wave)
        grid, txw = W.sensing_tx("cfg2", seed)              # input generation (QPSK grid + OFDM modulation) is not timed
        noise = W.std_normal_complex(txw.shape, seed + 1)
        cf = S.cfar2d_config(rp)
        t0 = time.time()
        rx = S.mono_static_sensing(txw, grid.shape, car, rp, cell["targetLoSConditions"], noise)
        t1 = time.time()
        S.fft2d(rp, cf, rx, grid)
        t_fft = time.time() - t1
    return time.time() - t0, t_fft


def _run_tasks(arg):
    tasks, seed = arg
    tot, fft = 0.0, 0.0
    for t in tasks:
        a, b = run_task(t, seed)
        tot += a
        fft += b
    return tot, fft


def _init_worker():
    """One BLAS / OpenMP thread per worker process: the pool already runs one process per core, nested threading only makes the
    processes fight over them."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass


def _warm(_):
    """Imports + one small report: the pool's processes are ready before the timed region."""
    from oracle import comm as OCm
    rng = np.random.default_rng(0)
    cfg = OCm.report_config(4, (2, 1), 24, 0, 1, "Subband", "Subband", 4)
    re_k, re_l = OCm.csirs_first_port_res(24, 1, 0)
    H = rng.standard_normal((288, 14, 2, 4)) + 1j * rng.standard_normal((288, 14, 2, 4))
    OCm.csi_report_vectorized(cfg, re_k, re_l, H, 0.1, _TABLE)
    return 0


def cpu_baseline():
    """One WHOLE cfg2 cell-frame on the host cores: its 77 tasks dealt round-robin to one process per core."""
    from concurrent.futures import ProcessPoolExecutor
    workers = max(1, min(os.cpu_count() or 1, 8))
    tasks = cell_frame_tasks()
    order = sorted(range(len(tasks)), key=lambda i: (tasks[i][0] != "sense", tasks[i][0] != "csi"))   # long tasks first
    parts = [[tasks[i] for i in order[w::workers]] for w in range(workers)]
    with ProcessPoolExecutor(max_workers=workers, initializer=_init_worker) as ex:
        list(ex.map(_warm, range(workers)))
        t0 = time.time()
        res = list(ex.map(_run_tasks, [(pt, 11) for pt in parts]))
        wall = time.time() - t0
    t_fft = sum(r[1] for r in res)
    return {"value": round(SUBFRAMES_PER_FRAME / wall, 3), "unit": "cell-subframes/s", "cores": workers, "kind": "port",
            "rd_map_sets_per_sec": round(1.0 / t_fft, 3) if t_fft > 0 else None,
            "sample": f"1 whole cfg2 cell-frame (32 CSI reports, 20 SRS reports, 24 precoding calls, sensing pass) in {wall:.1f} s wall "
                      f"on {workers} processes ({sum(r[0] for r in res):.1f} core-seconds; fft2D of its one map-set {t_fft:.1f} s on one "
                      "core); NumPy float64 restatement of the reference (MATLAB cannot run here), nothing extrapolated"}


def run_reference(args):
    """--impl reference: `workers` whole cell-frames over the K timed steps -- step s executes the s-th K-quantile of every
    frame's task list (in slot order), so the run as a whole executes complete cell-frames and nothing is extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    workers = max(1, min(os.cpu_count() or 1, 8))
    tasks = cell_frame_tasks()
    K = max(1, args.steps)
    bounds = [round(i * len(tasks) / K) for i in range(K + 1)]
    with ProcessPoolExecutor(max_workers=workers, initializer=_init_worker) as ex:
        for _ in range(args.warmup):
            list(ex.map(_warm, range(workers)))
        t0 = time.time()
        core_s, t_fft = 0.0, 0.0
        for sidx in range(K):
            chunk = tasks[bounds[sidx]: bounds[sidx + 1]]
            res = list(ex.map(_run_tasks, [(chunk, 100 + w) for w in range(workers)]))
            core_s += sum(r[0] for r in res)
            t_fft += sum(r[1] for r in res)
        wall = time.time() - t0
    value = workers * SUBFRAMES_PER_FRAME / wall
    cells = args.cells_per_gpu
    line = {"impl": "reference", "metric": "cell_subframes_per_sec", "value": round(value, 3), "unit": "cell-subframes/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(wall / K * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(cells, args.frames_per_step),
            "details": {"cell_frames_executed": workers, "core_seconds": round(core_s, 1),
                        "rd_map_sets_per_sec_one_core": round(workers / t_fft, 3) if t_fft > 0 else None,
                        "note": "bounded sample of the workload: %d whole cell-frames (one per process) spread over the %d timed steps; "
                                "value = cell-frames x 10 subframes / measured wall time" % (workers, K)},
            "cpu_baseline": {"value": round(value, 3), "unit": "cell-subframes/s", "cores": workers, "kind": "port",
                             "sample": f"{workers} whole cfg2 cell-frames over {workers} processes in {wall:.1f} s wall; NumPy float64 "
                                       "restatement of the reference (MATLAB cannot run here), nothing extrapolated"},
            "e2e": {"value": round(value, 3), "unit": "cell-subframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells-per-gpu", type=int, default=4)
    ap.add_argument("--frames-per-step", type=int, default=10,
                    help="frames of every cell per step (cfg2): 10 frames x 4 cells = 400 cell-subframes per step, a >= 1 s timed region at K = 20")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg5"],
                    help="cfg2 (default): weak scaling, cells_per_gpu cells per GPU; cfg3 / cfg5: the 7-cell 32-port and the 19-cell "
                         "openStreetMapCity scenarios, strong scaling (cells sharded over the GPUs)")
    ap.add_argument("--cfg5-radio", default="shipped", choices=["shipped", "small"])
    ap.add_argument("--timeline", default=None, help="write the CUDA-event timeline of the timed region (kernel groups and idle gaps) to this file")
    ap.add_argument("--skip-host-h", action="store_true", help="skip the host-resident-H occasion measurement of the e2e object")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--sense-ctx", default="own", choices=["own", "shared"],
                    help="library context of the sensing chain: its own (second stream in the e2e leg) or the COMM one")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload in ("cfg3", "cfg5"):
        run_cfg5(args)
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or int(os.environ.get("RANK", "0")) > 0:
        args.cpu_baseline = args.cpu_baseline and int(os.environ.get("WORLD_SIZE", "1")) == 1
    run_b200(args)


if __name__ == "__main__":
    main()
