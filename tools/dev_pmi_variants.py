"""Dev micro-benchmark of the CSI report (32 UEs, cfg2 shapes) with CUDA events; run once per kernel variant
(ISAC_PMI_FUSED / ISAC_PAIR_G / ISAC_PAIR_T are read when the library first launches)."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
P = importlib.import_module(PKG)
sys.argv = [sys.argv[0]]
import bench

comm = bench.CommWorkload(P, 4, 0)
lib, ctx, C = comm.ctx.lib, comm.ctx, comm.C
ptr, check = comm._lib.ptr, comm._lib.check
ctx.use_torch_stream()
check(lib.isac_cdl_generate_batch_dev(comm.dl_handles, comm.nb, comm.K, comm.SCS, 14, ptr(comm.sym_t), ptr(comm.t0_dl), ptr(comm.H)), ctx.handle)


def report():
    check(lib.isac_csi_report_dev(comm.csi_plan, ptr(comm.H), ptr(comm.nvar), comm.nb, ptr(comm.table), comm.table.size, 4,
                                  ptr(comm.RI), ptr(comm.i1), ptr(comm.i2), ptr(comm.cqi), C.byref(comm.rows)), ctx.handle)


def enqueue():
    check(lib.isac_csi_report_enqueue_dev(comm.csi_plan, ptr(comm.H), ptr(comm.nvar), comm.nb), ctx.handle)


def finish():
    check(lib.isac_csi_report_finish(comm.csi_plan, ptr(comm.table), comm.table.size, 4, ptr(comm.RI), ptr(comm.i1), ptr(comm.i2),
                                     ptr(comm.cqi), C.byref(comm.rows)), ctx.handle)


for _ in range(3):
    report()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    enqueue()
    b.record()
    finish()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
tag = {k: os.environ.get(k) for k in ("ISAC_PMI_FUSED", "ISAC_PAIR_G", "ISAC_PAIR_T", "ISAC_PAIR_BP")}
print(tag, "report kernels: median %.1f us, min %.1f us (32 UEs)" % (float(np.median(ts)), min(ts)))
print("RI", comm.RI[:8].tolist(), "i1", comm.i1[:6].tolist(), "cqi", comm.cqi[:4].tolist())
