"""Dev micro-benchmark: one CSI-RS occasion (batched CDL + fused RI/PMI/CQI report) at the bench's cfg2 shapes."""
import importlib
import sys
import time

import torch

sys.path.insert(0, ".")
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
P = importlib.import_module(PKG)
sys.argv = [sys.argv[0]]
import bench

comm = bench.CommWorkload(P, 4, 0)
lib, ctx, C = comm.ctx.lib, comm.ctx, comm.C
ptr, check = comm._lib.ptr, comm._lib.check


def occasion(gen=True):
    if gen:
        check(lib.isac_cdl_generate_batch_dev(comm.dl_handles, comm.nb, comm.K, comm.SCS, 14, ptr(comm.sym_t), ptr(comm.t0_dl), ptr(comm.H)), ctx.handle)
    check(lib.isac_csi_report_dev(comm.csi_plan, ptr(comm.H), ptr(comm.nvar), comm.nb, ptr(comm.table), comm.table.size, 4,
                                  ptr(comm.RI), ptr(comm.i1), ptr(comm.i2), ptr(comm.cqi), C.byref(comm.rows)), ctx.handle)


ctx.use_torch_stream()
for _ in range(3):
    occasion()
torch.cuda.synchronize()
for gen in (True, False):
    t0 = time.time()
    for _ in range(10):
        occasion(gen)
    torch.cuda.synchronize()
    print("cdl+report" if gen else "report only", round((time.time() - t0) / 10 * 1e6, 1), "us per occasion (32 UEs)")
print("RI", comm.RI[:8])
