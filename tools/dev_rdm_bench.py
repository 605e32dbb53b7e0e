"""Developer micro-benchmark of the RDM+CFAR pipeline (not the contract bench; see bench.py)."""
import importlib
import sys
import time

import numpy as np
import torch

PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
sys.path.insert(0, ".")
rdm = importlib.import_module(PKG + ".sensing._rdm")


def run(nSc, nSym, nAnts, nIFFT, nFFT, rows, cols, B, iters=20, variant=0):
    plan = rdm.RangeDopplerPlan(nSc, nSym, nAnts, nIFFT, nFFT, rows, cols, 1e-9, max_batch=B)
    plan.set_variant(variant)
    g = torch.Generator(device="cuda").manual_seed(0)
    rx = torch.view_as_complex(torch.randn(B, nAnts, nSym, nSc, 2, device="cuda", generator=g))
    tx = torch.view_as_complex(torch.randn(B, nAnts, nSym, nSc, 2, device="cuda", generator=g))
    pw = torch.empty(B, nAnts, nFFT, nIFFT, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    for _ in range(3):
        plan.run_dev(rx, tx, B, pw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run_dev(rx, tx, B, pw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = np.array(ts)
    alg = (16 * nSc * nSym * nAnts + 4 * nIFFT * nFFT * nAnts) * B
    ms = np.median(ts)
    print(f"variant={variant} B={B} {nSc}x{nSym}x{nAnts}->{nIFFT}x{nFFT}: median {ms*1e3:.1f} us  min {ts.min()*1e3:.1f} us  "
          f"alg {alg/1e6:.1f} MB  {alg/ms/1e6:.0f} GB/s  ({alg/ms/1e6/6570:.2%} of 6570)  maps/s {B/ms*1e3:.0f}")
    plan.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    variants = [int(a) for a in sys.argv[1:]] or [0, 3, 1]
    for variant in variants:
        for B in (1, 4, 16):
            run(3276, 168, 8, 4096, 256, (42, 411), (118, 140), B, variant=variant)
    run(624, 840, 4, 1024, 1024, (6, 52), (427, 599), 1)
    run(624, 840, 4, 1024, 1024, (6, 52), (427, 599), 8)
