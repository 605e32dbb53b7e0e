"""Short driver for ncu captures of the COMM kernels (cfg2: 8 ports [2 2], 8 rx, 273 PRB, batch of 8 UEs)."""
import importlib
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
P = importlib.import_module(PKG)
sys.argv = [sys.argv[0]]
import bench

comm = bench.CommWorkload(P, 4, 0)
for i in range(2):
    comm.step(i)
torch.cuda.synchronize()
