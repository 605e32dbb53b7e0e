"""Short driver for ncu captures of the RDM kernels (cfg2 shapes, batch 4)."""
import importlib
import sys

import torch

sys.path.insert(0, ".")
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
rdm = importlib.import_module(PKG + ".sensing._rdm")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
plan = rdm.RangeDopplerPlan(3276, 168, 8, 4096, 256, (42, 411), (118, 140), 1e-9, max_batch=B)
plan.set_variant(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
g = torch.Generator(device="cuda").manual_seed(0)
rx = torch.view_as_complex(torch.randn(B, 8, 168, 3276, 2, device="cuda", generator=g))
tx = torch.view_as_complex(torch.randn(B, 8, 168, 3276, 2, device="cuda", generator=g))
pw = torch.empty(B, 8, 256, 4096, device="cuda")
for _ in range(4):
    plan.run_dev(rx, tx, B, pw)
torch.cuda.synchronize()
