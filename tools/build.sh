#!/bin/bash
# rebuild lib/libisac_b200.so from the repo root (what __graft_entry__.build() does)
cd "$(dirname "$0")/.." && python -c "
import importlib; b=importlib.import_module('5g_based_system_level_integrated_sensing_and_communication_simulator_b200.build'); b.build_library(verbose=True)" 2>&1 | tail -${1:-3}
