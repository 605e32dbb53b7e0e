"""Import alias: ``import isac_b200`` == the package whose directory name starts with a digit."""
import importlib
import sys

_pkg = importlib.import_module("5g_based_system_level_integrated_sensing_and_communication_simulator_b200")
sys.modules[__name__] = _pkg
