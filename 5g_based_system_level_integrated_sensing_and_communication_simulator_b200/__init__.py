"""B200-native (sm_100a) hot path of the 5G system-level ISAC simulator.

Host-side mirror of the reference's MATLAB package API for the one data-parallel path this
repository accelerates:

    sensing.radarParams, sensing.monoStaticSensing, sensing.channelModels.basicRadarChannel,
    sensing.detection.cfar2D, sensing.estimation.fft2D / music2D / doaEstimation.{music, mvdrBF, digitalBF},
    sensing.ofdmModulate, sensing.postProcessing.getRMSE, networkTopology.blockages.city.checkLoS,
    communication.phyLayer.{dlPMISelect, riSelect, cqiSelect, pmiSelect, precodedSINR,
    sinrPerSubband, prgPrecode}, communication.pmiType1SinglePanelCodebook,
    simulation.cellSimulation (hot-path driver).

Every numeric kernel runs in hand-written CUDA behind the C ABI of include/isac_b200.h
(lib/libisac_b200.so, bound with ctypes in _lib.py).  There is NO CPU fallback: importing the
package is cheap, but the first call that needs the library raises if it is missing or if no
CUDA device is present.  The package never imports ``oracle/``.

The directory name starts with a digit, so import it with
``importlib.import_module("5g_based_system_level_integrated_sensing_and_communication_simulator_b200")``
(or through the root-level ``isac_b200`` alias module).
"""
from . import _lib  # noqa: F401  (does not load the .so until first use)

__all__ = ["_lib", "sensing", "communication", "simulation", "networkTopology", "workloads"]
__version__ = "0.1.0"


def __getattr__(name):
    import importlib
    if name in ("sensing", "communication", "simulation", "networkTopology", "workloads", "build"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
