"""The MEX gateways linked and executed against a functional mock of MATLAB's mx/mex API (tests/mexmock.py): on a machine
without a GPU every gateway must build, link against libisac_b200.so, marshal its inputs and surface the library's
"no CUDA device" status as the MATLAB error the .m shims (and the reference's try/catch, cellSimulation.m:196-202) expect."""
import glob
import os

import numpy as np
import pytest

import mexmock as M

GATEWAYS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(M.ROOT, "matlab", "mex", "*_mex.cpp")))


@pytest.mark.parametrize("name", GATEWAYS)
def test_gateway_builds_and_links(name):
    lib = M.build(name)
    assert hasattr(lib, "mexFunction")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_argument_errors_come_back_as_matlab_errors():
    with pytest.raises(M.MexError) as e:
        M.call("isac_ul_pmi_mex", 3, 1.0, np.zeros((24, 1, 2, 2), np.complex64))          # too few inputs
    assert e.value.identifier == "isac:pmiSelect:nargin"
    with pytest.raises(M.MexError) as e:
        M.call("isac_ul_pmi_mex", 3, 1.0, np.zeros((24, 1, 2, 2), np.complex128), 0.1, 4.0)   # double instead of single
    assert e.value.identifier == "isac:pmiSelect:type"
    with pytest.raises(M.MexError) as e:                                                      # portind does not match portsym
        M.call("isac_prg_precode_mex", 2, np.array([24.0, 14.0, 4.0]), 0.0, np.zeros((10, 2), np.complex64),
               np.zeros((9, 2), np.int32), np.zeros((2, 4, 1), np.complex64))
    assert e.value.identifier == "isac:prgPrecode:size"
    with pytest.raises(M.MexError) as e:
        M.call("isac_fft2d_mex", 1, {"nIFFT": 64.0}, np.zeros((4, 4, 2), np.complex64), np.zeros((4, 4, 2), np.complex64))
    assert e.value.identifier == "isac:mex:missingField"


@pytest.mark.skipif(not _no_gpu(), reason="needs a machine without a CUDA device")
def test_no_device_is_a_matlab_error_not_a_fallback():
    """There is no CPU fallback: without a GPU the first gateway call fails with isac:create:noDevice."""
    with pytest.raises(M.MexError) as e:
        M.call("isac_ul_pmi_mex", 3, 1.0, np.ones((24, 1, 2, 2), np.complex64), 0.1, 4.0)
    assert e.value.identifier == "isac:create:noDevice"
    assert "no CPU fallback" in str(e.value)
