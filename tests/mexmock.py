"""Test helper: build a MEX gateway of matlab/mex together with the functional mx/mex mock (tests/mex_stub/mex_mock.cpp) into
a shared object linked against libisac_b200.so, and call its mexFunction from Python with NumPy arrays / dicts standing in
for MATLAB arrays / structs.  Test infrastructure only (MATLAB does not exist in the build image)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "5g_based_system_level_integrated_sensing_and_communication_simulator_b200"
LIBDIR = os.path.join(ROOT, PKG, "lib")
OUT = os.path.join(ROOT, "tests", "_mexbuild")
CLS = {np.dtype(np.float64): 6, np.dtype(np.float32): 7, np.dtype(np.uint8): 9, np.dtype(np.int32): 12, np.dtype(np.uint64): 15,
       np.dtype(np.complex128): 6, np.dtype(np.complex64): 7, np.dtype(np.bool_): 3}
DT = {(6, False): np.float64, (7, False): np.float32, (9, False): np.uint8, (12, False): np.int32, (15, False): np.uint64,
      (6, True): np.complex128, (7, True): np.complex64, (3, False): np.bool_}


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.identifier = ident


_cache = {}


def build(name):
    """-> ctypes.CDLL of <name>.cpp + the mock (rebuilt when a source is newer)."""
    if name in _cache:
        return _cache[name]
    gxx = shutil.which("g++")
    assert gxx, "g++ not found"
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(ROOT, "matlab", "mex", name + ".cpp"), os.path.join(ROOT, "tests", "mex_stub", "mex_mock.cpp")]
    deps = srcs + [os.path.join(ROOT, "matlab", "mex", "isac_mex_common.h"), os.path.join(ROOT, "include", "isac_b200.h"),
                   os.path.join(ROOT, "tests", "mex_stub", "mex.h")]
    so = os.path.join(OUT, name + ".so")
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cmd = [gxx, "-std=c++17", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I",
               os.path.join(ROOT, "tests", "mex_stub"), *srcs, "-L", LIBDIR, "-lisac_b200", "-Wl,-rpath," + LIBDIR, "-o", so]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(so)
    vp = C.c_void_p
    lib.mxCreateNumericArray.restype = vp
    lib.mxCreateNumericArray.argtypes = [C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_int]
    lib.mxCreateStructMatrix.restype = vp
    lib.mxCreateStructMatrix.argtypes = [C.c_size_t, C.c_size_t, C.c_int, vp]
    lib.mxSetField.argtypes = [vp, C.c_size_t, C.c_char_p, vp]
    lib.mxGetField.restype = vp
    lib.mxGetField.argtypes = [vp, C.c_size_t, C.c_char_p]
    lib.mock_data.restype = vp
    lib.mock_data.argtypes = [vp]
    lib.mock_nbytes.restype = C.c_size_t
    lib.mock_nbytes.argtypes = [vp]
    lib.mock_class.argtypes = [vp]
    lib.mxIsComplex.restype = C.c_bool
    lib.mxIsComplex.argtypes = [vp]
    lib.mxIsStruct.restype = C.c_bool
    lib.mxIsStruct.argtypes = [vp]
    lib.mxGetNumberOfDimensions.restype = C.c_size_t
    lib.mxGetNumberOfDimensions.argtypes = [vp]
    lib.mxGetDimensions.restype = C.POINTER(C.c_size_t)
    lib.mxGetDimensions.argtypes = [vp]
    lib.mock_field_count.argtypes = [vp]
    lib.mock_field_name.restype = C.c_char_p
    lib.mock_field_name.argtypes = [vp, C.c_int]
    lib.mock_destroy.argtypes = [vp]
    lib.mock_mex_call.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_char_p, C.c_char_p, C.c_int]
    _cache[name] = lib
    return lib


def to_mx(lib, v):
    if isinstance(v, dict):
        s = lib.mxCreateStructMatrix(1, 1, 0, None)
        for k, x in v.items():
            lib.mxSetField(s, 0, k.encode(), to_mx(lib, x))
        return s
    a = np.asarray(v if v is not None else np.zeros((0, 0)))
    if a.dtype == np.int64:
        a = a.astype(np.float64)      # MATLAB numeric literals are double
    if a.dtype not in CLS:
        a = a.astype(np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)          # column vector
    a = np.asfortranarray(a)
    dims = (C.c_size_t * a.ndim)(*a.shape)
    m = lib.mxCreateNumericArray(a.ndim, dims, CLS[a.dtype], 1 if np.iscomplexobj(a) else 0)
    assert lib.mock_nbytes(m) == a.nbytes, (lib.mock_nbytes(m), a.nbytes)
    if a.nbytes:
        C.memmove(lib.mock_data(m), a.ctypes.data, a.nbytes)
    return m


def from_mx(lib, m):
    if not m:
        return None
    if lib.mxIsStruct(m):
        return {lib.mock_field_name(m, i).decode(): from_mx(lib, lib.mxGetField(m, 0, lib.mock_field_name(m, i)))
                for i in range(lib.mock_field_count(m))}
    nd = lib.mxGetNumberOfDimensions(m)
    d = lib.mxGetDimensions(m)
    shape = tuple(int(d[i]) for i in range(nd))
    dt = DT[(lib.mock_class(m), bool(lib.mxIsComplex(m)))]
    n = int(np.prod(shape))
    buf = (C.c_char * lib.mock_nbytes(m)).from_address(lib.mock_data(m)) if n else b""
    return np.frombuffer(bytes(buf), dtype=dt, count=n).reshape(shape, order="F").copy()


def call(name, nlhs, *args):
    """Run the gateway like MATLAB would: [out1, ...] = name(args...).  Raises MexError(identifier, message)."""
    lib = build(name)
    ins = [to_mx(lib, a) for a in args]
    prhs = (C.c_void_p * max(len(ins), 1))(*ins)
    plhs = (C.c_void_p * max(nlhs, 1))()
    eid, emsg = C.create_string_buffer(1024), C.create_string_buffer(1024)
    rc = lib.mock_mex_call(nlhs, plhs, len(ins), prhs, eid, emsg, 1024)
    for m in ins:
        lib.mock_destroy(m)
    if rc:
        raise MexError(eid.value.decode(), emsg.value.decode())
    outs = [from_mx(lib, plhs[i]) for i in range(max(nlhs, 1))]
    for i in range(max(nlhs, 1)):
        if plhs[i]:
            lib.mock_destroy(plhs[i])
    return outs if nlhs > 1 else outs[0]
