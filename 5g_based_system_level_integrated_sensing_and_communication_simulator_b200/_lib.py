"""ctypes binding of lib/libisac_b200.so (C ABI: include/isac_b200.h).

No fallback of any kind: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libisac_b200.so")

_lib = None
_lock = threading.Lock()
_ctxs = {}


class IsacError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"isac status {status}: {msg}")
        self.status = status


STATUS = {
    0: "OK", 1: "INVALID_ARG", 2: "CUDA", 3: "NO_DEVICE", 4: "UNSUPPORTED", 5: "CFAR_WINDOW",
    6: "NO_LOS_TARGET", 7: "NUM_DETS_ZERO", 8: "CAPACITY",
}


class RdmConfig(C.Structure):
    _fields_ = [
        ("nSc", C.c_int32), ("nSym", C.c_int32), ("nAnts", C.c_int32),
        ("nIFFT", C.c_int32), ("nFFT", C.c_int32),
        ("cutRow0", C.c_int32), ("cutRow1", C.c_int32), ("cutCol0", C.c_int32), ("cutCol1", C.c_int32),
        ("guardRows", C.c_int32), ("guardCols", C.c_int32), ("trainRows", C.c_int32), ("trainCols", C.c_int32),
        ("maxBatch", C.c_int32),
        ("pfa", C.c_double), ("kaiserBeta", C.c_double),
    ]


class DoaConfig(C.Structure):
    _fields_ = [
        ("isUpa", C.c_int32), ("nAnts", C.c_int32), ("nX", C.c_int32), ("nY", C.c_int32),
        ("d", C.c_double), ("aGran", C.c_double), ("aMax", C.c_double),
        ("eGran", C.c_double), ("eMax", C.c_double),
    ]


class Music2dConfig(C.Structure):
    _fields_ = [
        ("nSc", C.c_int32), ("nSym", C.c_int32), ("nAnts", C.c_int32),
        ("scsHz", C.c_double), ("fc", C.c_double), ("Tsri", C.c_double),
        ("rMax", C.c_double), ("vZone", C.c_double),
        ("doa", DoaConfig),
        ("numDetsOverride", C.c_int32),
    ]


class EchoConfig(C.Structure):
    _fields_ = [
        ("T", C.c_int64), ("nTx", C.c_int32), ("nTargets", C.c_int32),
        ("fc", C.c_double), ("fs", C.c_double), ("N0", C.c_double),
        ("range", C.c_void_p), ("velocity", C.c_void_p), ("largeScaleFading", C.c_void_p),
        ("steeringVec", C.c_void_p), ("los", C.c_void_p),
        ("nfft", C.c_int32), ("nSc", C.c_int32), ("nSymTx", C.c_int32), ("symbolsPerSubframe", C.c_int32),
        ("cpLengths", C.c_void_p),
    ]


class CsiConfig(C.Structure):
    _fields_ = [
        ("nPorts", C.c_int32), ("N1", C.c_int32), ("N2", C.c_int32), ("O1", C.c_int32), ("O2", C.c_int32),
        ("codebookMode", C.c_int32), ("nSizeBWP", C.c_int32), ("nStartBWP", C.c_int32), ("subbandSize", C.c_int32),
        ("pmiSubband", C.c_int32), ("cqiSubband", C.c_int32), ("K", C.c_int32), ("L", C.c_int32), ("nRx", C.c_int32),
        ("subsetRestriction", C.c_void_p), ("i2Restriction", C.c_void_p), ("riRestriction", C.c_uint8 * 8),
        ("nRE", C.c_int32), ("reK", C.c_void_p), ("reL", C.c_void_p), ("nPanels", C.c_int32),
    ]


class CdlConfig(C.Structure):
    _fields_ = [
        ("profile", C.c_int32), ("delaySpread", C.c_double), ("fc", C.c_double), ("maxDoppler", C.c_double),
        ("txSize", C.c_int32 * 3), ("rxSize", C.c_int32 * 3), ("txPattern38901", C.c_int32),
        ("rxPattern38901", C.c_int32), ("seed", C.c_uint64),
    ]


NOISE_NONE, NOISE_TENSOR, NOISE_PHILOX = 0, 1, 2
MAX_PEAKS = 64


def _declare(lib):
    vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
    P = C.POINTER
    sigs = {
        "isac_create": ([P(vp), C.c_int], C.c_int),
        "isac_destroy": ([vp], C.c_int),
        "isac_last_error": ([vp], C.c_char_p),
        "isac_set_stream": ([vp, vp], C.c_int),
        "isac_use_own_stream": ([vp], C.c_int),
        "isac_synchronize": ([vp], C.c_int),
        "isac_profile_enable": ([vp, i32], C.c_int),
        "isac_profile_collect": ([vp, vp, vp, vp], C.c_int),
        "isac_profile_timeline": ([vp, vp, i32, vp, vp, vp, P(i32)], C.c_int),
        "isac_version": ([], C.c_char_p),
        "isac_rdm_plan_create": ([vp, P(RdmConfig), P(vp)], C.c_int),
        "isac_rdm_plan_destroy": ([vp], C.c_int),
        "isac_rdm_plan_info": ([vp, P(f64), P(i32), P(i32)], C.c_int),
        "isac_rdm_plan_set_variant": ([vp, i32], C.c_int),
        "isac_rdm_cfar_dev": ([vp, vp, vp, i32, vp], C.c_int),
        "isac_cfar2d_dev": ([vp, vp, i32], C.c_int),
        "isac_rdm_get_detections": ([vp, i32, i32, vp, vp, vp], C.c_int),
        "isac_rdm_get_power": ([vp, i32, vp], C.c_int),
        "isac_rdm_cfar_host": ([vp, vp, vp, i32, i32, vp, vp, vp, vp], C.c_int),
        "isac_music_doa_host": ([vp, P(DoaConfig), vp, i32, P(i32), vp, P(i32), vp, vp], C.c_int),
        "isac_dev_malloc": ([vp, C.c_uint64, P(vp)], C.c_int),
        "isac_dev_free": ([vp, vp], C.c_int),
        "isac_memcpy_h2d": ([vp, vp, vp, C.c_uint64], C.c_int),
        "isac_memcpy_d2h": ([vp, vp, vp, C.c_uint64], C.c_int),
        "isac_chest_plan_create": ([vp, i32, i32, i32, i32, C.c_int64, vp, vp, i32, i32, i32, i32, i32, P(vp)], C.c_int),
        "isac_chest_plan_destroy": ([vp], C.c_int),
        "isac_channel_estimate_dev": ([vp, vp, i32, vp, vp], C.c_int),
        "isac_chest_get_nvar": ([vp, i32, vp], C.c_int),
        "isac_city_create": ([vp, i32, vp, vp, P(vp)], C.c_int),
        "isac_city_destroy": ([vp], C.c_int),
        "isac_city_check_los_host": ([vp, i32, vp, vp, i32, vp], C.c_int),
        "isac_city_check_los_dev": ([vp, i32, vp, vp, i32, vp], C.c_int),
        "isac_ofdm_modulate_dev": ([vp, vp, i32, i32, i32, i32, i32, vp, f64, vp, P(C.c_int64)], C.c_int),
        "isac_ofdm_modulate_ex_dev": ([vp, vp, i32, i32, i32, C.c_int64, i32, i32, vp, f64, i32, i32, vp, C.c_int64, C.c_int64,
                                       P(C.c_int64)], C.c_int),
        "isac_pathloss_host": ([vp, i32, f64, i32, vp, vp, vp, vp], C.c_int),
        "isac_link_budget_dev": ([vp, vp, C.c_int64, i32, vp, f64], C.c_int),
        "isac_thermal_noise_power": ([f64, f64, f64, P(f64)], C.c_int),
        "isac_dft_channel_matrix": ([i32, i32, vp], C.c_int),
        "isac_doa_scan_host": ([vp, P(DoaConfig), i32, vp, i32, P(i32), vp, P(i32), vp, vp], C.c_int),
        "isac_sense_plan_create": ([vp, P(RdmConfig), P(DoaConfig), f64, f64, P(vp)], C.c_int),
        "isac_sense_plan_destroy": ([vp], C.c_int),
        "isac_sense_plan_rdm": ([vp], vp),
        "isac_fft2d_dev": ([vp, vp, vp, i32, vp], C.c_int),
        "isac_fft2d_collect": ([vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
        "isac_fft2d_get_spectrum": ([vp, i32, vp], C.c_int),
        "isac_fft2d_host": ([vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
        "isac_music2d_dev": ([vp, P(Music2dConfig), vp, vp, P(i32), vp, P(i32), vp, P(i32), vp, P(i32), vp, vp,
                              P(i32)], C.c_int),
        "isac_antenna_covariance_dev": ([vp, vp, C.c_int64, i32, vp], C.c_int),
        "isac_type1sp_codebook": ([P(CsiConfig), i32, i32, P(i32), vp], C.c_int),
        "isac_type1mp_codebook": ([P(CsiConfig), i32, i32, P(i32), vp], C.c_int),
        "isac_type1mp_codebook_from_table": ([P(CsiConfig), i32, i32, P(i32), vp], C.c_int),
        "isac_pusch_codebook": ([i32, i32, P(i32), vp], C.c_int),
        "isac_pmi_plan_create": ([vp, P(CsiConfig), i32, i32, P(vp)], C.c_int),
        "isac_pmi_plan_destroy": ([vp], C.c_int),
        "isac_precoded_sinr_host": ([vp, vp, i32, i32, f64, vp, i32, i32, vp], C.c_int),
        "isac_pmi_plan_set_kernel": ([vp, i32], C.c_int),
        "isac_csi_plan_set_kernel": ([vp, i32], C.c_int),
        "isac_csi_plan_mp_dims": ([vp, i32, P(i32)], C.c_int),
        "isac_pmi_plan_mp_dims": ([vp, P(i32)], C.c_int),
        "isac_pmi_plan_info": ([vp, P(i32), P(i32), P(i32), P(i32), vp, vp], C.c_int),
        "isac_dl_pmi_select_dev": ([vp, vp, vp, i32], C.c_int),
        "isac_dl_pmi_collect": ([vp, i32, vp, vp, vp], C.c_int),
        "isac_dl_pmi_get_info": ([vp, i32, vp, vp], C.c_int),
        "isac_csi_plan_create": ([vp, P(CsiConfig), i32, P(vp)], C.c_int),
        "isac_csi_plan_destroy": ([vp], C.c_int),
        "isac_ri_select_dev": ([vp, vp, vp, i32, vp, vp, vp], C.c_int),
        "isac_cqi_select_dev": ([vp, i32, vp, vp, i32, vp, i32, vp, P(i32), vp, vp, vp], C.c_int),
        "isac_csi_report_dev": ([vp, vp, vp, i32, vp, i32, i32, vp, vp, vp, vp, P(i32)], C.c_int),
        "isac_csi_report_enqueue_dev": ([vp, vp, vp, i32], C.c_int),
        "isac_csi_report_finish": ([vp, vp, i32, i32, vp, vp, vp, vp, P(i32)], C.c_int),
        "isac_ul_pmi_select_dev": ([vp, i32, vp, i32, i32, i32, i32, f64, i32, i32, vp, vp, vp, P(i32), P(i32), P(i32)],
                                   C.c_int),
        "isac_ul_pmi_select_batch_dev": ([vp, i32, vp, i32, i32, i32, i32, f64, i32, i32, i32, vp, vp, P(i32), P(i32), vp], C.c_int),
        "isac_ul_pmi_select_batch_enqueue_dev": ([vp, i32, vp, i32, i32, i32, i32, f64, i32, i32], C.c_int),
        "isac_ul_pmi_select_batch_finish": ([vp, i32, vp, vp, P(i32), P(i32), vp], C.c_int),
        "isac_cdl_generate_batch_dev": ([vp, i32, i32, f64, i32, vp, vp, vp], C.c_int),
        "isac_prg_precode_dev": ([vp, i32, i32, i32, vp, vp, i32, i32, vp, i32, i32, vp, vp], C.c_int),
        "isac_prg_precode_batch_dev": ([vp, i32, i32, i32, vp, vp, i32, i32, vp, i32, i32, i32, vp, vp], C.c_int),
        "isac_cdl_create": ([vp, P(CdlConfig), P(vp)], C.c_int),
        "isac_cdl_destroy": ([vp], C.c_int),
        "isac_cdl_set_kernel": ([vp, i32], C.c_int),
        "isac_cdl_get_rays": ([vp, P(i32), P(i32), P(i32), P(i32), vp, vp, vp, vp], C.c_int),
        "isac_cdl_generate_dev": ([vp, i32, f64, i32, vp, f64, vp], C.c_int),
        "isac_radar_channel_dev": ([vp, P(EchoConfig), vp, vp, i32, C.c_uint64, vp], C.c_int),
        "isac_mono_static_sensing_dev": ([vp, P(EchoConfig), vp, vp, i32, C.c_uint64, vp, P(i32)], C.c_int),
        "isac_mono_static_sensing_host": ([vp, P(EchoConfig), vp, vp, i32, C.c_uint64, vp, P(i32)], C.c_int),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    return sigs


def exported_symbols():
    """Names the header declares (used by the CPU-only symbol test)."""
    import re
    hdr = os.path.join(os.path.dirname(PKG_DIR), "include", "isac_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"^(?:int|const char\*|isac_[a-z0-9_]+\s*\*)\s*(isac_[a-z0-9_]+)\s*\(", txt, flags=re.M)))


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                    "(python -m 5g_..._b200.build). There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            _declare(lib)
            _lib = lib
    return _lib


def check(status, ctx_handle=None):
    if status != 0:
        lib = load()
        msg = lib.isac_last_error(ctx_handle)
        raise IsacError(status, f"{STATUS.get(status, '?')}: {msg.decode() if msg else ''}")


class Context:
    """One library context (device + stream)."""

    def __init__(self, device=0):
        lib = load()
        h = C.c_void_p()
        st = lib.isac_create(C.byref(h), int(device))
        if st != 0:
            msg = lib.isac_last_error(None)
            raise IsacError(st, f"{STATUS.get(st, '?')}: {msg.decode() if msg else ''}")
        self.handle = h
        self.device = int(device)
        self.lib = lib

    def set_stream(self, cuda_stream_ptr):
        check(self.lib.isac_set_stream(self.handle, C.c_void_p(cuda_stream_ptr or 0)), self.handle)

    def use_own_stream(self):
        check(self.lib.isac_use_own_stream(self.handle), self.handle)

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def synchronize(self):
        check(self.lib.isac_synchronize(self.handle), self.handle)

    PROF_SLOTS = ("rdm_range", "rdm_doppler", "cfar", "echo_demod", "covariance", "music", "pmi_sinr", "cdl",
                  "prg_precode", "ul_tpmi", "ofdm_modulate", "channel_estimate")

    def profile_enable(self, on=True):
        check(self.lib.isac_profile_enable(self.handle, 1 if on else 0), self.handle)

    def profile_collect(self):
        """-> ({slot: (total_ms, count)}, launches) since the last collect; synchronises the stream."""
        ms = np.zeros(16)
        cnt = np.zeros(16, dtype=np.int32)
        n = C.c_int64()
        check(self.lib.isac_profile_collect(self.handle, ptr(ms), ptr(cnt), C.byref(n)), self.handle)
        out = {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(self.PROF_SLOTS) if cnt[i]}
        return out, int(n.value)

    def profile_timeline(self, base_event, max_rec=65536):
        """[(slot name, begin_ms, end_ms)] of the kernel groups recorded since the last collect, relative to the torch CUDA
        event ``base_event`` (call before ``profile_collect``, which clears the records)."""
        slots = np.zeros(max_rec, dtype=np.int32)
        t0 = np.zeros(max_rec)
        t1 = np.zeros(max_rec)
        n = C.c_int32()
        check(self.lib.isac_profile_timeline(self.handle, C.c_void_p(base_event.cuda_event), max_rec, ptr(slots), ptr(t0), ptr(t1),
                                             C.byref(n)), self.handle)
        return [(self.PROF_SLOTS[slots[i]], float(t0[i]), float(t1[i])) for i in range(n.value)]

    def close(self):
        if self.handle:
            self.lib.isac_destroy(self.handle)
            self.handle = None


def get_context(device=None) -> Context:
    """Process-wide context for a device (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _lock:
        ctx = _ctxs.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _ctxs[device] = ctx
    return ctx


def as_c64(a):
    """Host complex64, Fortran (MATLAB column-major) order, no copy when already so."""
    return np.asfortranarray(np.asarray(a, dtype=np.complex64))


def ptr(a):
    """Raw pointer of a numpy array / torch tensor (device or host) / None."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())
