"""The MEX gateways under matlab/mex cannot be built here (no MATLAB): type-check each of them with g++ -fsyntax-only
against include/isac_b200.h and the declaration-only mex.h stand-in of tests/mex_stub, so that every call they make into the
C ABI matches the header, and check that every gateway INTEGRATION.md names exists and that every .m shim calls one."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEX = os.path.join(ROOT, "matlab", "mex")
GATEWAYS = sorted(glob.glob(os.path.join(MEX, "*_mex.cpp")))


def test_gateways_present():
    names = {os.path.basename(g) for g in GATEWAYS}
    assert {"isac_fft2d_mex.cpp", "isac_mono_static_mex.cpp", "isac_dl_pmi_mex.cpp", "isac_csi_report_mex.cpp",
            "isac_ul_pmi_mex.cpp", "isac_prg_precode_mex.cpp", "isac_doa_mex.cpp", "isac_music2d_mex.cpp"} <= names
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for n in re.findall(r"isac_\w+_mex\.cpp", text):
        assert n in names, f"INTEGRATION.md names {n}, which does not exist"


@pytest.mark.parametrize("src", GATEWAYS, ids=[os.path.basename(g) for g in GATEWAYS])
def test_gateway_type_checks(src):
    gxx = shutil.which("g++")
    assert gxx, "g++ not found"
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration",
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "mex_stub"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_gateways_call_only_declared_symbols():
    header = open(os.path.join(ROOT, "include", "isac_b200.h")).read()
    declared = set(re.findall(r"\b(isac_\w+)\s*\(", header))
    for src in GATEWAYS + [os.path.join(MEX, "isac_mex_common.h")]:
        used = set(re.findall(r"\b(isac_\w+)\s*\(", open(src).read()))
        used -= {"isac_mex_ctx", "isac_mex_check", "isac_mex_cleanup"}
        used = {u for u in used if not u.endswith("_mex")}
        assert used <= declared, (os.path.basename(src), sorted(used - declared))


def test_shims_call_existing_gateways():
    names = {os.path.basename(g)[:-4] for g in GATEWAYS}
    shims = glob.glob(os.path.join(ROOT, "matlab", "+*", "**", "*.m"), recursive=True)
    assert shims
    for m in shims:
        if os.path.basename(m).startswith("isac"):      # configuration helpers shared by the shims
            continue
        called = set(re.findall(r"\b(isac_\w+_mex)\b", open(m).read()))
        assert called, f"{m} calls no gateway"
        assert called <= names, (m, sorted(called - names))
