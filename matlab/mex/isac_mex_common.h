/* Shared helpers of the MEX gateways (R2018a interleaved-complex API: mxComplexSingle == float2).
 * Build (on a machine with MATLAB + CUDA):  mex -R2018a -I../../include isac_fft2d_mex.cpp -L<repo>/.../lib -lisac_b200
 * The gateways only marshal; all arithmetic is behind the C ABI of include/isac_b200.h. */
#pragma once
#include "isac_b200.h"
#include "mex.h"
#include <string>

static isac_ctx* g_ctx = nullptr;

static void isac_mex_cleanup(void) {
    if (g_ctx) { isac_destroy(g_ctx); g_ctx = nullptr; }
}

static isac_ctx* isac_mex_ctx(void) {
    if (!g_ctx) {
        int st = isac_create(&g_ctx, 0);
        if (st) mexErrMsgIdAndTxt("isac:create:noDevice", "%s", isac_last_error(nullptr));
        mexAtExit(isac_mex_cleanup);
        mexLock();   /* keep the context (and its device buffers) alive between calls */
    }
    return g_ctx;
}

/* Map a failing status to a MATLAB error so the reference's try/catch (cellSimulation.m:196-202) keeps working. */
static void isac_mex_check(int st, const char* fn) {
    if (st == ISAC_OK) return;
    const std::string id = std::string("isac:") + fn + ":status" + std::to_string(st);
    mexErrMsgIdAndTxt(id.c_str(), "%s", isac_last_error(g_ctx));
}

static double field_scalar(const mxArray* s, const char* name) {
    const mxArray* f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("isac:mex:missingField", "missing field %s", name);
    return mxGetScalar(f);
}
