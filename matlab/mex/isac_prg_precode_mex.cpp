/* [antsym, antind] = isac_prg_precode_mex(siz, nstartgrid, portsym, portind, F)
 *   siz = [K L P]; portsym single complex [NRE x nLayers]; portind int32 [NRE x nLayers] (1-based); F single complex [nLayers x P x NPRG]
 *   antsym single complex [NRE x P]; antind uint32-compatible int32 [NRE x P]
 * Marshals communication.phyLayer.prgPrecode (+communication/+phyLayer/prgPrecode.m:53; call sites gNBPhy.m:822,826). */
#include "isac_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 5) mexErrMsgIdAndTxt("isac:prgPrecode:nargin", "five inputs required");
    const char* fn = "prgPrecode";
    const double* siz = mxGetDoubles(prhs[0]);
    const int K = (int)siz[0], L = (int)siz[1], nStart = (int)mxGetScalar(prhs[1]);
    const mxArray *sym = prhs[2], *ind = prhs[3], *F = prhs[4];
    require_csingle(sym, fn, "portsym");
    require_csingle(F, fn, "F");
    const int NRE = dim_of(sym, 0), nLayers = dim_of(sym, 1), P = dim_of(F, 1), NPRG = dim_of(F, 2);
    if (dim_of(F, 0) != nLayers || mxGetNumberOfElements(ind) != mxGetNumberOfElements(sym))
        mexErrMsgIdAndTxt("isac:prgPrecode:size", "portsym, portind and F disagree in size");
    const mwSize od[2] = {(mwSize)NRE, (mwSize)P};
    plhs[0] = mxCreateNumericArray(2, od, mxSINGLE_CLASS, mxCOMPLEX);
    mxArray* oi = mxCreateNumericArray(2, od, mxINT32_CLASS, mxREAL);
    if (NRE > 0) {
        const size_t nOut = (size_t)NRE * P;
        DevBuf ds(mxGetComplexSingles(sym), mxGetNumberOfElements(sym) * sizeof(mxComplexSingle), fn);
        DevBuf di(mxGetInt32s(ind), mxGetNumberOfElements(ind) * sizeof(int32_t), fn);
        DevBuf dF(mxGetComplexSingles(F), mxGetNumberOfElements(F) * sizeof(mxComplexSingle), fn);
        DevBuf os(nullptr, nOut * sizeof(mxComplexSingle), fn), oid(nullptr, nOut * sizeof(int32_t), fn);
        int rc = isac_prg_precode_dev(isac_mex_ctx(), K, L, nStart, ds.p, (const int32_t*)di.p, NRE, nLayers, dF.p, P, NPRG, os.p,
                                      (int32_t*)oid.p);
        if (!rc) rc = isac_memcpy_d2h(isac_mex_ctx(), mxGetComplexSingles(plhs[0]), os.p, nOut * sizeof(mxComplexSingle));
        if (!rc) rc = isac_memcpy_d2h(isac_mex_ctx(), mxGetInt32s(oi), oid.p, nOut * sizeof(int32_t));
        isac_mex_check(rc, fn);
    }
    if (nlhs > 1) plhs[1] = oi;
}
