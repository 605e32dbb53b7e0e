/* [i1, i2, sinrPerRE, sinrPerSubband, W, reK, reL, mpDims] = isac_dl_pmi_mex(cfg, nLayers, H, nVar)
 *   cfg : struct built by matlab/+communication/+phyLayer/dlPMISelect.m from the validated reportConfig
 *   H   : single complex [K x L x nRx x P];  nVar: double scalar
 *   i1 [3x1], i2 [nSB x 1] (1-based, NaN = not reported); sinrPerRE [nRE x nLayers x i2 x i11 x i12 x i13] at the CSI-RS REs
 *   (reK, reL: their 1-based subscripts); sinrPerSubband [nSB x nLayers x ...]; W complex double [P x nLayers x ...]
 *   Type1MultiPanel (cfg.nPanels = Ng >= 2): the 9-D index set [i20 i21 i22 | i11 i12 i13 i141 i142 i143] comes back flattened in
 *   MATLAB linear order -- i2 = i20,i21,i22 and i1(3) = i13,i141,i142,i143 combined, arrays [.. x i2 x i11 x i12 x i13'] -- and
 *   mpDims = [i20 i21 i22 i13 i141 i142 i143] lengths lets the .m shim reshape / ind2sub them (zeros for a single-panel report).
 * Marshals communication.phyLayer.dlPMISelect (+communication/+phyLayer/dlPMISelect.m:1). */
#include "isac_mex_common.h"

static PlanCache<isac_pmi_plan> g_plans(isac_pmi_plan_destroy);
static void drop_plans(void) { g_plans.clear(); }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 4) mexErrMsgIdAndTxt("isac:dlPMISelect:nargin", "four inputs required");
    const char* fn = "dlPMISelect";
    const mxArray* H = prhs[2];
    require_csingle(H, fn, "H");
    const int nLayers = (int)mxGetScalar(prhs[1]);
    const double nVar = mxGetScalar(prhs[3]);
    CsiCfg cs(prhs[0], dim_of(H, 2));
    if (dim_of(H, 0) != cs.c.K || dim_of(H, 1) != cs.c.L || dim_of(H, 3) != cs.c.nPorts)
        mexErrMsgIdAndTxt("nr5g:hDLPMISelect:InvalidChannelDims", "H must be K-by-L-by-nRxAnts-by-NumCSIRSPorts");
    isac_ctx* ctx = isac_mex_ctx();
    g_plan_cleanup = drop_plans;
    std::string key = cs.key();
    key_add(key, nLayers);
    isac_pmi_plan* plan = g_plans.find(key);          /* one plan per (report configuration, rank) */
    if (!plan) {
        isac_mex_check(isac_pmi_plan_create(ctx, &cs.c, nLayers, 1, &plan), fn);
        g_plans.put(key, plan);
    }
    int32_t dims[4] = {0, 0, 0, 0}, nSB = 0, nCqiSB = 0, nRE = 0;
    isac_mex_check(isac_pmi_plan_info(plan, dims, &nSB, &nCqiSB, &nRE, nullptr, nullptr), fn);
    std::vector<int32_t> reK(nRE), reL(nRE);
    isac_mex_check(isac_pmi_plan_info(plan, dims, &nSB, &nCqiSB, &nRE, reK.data(), reL.data()), fn);
    const size_t nCand = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
    std::vector<double> i1(3), i2(nSB), S((size_t)nRE * nLayers * nCand), Sb((size_t)nSB * nLayers * nCand);
    int32_t mp[7] = {0, 0, 0, 0, 0, 0, 0};
    isac_mex_check(isac_pmi_plan_mp_dims(plan, mp), fn);
    int rc;
    {
        DevBuf Hd(mxGetComplexSingles(H), mxGetNumberOfElements(H) * sizeof(mxComplexSingle), fn);
        rc = isac_dl_pmi_select_dev(plan, Hd.p, &nVar, 1);
        if (!rc) rc = isac_dl_pmi_collect(plan, 1, i1.data(), i2.data(), nullptr);
        if (!rc) rc = isac_dl_pmi_get_info(plan, 1, S.data(), Sb.data());
    }
    if (rc) g_plans.drop(plan);
    isac_mex_check(rc, fn);
    std::vector<double> W(2 * (size_t)cs.c.nPorts * nLayers * nCand);
    int32_t wd[9];
    if (cs.c.nPanels >= 2) isac_mex_check(isac_type1mp_codebook(&cs.c, cs.c.nPanels, nLayers, wd, W.data()), fn);   /* :1351 */
    else isac_mex_check(isac_type1sp_codebook(&cs.c, nLayers, 0, wd, W.data()), fn);   /* dlPMISelect.m:853 */
    const std::vector<mwSize> tail = {(mwSize)dims[0], (mwSize)dims[1], (mwSize)dims[2], (mwSize)dims[3]};
    auto with_tail = [&](mwSize a, mwSize b) { std::vector<mwSize> d = {a, b}; d.insert(d.end(), tail.begin(), tail.end()); return d; };
    plhs[0] = double_array({3, 1}, i1.data());
    if (nlhs > 1) plhs[1] = double_array({(mwSize)nSB, 1}, i2.data());
    if (nlhs > 2) plhs[2] = double_array(with_tail((mwSize)nRE, (mwSize)nLayers), S.data());
    if (nlhs > 3) plhs[3] = double_array(with_tail((mwSize)nSB, (mwSize)nLayers), Sb.data());
    if (nlhs > 4) plhs[4] = complex_double_array(with_tail((mwSize)cs.c.nPorts, (mwSize)nLayers), W.data());
    if (nlhs > 5) { std::vector<double> v(reK.begin(), reK.end()); plhs[5] = double_array({(mwSize)nRE, 1}, v.data()); }
    if (nlhs > 6) { std::vector<double> v(reL.begin(), reL.end()); plhs[6] = double_array({(mwSize)nRE, 1}, v.data()); }
    if (nlhs > 7) { std::vector<double> v(mp, mp + 7); plhs[7] = double_array({1, 7}, v.data()); }
}
