/* [RI, i1, i2, CQI, SINRPerSubbandPerCW, mpDims] = isac_csi_report_mex(cfg, H, nVar, SINRTable, rankCap, mode [, nLayers])
 *   mode 0: fused UE report (uePhy.m:900-907): RI = min(riSelect(..), rankCap), then cqiSelect at that rank
 *   mode 1: [RI,PMISet] = riSelect(carrier,csirs,reportConfig,H,nVar)            (riSelect.m:1; CQI = [])
 *   mode 2: [CQI,PMISet] = cqiSelect(carrier,csirs,reportConfig,nLayers,H,nVar,SINRTable) (cqiSelect.m:1; RI = nLayers)
 *   H: single complex [K x L x nRx x P]; CQI [cqiRows x 2] (second codeword column NaN when nLayers <= 4).
 *   SINRPerSubbandPerCW (mode 2 only, else []): [rows x 2] linear SINR per codeword, wideband value first when there is more
 *   than one CQI subband (cqiSelect.m:610-633).
 *   Type1MultiPanel (cfg.nPanels = Ng >= 2): i1(3) and i2 are linear indices into the flattened index sets of the reported rank;
 *   mpDims = [i20 i21 i22 i13 i141 i142 i143] lengths at that rank for ind2sub in the .m shim (zeros otherwise). */
#include "isac_mex_common.h"

static PlanCache<isac_csi_plan> g_plans(isac_csi_plan_destroy);
static void drop_plans(void) { g_plans.clear(); }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 6) mexErrMsgIdAndTxt("isac:csiReport:nargin", "six or seven inputs required");
    const char* fn = "csiReport";
    const mxArray* H = prhs[1];
    require_csingle(H, fn, "H");
    const double nVar = mxGetScalar(prhs[2]);
    const mxArray* tab = prhs[3];
    const int rankCap = (int)mxGetScalar(prhs[4]);
    const int mode = (int)mxGetScalar(prhs[5]);
    const int nLayers = nrhs > 6 ? (int)mxGetScalar(prhs[6]) : 0;
    CsiCfg cs(prhs[0], dim_of(H, 2));
    isac_ctx* ctx = isac_mex_ctx();
    g_plan_cleanup = drop_plans;
    const std::string key = cs.key();
    isac_csi_plan* plan = g_plans.find(key);          /* one plan per report configuration, reused by every later call */
    if (!plan) {
        isac_mex_check(isac_csi_plan_create(ctx, &cs.c, 1, &plan), fn);
        g_plans.put(key, plan);
    }
    const int nSBmax = cs.c.nSizeBWP + 1;          /* upper bound on the subband count */
    std::vector<double> RI(1, mxGetNaN()), i1(3), i2(nSBmax), cqi(2 * (size_t)(nSBmax + 1), mxGetNaN()), sinrCW(2 * (size_t)(nSBmax + 1));
    int32_t cqiRows = 0;
    int rc;
    {
        DevBuf Hd(mxGetComplexSingles(H), mxGetNumberOfElements(H) * sizeof(mxComplexSingle), fn);
        if (mode == 1)
            rc = isac_ri_select_dev(plan, Hd.p, &nVar, 1, RI.data(), i1.data(), i2.data());
        else if (mode == 2) {
            rc = isac_cqi_select_dev(plan, nLayers, Hd.p, &nVar, 1, mxGetDoubles(tab), (int32_t)mxGetNumberOfElements(tab), cqi.data(),
                                     &cqiRows, i1.data(), i2.data(), sinrCW.data());
            RI[0] = nLayers;
        } else
            rc = isac_csi_report_dev(plan, Hd.p, &nVar, 1, mxGetDoubles(tab), (int32_t)mxGetNumberOfElements(tab), rankCap, RI.data(),
                                     i1.data(), i2.data(), cqi.data(), &cqiRows);
    }
    if (rc) g_plans.drop(plan);
    isac_mex_check(rc, fn);
    int nSB = 1;                                   /* PMI subbands of the report (dlPMISelect.m:465-501) */
    if (cs.c.pmiSubband && cs.c.subbandSize > 0 && cs.c.nSizeBWP >= 24)   /* first subband ends on a SubbandSize boundary */
        nSB = (cs.c.nStartBWP % cs.c.subbandSize + cs.c.nSizeBWP + cs.c.subbandSize - 1) / cs.c.subbandSize;
    plhs[0] = mxCreateDoubleScalar(RI[0]);
    if (nlhs > 1) plhs[1] = double_array({3, 1}, i1.data());
    if (nlhs > 2) plhs[2] = double_array({(mwSize)nSB, 1}, i2.data());
    if (nlhs > 3) plhs[3] = double_array({(mwSize)cqiRows, 2}, cqi.data());
    if (nlhs > 4) {
        int nCqiSB = 1;                            /* CQI subbands: the same partition rule with the CQI mode */
        if (cs.c.cqiSubband && cs.c.subbandSize > 0 && cs.c.nSizeBWP >= 24)
            nCqiSB = (cs.c.nStartBWP % cs.c.subbandSize + cs.c.nSizeBWP + cs.c.subbandSize - 1) / cs.c.subbandSize;
        const int rowsFull = nCqiSB > 1 ? nCqiSB + 1 : 1;
        plhs[4] = mode == 2 ? double_array({(mwSize)rowsFull, 2}, sinrCW.data()) : mxCreateDoubleMatrix(0, 0, mxREAL);
    }
    if (nlhs > 5) {
        int32_t mp[7] = {0, 0, 0, 0, 0, 0, 0};
        const double r = RI[0];
        if (cs.c.nPanels >= 2 && r == r && r >= 1) isac_mex_check(isac_csi_plan_mp_dims(plan, (int32_t)r, mp), fn);
        std::vector<double> v(mp, mp + 7);
        plhs[5] = double_array({1, 7}, v.data());
    }
}
