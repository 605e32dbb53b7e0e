/* H = isac_cdl_mex(cfg, K, scsHz, symTimes, t0)
 *   cfg      : struct profile (0 CDL-A, 2 CDL-C, 3 CDL-D), delaySpread, fc, maxDoppler, txSize[3], rxSize[3],
 *              txPattern38901, rxPattern38901, seed  -- the properties cdl.m:56-65 sets on the nrCDLChannel object
 *   K        : subcarriers (12*NSizeGrid);  scsHz: subcarrier spacing in Hz
 *   symTimes : double [1 x L] start times of the OFDM symbols of the slot (s);  t0: start time of the slot (s)
 *   H        : single complex [K x L x nRx x nTx] -- the array nrChannelEstimate hands the CSI functions (uePhy.m:897)
 * Marshals the device CDL generator (isac_cdl_create / isac_cdl_generate_dev): the frequency-domain replacement of
 * nrCDLChannel filtering + nrChannelEstimate (uePhy.m:731,897; gNBPhy.m:840,1030).  Channels are cached per configuration
 * (ray tables are drawn once per seed, like the toolbox object). */
#include "isac_mex_common.h"

static PlanCache<isac_cdl_channel> g_channels(isac_cdl_destroy);
static void drop_channels(void) { g_channels.clear(); }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    const char* fn = "cdlChannelMatrix";
    if (nrhs != 5) mexErrMsgIdAndTxt("isac:cdlChannelMatrix:nargin", "five inputs required");
    const mxArray* cfg = prhs[0];
    isac_cdl_config c = {};
    c.profile = (int32_t)field_scalar(cfg, "profile"); c.delaySpread = field_scalar(cfg, "delaySpread");
    c.fc = field_scalar(cfg, "fc"); c.maxDoppler = field_scalar(cfg, "maxDoppler");
    const std::vector<int32_t> ts = field_int32s(cfg, "txSize"), rs = field_int32s(cfg, "rxSize");
    if (ts.size() < 3 || rs.size() < 3) mexErrMsgIdAndTxt("isac:cdlChannelMatrix:size", "txSize / rxSize need [M N P]");
    for (int i = 0; i < 3; ++i) { c.txSize[i] = ts[i]; c.rxSize[i] = rs[i]; }
    c.txPattern38901 = (int32_t)field_scalar(cfg, "txPattern38901"); c.rxPattern38901 = (int32_t)field_scalar(cfg, "rxPattern38901");
    c.seed = (uint64_t)field_scalar(cfg, "seed");
    const int K = (int)mxGetScalar(prhs[1]);
    const double scs = mxGetScalar(prhs[2]), t0 = mxGetScalar(prhs[4]);
    const int L = (int)mxGetNumberOfElements(prhs[3]);
    if (!mxIsDouble(prhs[3]) || L < 1 || K < 1) mexErrMsgIdAndTxt("isac:cdlChannelMatrix:type", "symTimes must be a non-empty double vector");
    isac_ctx* ctx = isac_mex_ctx();
    g_plan_cleanup = drop_channels;
    std::string key;
    key_add(key, c);
    isac_cdl_channel* ch = g_channels.find(key);
    if (!ch) {
        isac_mex_check(isac_cdl_create(ctx, &c, &ch), fn);
        g_channels.put(key, ch);
    }
    const int nRx = c.rxSize[0] * c.rxSize[1] * c.rxSize[2], nTx = c.txSize[0] * c.txSize[1] * c.txSize[2];
    const mwSize dims[4] = {(mwSize)K, (mwSize)L, (mwSize)nRx, (mwSize)nTx};
    plhs[0] = mxCreateNumericArray(4, dims, mxSINGLE_CLASS, mxCOMPLEX);
    const size_t bytes = (size_t)K * L * nRx * nTx * sizeof(mxComplexSingle);
    int rc;
    {
        DevBuf Hd(nullptr, bytes, fn);
        rc = isac_cdl_generate_dev(ch, K, scs, L, mxGetDoubles(prhs[3]), t0, Hd.p);
        if (!rc) rc = isac_memcpy_d2h(ctx, mxGetComplexSingles(plhs[0]), Hd.p, bytes);
    }
    isac_mex_check(rc, fn);
}
