/* [pmi, sinr, subbandIndices] = isac_ul_pmi_mex(nLayers, hest, noiseEst, bandSize)
 *   hest: single complex [K x nSym x nRx x nPorts] (zeros where there is no SRS estimate, pmiSelect.m:36-41)
 *   pmi [1 x nSB] 0-based TPMI (NaN), sinr [nSB x nTPMI], subbandIndices [nSB x 2]; scalar NaNs when no RE carries an estimate
 * Marshals communication.phyLayer.pmiSelect (+communication/+phyLayer/pmiSelect.m:28; call site gNBPhy.m:1035). */
#include "isac_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 4) mexErrMsgIdAndTxt("isac:pmiSelect:nargin", "four inputs required");
    const char* fn = "pmiSelect";
    const mxArray* h = prhs[1];
    require_csingle(h, fn, "hest");
    const int nLayers = (int)mxGetScalar(prhs[0]), bandSize = (int)mxGetScalar(prhs[3]);
    const double noiseEst = mxGetScalar(prhs[2]);
    const int K = dim_of(h, 0), nSym = dim_of(h, 1), nRx = dim_of(h, 2), nPorts = dim_of(h, 3);
    if (bandSize < 1) mexErrMsgIdAndTxt("isac:pmiSelect:bandSize", "bandSize must be positive");
    const int maxSB = (K / 12 + bandSize - 1) / bandSize + 1;
    int32_t nT = 0;
    isac_mex_check(isac_pusch_codebook(nLayers, nPorts, &nT, nullptr), fn);   /* TPMI count (pmiSelect.m:44) */
    std::vector<double> pmi(maxSB), sinr((size_t)maxSB * nT);
    std::vector<int32_t> sb(2 * (size_t)maxSB);
    int32_t nSB = 0, nTPMI = 0, none = 0;
    int rc;
    {
        DevBuf hd(mxGetComplexSingles(h), mxGetNumberOfElements(h) * sizeof(mxComplexSingle), fn);
        rc = isac_ul_pmi_select_dev(isac_mex_ctx(), nLayers, hd.p, K, nSym, nRx, nPorts, noiseEst, bandSize, maxSB, pmi.data(),
                                    sinr.data(), sb.data(), &nSB, &nTPMI, &none);
    }
    isac_mex_check(rc, fn);
    if (none) {                                       /* pmiSelect.m:60-64 */
        plhs[0] = mxCreateDoubleScalar(mxGetNaN());
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(mxGetNaN());
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(mxGetNaN());
        return;
    }
    plhs[0] = row_vector(pmi.data(), nSB);
    if (nlhs > 1) plhs[1] = double_array({(mwSize)nSB, (mwSize)nTPMI}, sinr.data());
    if (nlhs > 2) { std::vector<double> v(sb.begin(), sb.begin() + 2 * (size_t)nSB); plhs[2] = double_array({(mwSize)nSB, 2}, v.data()); }
}
