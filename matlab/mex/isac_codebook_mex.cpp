/* W = isac_codebook_mex(cfg, nLayers, variant)
 *   cfg     : struct nPorts,N1,N2,O1,O2,codebookMode,subsetRestriction,i2Restriction (+ the fields CsiCfg expects; the RE
 *             list may be empty)
 *   variant : 0 = the UE-side copy getPMIType1SinglePanelCodebook (dlPMISelect.m:853-1349),
 *             1 = the gNB-side copy communication.pmiType1SinglePanelCodebook (pmiType1SinglePanelCodebook.m:46-554,
 *                 consumer schedulerEntity.m:736-777), including its two deviations (:348/:358 and :225/:227)
 *   W       : complex double [nPorts x nLayers x i2 x i11 x i12 x i13], restricted precoders all zero
 * Pure host code behind isac_type1sp_codebook: works without a GPU. */
#include "isac_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    const char* fn = "pmiType1SinglePanelCodebook";
    if (nrhs != 3) mexErrMsgIdAndTxt("isac:pmiType1SinglePanelCodebook:nargin", "three inputs required");
    CsiCfg cs(prhs[0], 1);
    const int nLayers = (int)mxGetScalar(prhs[1]), variant = (int)mxGetScalar(prhs[2]);
    int32_t dims[4] = {0, 0, 0, 0};
    int rc = isac_type1sp_codebook(&cs.c, nLayers, variant, dims, nullptr);
    if (rc) mexErrMsgIdAndTxt("isac:pmiType1SinglePanelCodebook:config", "invalid codebook configuration (status %d)", rc);
    const size_t n = (size_t)cs.c.nPorts * nLayers * dims[0] * dims[1] * dims[2] * dims[3];
    std::vector<double> W(2 * n);
    rc = isac_type1sp_codebook(&cs.c, nLayers, variant, dims, W.data());
    if (rc) mexErrMsgIdAndTxt("isac:pmiType1SinglePanelCodebook:config", "invalid codebook configuration (status %d)", rc);
    plhs[0] = complex_double_array({(mwSize)cs.c.nPorts, (mwSize)nLayers, (mwSize)dims[0], (mwSize)dims[1], (mwSize)dims[2], (mwSize)dims[3]},
                                   W.data());
}
