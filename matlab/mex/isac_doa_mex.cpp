/* [L, aziEst, PdB] = isac_doa_mex(cfg, method, numDets, Ra)
 *   cfg: struct isUpa,nAnts,nX,nY,aGran,aMax,eGran,eMax;  method 0 MUSIC (music.m:1), 1 MVDR (mvdrBF.m:1), 2 beamscan (digitalBF.m:1)
 *   numDets: scalar or [] (MUSIC only: eigen-gap rule, music.m:109-125);  Ra: double (complex) [n x n]
 *   aziEst [1 x L] degrees (ULA; empty for UPA, whose peak picker tools.find2DPeaks is missing in the reference);
 *   PdB [1 x aSteps] (ULA) or [eSteps x aSteps] (UPA)
 * Marshals sensing.estimation.doaEstimation.{music,mvdrBF,digitalBF} onto isac_doa_scan_host. */
#include "isac_mex_common.h"
#include <cmath>

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 4) mexErrMsgIdAndTxt("isac:doa:nargin", "four inputs required");
    const char* fn = "doa";
    isac_doa_config a = doa_from_cfg(prhs[0]);
    const int method = (int)mxGetScalar(prhs[1]);
    const int numDets = mxIsEmpty(prhs[2]) ? -1 : (int)mxGetScalar(prhs[2]);
    const int n = a.isUpa ? a.nX * a.nY : a.nAnts;
    if ((int)mxGetM(prhs[3]) != n || (int)mxGetN(prhs[3]) != n) mexErrMsgIdAndTxt("isac:doa:size", "Ra must be nAnts-by-nAnts");
    std::vector<double> Ra = complex_doubles(prhs[3]);
    const int aSteps = (int)std::floor((a.aMax + 1) / a.aGran);            /* music.m:14 */
    const int eSteps = a.isUpa ? (int)std::floor((a.eMax + 1) / a.eGran) : 1;  /* music.m:15 */
    std::vector<double> azi(ISAC_MAX_PEAKS), PdB((size_t)aSteps * eSteps);
    int32_t L = 0, nA = 0;
    isac_mex_check(isac_doa_scan_host(isac_mex_ctx(), &a, method, Ra.data(), numDets, &L, azi.data(), &nA, PdB.data(), nullptr), fn);
    plhs[0] = mxCreateDoubleScalar((double)L);
    if (nlhs > 1) plhs[1] = row_vector(azi.data(), nA);
    if (nlhs > 2) plhs[2] = double_array({(mwSize)eSteps, (mwSize)aSteps}, PdB.data());
}
