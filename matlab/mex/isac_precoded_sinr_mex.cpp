/* sinr = isac_precoded_sinr_mex(H, sigma, W)
 *   H : double (complex) [nRx x nPorts x nRE];  sigma: noise standard deviation;  W: double (complex) [nPorts x nLayers]
 *   sinr : double [nRE x 1], LMMSE SINR summed over the layers of every RE
 * Marshals communication.phyLayer.precodedSINR (+communication/+phyLayer/precodedSINR.m:11-18; called per RE and TPMI from
 * pmiSelect.m:52) onto isac_precoded_sinr_host; a 3-D H evaluates a batch of REs that share W in one launch. */
#include "isac_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    const char* fn = "precodedSINR";
    if (nrhs != 3) mexErrMsgIdAndTxt("isac:precodedSINR:nargin", "three inputs required");
    const mxArray *H = prhs[0], *W = prhs[2];
    if (!mxIsDouble(H) || !mxIsDouble(W)) mexErrMsgIdAndTxt("isac:precodedSINR:type", "H and W must be double");
    const int R = dim_of(H, 0), P = dim_of(H, 1), B = dim_of(H, 2), nu = dim_of(W, 1);
    if (dim_of(W, 0) != P) mexErrMsgIdAndTxt("isac:precodedSINR:size", "W must be nPorts-by-nLayers");
    const std::vector<double> h = complex_doubles(H), w = complex_doubles(W);
    std::vector<double> out((size_t)B);
    isac_mex_check(isac_precoded_sinr_host(isac_mex_ctx(), h.data(), R, P, mxGetScalar(prhs[1]), w.data(), nu, B, out.data()), fn);
    plhs[0] = double_array({(mwSize)B, 1}, out.data());
}
