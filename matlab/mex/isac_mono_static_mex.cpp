/* echoGrid = isac_mono_static_mex(cfg, txWaveform, seed [, noise])
 *   cfg        : struct fc,fs,N0,range,velocity,largeScaleFading,steeringVec[nTx x nTargets],los,nfft,nSc,nSymTx,cpLengths
 *   txWaveform : single complex [T x nTx];  noise (optional): single complex [T x nTx] standard normal (re, im)
 *   echoGrid   : single complex [nSc x nSymOut x nTx]
 * Marshals sensing.monoStaticSensing (+sensing/monoStaticSensing.m:1) onto isac_mono_static_sensing_host. */
#include "isac_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    if (nrhs < 3 || nrhs > 4) mexErrMsgIdAndTxt("isac:monoStaticSensing:nargin", "three or four inputs required");
    const mxArray *cfg = prhs[0], *tx = prhs[1];
    require_csingle(tx, "monoStaticSensing", "txWaveform");
    std::vector<double> range = field_doubles(cfg, "range"), vel = field_doubles(cfg, "velocity"),
                        lsf = field_doubles(cfg, "largeScaleFading"), sv = complex_doubles(field(cfg, "steeringVec"));
    std::vector<int32_t> los = field_int32s(cfg, "los"), cp = field_int32s(cfg, "cpLengths");
    isac_echo_config e = {};
    e.T = (int64_t)mxGetM(tx); e.nTx = dim_of(tx, 1); e.nTargets = (int32_t)range.size();
    if (vel.size() != range.size() || lsf.size() != range.size() || sv.size() != 2 * (size_t)e.nTx * range.size() ||
        (!los.empty() && los.size() != range.size()))
        mexErrMsgIdAndTxt("isac:monoStaticSensing:size", "per-target fields disagree in length");
    e.fc = field_scalar(cfg, "fc"); e.fs = field_scalar(cfg, "fs"); e.N0 = field_scalar(cfg, "N0");
    e.range = range.data(); e.velocity = vel.data(); e.largeScaleFading = lsf.data(); e.steeringVec = sv.data();
    e.los = los.empty() ? nullptr : los.data();
    e.nfft = (int32_t)field_scalar(cfg, "nfft"); e.nSc = (int32_t)field_scalar(cfg, "nSc");
    e.nSymTx = (int32_t)field_scalar(cfg, "nSymTx");
    e.symbolsPerSubframe = (int32_t)cp.size(); e.cpLengths = cp.data();
    const uint64_t seed = (uint64_t)mxGetScalar(prhs[2]);
    const void* noise = nullptr;
    int32_t mode = ISAC_NOISE_PHILOX;
    if (nrhs == 4 && !mxIsEmpty(prhs[3])) {
        require_csingle(prhs[3], "monoStaticSensing", "noise");
        if (mxGetNumberOfElements(prhs[3]) != mxGetNumberOfElements(tx))
            mexErrMsgIdAndTxt("isac:monoStaticSensing:size", "noise must have the size of txWaveform");
        noise = mxGetComplexSingles(prhs[3]);
        mode = ISAC_NOISE_TENSOR;
    }
    int32_t nSymOut = 0;   /* query, then run */
    isac_mex_check(isac_mono_static_sensing_host(isac_mex_ctx(), &e, mxGetComplexSingles(tx), noise, mode, seed, nullptr, &nSymOut),
                   "monoStaticSensing");
    const mwSize dims[3] = {(mwSize)e.nSc, (mwSize)nSymOut, (mwSize)e.nTx};
    plhs[0] = mxCreateNumericArray(3, dims, mxSINGLE_CLASS, mxCOMPLEX);
    isac_mex_check(isac_mono_static_sensing_host(isac_mex_ctx(), &e, mxGetComplexSingles(tx), noise, mode, seed,
                                                 mxGetComplexSingles(plhs[0]), &nSymOut), "monoStaticSensing");
}
