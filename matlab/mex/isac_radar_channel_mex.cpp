/* rxWaveform = isac_radar_channel_mex(cfg, txWaveform, seed [, noise])
 *   cfg        : struct fc,fs,N0,range,velocity,largeScaleFading,steeringVec[nTx x nTargets],los
 *   txWaveform : single complex [T x nTx];  noise (optional): single complex [T x nTx] standard normal (re, im)
 *   rxWaveform : single complex [T x nTx]
 * Marshals sensing.channelModels.basicRadarChannel (+sensing/+channelModels/basicRadarChannel.m:1) onto
 * isac_radar_channel_dev.  Every target NLoS -> error isac:basicRadarChannel:status6 (the reference returns an empty
 * waveform there, basicRadarChannel.m:59-64, which monoStaticSensing cannot demodulate). */
#include "isac_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    const char* fn = "basicRadarChannel";
    if (nrhs < 3 || nrhs > 4) mexErrMsgIdAndTxt("isac:basicRadarChannel:nargin", "three or four inputs required");
    const mxArray *cfg = prhs[0], *tx = prhs[1];
    require_csingle(tx, fn, "txWaveform");
    std::vector<double> range = field_doubles(cfg, "range"), vel = field_doubles(cfg, "velocity"),
                        lsf = field_doubles(cfg, "largeScaleFading"), sv = complex_doubles(field(cfg, "steeringVec"));
    std::vector<int32_t> los = field_int32s(cfg, "los");
    isac_echo_config e = {};
    e.T = (int64_t)mxGetM(tx); e.nTx = dim_of(tx, 1); e.nTargets = (int32_t)range.size();
    if (vel.size() != range.size() || lsf.size() != range.size() || sv.size() != 2 * (size_t)e.nTx * range.size() ||
        (!los.empty() && los.size() != range.size()))
        mexErrMsgIdAndTxt("isac:basicRadarChannel:size", "per-target fields disagree in length");
    e.fc = field_scalar(cfg, "fc"); e.fs = field_scalar(cfg, "fs"); e.N0 = field_scalar(cfg, "N0");
    e.range = range.data(); e.velocity = vel.data(); e.largeScaleFading = lsf.data(); e.steeringVec = sv.data();
    e.los = los.empty() ? nullptr : los.data();
    const uint64_t seed = (uint64_t)mxGetScalar(prhs[2]);
    const size_t bytes = (size_t)e.T * e.nTx * sizeof(mxComplexSingle);
    int32_t mode = ISAC_NOISE_PHILOX;
    const mxArray* nz = nullptr;
    if (nrhs == 4 && !mxIsEmpty(prhs[3])) {
        nz = prhs[3];
        require_csingle(nz, fn, "noise");
        if (mxGetNumberOfElements(nz) != (size_t)e.T * e.nTx) mexErrMsgIdAndTxt("isac:basicRadarChannel:size", "noise must be T-by-nTx");
        mode = ISAC_NOISE_TENSOR;
    }
    const mwSize dims[2] = {(mwSize)e.T, (mwSize)e.nTx};
    plhs[0] = mxCreateNumericArray(2, dims, mxSINGLE_CLASS, mxCOMPLEX);
    int rc;
    {
        DevBuf txd(mxGetComplexSingles(tx), bytes, fn), rxd(nullptr, bytes, fn);
        DevBuf nzd(nz ? mxGetComplexSingles(nz) : nullptr, nz ? bytes : 16, fn);
        rc = isac_radar_channel_dev(isac_mex_ctx(), &e, txd.p, nz ? nzd.p : nullptr, mode, seed, rxd.p);
        if (!rc) rc = isac_memcpy_d2h(isac_mex_ctx(), mxGetComplexSingles(plhs[0]), rxd.p, bytes);
    }
    isac_mex_check(rc, fn);
}
