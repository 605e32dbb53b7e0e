/* estResults = isac_music2d_mex(cfg, rxGrid, txGrid)
 *   cfg  : struct scsHz,fc,Tsri,rMax,vZone,isUpa,nAnts,nX,nY,aGran,aMax,eGran,eMax
 *   grids: single complex [nSc x nSym x nAnts]
 *   estResults: struct rngEst, velEst, aziEst, eleEst, PrmusicdB, PvmusicdB
 * Marshals sensing.estimation.music2D (+sensing/+estimation/music2D.m:1) onto isac_music2d_dev. */
#include "isac_mex_common.h"
#include <cmath>

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    (void)nlhs;
    if (nrhs != 3) mexErrMsgIdAndTxt("isac:music2D:nargin", "three inputs required");
    const char* fn = "music2D";
    const mxArray *cfg = prhs[0], *rx = prhs[1], *tx = prhs[2];
    require_csingle(rx, fn, "rxGrid");
    require_csingle(tx, fn, "txGrid");
    if (mxGetNumberOfElements(rx) != mxGetNumberOfElements(tx)) mexErrMsgIdAndTxt("isac:music2D:size", "grids disagree in size");
    isac_music2d_config m = {};
    m.nSc = dim_of(rx, 0); m.nSym = dim_of(rx, 1); m.nAnts = dim_of(rx, 2);
    m.scsHz = field_scalar(cfg, "scsHz"); m.fc = field_scalar(cfg, "fc"); m.Tsri = field_scalar(cfg, "Tsri");
    m.rMax = field_scalar(cfg, "rMax"); m.vZone = field_scalar(cfg, "vZone");
    m.doa = doa_from_cfg(cfg);
    m.numDetsOverride = 0;
    const int rSteps = (int)std::floor((m.rMax + 1) / 0.5), vSteps = (int)std::floor((2 * m.vZone + 1) / 0.5);   /* music2D.m:45-46 */
    std::vector<double> azi(ISAC_MAX_PEAKS), rng(ISAC_MAX_PEAKS), vel(ISAC_MAX_PEAKS), Pr(rSteps), Pv(vSteps);
    int32_t L = 0, nA = 0, nR = 0, nV = 0, sweeps = 0;
    int rc;
    {
        const size_t bytes = mxGetNumberOfElements(rx) * sizeof(mxComplexSingle);
        DevBuf drx(mxGetComplexSingles(rx), bytes, fn), dtx(mxGetComplexSingles(tx), bytes, fn);
        rc = isac_music2d_dev(isac_mex_ctx(), &m, drx.p, dtx.p, &L, azi.data(), &nA, rng.data(), &nR, vel.data(), &nV, Pr.data(),
                              Pv.data(), &sweeps);
    }
    isac_mex_check(rc, fn);
    const char* names[] = {"rngEst", "velEst", "aziEst", "eleEst", "PrmusicdB", "PvmusicdB"};
    plhs[0] = mxCreateStructMatrix(1, 1, 6, names);
    mxSetField(plhs[0], 0, "rngEst", row_vector(rng.data(), nR));
    mxSetField(plhs[0], 0, "velEst", row_vector(vel.data(), nV));
    mxSetField(plhs[0], 0, "aziEst", row_vector(azi.data(), nA));
    mxArray* ele = mxCreateDoubleMatrix(1, nA, mxREAL);
    for (int i = 0; i < nA; ++i) mxGetDoubles(ele)[i] = mxGetNaN();   /* music.m:104 */
    mxSetField(plhs[0], 0, "eleEst", ele);
    mxSetField(plhs[0], 0, "PrmusicdB", row_vector(Pr.data(), rSteps));
    mxSetField(plhs[0], 0, "PvmusicdB", row_vector(Pv.data(), vSteps));
}
