/* estResults = isac_fft2d_mex(cfg, rxGrid, txGrid)
 *   cfg   : struct nIFFT,nFFT,rRes,vRes,cutRows[2],cutCols[2],Pfa,isUpa,nAnts,nX,nY,aGran,aMax,eGran,eMax
 *   grids : single complex [nSc x nSym x nAnts]
 * Marshals sensing.estimation.fft2D (+sensing/+estimation/fft2D.m:1) onto isac_fft2d_host. */
#include "isac_mex_common.h"

static PlanCache<isac_sense_plan> g_plans(isac_sense_plan_destroy);
static void drop_plans(void) { g_plans.clear(); }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 3) mexErrMsgIdAndTxt("isac:fft2d:nargin", "three inputs required");
    const mxArray *cfg = prhs[0], *rx = prhs[1], *tx = prhs[2];
    if (!mxIsSingle(rx) || !mxIsComplex(rx) || !mxIsSingle(tx) || !mxIsComplex(tx))
        mexErrMsgIdAndTxt("isac:fft2d:type", "grids must be complex single");
    const mwSize* d = mxGetDimensions(rx);
    const int nd = (int)mxGetNumberOfDimensions(rx);
    isac_rdm_config r = {};
    r.nSc = (int32_t)d[0]; r.nSym = (int32_t)d[1]; r.nAnts = nd > 2 ? (int32_t)d[2] : 1;
    r.nIFFT = (int32_t)field_scalar(cfg, "nIFFT"); r.nFFT = (int32_t)field_scalar(cfg, "nFFT");
    const double* cr = mxGetDoubles(mxGetField(cfg, 0, "cutRows"));
    const double* cc = mxGetDoubles(mxGetField(cfg, 0, "cutCols"));
    r.cutRow0 = (int32_t)cr[0]; r.cutRow1 = (int32_t)cr[1]; r.cutCol0 = (int32_t)cc[0]; r.cutCol1 = (int32_t)cc[1];
    /* guard / training band sizes of the detector object built by sensing.detection.cfar2D (cfar2D.m:27-33; the shim reads
     * cfar.cfarDetector2D.GuardBandSize / TrainingBandSize); the reference's values when the fields are absent */
    const mxArray* gb = mxGetField(cfg, 0, "guardBand");
    const mxArray* tb = mxGetField(cfg, 0, "trainBand");
    r.guardRows = gb ? (int32_t)mxGetDoubles(gb)[0] : 2; r.guardCols = gb ? (int32_t)mxGetDoubles(gb)[mxGetNumberOfElements(gb) > 1] : 2;
    r.trainRows = tb ? (int32_t)mxGetDoubles(tb)[0] : 1; r.trainCols = tb ? (int32_t)mxGetDoubles(tb)[mxGetNumberOfElements(tb) > 1] : 1;
    r.maxBatch = 1; r.pfa = field_scalar(cfg, "Pfa"); r.kaiserBeta = 3.0;   /* fft2D.m:135 */
    isac_doa_config a = doa_from_cfg(cfg);
    isac_ctx* ctx = isac_mex_ctx();
    g_plan_cleanup = drop_plans;
    const double rRes = field_scalar(cfg, "rRes"), vRes = field_scalar(cfg, "vRes");
    std::string key;
    key_add(key, r); key_add(key, a); key_add(key, rRes); key_add(key, vRes);
    isac_sense_plan* plan = g_plans.find(key);        /* windows, twiddles, tensor map and device arenas are built once per shape */
    if (!plan) {
        isac_mex_check(isac_sense_plan_create(ctx, &r, &a, rRes, vRes, &plan), "fft2D");
        g_plans.put(key, plan);
    }
    const int maxOut = 4096;
    std::vector<double> rng(maxOut), vel(maxOut), azi(ISAC_MAX_PEAKS);
    int32_t nR = 0, nV = 0, nA = 0, L = 0, st = 0;
    int rc = isac_fft2d_host(plan, mxGetComplexSingles(rx), mxGetComplexSingles(tx), 1, maxOut, rng.data(), &nR, vel.data(), &nV,
                             azi.data(), &nA, &L, &st);
    if (rc) g_plans.drop(plan);
    isac_mex_check(rc, "fft2D");
    isac_mex_check(st, "fft2D");   /* zero detections -> error, like findpeaks(...,'NPeaks',0) (music.m:102) */
    const char* names[] = {"rngEst", "velEst", "aziEst", "eleEst"};
    plhs[0] = mxCreateStructMatrix(1, 1, 4, names);
    mxSetField(plhs[0], 0, "rngEst", row_vector(rng.data(), nR));
    mxSetField(plhs[0], 0, "velEst", row_vector(vel.data(), nV));
    mxSetField(plhs[0], 0, "aziEst", row_vector(azi.data(), nA));
    mxArray* ele = mxCreateDoubleMatrix(1, nA, mxREAL);
    for (int i = 0; i < nA; ++i) mxGetDoubles(ele)[i] = mxGetNaN();   /* music.m:104 */
    mxSetField(plhs[0], 0, "eleEst", ele);
}
