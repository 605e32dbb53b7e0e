/* isac_b200.h — C ABI of the B200-native ISAC hot-path library (libisac_b200.so).
 *
 * Drop-in boundary for the data-parallel inner loop of
 * xds0112/5G_based_System_level_Integrated_Sensing_and_Communication_Simulator.
 * The reference has no FFI layer: its boundary is the MATLAB package namespace, so every entry
 * point below names the reference function (file:line, relative to the reference root) whose
 * arithmetic it replaces.  MEX gateways (matlab/mex/*.cpp) and the Python host mirror
 * (5g_based_..._b200/) bind exactly these symbols.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns an isac_status (0 = OK);
 *     isac_last_error(ctx) returns the message of the last failure on that context.
 *   - arrays are MATLAB column-major; complex = interleaved float32 pairs (mxComplexSingle);
 *     indices RETURNED to the caller are 1-based like the reference's.
 *   - `_dev` entry points take device pointers and enqueue on the context's stream without
 *     synchronising; `_host` entry points take host pointers, stage through pinned memory and
 *     return after the results are on the host.
 *   - there is no CPU fallback: without a CUDA device isac_create fails with ISAC_ERR_NO_DEVICE.
 */
#ifndef ISAC_B200_H
#define ISAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ISAC_OK = 0,
    ISAC_ERR_INVALID_ARG = 1,
    ISAC_ERR_CUDA = 2,
    ISAC_ERR_NO_DEVICE = 3,
    ISAC_ERR_UNSUPPORTED = 4,
    ISAC_ERR_CFAR_WINDOW = 5,   /* CUT training window leaves the map (CFARDetector2D errors; caught at cellSimulation.m:196-202) */
    ISAC_ERR_NO_LOS_TARGET = 6, /* all targets NLoS (basicRadarChannel.m:59,64 -> empty waveform) */
    ISAC_ERR_NUM_DETS_ZERO = 7, /* MUSIC with zero sources (findpeaks NPeaks=0 errors, music.m:102) */
    ISAC_ERR_CAPACITY = 8
} isac_status;

typedef struct isac_ctx isac_ctx;
typedef struct isac_rdm_plan isac_rdm_plan;

/* ---- context ------------------------------------------------------------------------------- */
int isac_create(isac_ctx** ctx, int device);
int isac_destroy(isac_ctx* ctx);
const char* isac_last_error(const isac_ctx* ctx);
/* Enqueue on an existing cudaStream_t (e.g. torch's current stream); NULL = CUDA legacy default stream. */
int isac_set_stream(isac_ctx* ctx, void* cuda_stream);
/* Go back to the context's private non-blocking stream (the default after isac_create). */
int isac_use_own_stream(isac_ctx* ctx);
int isac_synchronize(isac_ctx* ctx);
const char* isac_version(void);

/* ---- K3+K4: 2D-FFT range-Doppler map + 2D CA-CFAR ---------------------------------------------
 * Replaces sensing.estimation.fft2D's vectorised core (+sensing/+estimation/fft2D.m:37-46, :59-63)
 * and the phased.CFARDetector2D step built by sensing.detection.cfar2D
 * (+sensing/+detection/cfar2D.m:15-33). */
typedef struct {
    int32_t nSc, nSym, nAnts;    /* size(rxGrid)                      fft2D.m:32 */
    int32_t nIFFT, nFFT;         /* radarEstParams.nIFFT / .nFFT      radarParams.m:69,75 */
    int32_t cutRow0, cutRow1;    /* rngIdx(1):rngIdx(2), 1-based      cfar2D.m:21,23 */
    int32_t cutCol0, cutCol1;    /* dopIdx(1):dopIdx(2), 1-based      cfar2D.m:22,23 */
    int32_t guardRows, guardCols;  /* GuardBandSize    = [2 2]        cfar2D.m:32 */
    int32_t trainRows, trainCols;  /* TrainingBandSize = [1 1]        cfar2D.m:33 */
    int32_t maxBatch;            /* map-sets (cells / CPIs) per call */
    double pfa;                  /* ProbabilityFalseAlarm             cfar2D.m:30 */
    double kaiserBeta;           /* kaiser(n,3)                       fft2D.m:135 */
} isac_rdm_config;

int isac_rdm_plan_create(isac_ctx* ctx, const isac_rdm_config* cfg, isac_rdm_plan** plan);
int isac_rdm_plan_destroy(isac_rdm_plan* plan);
/* CA-CFAR threshold factor alpha = N (Pfa^(-1/N) - 1) the plan uses, and its training-cell count N */
int isac_rdm_plan_info(const isac_rdm_plan* plan, double* alpha, int32_t* nTrain, int32_t* nCut);

/* rxGrid / txGrid: device, complex64 [nSc x nSym x nAnts x batch].
 * rdPower: device float32 [nIFFT x nFFT x nAnts x batch] or NULL (plan-owned buffer is used).
 * Runs range IFFT, Doppler FFT, |.|^2, CFAR and ordered compaction; results stay on the device. */
int isac_rdm_cfar_dev(isac_rdm_plan* plan, const void* rxGrid, const void* txGrid, int32_t batch,
                      float* rdPower);
/* CFAR + compaction only, on a caller-supplied device power map (abs(rdm).^2, fft2D.m:61-62). */
int isac_cfar2d_dev(isac_rdm_plan* plan, const float* rdPower, int32_t batch);
/* Copy the last run's detections to the host (synchronises the stream).
 * detCount [nAnts x batch]; detRowCol [2 x maxDet x nAnts x batch] = CFARDetector2D 'Detection
 * index' columns ([row; col], 1-based, CUT order); peaks [maxDet x nAnts x batch] = rdResponse
 * at each detection (fft2D.m:74).  Entries beyond detCount are untouched. */
int isac_rdm_get_detections(isac_rdm_plan* plan, int32_t batch, int32_t maxDet, int32_t* detCount,
                            int32_t* detRowCol, float* peaks);
/* Copy the last run's power map to the host (synchronises). */
int isac_rdm_get_power(isac_rdm_plan* plan, int32_t batch, float* rdPowerHost);
/* Host-buffer variant of isac_rdm_cfar_dev: H2D of both grids, run, D2H of detections
 * (and of the power map when rdPowerHost != NULL). */
int isac_rdm_cfar_host(isac_rdm_plan* plan, const void* rxGridHost, const void* txGridHost,
                       int32_t batch, int32_t maxDet, int32_t* detCount, int32_t* detRowCol,
                       float* peaks, float* rdPowerHost);

#ifdef __cplusplus
}
#endif
#endif /* ISAC_B200_H */
