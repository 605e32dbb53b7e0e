/* isac_b200.h — C ABI of the B200-native ISAC hot-path library (libisac_b200.so).
 *
 * Drop-in boundary for the data-parallel inner loop of
 * xds0112/5G_based_System_level_Integrated_Sensing_and_Communication_Simulator.
 * The reference has no FFI layer: its boundary is the MATLAB package namespace, so every entry
 * point below names the reference function (file:line, relative to the reference root) whose
 * arithmetic it replaces.  MEX gateways (the .cpp files of matlab/mex) and the Python host mirror
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
/* CUDA-event timing of kernel groups on the launching stream (bench.py's live roofline) and the number
 * of kernels this library launched.  Slots: 0 rdm_range, 1 rdm_doppler, 2 cfar, 3 echo_demod, 4 covariance,
 * 5 music, 6 pmi_sinr, 7 cdl, 8 prg_precode, 9 ul_tpmi, 10 ofdm_modulate, 11 channel_estimate (16 slots).  collect() synchronises and resets. */
#define ISAC_PROF_SLOTS 16
int isac_profile_enable(isac_ctx* ctx, int32_t on);
int isac_profile_collect(isac_ctx* ctx, double* msPerSlot, int32_t* countPerSlot, int64_t* launches);
/* Tooling: begin / end time stamps (ms after `baseEvent`, a cudaEvent_t recorded on the same device) of every kernel group
 * recorded since the last isac_profile_collect, in launch order -- the per-frame timeline bench.py --timeline prints. */
int isac_profile_timeline(isac_ctx* ctx, void* baseEvent, int32_t maxRec, int32_t* slots, double* beginMs, double* endMs, int32_t* nRec);
const char* isac_version(void);
/* Device-memory helpers for gateways that drive `_dev` entry points without linking the CUDA runtime themselves (the MEX
 * files of matlab/mex): allocation on the context's device, blocking copies ordered on the context's stream. */
int isac_dev_malloc(isac_ctx* ctx, uint64_t bytes, void** devPtr);
int isac_dev_free(isac_ctx* ctx, void* devPtr);
int isac_memcpy_h2d(isac_ctx* ctx, void* devDst, const void* hostSrc, uint64_t bytes);
int isac_memcpy_d2h(isac_ctx* ctx, void* hostDst, const void* devSrc, uint64_t bytes);

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
/* Range-kernel selection for nIFFT = 4096 (same results to rounding; exposed for A/B timing and the parity tests):
 * 0 = lean persistent TMA range kernel (raw IFFT) + bulk-staged persistent Doppler kernel (default), 1 = first TMA
 * range kernel, 2 = one CTA per column, 3 = lean range kernel + one-tile-per-CTA Doppler kernel. */
int isac_rdm_plan_set_variant(isac_rdm_plan* plan, int32_t variant);

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

/* ---- K5+K6: MUSIC DoA / fft2D estimator / music2D --------------------------------------------- */
typedef struct {
    int32_t isUpa;          /* isa(array,'parameters.baseStation.antenna.upa')   music.m:31 */
    int32_t nAnts;          /* ULA: array.numElements                             music.m:76 */
    int32_t nX, nY;         /* UPA: array.nV, array.nH                            music.m:34-35 */
    double d;               /* element spacing / lambda = 0.5                     music.m:12 */
    double aGran, aMax;     /* azimuthScanGranularity / azimuthScanScale          radarParams.m:120,122 */
    double eGran, eMax;     /* elevationScanGranularity / elevationScanScale      radarParams.m:121,123 */
} isac_doa_config;

#define ISAC_MAX_PEAKS 64   /* capacity of findpeaks(...,'NPeaks',L) outputs */

/* [L, aziEst, eleEst] = sensing.estimation.doaEstimation.music(numDets, radarEstParams, Ra)
 * (+sensing/+estimation/+doaEstimation/music.m:1).  Ra: host complex128 [n x n] column-major.
 * numDets < 0 means [] (eigen-gap rule, music.m:109-125).  ULA: aziEst[<=ISAC_MAX_PEAKS] degrees,
 * PmusicdB[aSteps].  UPA: the reference's peak picker does not exist (tools.find2DPeaks), so only
 * PmusicdB[eSteps x aSteps] (column-major) is returned and *nAzi = 0.  Pmusic may be NULL. */
int isac_music_doa_host(isac_ctx* ctx, const isac_doa_config* doa, const double* Ra, int32_t numDets,
                        int32_t* L, double* aziEst, int32_t* nAzi, double* PmusicdB, double* Pmusic);

/* The three DoA scanners of +sensing/+estimation/+doaEstimation share one entry point (same steering vectors, scan
 * grids, normalisation and findpeaks tail; they differ in the scanned quadratic form):
 *   ISAC_DOA_MUSIC  1/(a' Un Un' a + eps)   music.m:1      (numDets < 0 -> eigen-gap rule)
 *   ISAC_DOA_MVDR   1/(a' Ra^-1 a + eps)    [aziEst, eleEst] = doaEstimation.mvdrBF(numDets, radarEstParams, Ra)    mvdrBF.m:1
 *   ISAC_DOA_DBF    a' Ra a                 [aziEst, eleEst] = doaEstimation.digitalBF(numDets, radarEstParams, Ra) digitalBF.m:1
 * MVDR / DBF need numDets >= 1 (it is findpeaks' NPeaks); *L returns the source count used.  Outputs as
 * isac_music_doa_host: the dB spectrum in PdB, the absolute spectrum in P (may be NULL). */
#define ISAC_DOA_MUSIC 0
#define ISAC_DOA_MVDR 1
#define ISAC_DOA_DBF 2
int isac_doa_scan_host(isac_ctx* ctx, const isac_doa_config* doa, int32_t method, const double* Ra, int32_t numDets,
                       int32_t* L, double* aziEst, int32_t* nAzi, double* PdB, double* P);

typedef struct isac_sense_plan isac_sense_plan;
/* estResults = sensing.estimation.fft2D(radarEstParams, cfar, rxGrid, txGrid)
 * (+sensing/+estimation/fft2D.m:1): RDM + CFAR (above) + Ra (fft2D.m:106-107) + MUSIC (fft2D.m:111). */
int isac_sense_plan_create(isac_ctx* ctx, const isac_rdm_config* rdm, const isac_doa_config* doa,
                           double rRes, double vRes, isac_sense_plan** plan);
int isac_sense_plan_destroy(isac_sense_plan* plan);
/* the embedded RDM plan (for isac_rdm_get_detections / isac_rdm_get_power); owned by the sense plan */
isac_rdm_plan* isac_sense_plan_rdm(isac_sense_plan* plan);
/* enqueue the chain on device grids [nSc x nSym x nAnts x batch]; no host sync for ULA arrays */
int isac_fft2d_dev(isac_sense_plan* plan, const void* rxGrid, const void* txGrid, int32_t batch, float* rdPower);
/* fetch estResults of the last run: for map-set b, rngEst[b*maxOut + i], i < nRng[b] (unique,'stable'
 * of the per-antenna peak-sorted lists, fft2D.m:89-102); velEst likewise; aziEst[b*ISAC_MAX_PEAKS+i];
 * status[b] = ISAC_ERR_NUM_DETS_ZERO when nothing was detected (reference: error -> senResults = NaN).
 * Waits only for the result copies staged behind that run, so OTHER work (another plan's kernels, CSI reports) may
 * already be enqueued behind it; the next isac_fft2d_dev of the SAME plan must come after this call (one staging
 * buffer and one detection list per plan). */
int isac_fft2d_collect(isac_sense_plan* plan, int32_t batch, int32_t maxOut, double* rngEst, int32_t* nRng,
                       double* velEst, int32_t* nVel, double* aziEst, int32_t* nAzi, int32_t* L,
                       int32_t* status);
/* copy the DoA pseudo-spectrum (dB) of the last run: [specLen x batch], specLen = aSteps (ULA) or eSteps*aSteps */
int isac_fft2d_get_spectrum(isac_sense_plan* plan, int32_t batch, double* PmusicdB);
/* host-buffer convenience: H2D, isac_fft2d_dev, isac_fft2d_collect */
int isac_fft2d_host(isac_sense_plan* plan, const void* rxGridHost, const void* txGridHost, int32_t batch,
                    int32_t maxOut, double* rngEst, int32_t* nRng, double* velEst, int32_t* nVel,
                    double* aziEst, int32_t* nAzi, int32_t* L, int32_t* status);

typedef struct {
    int32_t nSc, nSym, nAnts;   /* size(rxGrid)                               music2D.m:34 */
    double scsHz;               /* bsParams.scs*1e3                           music2D.m:35 */
    double fc, Tsri;            /* rdrEstParams.fc, .Tsri                     music2D.m:37,39 */
    double rMax, vZone;         /* cfarEstZone(1,2), cfarEstZone(2,2)         music2D.m:42-43 */
    isac_doa_config doa;
    int32_t numDetsOverride;    /* <= 0: reference behaviour (L from the eigen-gap rule on Ra) */
} isac_music2d_config;
/* estResults = sensing.estimation.music2D(rdrEstParams, bsParams, rxGrid, txGrid) (music2D.m:1).
 * rx/tx: device complex64 [nSc x nSym x nAnts].  Outputs (host): L, aziEst/rngEst/velEst
 * [<=ISAC_MAX_PEAKS each], PrmusicdB[rSteps], PvmusicdB[vSteps] (may be NULL). */
int isac_music2d_dev(isac_ctx* ctx, const isac_music2d_config* cfg, const void* rxGrid, const void* txGrid,
                     int32_t* L, double* aziEst, int32_t* nAzi, double* rngEst, int32_t* nRng, double* velEst,
                     int32_t* nVel, double* PrmusicdB, double* PvmusicdB, int32_t* sweeps);
/* antenna covariance Ra (fft2D.m:106-107) of a device grid -> host complex128 [nAnts x nAnts] */
int isac_antenna_covariance_dev(isac_ctx* ctx, const void* rxGrid, int64_t nScSym, int32_t nAnts, double* RaHost);

/* ---- K1+K2: mono-static echo synthesis + OFDM demodulation --------------------------------------
 * rxWaveform = sensing.channelModels.basicRadarChannel(txWaveform, radarParams, targetLoSConditions)
 *   (+sensing/+channelModels/basicRadarChannel.m:1) and
 * echoGrid = sensing.monoStaticSensing(txWaveform, txDimension, carrierInfo, radarParams, targetLoSConditions)
 *   (+sensing/monoStaticSensing.m:1; nrOFDMDemodulate at :16, zero padding at :19-21). */
typedef struct {
    int64_t T;                      /* size(txWaveform,1)                         basicRadarChannel.m:9 */
    int32_t nTx;                    /* size(txWaveform,2); Rx array == Tx array   radarParams.m:88,105 */
    int32_t nTargets;               /* radarParams.nTargets                       basicRadarChannel.m:18 */
    double fc, fs, N0;              /* radarParams.fc / .fs / .N0                 :13,:15,:67 */
    const double* range;            /* [nTargets] radarParams.range               :21 */
    const double* velocity;         /* [nTargets] radarParams.velocity            :25 */
    const double* largeScaleFading; /* [nTargets]                                 :34 */
    const double* steeringVec;      /* complex128 [nTx x nTargets] RxSteeringVec  :35 */
    const int32_t* los;             /* [nTargets] targetLoSConditions (NULL = all 1) :40 */
    int32_t nfft, nSc;              /* nrOFDMInfo.Nfft, 12*NRBsDL                 monoStaticSensing.m:9-11 */
    int32_t nSymTx;                 /* txDimension(2)                             monoStaticSensing.m:19 */
    int32_t symbolsPerSubframe;     /* length of cpLengths */
    const int32_t* cpLengths;       /* nrOFDMInfo.CyclicPrefixLengths of one subframe */
} isac_echo_config;

enum { ISAC_NOISE_NONE = 0,    /* noiseless */
       ISAC_NOISE_TENSOR = 1,  /* noise = randn+1j*randn supplied by the caller, complex64 [T x nTx] (:68) */
       ISAC_NOISE_PHILOX = 2   /* generated on the device from `seed` (statistically equivalent) */ };

int isac_radar_channel_dev(isac_ctx* ctx, const isac_echo_config* cfg, const void* txWaveform, const void* noise,
                           int32_t noiseMode, uint64_t seed, void* rxWaveform);
/* echoGrid: device complex64 [nSc x nSymOut x nTx]; *nSymOut = max(whole symbols in T, nSymTx).
 * Call with echoGrid == NULL to query nSymOut. */
int isac_mono_static_sensing_dev(isac_ctx* ctx, const isac_echo_config* cfg, const void* txWaveform,
                                 const void* noise, int32_t noiseMode, uint64_t seed, void* echoGrid,
                                 int32_t* nSymOut);
/* host buffers; echoGridHost == NULL queries *nSymOut as above */
int isac_mono_static_sensing_host(isac_ctx* ctx, const isac_echo_config* cfg, const void* txWaveformHost,
                                  const void* noiseHost, int32_t noiseMode, uint64_t seed, void* echoGridHost,
                                  int32_t* nSymOut);

/* txWaveform = scale * nrOFDMModulate(carrier, txGrid)  (+phyLayer/gNBPhy.m:599; the grid / waveform pair the gNB PHY
 * accumulates for sensing, gNBPhy.m:604-612) -- SURVEY 8(f) row 2: the step immediately before the sensing hot path.
 * txGrid: device complex64 [nSc x nSym x nAnts]; txWaveform: device complex64 [T x nAnts] with
 * T = sum_s (cpLengths[s mod symbolsPerSubframe] + nfft), returned in *T (call with txWaveform == NULL to query it).
 * Plain CP-OFDM (IFFT + cyclic prefix); isac_ofdm_modulate_ex_dev adds the raised-cosine windowing and block placement. */
int isac_ofdm_modulate_dev(isac_ctx* ctx, const void* txGrid, int32_t nSc, int32_t nSym, int32_t nAnts, int32_t nfft,
                           int32_t symbolsPerSubframe, const int32_t* cpLengths, double scale, void* txWaveform,
                           int64_t* T);

/* ---- LoS / blockage geometry of the city layout ---------------------------------------------------
 * losDecision = simuLayout.checkLoS(uePos, antPos)  (+networkTopology/+blockages/openStreetMapCity.m:67; call sites
 * +simulation/networkSimulation.m:138 (UEs) and :154 (targets), one MATLAB call per link) for a whole batch of links.
 * A city is the list of wall polygons of its buildings (building.m:61-73: one 4-corner wall per floor-plan edge plus the
 * ceiling); corners: host double [3 x nCorners] column-major, walls concatenated, wallOffsets[nWalls+1] (first = 0).
 * Link i is blocked when the winding number of the user projected along the link onto any wall plane exceeds 0.1
 * (wallBlockage.m:121-127, :178-222).  uePos [3 x nLinks]; antPos [3 x nAnt], nAnt == nLinks (element-wise pairs) or 1
 * (one antenna for all links); los[i] = 1 for line of sight, 0 for blocked. */
typedef struct isac_city isac_city;
int isac_city_create(isac_ctx* ctx, int32_t nWalls, const int32_t* wallOffsets, const double* corners, isac_city** city);
int isac_city_destroy(isac_city* city);
int isac_city_check_los_host(isac_city* city, int32_t nLinks, const double* uePos, const double* antPos, int32_t nAnt,
                             int32_t* los);
int isac_city_check_los_dev(isac_city* city, int32_t nLinks, const double* uePos, const double* antPos, int32_t nAnt,
                            int32_t* los);   /* device pointers, results stay on the device */

/* ---- K7-K10, K12: Type-I codebook, PMI / RI / CQI selection, UL TPMI selection, PRG precoding --------- */
typedef struct {
    int32_t nPorts;                    /* csirs.NumCSIRSPorts                                  dlPMISelect.m:326 */
    int32_t N1, N2, O1, O2;            /* PanelDimensions, OverSamplingFactors (TS 38.214 Table 5.2.2.2.1-2)  :604-644 */
    int32_t codebookMode;              /* reportConfig.CodebookMode (1|2)                       :583-590 */
    int32_t nSizeBWP, nStartBWP;       /* reportConfig.NSizeBWP / NStartBWP (CRB index; subbands split at NStartBWP mod SubbandSize)  :536-569 */
    int32_t subbandSize;               /* reportConfig.SubbandSize (0 when not applicable)      :705-742 */
    int32_t pmiSubband, cqiSubband;    /* PMIMode / CQIMode == 'Subband'                         :683-688, cqiSelect.m:865 */
    int32_t K, L;                      /* carrier.NSizeGrid*12, carrier.SymbolsPerSlot          :829-831 */
    int32_t nRx;                       /* size(H,3) */
    const uint8_t* subsetRestriction;  /* CodebookSubsetRestriction bits, NULL = all ones       :744-764 */
    const uint8_t* i2Restriction;      /* 16 bits, NULL = all ones                              :766-775 */
    uint8_t riRestriction[8];          /* RIRestriction                                         riSelect.m:440-447 */
    int32_t nRE;                       /* CSI-RS REs kept by validateInputs (port 1, lowest RE of each CDM group)  :797-833 */
    const int32_t* reK;                /* 1-based subcarrier subscripts relative to the BWP     :352-356 */
    const int32_t* reL;                /* 1-based OFDM symbol subscripts */
    int32_t nPanels;                   /* Type1MultiPanel: Ng of PanelDimensions = [Ng N1 N2] (nPorts = 2*Ng*N1*N2); 0 / 1 = Type1SinglePanel */
} isac_csi_config;

/* W = getPMIType1SinglePanelCodebook(reportConfig,nLayers) (dlPMISelect.m:853; variant 0) or
 * communication.pmiType1SinglePanelCodebook(reportConfig,nLayers) (pmiType1SinglePanelCodebook.m:1; variant 1,
 * including that copy's deviations).  dims = [i2 i11 i12 i13] lengths; W complex128 [P x nLayers x prod(dims)]
 * (NULL to query dims).  Pure host code: works without a GPU. */
int isac_type1sp_codebook(const isac_csi_config* cfg, int32_t nLayers, int32_t variant, int32_t dims[4], double* W);
/* Wmp = getPMIType1MultiPanelCodebook(reportConfig,nLayers) (dlPMISelect.m:1351; TS 38.214 Tables 5.2.2.2.2-1..-6) for
 * PanelDimensions = [nPanels N1 N2] (nPanels = Ng in {2,4}; codebook mode 2 only with Ng = 2; nLayers <= 4).
 * dims = [i20 i21 i22 i11 i12 i13 i141 i142 i143] lengths; W complex128 [P x nLayers x prod(dims)] with P = 2*Ng*N1*N2
 * (NULL to query dims).  Uses cfg->N1, N2, O1, O2, codebookMode, subsetRestriction.  Pure host code.  Selection over it:
 * isac_pmi_plan_create with cfg->nPanels = Ng (dlPMISelect); the RI / CQI report entry points cover Type1SinglePanel only. */
int isac_type1mp_codebook(const isac_csi_config* cfg, int32_t nPanels, int32_t nLayers, int32_t dims[9], double* W);
/* The same array materialised from the beam / co-phasing table the SINR kernels read (consistency check of that table). */
int isac_type1mp_codebook_from_table(const isac_csi_config* cfg, int32_t nPanels, int32_t nLayers, int32_t dims[9], double* W);
/* nrPUSCHCodebook(nlayers,nports,tpmi).' for tpmi = 0..maxTPMI (pmiSelect.m:45; TS 38.211 Tables 6.3.1.5-1..7);
 * W complex128 [nPorts x nLayers x nTPMI] (NULL to query nTPMI). */
int isac_pusch_codebook(int32_t nLayers, int32_t nPorts, int32_t* nTPMI, double* W);

typedef struct isac_pmi_plan isac_pmi_plan;
/* [PMISet,info] = communication.phyLayer.dlPMISelect(carrier,csirs,reportConfig,nLayers,H,nVar) (dlPMISelect.m:1) */
int isac_pmi_plan_create(isac_ctx* ctx, const isac_csi_config* cfg, int32_t nLayers, int32_t maxBatch, isac_pmi_plan** plan);
int isac_pmi_plan_destroy(isac_pmi_plan* plan);
/* SINR kernel of the plan: 0 (default) = Gram-pair form (beam-response inner products shared by all candidates),
 * 1 = direct form (H*W per candidate).  Same results to rounding; the direct form is the fallback for codebooks whose
 * pair dictionary does not fit in shared memory. */
int isac_pmi_plan_set_kernel(isac_pmi_plan* plan, int32_t direct);
/* Type1MultiPanel plans (cfg->nPanels = Ng >= 2, nLayers <= 4, dlPMISelect.m:1351-1772) keep the reference's 9-D index set
 * [i20 i21 i22 | i11 i12 i13 i141 i142 i143] flattened in MATLAB linear order: dims[0] = i20*i21*i22, dims[3] =
 * i13*i141*i142*i143, and PMISet comes back flattened likewise (i2 per subband; i1 = [i11 i12 i13']); mpDims = {i20, i21, i22,
 * i13, i141, i142, i143} lengths un-flatten it (all zero for a single-panel plan).  They run the direct SINR kernel. */
int isac_pmi_plan_mp_dims(const isac_pmi_plan* plan, int32_t mpDims[7]);
/* dims = [i2 i11 i12 i13]; REs sorted by subcarrier as the plan stores them (reKs/reLs may be NULL) */
int isac_pmi_plan_info(const isac_pmi_plan* plan, int32_t dims[4], int32_t* nSB, int32_t* nCqiSB, int32_t* nRE,
                       int32_t* reKs, int32_t* reLs);
/* H: device complex64 [K x L x nRx x nPorts x batch]; nVar: host double[batch] */
int isac_dl_pmi_select_dev(isac_pmi_plan* plan, const void* H, const double* nVar, int32_t batch);
/* PMISet of the last run: i1 [3 x batch], i2 [nSB x batch] (1-based, NaN = not reported);
 * sinrAtPmi [nSB x nLayers x batch] = info.SINRPerSubband at the reported indices (may be NULL) */
int isac_dl_pmi_collect(isac_pmi_plan* plan, int32_t batch, double* i1, double* i2, double* sinrAtPmi);
/* info.SINRPerRE at the CSI-RS REs [nRE x nLayers x nCand x batch] and info.SINRPerSubband
 * [nSB x nLayers x nCand x batch] (nCand = prod(dims), MATLAB index order); either may be NULL */
int isac_dl_pmi_get_info(isac_pmi_plan* plan, int32_t batch, double* sinrPerRE, double* sinrPerSubband);

typedef struct isac_csi_plan isac_csi_plan;
/* One plan per report configuration: holds the per-rank PMI plans (riSelect.m:254 loops over the valid ranks). */
int isac_csi_plan_create(isac_ctx* ctx, const isac_csi_config* cfg, int32_t maxBatch, isac_csi_plan** plan);
int isac_csi_plan_destroy(isac_csi_plan* plan);
/* Type1MultiPanel reports (cfg->nPanels = Ng >= 2; ranks 1-4, riSelect.m:222-231): i1[2] and i2 come back as linear indices
 * into the flattened index sets [i13 i141 i142 i143] and [i20 i21 i22] of the reported rank (MATLAB order); mpDims = their
 * lengths [i20 i21 i22 i13 i141 i142 i143] at `nLayers`, for ind2sub on the caller's side (zeros for a single-panel plan). */
int isac_csi_plan_mp_dims(const isac_csi_plan* plan, int32_t nLayers, int32_t mpDims[7]);
int isac_csi_plan_set_kernel(isac_csi_plan* plan, int32_t direct);   /* as isac_pmi_plan_set_kernel, all ranks */
/* [RI,PMISet] = communication.phyLayer.riSelect(carrier,csirs,reportConfig,H,nVar) (riSelect.m:1).
 * RI [batch] (NaN when nothing is reportable), i1 [3 x batch], i2 [nSB x batch]. */
int isac_ri_select_dev(isac_csi_plan* plan, const void* H, const double* nVar, int32_t batch, double* RI, double* i1,
                       double* i2);
/* [CQI,PMISet,..] = communication.phyLayer.cqiSelect(carrier,csirs,reportConfig,nLayers,H,nVar,SINRTable)
 * (cqiSelect.m:1).  cqi [cqiRows x 2 x batch] (second codeword column NaN when nLayers <= 4), *cqiRows =
 * nCqiSB+1 in subband CQI mode else 1; sinrPerSubbandPerCW same shape with (nCqiSB+1) rows when nCqiSB > 1. */
int isac_cqi_select_dev(isac_csi_plan* plan, int32_t nLayers, const void* H, const double* nVar, int32_t batch,
                        const double* sinrTable, int32_t tableLen, double* cqi, int32_t* cqiRows, double* i1, double* i2,
                        double* sinrPerSubbandPerCW);
/* Fused UE CSI report (uePhy.m:900-907): rank = min(riSelect(..), rankCap) then cqiSelect at that rank, without
 * re-evaluating the rank already scored by the RI loop.  rankCap <= 0 disables the cap. */
int isac_csi_report_dev(isac_csi_plan* plan, const void* H, const double* nVar, int32_t batch, const double* sinrTable,
                        int32_t tableLen, int32_t rankCap, double* RI, double* i1, double* i2, double* cqi,
                        int32_t* cqiRows);

/* The same report in two halves, so that the caller can enqueue independent work (the next slot's precoding, the sensing
 * chain, another cell) behind the report's kernels before blocking: _enqueue_dev launches the kernels of every valid rank and
 * the D2H copy of the selection results without synchronising; _finish waits for that copy only (not for work enqueued
 * later on the stream) and runs the host-side RI / CQI tails.  One report may be pending per plan; H must stay valid until
 * _finish returns. */
int isac_csi_report_enqueue_dev(isac_csi_plan* plan, const void* H, const double* nVar, int32_t batch);
int isac_csi_report_finish(isac_csi_plan* plan, const double* sinrTable, int32_t tableLen, int32_t rankCap, double* RI,
                           double* i1, double* i2, double* cqi, int32_t* cqiRows);

/* sinr = communication.phyLayer.precodedSINR(H,sigma,W) (precodedSINR.m:11-18) for `batch` REs that share W:
 * H host complex128 [nRx x nPorts x batch], W host complex128 [nPorts x nLayers], sinr host double [batch]
 * (LMMSE SINR summed over the layers).  nLayers <= 8, sigma > 0. */
int isac_precoded_sinr_host(isac_ctx* ctx, const void* H, int32_t nRx, int32_t nPorts, double sigma, const void* W,
                            int32_t nLayers, int32_t batch, double* sinr);
/* [pmi,sinr,subbandIndices] = communication.phyLayer.pmiSelect(nlayers,hest,noiseest,bandSize) (pmiSelect.m:28).
 * hest: device complex64 [K x nSym x nRx x nPorts].  pmi [nSB] 0-based TPMI (NaN), sinr [nSB x nTPMI],
 * subbandIndices [nSB x 2].  *none = 1 reproduces the reference's scalar-NaN outputs (pmiSelect.m:60-64). */
int isac_ul_pmi_select_dev(isac_ctx* ctx, int32_t nLayers, const void* hest, int32_t K, int32_t nSym, int32_t nRx,
                           int32_t nPorts, double noiseEst, int32_t bandSize, int32_t maxSB, double* pmi, double* sinr,
                           int32_t* subbandIndices, int32_t* nSB, int32_t* nTPMI, int32_t* none);
/* Batched form: hest [K x nSym x nRx x nPorts x batch]; pmi [maxSB x batch], sinr [maxSB*nTPMI x batch] (each column
 * packed as [nSB x nTPMI]), none [batch].  One stream synchronisation for the whole batch. */
int isac_ul_pmi_select_batch_dev(isac_ctx* ctx, int32_t nLayers, const void* hest, int32_t K, int32_t nSym, int32_t nRx,
                                 int32_t nPorts, double noiseEst, int32_t bandSize, int32_t batch, int32_t maxSB, double* pmi,
                                 double* sinr, int32_t* nSB, int32_t* nTPMI, int32_t* none);
/* The batched report in two halves (as isac_csi_report_enqueue_dev / _finish): _enqueue_dev launches the kernels and the copy
 * of the band SINRs without synchronising, _finish waits for that copy only and runs the host-side TPMI selection.  One
 * report may be pending per context. */
int isac_ul_pmi_select_batch_enqueue_dev(isac_ctx* ctx, int32_t nLayers, const void* hest, int32_t K, int32_t nSym, int32_t nRx,
                                         int32_t nPorts, double noiseEst, int32_t bandSize, int32_t batch);
int isac_ul_pmi_select_batch_finish(isac_ctx* ctx, int32_t maxSB, double* pmi, double* sinr, int32_t* nSB, int32_t* nTPMI,
                                    int32_t* none);
/* [antsym,antind] = communication.phyLayer.prgPrecode(siz,nstartgrid,portsym,portind,F) (prgPrecode.m:53).
 * portsym complex64 / portind int32 (1-based) [NRE x nLayers], F complex64 [nLayers x P x NPRG] (device);
 * antsym complex64 / antind int32 [NRE x P] (device). */
int isac_prg_precode_dev(isac_ctx* ctx, int32_t K, int32_t Lsym, int32_t nStartGrid, const void* portsym,
                         const int32_t* portind, int32_t NRE, int32_t nLayers, const void* F, int32_t P, int32_t NPRG,
                         void* antsym, int32_t* antind);
/* `batch` independent allocations of the same size in one launch (e.g. the cells of one slot): every array gains a
 * trailing batch dimension (portsym/portind [NRE x nLayers x batch], F [nLayers x P x NPRG x batch], outputs [NRE x P x batch]). */
int isac_prg_precode_batch_dev(isac_ctx* ctx, int32_t K, int32_t Lsym, int32_t nStartGrid, const void* portsym,
                               const int32_t* portind, int32_t NRE, int32_t nLayers, const void* F, int32_t P, int32_t NPRG,
                               int32_t batch, void* antsym, int32_t* antind);

/* ---- K11: CDL channel (TR 38.901 7.7.1) as a frequency-domain channel matrix ---------------------
 * Replaces nrCDLChannel filtering + nrChannelEstimate (uePhy.m:731,897; gNBPhy.m:840,1030; objects
 * configured at +parameters/+channelModels/+communication/cdl.m:48-88, profile chosen by
 * communication.channelModels.updateCDLModels.m:7-15).  Statistical parity only (toolbox RNG stream). */
typedef struct {
    int32_t profile;         /* 0 CDL-A, 1 CDL-B, 2 CDL-C, 3 CDL-D, 4 CDL-E      cdl.m:58 / updateCDLModels.m:11-13 */
    double delaySpread;      /* channel.DelaySpread = 300e-9                 cdl.m:59 */
    double fc;               /* channel.CarrierFrequency                     cdl.m:60 */
    double maxDoppler;       /* MaximumDopplerShift (toolbox default 5 Hz) */
    int32_t txSize[3];       /* TransmitAntennaArray.Size(1:3) = [M N P]     cdl.m:61 */
    int32_t rxSize[3];       /* ReceiveAntennaArray.Size(1:3)                cdl.m:62 */
    int32_t txPattern38901;  /* 1: '38.901' element (toolbox Tx default), 0: isotropic */
    int32_t rxPattern38901;  /* toolbox Rx default: isotropic (0) */
    uint64_t seed;
} isac_cdl_config;
typedef struct isac_cdl_channel isac_cdl_channel;
int isac_cdl_create(isac_ctx* ctx, const isac_cdl_config* cfg, isac_cdl_channel** ch);
int isac_cdl_destroy(isac_cdl_channel* ch);
/* ray tables of the channel: nClusters, nRays (incl. the LOS ray), tau[nClusters], nu[nRays], cluster[nRays],
 * g complex128 [nTx x nRx x nRays] (s fastest); any pointer may be NULL */
int isac_cdl_get_rays(const isac_cdl_channel* ch, int32_t* nClusters, int32_t* nRays, int32_t* nRx, int32_t* nTx,
                      double* tau, double* nu, int32_t* cluster, double* g);
/* H: device complex64 [K x L x nRx x nTx] for subcarrier spacing scsHz, symbol times t0 + symTime[l] (seconds) */
int isac_cdl_generate_dev(isac_cdl_channel* ch, int32_t K, double scsHz, int32_t L, const double* symTime, double t0,
                          void* H);
/* n channels (all with the same nRx x nTx), outputs stacked: H [K x L x nRx x nTx x n], start times t0[n] */
/* Response kernel of the channel: 0 (default) tcgen05.mma kind::tf32 with TMEM accumulators, 1 legacy mma.sync.
 * Same 3xTF32 arithmetic; a batch uses the setting of its first channel. */
int isac_cdl_set_kernel(isac_cdl_channel* ch, int32_t legacyMma);
int isac_cdl_generate_batch_dev(isac_cdl_channel* const* ch, int32_t n, int32_t K, double scsHz, int32_t L,
                                const double* symTime, const double* t0, void* H);

/* The same step for a BLOCK of symbols written into resident buffers -- the device form of the sensing tap of the gNB PHY,
 * which appends every DL slot's grid and waveform to senTxGrid / senTxWave (gNBPhy.m:604-612: cat per slot, O(n^2) copying):
 * txGrid: device complex64, antenna pages `gridStride` symbols apart (0: nSym), block starts at the pointer passed;
 * the block's waveform goes to txWaveform[sampleOffset ...] of a buffer with `waveStride` samples per antenna (0: the block's
 * own length); symPhase = index of the block's first symbol in the subframe's CP pattern (14 * slot-in-subframe for whole
 * slots); windowing = N >= 0 samples of raised-cosine shaping / overlap with the symbol in front (nrOFDMModulate 'Windowing';
 * the head of the block's first symbol is folded into whatever the buffer already holds in front of it).  *T = block length. */
int isac_ofdm_modulate_ex_dev(isac_ctx* ctx, const void* txGrid, int32_t nSc, int32_t nSym, int32_t nAnts, int64_t gridStride,
                              int32_t nfft, int32_t symbolsPerSubframe, const int32_t* cpLengths, double scale, int32_t windowing,
                              int32_t symPhase, void* txWaveform, int64_t waveStride, int64_t sampleOffset, int64_t* T);

/* ---- link budget of the channel application step (SURVEY 8(a) row a16 tail, 8(f) row 4) --------------------------------
 * pathLoss = communication.pathlossModels.config5GNRModels(scenario, fc, los, bsPosition, uePosition)
 *   (+communication/+pathlossModels/config5GNRModels.m:27-36 -> nrPathLoss: TR 38.901 Table 7.4.1-1, no shadow fading,
 *   EnvironmentHeight 1 m, BuildingHeight 5 m, StreetWidth 20 m) and configFreeSpaceModel (configFreeSpaceModel.m:1-8),
 * for nLinks links at once (call sites uePhy.m:743-747, gNBPhy.m:852-856).  bsPos / uePos: host double [nLinks x 3] row-major
 * (x, y, z of each link); los: host int32 [nLinks] (ignored for ISAC_PL_FSPL); plDb: host double [nLinks].
 * Identical positions give 0 dB (config5GNRModels.m:32-33).  The InF-* scenarios are not built (ISAC_ERR_INVALID_ARG). */
enum { ISAC_PL_UMA = 0, ISAC_PL_UMI = 1, ISAC_PL_RMA = 2, ISAC_PL_INH = 3, ISAC_PL_FSPL = 4 };
int isac_pathloss_host(isac_ctx* ctx, int32_t scenario, double fcHz, int32_t nLinks, const double* bsPos, const double* uePos,
                       const int32_t* los, double* plDb);
/* rxWaveform = db2mag(-pathLoss)*rxWaveform; applyRxGain (uePhy.m:748-751, :935-940) on the frequency-domain channel matrices
 * of nLinks links: H device complex64 [elemsPerLink x nLinks], scaled in place by 10^((rxGainDb - plDb[link])/20). */
int isac_link_budget_dev(isac_ctx* ctx, void* H, int64_t elemsPerLink, int32_t nLinks, const double* plDb, double rxGainDb);
/* applyThermalNoise (uePhy.m:942-950, gNBPhy.m:1100-1108): Nt = k (T + 290 (10^(NF/10) - 1)) fs in watts -- the noise variance
 * per complex sample (and, with the unitary OFDM demodulation scaling, per resource element).  Pure host code. */
int isac_thermal_noise_power(double noiseFigureDb, double temperatureK, double sampleRate, double* Nt);
/* Channel matrix used when no CDL model is attached (uePhy.m:735-739): H = fft(eye(max(nTx,nRx))); H = H(1:nTx,1:nRx)/norm(H).
 * H: host complex128 [nTx x nRx] column-major.  Pure host code. */
int isac_dft_channel_matrix(int32_t nTx, int32_t nRx, double* H);

/* ---- channel estimation from reference signals ------------------------------------------------
 * [Hest, nVar] = nrChannelEstimate(rxGrid, refInd, refSym, 'CDMLengths', [FD TD] [, 'AveragingWindow', [F T]]) -- the 5G
 * Toolbox call the reference makes right before the COMM hot path (+communication/+phyLayer/uePhy.m:897 CSI-RS with
 * cdmLen from :889-895; gNBPhy.m:1030 SRS with 'AveragingWindow',[0 7]; uePhy.m:836 / gNBPhy.m:935 DM-RS).
 * LS estimates at the reference REs, CDM despreading (block means), optional F x T moving average over blocks (0 or 1 =
 * none; the toolbox's automatic window is not reproduced), linear interpolation with constant extrapolation in frequency
 * then time, nVar from second differences of the despread estimates (algorithm: oracle/chest.py; PARITY-UNPINNED).
 * refInd: host int32 [nRef], 1-based column-major linear indices into the K x L x nPorts grid; refSym: host complex64
 * [nRef].  Every port needs the same number of reference subcarriers on each of its reference symbols. */
typedef struct isac_chest_plan isac_chest_plan;
int isac_chest_plan_create(isac_ctx* ctx, int32_t K, int32_t L, int32_t nRx, int32_t nPorts, int64_t nRef,
                           const int32_t* refInd, const void* refSym, int32_t cdmFd, int32_t cdmTd, int32_t avgF,
                           int32_t avgT, int32_t maxBatch, isac_chest_plan** plan);
int isac_chest_plan_destroy(isac_chest_plan* plan);
/* rxGrid: device complex64 [K x L x nRx x batch]; Hest: device complex64 [K x L x nRx x nPorts x batch] (the layout
 * isac_dl_pmi_select_dev / isac_csi_report_dev / isac_ul_pmi_select_batch_dev consume); nVar: host double [batch], or NULL
 * to enqueue without synchronising (fetch later with isac_chest_get_nvar). */
int isac_channel_estimate_dev(isac_chest_plan* plan, const void* rxGrid, int32_t batch, void* Hest, double* nVar);
int isac_chest_get_nvar(isac_chest_plan* plan, int32_t batch, double* nVar);   /* synchronises the stream */

#ifdef __cplusplus
}
#endif
#endif /* ISAC_B200_H */
