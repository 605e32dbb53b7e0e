// Sensing estimator orchestration: fft2D (RDM + CFAR + covariance + MUSIC DoA) and music2D.
#pragma once
#include "music.cuh"
#include "rdm.cuh"
#include <vector>

namespace isac {

// fft2D estimator = RdmPlan + DoA stage buffers (per map-set)
struct SensePlan {
    RdmPlan* rdm = nullptr;
    DoaConfig doa{};
    double rRes = 0.0, vRes = 0.0;  // radarEstParams.rRes / .vRes (fft2D.m:81-82)
    int aSteps = 0, eSteps = 0;
    double2* d_Ra = nullptr;   // [nAnts^2 x maxBatch]
    double* d_w = nullptr;     // [nAnts x maxBatch]  eigenvalues, descending
    double2* d_V = nullptr;    // [nAnts^2 x maxBatch]
    int* d_L = nullptr;        // [maxBatch]
    double* d_P = nullptr;     // [specLen x maxBatch]
    double* d_PdB = nullptr;   // [specLen x maxBatch]
    int* d_peakLoc = nullptr;  // [kMaxPeaks x maxBatch]
    int* d_nPeaks = nullptr;   // [maxBatch]
    int* d_status = nullptr;   // [maxBatch]
    int* d_order = nullptr;    // large-array eigen order [nAnts]
    int specLen = 0;
    // Results of the last run staged in pinned memory by copies enqueued behind the chain (sense_fft2d_run), so that
    // sense_fft2d_collect waits on `ready` only -- not on work enqueued later on the stream -- and never copies again:
    //   counts [pages] | L [B] | nPeaks [B] | status [B] | peakLoc [kMaxPeaks x B] | det [detCap x pages] | peak [detCap x pages]
    // Only the first detCap detections of every (antenna, map-set) page are staged; a page with more falls back to a
    // direct copy in collect -- valid because the next run of the SAME plan may only be enqueued after collect (one staging
    // buffer, one device detection list per plan; include/isac_b200.h states the contract).
    char* h_stage = nullptr;
    int detCap = 0;
    cudaEvent_t ready = nullptr;
    int stagedBatch = 0;       // batch of the staged run (0 = nothing staged)
};

int sense_plan_create(Ctx* ctx, const RdmConfig& rc, const DoaConfig& doa, double rRes, double vRes, SensePlan** out);
void sense_plan_destroy(SensePlan* p);
// enqueue the whole fft2D chain for `batch` map-sets (device grids); no host synchronisation for ULA arrays
int sense_fft2d_run(SensePlan* p, const float2* rx, const float2* tx, int batch, float* powOut, cudaStream_t st);

struct Fft2dResult {          // estResults of fft2D.m for one map-set
    std::vector<double> rngEst, velEst, aziEst;
    int L = 0;
    int status = 0;
};
// D2H + host tail (per-antenna stable sort by peak, unique-stable: fft2D.m:64-102)
int sense_fft2d_collect(SensePlan* p, int batch, std::vector<Fft2dResult>& out);

// doaEstimation.music / mvdrBF / digitalBF (method = DoaMethod) on a caller-supplied covariance (device double2 [n x n]);
// MUSIC: numDets <= 0 -> eigen-gap rule
int music_doa_run(Ctx* ctx, const DoaConfig& doa, const double2* dRa, int numDets, int* L, std::vector<double>& aziEst,
                  std::vector<double>& PdB, std::vector<double>& P, cudaStream_t st, int method = kDoaMusic);

struct Music2dConfig {
    int nSc, nSym, nAnts;
    double scsHz, fc, Tsri;     // bsParams.scs*1e3, rdrEstParams.fc, .Tsri  (music2D.m:35-39)
    double rMax, vZone;         // cfarEstZone(1,2), cfarEstZone(2,2)       (music2D.m:42-43)
    DoaConfig doa;
    int numDetsOverride;        // <= 0: reference behaviour (eigen-gap rule on Ra)
};
struct Music2dResult {
    int L = 0;
    std::vector<double> aziEst, rngEst, velEst, PrdB, PvdB, Pr, Pv, PdoadB;
    int sweeps = 0;
};
int music2d_run(Ctx* ctx, const Music2dConfig& c, const float2* rx, const float2* tx, Music2dResult& out,
                cudaStream_t st);

}  // namespace isac
