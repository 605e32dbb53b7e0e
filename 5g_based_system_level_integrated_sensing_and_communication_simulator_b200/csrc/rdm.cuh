// Range-Doppler map pipeline (K3 + K4): declarations shared between rdm.cu, music.cu and capi.cu.
#pragma once
#include "isac_common.cuh"
#include <cuda.h>
#include <vector>

namespace isac {

// Mirrors isac_rdm_config in include/isac_b200.h (kept POD, same field order).
struct RdmConfig {
    int32_t nSc, nSym, nAnts, nIFFT, nFFT;
    int32_t cutRow0, cutRow1, cutCol0, cutCol1;  // 1-based inclusive CUT rectangle (cfar2D.m:21-24)
    int32_t guardRows, guardCols, trainRows, trainCols;
    int32_t maxBatch;
    double pfa;
    double kaiserBeta;
};

struct RdmPlan {
    Ctx* ctx = nullptr;
    RdmConfig cfg{};
    int M = 0;            // symbols entering the Doppler FFT = min(nSym, nFFT) (fft2D.m:46 truncates / pads)
    int nCut = 0, nCutRows = 0, nCutCols = 0;
    int nTrain = 0;
    double alpha = 0.0;   // CA-CFAR threshold factor N (Pfa^(-1/N) - 1)
    int rowWords = 0;     // 32-bit words of the detected-row bitmap per map-set
    // device buffers (plan-owned)
    float* d_win1 = nullptr;     // kaiser(nSc)                         (fft2D.m:43)
    float* d_win2 = nullptr;     // kaiser(nIFFT)[(n-N/2) mod N]/sqrt(N) (fft2D.m:44-45 folded)
    float* d_rowScale = nullptr; // w2[(n-N/2) mod N]^2 / (nIFFT nFFT): power scale per range row (raw-IFFT pipeline)
    int* d_tickets = nullptr;    // [2 x maxBatch] work counters of the persistent range / Doppler kernels
    float2* d_inter = nullptr;   // range profiles [nIFFT x M x nAnts]: ONE map-set, reused so it stays in L2
    float* d_pow = nullptr;      // |RDM|^2 [nIFFT x nFFT x nAnts x maxBatch]
    uint8_t* d_flags = nullptr;  // CFAR decisions [nCut x nAnts x maxBatch]
    uint32_t* d_rowmask = nullptr;  // [rowWords x maxBatch]
    int32_t* d_detCount = nullptr;  // [nAnts x maxBatch]
    int2* d_det = nullptr;          // [nCut x nAnts x maxBatch] (row, col) 1-based, CUT order
    float* d_peak = nullptr;        // [nCut x nAnts x maxBatch]
    CUtensorMap interMap{};      // 2-D TMA view of d_inter for the bulk-staged Doppler kernel (nFFT = 256)
    bool hasInterMap = false;
    bool pdl = true;             // programmatic dependent launch between the range and Doppler kernels of a batch (ISAC_RDM_PDL=0 disables)
    int hints = 0x149;           // L2 eviction priorities of the lean pipeline (RdmDev::hints; ISAC_RDM_HINTS overrides): rx/tx loads and
                                 // power-map stores evict_first, range-profile stores evict_last, staged range-profile lines discarded
                                 // (measured 155.7 -> 147.3 us per chain of 4 cfg2 map-sets; gpurun_out/c15_hints.log)
    const float* lastPow = nullptr; // power map used by the last run (plan-owned or caller's)
    int lastBatch = 0;
    int variant = 0;             // N = 4096 pipelines: 0 lean persistent TMA range kernel (raw IFFT) + bulk-staged
                                 // persistent Doppler kernel (F = 256; other F as 3), 1 first TMA range kernel,
                                 // 2 one-CTA-per-column range kernel, 3 lean range kernel + one-tile-per-CTA Doppler kernel
};

int rdm_plan_create(Ctx* ctx, const RdmConfig& cfg, RdmPlan** out);
void rdm_plan_destroy(RdmPlan* plan);
// rx/tx: device pointers, [nSc x nSym x nAnts x batch] interleaved complex64, column-major.
// powOut: optional device pointer [nIFFT x nFFT x nAnts x batch] float32 (nullptr -> plan-owned buffer).
int rdm_run(RdmPlan* plan, const float2* rx, const float2* tx, int batch, float* powOut, cudaStream_t stream);
// CFAR + compaction only, on an existing power map (device)
int rdm_cfar_only(RdmPlan* plan, const float* pow, int batch, cudaStream_t stream);

}  // namespace isac
