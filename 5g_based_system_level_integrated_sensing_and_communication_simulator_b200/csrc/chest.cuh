// Pilot-based channel estimation: LS + CDM despreading + interpolation + noise estimate.
#pragma once
#include "isac_common.cuh"
#include <vector>

namespace isac {

struct ChestConfig {
    int32_t K, L, nRx, nPorts;
    int32_t cdmFd, cdmTd;  // 'CDMLengths' = [FD TD]
    int32_t avgF, avgT;    // 'AveragingWindow' = [F T] (0 or 1 = none)
    int32_t maxBatch;
};

struct ChestPlan {
    Ctx* ctx = nullptr;
    ChestConfig cfg{};
    int nK = 0, nL = 0;    // reference subcarriers / symbols per port
    int nBf = 0, nBt = 0;  // CDM blocks per port along frequency / time
    // device tables (plan-owned)
    int32_t* d_refK = nullptr;   // [nK x P] 0-based reference subcarriers of each port (ascending)
    int32_t* d_refL = nullptr;   // [nL x P] 0-based reference symbols
    float2* d_inv = nullptr;     // [nK x nL x P] conj(s)/|s|^2
    int32_t* d_flo = nullptr;    // [K x P] lower block of the frequency interpolation
    float* d_fw = nullptr;       // [K x P] weight of the upper block
    int32_t* d_tlo = nullptr;    // [L x P]
    float* d_tw = nullptr;       // [L x P]
    float2* d_D = nullptr;       // despread estimates [nBf x nBt x nRx x P x maxBatch]
    float2* d_A = nullptr;       // after the averaging window (same shape; only when a window is set)
    double* d_nvar = nullptr;    // [maxBatch]
    double* h_nvar = nullptr;    // pinned [maxBatch]
};

int chest_plan_create(Ctx* ctx, const ChestConfig& cfg, long long nRef, const int32_t* refInd, const float2* refSym,
                      ChestPlan** out);
void chest_plan_destroy(ChestPlan* p);
// rx: device [K x L x nRx x batch]; H: device [K x L x nRx x P x batch]; nVarHost: host [batch] or nullptr (no sync)
int chest_run(ChestPlan* p, const float2* rx, int batch, float2* H, double* nVarHost, cudaStream_t st);

}  // namespace isac
