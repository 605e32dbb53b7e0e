// Link budget of the channel application step (uePhy.applyChannelModel / gNBPhy.applyChannelModel tails): TR 38.901 path loss,
// receive gain, thermal noise power and the DFT-matrix fallback channel.
#pragma once
#include "isac_common.cuh"

namespace isac {

enum PathLossScenario : int { kPlUMa = 0, kPlUMi = 1, kPlRMa = 2, kPlInH = 3, kPlFspl = 4 };

// host arrays in, host array out (one launch for the batch); bs / ue [n x 3] row-major, los [n] (ignored for kPlFspl)
int pathloss_run(Ctx* ctx, int scenario, double fcHz, int nLinks, const double* bsPos, const double* uePos, const int* los,
                 double* plDb, cudaStream_t st);
// H[link][elems] *= 10^((rxGainDb - plDb[link]) / 20)   (device complex64, in place)
int link_scale_run(Ctx* ctx, float2* H, long long elemsPerLink, int nLinks, const double* plDb, double rxGainDb, cudaStream_t st);

}  // namespace isac
