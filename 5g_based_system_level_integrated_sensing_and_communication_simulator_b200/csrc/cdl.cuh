// K11: TR 38.901 7.7.1 CDL channel, generated directly as the frequency response H[K x L x nRx x nTx].
#pragma once
#include "isac_common.cuh"
#include <complex>
#include <vector>

namespace isac {

// Mirrors isac_cdl_config (include/isac_b200.h): the nrCDLChannel properties the reference sets
// (+parameters/+channelModels/+communication/cdl.m:56-63) plus the toolbox defaults it relies on.
struct CdlConfig {
    int profile;              // 0 = CDL-A, 2 = CDL-C, 3 = CDL-D (B, E: not tabulated here)
    double delaySpread;       // channel.DelaySpread = 300e-9
    double fc;                // channel.CarrierFrequency
    double maxDoppler;        // MaximumDopplerShift (toolbox default 5 Hz)
    int txSize[3];            // TransmitAntennaArray.Size(1:3) = [M N P]
    int rxSize[3];            // ReceiveAntennaArray.Size(1:3)
    int txPattern38901;       // Transmit element '38.901' (default) vs isotropic
    int rxPattern38901;       // Receive element: 'isotropic' by default
    unsigned long long seed;  // replaces Seed = 73 of the toolbox's mt19937ar stream (statistical parity only)
};

struct CdlRays {
    int nCl = 0, nRay = 0, nRx = 0, nTx = 0;
    bool los = false;
    bool legacyMma = false;                    // true: mma.sync kernel instead of the tcgen05 one (isac_cdl_set_kernel)
    std::vector<double> tau;                   // [nCl] seconds
    std::vector<double> power;                 // [nCl] linear (after normalisation, NLOS part of cluster 1 for CDL-D)
    std::vector<double> nu;                    // [nCl*nRay (+1 LOS)] Doppler of each ray (Hz)
    std::vector<int> cluster;                  // cluster of each ray
    std::vector<std::complex<double>> g;       // [ray][u][s] static coefficient
    double2* d_g = nullptr;                    // device copies (uploaded on first use)
    double* d_nu = nullptr;
    double* d_tau = nullptr;
};

int cdl_build_rays(Ctx* ctx, const CdlConfig& c, CdlRays& r);

// H[k,l,u,s] = sum_n exp(-2 pi j f_k tau_n) * sum_{m in n} g_m[u,s] exp(2 pi j nu_m t_l),
// f_k = (k - K/2)*scs, t_l = t0 + symTime[l].  H: device complex64 [K x L x nRx x nTx].
void cdl_free(CdlRays& rays);
int cdl_generate(Ctx* ctx, CdlRays& rays, int K, double scsHz, int L, const double* symTime, double t0, float2* H,
                 cudaStream_t st);

// n channels of identical geometry in two launches (blockIdx.z = channel); H stacked [n][K x L x nRx x nTx]
int cdl_generate_batch(Ctx* ctx, CdlRays* const* rays, int n, int K, double scsHz, int L, const double* symTime,
                       const double* t0, float2* H, cudaStream_t st);

}  // namespace isac
