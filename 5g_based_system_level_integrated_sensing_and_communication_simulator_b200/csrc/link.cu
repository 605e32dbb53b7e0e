// Path loss / link budget (SURVEY 8(a) row a16 tail, 8(f) row 4).
//
// Replaces, for a batch of links, communication.pathlossModels.config5GNRModels (config5GNRModels.m:27-36: nrPathLoss with an
// nrPathLossConfig of the named scenario -- TR 38.901 Table 7.4.1-1 without shadow fading, toolbox defaults EnvironmentHeight
// 1 m, BuildingHeight 5 m, StreetWidth 20 m) and configFreeSpaceModel (configFreeSpaceModel.m:1-8: fspl), and the scaling
// rxWaveform = db2mag(-pathLoss) * rxWaveform followed by applyRxGain (uePhy.m:743-751, :935-940; gNBPhy.m:852-860) applied
// to the frequency-domain channel matrix the device CDL generator produces.  The reference evaluates these once per packet and
// link in scalar MATLAB; with moving targets / a 19-cell layout the link set is re-evaluated every CPI.
#include "link.cuh"
#include "ctx.cuh"
#include <cmath>

namespace isac {

constexpr double kLight = 299792458.0;

__device__ double pl_uma_umi(bool uma, double fc, bool los, double d2, double d3, double hBS, double hUT) {
    const double fG = fc * 1e-9, hE = 1.0;
    const double dBP = 4.0 * (hBS - hE) * (hUT - hE) * fc / kLight;
    const double a = uma ? 28.0 : 32.4, b1 = uma ? 22.0 : 21.0, c2 = uma ? 9.0 : 9.5;
    const double lf = 20.0 * log10(fG);
    double plLos;
    if (d2 <= dBP) plLos = a + b1 * log10(d3) + lf;
    else plLos = a + 40.0 * log10(d3) + lf - c2 * log10(dBP * dBP + (hBS - hUT) * (hBS - hUT));
    if (los) return plLos;
    const double plN = uma ? 13.54 + 39.08 * log10(d3) + lf - 0.6 * (hUT - 1.5)
                           : 35.3 * log10(d3) + 22.4 + 21.3 * log10(fG) - 0.3 * (hUT - 1.5);
    return fmax(plLos, plN);
}

__device__ double pl_rma(double fc, bool los, double d2, double d3, double hBS, double hUT) {
    const double fG = fc * 1e-9, h = 5.0, W = 20.0, pi = 3.14159265358979323846;
    const double dBP = 2.0 * pi * hBS * hUT * fc / kLight;
    auto pl1 = [&](double d) {
        return 20.0 * log10(40.0 * pi * d * fG / 3.0) + fmin(0.03 * pow(h, 1.72), 10.0) * log10(d) -
               fmin(0.044 * pow(h, 1.72), 14.77) + 0.002 * log10(h) * d;
    };
    const double plLos = d2 <= dBP ? pl1(d3) : pl1(dBP) + 40.0 * log10(d3 / dBP);
    if (los) return plLos;
    const double t = log10(11.75 * hUT);
    const double plN = 161.04 - 7.1 * log10(W) + 7.5 * log10(h) - (24.37 - 3.7 * (h / hBS) * (h / hBS)) * log10(hBS) +
                       (43.42 - 3.1 * log10(hBS)) * (log10(d3) - 3.0) + 20.0 * log10(fG) - (3.2 * t * t - 4.97);
    return fmax(plLos, plN);
}

__global__ void __launch_bounds__(128)
pathloss_kernel(int scenario, double fc, int n, const double* __restrict__ bs, const double* __restrict__ ue, const int* __restrict__ los,
                double* __restrict__ pl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double dx = ue[3 * i] - bs[3 * i], dy = ue[3 * i + 1] - bs[3 * i + 1], dz = ue[3 * i + 2] - bs[3 * i + 2];
    const double d2 = sqrt(dx * dx + dy * dy), d3 = sqrt(dx * dx + dy * dy + dz * dz);
    if (d3 == 0.0) { pl[i] = 0.0; return; }   // identical positions: 0 dB instead of -Inf (config5GNRModels.m:32-33)
    const double hBS = bs[3 * i + 2], hUT = ue[3 * i + 2];
    const bool l = scenario == kPlFspl ? true : los[i] != 0;
    double v;
    switch (scenario) {
        case kPlUMa: v = pl_uma_umi(true, fc, l, d2, d3, hBS, hUT); break;
        case kPlUMi: v = pl_uma_umi(false, fc, l, d2, d3, hBS, hUT); break;
        case kPlRMa: v = pl_rma(fc, l, d2, d3, hBS, hUT); break;
        case kPlInH: {
            const double fG = fc * 1e-9, pL = 32.4 + 17.3 * log10(d3) + 20.0 * log10(fG);
            v = l ? pL : fmax(pL, 38.3 * log10(d3) + 17.30 + 24.9 * log10(fG));
            break;
        }
        default: {   // fspl(R, lambda) = 20 log10(4 pi R / lambda), clipped at 0 dB like the toolbox function
            v = fmax(20.0 * log10(4.0 * 3.14159265358979323846 * d3 * fc / kLight), 0.0);
            break;
        }
    }
    pl[i] = v;
}

__global__ void __launch_bounds__(256)
link_scale_kernel(float2* __restrict__ H, long long elems, const double* __restrict__ pl, double rxGainDb) {
    const float s = (float)pow(10.0, (rxGainDb - pl[blockIdx.y]) / 20.0);   // db2mag(-pathLoss) * 10^(RxGain/20)
    float2* __restrict__ h = H + (long long)blockIdx.y * elems;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
        float2 v = h[i];
        v.x *= s;
        v.y *= s;
        h[i] = v;
    }
}

int pathloss_run(Ctx* ctx, int scenario, double fcHz, int n, const double* bsPos, const double* uePos, const int* los, double* plDb,
                 cudaStream_t st) {
    if (scenario < kPlUMa || scenario > kPlFspl || !(fcHz > 0.0) || n < 0 || !bsPos || !uePos || !plDb || (scenario != kPlFspl && !los)) {
        set_error(ctx, "pathloss: invalid argument (scenario 0 UMa, 1 UMi, 2 RMa, 3 InH, 4 fspl; the InF-* scenarios are not built)");
        return kErrInvalidArg;
    }
    if (n == 0) return kOk;
    void *dB = nullptr, *dU = nullptr, *dL = nullptr, *dP = nullptr;
    int s;
    if ((s = ctx_scratch(ctx, 20, sizeof(double) * 3 * n, &dB)) || (s = ctx_scratch(ctx, 21, sizeof(double) * 3 * n, &dU)) ||
        (s = ctx_scratch(ctx, 22, sizeof(int) * n, &dL)) || (s = ctx_scratch(ctx, 23, sizeof(double) * n, &dP)))
        return s;
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(dB, bsPos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(dU, uePos, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    if (los) ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(dL, los, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    pathloss_kernel<<<(n + 127) / 128, 128, 0, st>>>(scenario, fcHz, n, (const double*)dB, (const double*)dU, (const int*)dL, (double*)dP);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(plDb, dP, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    return kOk;
}

int link_scale_run(Ctx* ctx, float2* H, long long elems, int nLinks, const double* plDb, double rxGainDb, cudaStream_t st) {
    if (!H || elems < 1 || nLinks < 1 || nLinks > 65535 || !plDb) {
        set_error(ctx, "link budget: invalid argument");
        return kErrInvalidArg;
    }
    void* dP = nullptr;
    int s;
    if ((s = ctx_scratch(ctx, 23, sizeof(double) * nLinks, &dP))) return s;
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(dP, plDb, sizeof(double) * nLinks, cudaMemcpyHostToDevice, st));
    long long bx = (elems + 255) / 256;
    const long long cap = 8LL * ctx_num_sms(ctx);
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, nLinks);
    link_scale_kernel<<<grid, 256, 0, st>>>(H, elems, (const double*)dP, rxGainDb);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    count_launches(ctx, 1);
    return kOk;
}

}  // namespace isac
