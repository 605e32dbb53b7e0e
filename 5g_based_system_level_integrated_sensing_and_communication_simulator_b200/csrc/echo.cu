// K1 + K2: mono-static radar echo synthesis fused with CP-OFDM demodulation.
//
// Replaces sensing.channelModels.basicRadarChannel (+sensing/+channelModels/basicRadarChannel.m:8-74)
// and the nrOFDMDemodulate call of sensing.monoStaticSensing (+sensing/monoStaticSensing.m:13-21).
//
// basicRadarChannel, restated analytically (the up-converted waveform never exists):
//   rx[n,r] = sum_i a_i[r] * w_i[n] + sqrt(N0/2) * z[n,r],                        0 <= n < T
//   w_i[n]  = beta_i * exp(2 pi j fd_i Ts n) * u_i[n - s_i],   u_i[m] = sum_t tx[m,t] a_i[t]  (0 for m < 0)
//   s_i = ceil(2 R_i / (c Ts)) (:21-22), fd_i = 2 v_i / lambda (:25),
//   beta_i = largeScaleFading_i * exp(-2 pi j fc Ts s_i): the carrier terms of :30, :42, :73 cancel to
//   this constant, evaluated in float64 on the host (2 pi fc t itself does not fit float32).
// OFDM demodulation is linear, so per OFDM symbol the kernel transforms the nTargets streams w_i
// (not the nAnts antenna streams) and combines them per antenna in the frequency domain:
//   echoGrid[k,s,r] = ramp_s[k] * ( sum_i a_i[r] FFT{w_i}[bin(k)] + noise term ).
// Noise: (a) explicit time-domain standard-normal tensor z (parity mode; one extra FFT per antenna),
//        (b) counter-based Philox generated directly in the frequency domain (white Gaussian noise is
//            invariant under the unitary DFT up to the sqrt(Nfft) scale), or (c) none.
#include "echo.cuh"
#include "ctx.cuh"
#include "fft_core.cuh"
#include <cmath>
#include <vector>

namespace isac {

struct EchoDev {
    const float2* tx;      // [T x nTx]
    const float2* noise;   // [T x nAnts] standard normals, or nullptr
    const float2* steer;   // [nAnts x nTgt] float2 (device) when it does not fit the inline table
    const float2* beamed;  // [T x nTgt] u_i[m] = sum_t tx[m,t] a_i[t] from echo_beamform_kernel, or nullptr (gather in place)
    FftTw tw;
    float2* out;           // [nSc x nSymOut x nAnts]
    long long T;
    int nTx, nAnts, nTgt, nSymRx, nSymOut, nfft, nSc;
    int noiseMode;         // 0 none, 1 explicit time-domain tensor, 2 generated (frequency domain)
    float noiseSigma;      // sqrt(N0/2)
    unsigned long long seed;
    double fcTsFrac;       // frac(fc*Ts): the receiver down-conversion also rotates the noise (:68-74)
    int shift[kEchoMaxTargets];
    float2 beta[kEchoMaxTargets];
    double fdTs[kEchoMaxTargets];
    // symbol timing is periodic per subframe: start(s) = (s / symPer) * subframeLen + startTab[s % symPer]
    int symPer, subframeLen;
    int cpTab[kEchoMaxSymPerSubframe];
    int startTab[kEchoMaxSymPerSubframe];
    int steerInline;       // 1: steering vectors travel in steerTab (no upload, no synchronisation)
    float2 steerTab[kEchoInlineSteer];
};

// Philox4x32-10
template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < ROUNDS; ++i) {
        const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float2 gauss_pair(unsigned a, unsigned b) {
    const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // (0,1)
    const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}
// Box-Muller on the special-function unit (lg2 / sqrt / sin / cos approximations, absolute error ~1e-6 of a unit-variance
// sample): the frequency-domain noise of the production path draws 2 nAnts normals per resource element, which made
// echo_combine_kernel ALU-bound with the accurate forms above.
__device__ __forceinline__ float2 gauss_pair_fast(unsigned a, unsigned b) {
    const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // (0,1)
    const float ang = ((float)b + 0.5f) * 1.4629180792671596e-9f;   // 2 pi u2
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * __log2f(u1)));   // sqrt(-2 ln u1)
    return make_float2(r * __cosf(ang), r * __sinf(ang));
}

// noise sample after the receive mixer: z[n] * exp(-2 pi j fc Ts n)   (basicRadarChannel.m:69,73-74)
__device__ __forceinline__ float2 mixed_noise(const EchoDev& p, const float2* __restrict__ z, long long n) {
    const double c = p.fcTsFrac * (double)n;
    float sn, cs;
    sincospif(-2.0f * (float)(c - floor(c)), &sn, &cs);
    return cmul(__ldg(z + n), make_float2(cs, sn));
}

// stream sample w_i[n] (see file header)
__device__ __forceinline__ float2 target_sample(const EchoDev& p, int i, long long n, long long n0, double base,
                                                const float2* __restrict__ a /*steer column i*/) {
    const long long m = n - p.shift[i];
    if (n >= p.T || m < 0) return make_float2(0.f, 0.f);
    float2 u = make_float2(0.f, 0.f);
    if (p.beamed) u = __ldg(p.beamed + (long long)i * p.T + m);
    else for (int t = 0; t < p.nTx; ++t) {
        const float2 x = __ldg(p.tx + (long long)t * p.T + m);
        const float2 at = a[t];
        u.x += x.x * at.x - x.y * at.y;
        u.y += x.x * at.y + x.y * at.x;
    }
    const float ph = (float)base + (float)p.fdTs[i] * (float)(n - n0);  // cycles
    float s, c;
    sincospif(2.0f * ph, &s, &c);
    return cmul(cmul(u, make_float2(c, s)), p.beta[i]);
}

// Pass 0: transmit beamforming of every target stream, u_i[m] = sum_t tx[m,t] a_i[t] (basicRadarChannel.m:51 rank-1 spatial
// response, a_t == a_r).  One thread per sample: the nTx antenna streams are read ONCE (coalesced along m) for all targets;
// pass 1 then reads one beamformed stream per (symbol, target) instead of gathering nTx streams per target.
template <int NTG>   // NTG = nTgt rounded up to a power of two: the per-target loops unroll without dead predicated work
__global__ void __launch_bounds__(256)
echo_beamform_kernel(const EchoDev p, float2* __restrict__ U) {
    __shared__ float2 steerS[kEchoInlineSteer];
    for (int i = threadIdx.x; i < p.nAnts * p.nTgt && i < kEchoInlineSteer; i += blockDim.x)
        steerS[i] = p.steerInline ? p.steerTab[i] : p.steer[i];
    __syncthreads();
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= p.T) return;
    float2 u[NTG];
#pragma unroll
    for (int i = 0; i < NTG; ++i) u[i] = make_float2(0.f, 0.f);
    for (int t = 0; t < p.nTx; ++t) {
        const float2 x = ld_stream(p.tx + (long long)t * p.T + m);
#pragma unroll
        for (int i = 0; i < NTG; ++i)
            if (i < p.nTgt) {   // same accumulation order over t as the in-place gather of target_sample
                const float2 at = steerS[i * p.nAnts + t];
                u[i].x += x.x * at.x - x.y * at.y;
                u[i].y += x.x * at.y + x.y * at.x;
            }
    }
#pragma unroll
    for (int i = 0; i < NTG; ++i)
        if (i < p.nTgt) U[(long long)i * p.T + m] = u[i];
}

// Pass 1: one CTA per (OFDM symbol, stream).  Streams 0..nTgt-1 are the target waveforms w_i; with an explicit noise
// tensor (parity mode) streams nTgt..nTgt+nAnts-1 are the mixed noise of each antenna.  The spectrum of the stream at the
// nSc occupied subcarriers goes to W[(stream*nSymRx + s)*nSc + k] (17.6 MB at cfg2: stays in L2 for pass 2).
template <int R1, int R2>
__global__ void __launch_bounds__(R1 * R2)
echo_stream_fft_kernel(const EchoDev p, float2* __restrict__ W) {
    using G = FftGeom<R1, R2, true>;
    constexpr int NT = G::NT, NF = G::N;
    extern __shared__ float2 smem[];
    float2* fftbuf = smem;
    float2* steerS = smem + G::kElems;   // [nAnts] steering vector of this target
    const int s = blockIdx.x, q = blockIdx.y, tf = threadIdx.x;
    const int cp = p.cpTab[s % p.symPer];
    const int off = cp / 2;                       // fix(cp * CyclicPrefixFraction), fraction 0.5
    const long long n0 = (long long)(s / p.symPer) * p.subframeLen + p.startTab[s % p.symPer] + off;
    const int half = p.nSc / 2;
    float2 v[16];
    if (q < p.nTgt) {
        for (int i = threadIdx.x; i < p.nAnts; i += blockDim.x)
            steerS[i] = p.steerInline ? p.steerTab[q * p.nAnts + i] : p.steer[q * p.nAnts + i];   // a_t == a_r (:36)
        __syncthreads();
        const double c = p.fdTs[q] * (double)n0;
        const double base = c - floor(c);
        auto load = [&](int n) -> float2 { return target_sample(p, q, n0 + n, n0, base, steerS); };
        block_fft<R1, R2, -1, true>(v, fftbuf, 1, tf, p.tw, load);
    } else {
        const float2* __restrict__ z = p.noise + (long long)(q - p.nTgt) * p.T;
        auto load = [&](int n) -> float2 {
            const long long nn = n0 + n;
            if (nn >= p.T) return make_float2(0.f, 0.f);
            return mixed_noise(p, z, nn);
        };
        block_fft<R1, R2, -1, true>(v, fftbuf, 1, tf, p.tw, load);
    }
    float2* __restrict__ Wq = W + ((size_t)q * p.nSymRx + s) * p.nSc;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        const int bin = tf + NT * d;
        int k = -1;
        if (bin < p.nSc - half) k = bin + half;
        else if (bin >= NF - half) k = bin - (NF - half);
        if (k >= 0) Wq[k] = v[d];
    }
}

// Pass 2: echoGrid[k,s,r] = ramp_s[k] * ( sum_i a_i[r] W_i[k,s] + noise ), zero beyond the demodulated symbols.
// One thread per (subcarrier, symbol) loops over the receive antennas: the nTgt stream spectra and the phase ramp are
// fetched / evaluated once for all antennas, and one Philox call feeds two antennas.
template <int NTG>
__global__ void __launch_bounds__(256)
echo_combine_kernel(const EchoDev p, const float2* __restrict__ W, int NF) {
    __shared__ float2 steerS[kEchoInlineSteer];
    const bool steerInSmem = p.steerInline || p.nAnts * p.nTgt <= kEchoInlineSteer;
    for (int i = threadIdx.x; i < p.nAnts * p.nTgt && i < kEchoInlineSteer; i += blockDim.x)
        steerS[i] = p.steerInline ? p.steerTab[i] : p.steer[i];
    __syncthreads();
    const int s = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.nSc) return;
    float2* __restrict__ o = p.out + (long long)s * p.nSc + k;      // + r * nSymOut * nSc
    const long long page = (long long)p.nSymOut * p.nSc;
    if (s >= p.nSymRx) {  // zero padding up to txDimension(2) (monoStaticSensing.m:19-21)
        for (int r = 0; r < p.nAnts; ++r) o[r * page] = make_float2(0.f, 0.f);
        return;
    }
    float2 w[NTG];
#pragma unroll
    for (int i = 0; i < NTG; ++i)
        w[i] = i < p.nTgt ? __ldcs(W + ((size_t)i * p.nSymRx + s) * p.nSc + k) : make_float2(0.f, 0.f);
    const int cp = p.cpTab[s % p.symPer];
    const int off = cp / 2, half = p.nSc / 2;
    // undo the early FFT window start: exp(+2 pi j kk (cp - off)/Nfft), kk = k - nSc/2
    const float rampStep = (float)(cp - off) / (float)NF;  // cycles per subcarrier index
    float sn, cs;
    sincospif(2.0f * rampStep * (float)(k - half), &sn, &cs);
    const float2 ramp = make_float2(cs, sn);
    const float sc = p.noiseSigma * sqrtf((float)NF);
    uint4 rn = make_uint4(0u, 0u, 0u, 0u);
    for (int r = 0; r < p.nAnts; ++r) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NTG; ++i)
            if (i < p.nTgt) {
                const float2 ar = steerInSmem ? steerS[i * p.nAnts + r] : p.steer[i * p.nAnts + r];
                acc.x += w[i].x * ar.x - w[i].y * ar.y;
                acc.y += w[i].x * ar.y + w[i].y * ar.x;
            }
        if (p.noiseMode == 1) {
            const float2 z = __ldcs(W + ((size_t)(p.nTgt + r) * p.nSymRx + s) * p.nSc + k);
            acc.x += p.noiseSigma * z.x;
            acc.y += p.noiseSigma * z.y;
        } else if (p.noiseMode == 2) {
            if ((r & 1) == 0)   // one counter block = four uniforms = two complex normals (antennas r, r+1)
                rn = philox4x32<7>(make_uint4((unsigned)k, (unsigned)s, (unsigned)(r >> 1), 0x15ACu),   // Philox4x32-7 (crush-resistant)
                                   make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32)));
            const float2 g = (r & 1) ? gauss_pair_fast(rn.z, rn.w) : gauss_pair_fast(rn.x, rn.y);
            acc.x += sc * g.x;
            acc.y += sc * g.y;
        }
        o[r * page] = cmul(acc, ramp);
    }
}

// basicRadarChannel alone: rxWaveform [T x nAnts]
__global__ void __launch_bounds__(256)
radar_channel_kernel(const EchoDev p, float2* __restrict__ rxWave) {
    extern __shared__ float2 steerS[];
    for (int i = threadIdx.x; i < p.nAnts * p.nTgt; i += blockDim.x) steerS[i] = p.steerInline ? p.steerTab[i] : p.steer[i];
    __syncthreads();
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.T) return;
    float2 w[kEchoMaxTargets];
    for (int i = 0; i < p.nTgt; ++i) {
        const double c = p.fdTs[i] * (double)n;
        w[i] = target_sample(p, i, n, n, c - floor(c), steerS + i * p.nAnts);
    }
    for (int r = 0; r < p.nAnts; ++r) {
        float2 acc = make_float2(0.f, 0.f);
        for (int i = 0; i < p.nTgt; ++i) {
            const float2 ar = steerS[i * p.nAnts + r];
            acc.x += w[i].x * ar.x - w[i].y * ar.y;
            acc.y += w[i].x * ar.y + w[i].y * ar.x;
        }
        if (p.noiseMode == 1) {
            const float2 z = mixed_noise(p, p.noise + (long long)r * p.T, n);
            acc.x += p.noiseSigma * z.x;
            acc.y += p.noiseSigma * z.y;
        } else if (p.noiseMode == 2) {
            const uint4 ctr = make_uint4((unsigned)(n & 0xffffffffu), (unsigned)(n >> 32), (unsigned)r, 0x7D0Au);
            const uint4 rn = philox4x32(ctr, make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32)));
            const float2 g = gauss_pair(rn.x, rn.y);
            acc.x += p.noiseSigma * g.x;
            acc.y += p.noiseSigma * g.y;
        }
        rxWave[(long long)r * p.T + n] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int fill_targets(Ctx* ctx, const EchoConfig& c, EchoDev& d, std::vector<float2>& steerHost) {
    const double kC = 299792458.0;
    const double lambda = kC / c.fc, Ts = 1.0 / c.fs;
    int nT = 0;
    steerHost.clear();
    for (int i = 0; i < c.nTargets; ++i) {
        if (c.los && c.los[i] != 1) continue;  // NLoS targets reflect nothing (basicRadarChannel.m:57-58)
        if (nT >= kEchoMaxTargets) {
            set_error(ctx, "echo: more LoS targets than kEchoMaxTargets");
            return kErrCapacity;
        }
        const double delay = 2.0 * c.range[i] / kC;                    // :21
        const long long sh = (long long)std::ceil(delay / Ts);         // :22
        const double fd = 2.0 * c.velocity[i] / lambda;                // :25
        double cyc = c.fc * Ts * (double)sh;                           // carrier phase of the delay, in cycles
        cyc -= std::floor(cyc);
        const double ang = -2.0 * M_PI * cyc;
        d.shift[nT] = (int)sh;
        d.beta[nT] = make_float2((float)(c.largeScaleFading[i] * std::cos(ang)), (float)(c.largeScaleFading[i] * std::sin(ang)));
        d.fdTs[nT] = fd * Ts;
        for (int a = 0; a < c.nTx; ++a)
            steerHost.push_back(make_float2((float)c.steeringVec[2 * ((size_t)i * c.nTx + a)],
                                            (float)c.steeringVec[2 * ((size_t)i * c.nTx + a) + 1]));
        ++nT;
    }
    d.nTgt = nT;
    if (nT == 0) {
        set_error(ctx, "basicRadarChannel: no LoS target (the reference produces an empty waveform)");
        return kErrNoLosTarget;
    }
    return kOk;
}

static int common_dev(Ctx* ctx, const EchoConfig& c, const float2* tx, const float2* noise, int noiseMode,
                      unsigned long long seed, EchoDev& d, cudaStream_t st) {
    if (!tx || c.T < 1 || c.nTx < 1 || c.nTargets < 0) {
        set_error(ctx, "echo: invalid argument");
        return kErrInvalidArg;
    }
    if (noiseMode == 1 && !noise) {
        set_error(ctx, "echo: noise mode 1 needs the standard-normal tensor");
        return kErrInvalidArg;
    }
    d = EchoDev{};
    std::vector<float2> steerHost;
    int s = fill_targets(ctx, c, d, steerHost);
    if (s) return s;
    if (steerHost.size() <= (size_t)kEchoInlineSteer) {
        d.steerInline = 1;
        for (size_t i = 0; i < steerHost.size(); ++i) d.steerTab[i] = steerHost[i];
        d.steer = nullptr;
    } else {
        void* dSteer = nullptr;
        if ((s = ctx_scratch(ctx, 15, sizeof(float2) * steerHost.size(), &dSteer))) return s;
        void* pin = nullptr;
        if ((s = ctx_pinned(ctx, 7, sizeof(float2) * steerHost.size(), &pin))) return s;
        ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));  // the pinned staging buffer may still be in flight
        std::memcpy(pin, steerHost.data(), sizeof(float2) * steerHost.size());
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(dSteer, pin, sizeof(float2) * steerHost.size(), cudaMemcpyHostToDevice, st));
        d.steerInline = 0;
        d.steer = (const float2*)dSteer;
    }
    d.tx = tx;
    d.noise = noise;
    d.T = c.T;
    d.nTx = c.nTx;
    d.nAnts = c.nTx;  // Rx and Tx share the array (radarParams.m:88,105)
    d.noiseMode = noiseMode;
    d.noiseSigma = (float)std::sqrt(c.N0 / 2.0);  // basicRadarChannel.m:67
    d.seed = seed;
    const double ft = c.fc / c.fs;
    d.fcTsFrac = ft - std::floor(ft);
    return kOk;
}

int radar_channel_run(Ctx* ctx, const EchoConfig& c, const float2* tx, const float2* noise, int noiseMode,
                      unsigned long long seed, float2* rxWave, cudaStream_t st) {
    EchoDev d;
    int s = common_dev(ctx, c, tx, noise, noiseMode, seed, d, st);
    if (s) return s;
    const size_t smem = sizeof(float2) * (size_t)d.nAnts * d.nTgt;
    radar_channel_kernel<<<(unsigned)((c.T + 255) / 256), 256, smem, st>>>(d, rxWave);
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

template <int R1, int R2>
static cudaError_t launch_echo(const EchoDev& d, float2* W, cudaStream_t st) {
    using G = FftGeom<R1, R2, true>;
    const size_t smem = sizeof(float2) * ((size_t)G::kElems + d.nAnts);
    auto k = echo_stream_fft_kernel<R1, R2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int streams = d.nTgt + (d.noiseMode == 1 ? d.nAnts : 0);
    const dim3 gB((unsigned)((d.T + 255) / 256)), gC((d.nSc + 255) / 256, d.nSymOut);
    float2* U = const_cast<float2*>(d.beamed);
#define ISAC_ECHO_NTG(NTG_)                                                                   \
    do {                                                                                   \
        if (U) echo_beamform_kernel<NTG_><<<gB, 256, 0, st>>>(d, U);                           \
        if (d.nSymRx > 0 && streams > 0) k<<<dim3(d.nSymRx, streams), G::NT, smem, st>>>(d, W); \
        echo_combine_kernel<NTG_><<<gC, 256, 0, st>>>(d, W, G::N);                             \
    } while (0)
    if (d.nTgt <= 1) ISAC_ECHO_NTG(1);
    else if (d.nTgt <= 2) ISAC_ECHO_NTG(2);
    else if (d.nTgt <= 4) ISAC_ECHO_NTG(4);
    else if (d.nTgt <= 8) ISAC_ECHO_NTG(8);
    else ISAC_ECHO_NTG(16);
#undef ISAC_ECHO_NTG
    return cudaGetLastError();
}

int mono_static_sensing_run(Ctx* ctx, const EchoConfig& c, const float2* tx, const float2* noise, int noiseMode,
                            unsigned long long seed, float2* echoGrid, int* nSymRxOut, cudaStream_t st) {
    if (c.nfft < 128 || c.nfft > 4096 || (c.nfft & (c.nfft - 1)) || c.nSc < 2 || c.nSc > c.nfft || c.symbolsPerSubframe < 1) {
        set_error(ctx, "monoStaticSensing: unsupported numerology");
        return kErrUnsupported;
    }
    if (c.symbolsPerSubframe > kEchoMaxSymPerSubframe) {
        set_error(ctx, "monoStaticSensing: more than 56 symbols per subframe (subcarrier spacing > 60 kHz) is not supported");
        return kErrUnsupported;
    }
    // whole symbols contained in the waveform (nrOFDMDemodulate)
    int nSymRx = 0;
    long long acc = 0;
    for (int s = 0;; ++s) {
        const int len = c.cpLengths[s % c.symbolsPerSubframe] + c.nfft;
        if (acc + len > c.T) break;
        acc += len;
        ++nSymRx;
    }
    const int nSymOut = nSymRx > c.nSymTx ? nSymRx : c.nSymTx;  // pad up to txDimension(2) (:19-21)
    if (nSymRxOut) *nSymRxOut = nSymOut;
    if (!echoGrid) return kOk;  // size query
    EchoDev d;
    int s = common_dev(ctx, c, tx, noise, noiseMode, seed, d, st);
    if (s) return s;
    d.symPer = c.symbolsPerSubframe;
    int off = 0;
    for (int q = 0; q < c.symbolsPerSubframe; ++q) {
        d.cpTab[q] = c.cpLengths[q];
        d.startTab[q] = off;
        off += c.cpLengths[q] + c.nfft;
    }
    d.subframeLen = off;
    d.nSymRx = nSymRx;
    d.nSymOut = nSymOut;
    d.nfft = c.nfft;
    d.nSc = c.nSc;
    d.out = echoGrid;
    ctx_fft_tw(ctx, c.nfft, &d.tw.tw1, &d.tw.tw2);
    void* dW = nullptr;   // stream spectra between the two passes
    const size_t streams = (size_t)d.nTgt + (noiseMode == 1 ? d.nAnts : 0);
    if ((s = ctx_scratch(ctx, 17, sizeof(float2) * (streams ? streams : 1) * (nSymRx ? nSymRx : 1) * c.nSc, &dW))) return s;
    void* dU = nullptr;   // beamformed target streams (pass 0) when the steering table fits shared memory
    if (d.nAnts * d.nTgt <= kEchoInlineSteer) {
        if ((s = ctx_scratch(ctx, 18, sizeof(float2) * (size_t)d.nTgt * c.T, &dU))) return s;
    }
    d.beamed = (const float2*)dU;
    cudaError_t e;
    const int pr = prof_begin(ctx, kProfEcho, st);
    switch (c.nfft) {
        case 128: e = launch_echo<1, 8>(d, (float2*)dW, st); break;
        case 256: e = launch_echo<1, 16>(d, (float2*)dW, st); break;
        case 512: e = launch_echo<2, 16>(d, (float2*)dW, st); break;
        case 1024: e = launch_echo<4, 16>(d, (float2*)dW, st); break;
        case 2048: e = launch_echo<8, 16>(d, (float2*)dW, st); break;
        default: e = launch_echo<16, 16>(d, (float2*)dW, st); break;
    }
    prof_end(ctx, pr, st);
    count_launches(ctx, dU ? 3 : 2);
    ISAC_CUDA_CHECK(ctx, e);
    return kOk;
}

}  // namespace isac
