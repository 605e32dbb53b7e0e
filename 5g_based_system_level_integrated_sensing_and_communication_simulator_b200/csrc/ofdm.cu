// CP-OFDM modulation of the sensing transmit grid on the device.
//
// Replaces the `txWaveform = signalAmp * nrOFDMModulate(carrier, txGrid)` step of the gNB PHY
// (+phyLayer/gNBPhy.m:599, grid / waveform accumulation for sensing at :604-612) so that the waveform handed to
// sensing.monoStaticSensing never has to exist on the host: the grid (8*nSc bytes per symbol and antenna) is uploaded
// once, the waveform (8*(Nfft+CP) bytes) is produced in HBM.
//
//   x_s[n] = (scale / Nfft) * sum_k grid[k, s, r] * exp(+2 pi i bin(k) n / Nfft),   bin(k) = (k - nSc/2) mod Nfft
//   wave[start_s + m, r] = x_s[(m - cp_s) mod Nfft],   0 <= m < cp_s + Nfft          (TS 38.211 5.3.1, CP = tail copy)
//
// Windowing (optional, `windowing` = N samples; nrOFDMModulate's 'Windowing' argument, PARITY-UNPINNED toolbox internals
// restated from the documented W-OLA scheme): every symbol's cyclic extension grows by N samples, its first N and its last N
// samples are shaped by the raised cosine p[i] = 0.5 (1 - sin(pi (N + 1 - 2 i) / (2 N))), i = 1..N, and the rising head
// overlaps (adds to) the falling tail of the symbol before it, so the waveform length does not change.  N = 0 gives the plain
// CP-OFDM waveform, the exact inverse of the demodulator of echo.cu.
// One 16-values-per-thread block FFT (fft_core.cuh, packed FP32x2 butterflies) per (symbol, antenna); loads and
// stores are coalesced float2 runs; the cyclic prefix is written from the same registers.
#include "ofdm.cuh"
#include "ctx.cuh"
#include "fft_core.cuh"

namespace isac {

struct OfdmDev {
    const float2* grid;  // [nSc x nSym x nAnts]
    float2* wave;        // [T x nAnts]
    FftTw tw;
    long long T;
    int nSc, nSym, nAnts, nfft;
    float scale;         // caller's scale / Nfft
    int symPer, subframeLen;
    int nWin, symPhase;  // window length; position of symbol 0 in the CP pattern
    long long gridStride, sampleOffset;
    float2* head;        // [nSym][nAnts][nWin] rising-head samples of every symbol (windowing only)
    int cpTab[kOfdmMaxSymPerSubframe];
    int startTab[kOfdmMaxSymPerSubframe];
};

template <int R1, int R2>
__global__ void __launch_bounds__(R1 * R2)
ofdm_modulate_kernel(const OfdmDev p) {
    using G = FftGeom<R1, R2, true>;
    extern __shared__ float2 smem[];
    const int tf = threadIdx.x;
    const int s = blockIdx.x, r = blockIdx.y;
    const int N = G::N, nSc = p.nSc, half = nSc / 2;
    const float2* __restrict__ g = p.grid + ((long long)r * p.gridStride + s) * nSc;
    auto load = [&](int n) -> float2 {  // IFFT input bin n holds subcarrier k = n + nSc/2 (n < ceil(nSc/2)) or n - (N - nSc/2)
        int k = -1;
        if (n < nSc - half) k = n + half;
        else if (n >= N - half) k = n - (N - half);
        return k >= 0 ? __ldg(g + k) : make_float2(0.f, 0.f);
    };
    float2 v[16];
    block_fft<R1, R2, +1, true>(v, smem, 1, tf, p.tw, load);
    const int sa = s + p.symPhase;                 // position in the subframe's CP pattern
    const int q = sa % p.symPer;
    const int cp = p.cpTab[q];
    const long long start = (long long)(sa / p.symPer) * p.subframeLen + p.startTab[q] -
                            ((long long)(p.symPhase / p.symPer) * p.subframeLen + p.startTab[p.symPhase % p.symPer]) + p.sampleOffset;
    float2* __restrict__ out = p.wave + (long long)r * p.T + start;
    float2* __restrict__ hd = p.nWin ? p.head + ((long long)s * p.nAnts + r) * p.nWin : nullptr;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        const int n = tf + G::NT * d;
        const float2 x = pk_scale(v[d], p.scale);
        out[cp + n] = x;
        if (n >= N - cp) out[n - (N - cp)] = x;  // cyclic prefix = the last cp samples of the symbol
        if (p.nWin && n >= N - cp - p.nWin && n < N - cp) hd[n - (N - cp - p.nWin)] = x;   // the N samples before the prefix
    }
}

// Windowing fix-up: the N samples in front of every symbol's prefix (the tail of the symbol before it, possibly written by an
// earlier call into the same resident buffer) become fall * tail + rise * head; the first N samples of a symbol's own prefix+body
// are NOT touched (the documented scheme shapes the extension, not the nominal prefix).  One thread per (sample, symbol, antenna).
__global__ void __launch_bounds__(128)
ofdm_window_kernel(const OfdmDev p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y, r = blockIdx.z;
    if (j >= p.nWin) return;
    const int sa = s + p.symPhase, q = sa % p.symPer;
    const long long start = (long long)(sa / p.symPer) * p.subframeLen + p.startTab[q] -
                            ((long long)(p.symPhase / p.symPer) * p.subframeLen + p.startTab[p.symPhase % p.symPer]) + p.sampleOffset;
    if (start < p.nWin) return;                   // first symbol of the buffer: nothing in front of it
    const float rise = 0.5f * (1.0f - sinpif((float)(p.nWin + 1 - 2 * (j + 1)) / (float)(2 * p.nWin)));
    float2* __restrict__ o = p.wave + (long long)r * p.T + start - p.nWin + j;
    const float2 h = p.head[((long long)s * p.nAnts + r) * p.nWin + j];
    const float2 t = *o;
    *o = make_float2((1.0f - rise) * t.x + rise * h.x, (1.0f - rise) * t.y + rise * h.y);   // fall[j] = rise[N-1-j] = 1 - rise[j]
}

template <int R1, int R2>
static cudaError_t launch_mod(const OfdmDev& d, cudaStream_t st) {
    using G = FftGeom<R1, R2, true>;
    const size_t smem = sizeof(float2) * (size_t)G::kElems;
    auto k = ofdm_modulate_kernel<R1, R2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<dim3(d.nSym, d.nAnts), G::NT, smem, st>>>(d);
    return cudaGetLastError();
}

long long ofdm_waveform_length(const OfdmConfig& c) {
    long long T = 0;
    for (int s = 0; s < c.nSym; ++s) T += c.cpLengths[(s + c.symPhase) % c.symbolsPerSubframe] + c.nfft;
    return T;
}

int ofdm_modulate_run(Ctx* ctx, const OfdmConfig& c, const float2* grid, float2* wave, cudaStream_t st) {
    if (c.windowing < 0 || c.symPhase < 0 || c.sampleOffset < 0) {
        set_error(ctx, "ofdmModulate: negative windowing / offset");
        return kErrInvalidArg;
    }
    if (c.nfft < 128 || c.nfft > 4096 || (c.nfft & (c.nfft - 1)) || c.nSc < 2 || c.nSc > c.nfft || c.nSym < 1 || c.nAnts < 1 ||
        c.symbolsPerSubframe < 1 || c.symbolsPerSubframe > kOfdmMaxSymPerSubframe || !c.cpLengths) {
        set_error(ctx, "ofdmModulate: unsupported numerology");
        return kErrUnsupported;
    }
    if (!grid || !wave) {
        set_error(ctx, "ofdmModulate: null pointer");
        return kErrInvalidArg;
    }
    OfdmDev d{};
    d.grid = grid;
    d.wave = wave;
    d.T = c.waveStride > 0 ? c.waveStride : ofdm_waveform_length(c);
    d.gridStride = c.gridStride > 0 ? c.gridStride : c.nSym;
    d.sampleOffset = c.sampleOffset;
    d.nWin = c.windowing;
    d.symPhase = c.symPhase;
    d.nSc = c.nSc;
    d.nSym = c.nSym;
    d.nAnts = c.nAnts;
    d.nfft = c.nfft;
    d.scale = (float)(c.scale / (double)c.nfft);
    d.symPer = c.symbolsPerSubframe;
    int off = 0;
    for (int q = 0; q < c.symbolsPerSubframe; ++q) {
        if (c.cpLengths[q] < 0 || c.cpLengths[q] > c.nfft) {
            set_error(ctx, "ofdmModulate: cyclic prefix longer than the symbol");
            return kErrInvalidArg;
        }
        d.cpTab[q] = c.cpLengths[q];
        d.startTab[q] = off;
        off += c.cpLengths[q] + c.nfft;
    }
    d.subframeLen = off;
    if (d.nWin) {
        for (int q = 0; q < c.symbolsPerSubframe; ++q)
            if (d.nWin > c.cpLengths[q] || d.nWin > c.nfft - c.cpLengths[q]) {
                set_error(ctx, "ofdmModulate: windowing must not exceed the cyclic prefix");
                return kErrInvalidArg;
            }
        void* hd = nullptr;
        int s = ctx_scratch(ctx, 19, sizeof(float2) * (size_t)c.nSym * c.nAnts * d.nWin, &hd);
        if (s) return s;
        d.head = (float2*)hd;
    }
    ctx_fft_tw(ctx, c.nfft, &d.tw.tw1, &d.tw.tw2);
    cudaError_t e;
    const int pr = prof_begin(ctx, kProfOfdmMod, st);
    switch (c.nfft) {
        case 128: e = launch_mod<1, 8>(d, st); break;
        case 256: e = launch_mod<1, 16>(d, st); break;
        case 512: e = launch_mod<2, 16>(d, st); break;
        case 1024: e = launch_mod<4, 16>(d, st); break;
        case 2048: e = launch_mod<8, 16>(d, st); break;
        default: e = launch_mod<16, 16>(d, st); break;
    }
    if (e == cudaSuccess && d.nWin) {
        ofdm_window_kernel<<<dim3((d.nWin + 127) / 128, d.nSym, d.nAnts), 128, 0, st>>>(d);
        e = cudaGetLastError();
        count_launches(ctx, 1);
    }
    prof_end(ctx, pr, st);
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, e);
    return kOk;
}

}  // namespace isac
