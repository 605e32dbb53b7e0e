// K3 + K4: windowed 2D-FFT range-Doppler map and 2D CA-CFAR on sm_100a.
//
// Replaces the vectorised core of sensing.estimation.fft2D (reference
// +sensing/+estimation/fft2D.m:37-46, :61-63) and the phased.CFARDetector2D step configured by
// sensing.detection.cfar2D (+sensing/+detection/cfar2D.m:15-33).
//
// Closed form implemented (derivation in DESIGN.md; N = nIFFT, F = nFFT, M = min(nSym, F)):
//   y[n,s,r]   = (1/sqrt(N)) * sum_k rx[k,s,r] conj(tx[k,s,r]) w1[k] e^{+2 pi i k n / N}
//   RDM[n,q,r] = (1/sqrt(F)) * w2[(n - N/2) mod N] *
//                sum_{s'<M} (-1)^{s'} y[n, (s' + floor(nSym/2)) mod nSym, r] e^{-2 pi i s' q / F}
//   rdResponse = |RDM|^2
// i.e. the reference's dimension-less ifftshift/fftshift cancel on the range and antenna axes,
// leave the second Kaiser window applied at the *shifted* range index, rotate the symbol axis
// before the zero-padded / truncated Doppler FFT, and centre the Doppler axis.
//
// Kernel A (range): one 256-thread group per (s', r, map) column: fused rx*conj(tx)*w1 prologue,
//   N-point inverse FFT in shared memory, w2/sqrt(N)/(-1)^s' epilogue, coalesced float2 stores.
// Kernel B (Doppler): one CTA per tile of RT consecutive range rows of one antenna page; loads are
//   coalesced along the range axis, the F-point FFTs run interleaved in shared memory, the
//   epilogue writes |.|^2 (the Doppler fftshift is the (-1)^s' modulation applied by kernel A).
// Kernel C (CFAR): one thread per cell under test, float64 training sum in a fixed order,
//   strict `>` against alpha*mean, flags + detected-row bitmap.
// Kernel D (compaction): one CTA per (antenna, map): ordered stream compaction of the flags into
//   the detector's 'Detection index' output (CUT order = range fastest) plus peak powers.
#include "rdm.cuh"
#include "fft_core.cuh"
#include <cmath>
#include <vector>

namespace isac {

struct RdmDev {
    const float2* rx;
    const float2* tx;
    const float* win1;
    const float* win2;
    FftTw twR;  // tables of the range IFFT size
    FftTw twD;  // tables of the Doppler FFT size
    float2* inter;
    float* pow;
    int nSc, nSym, nAnts, nIFFT, nFFT, M;
    long long totalCols;  // M * nAnts * batch
};

// ------------------------------------------------------------------------------------------
// Kernel A: range IFFT
// ------------------------------------------------------------------------------------------
template <int R1, int R2>
__global__ void __launch_bounds__((R1 * R2 >= 256 ? R1 * R2 : 256))
rdm_range_ifft_kernel(const RdmDev p) {
    using G = FftGeom<R1, R2, true>;
    extern __shared__ float2 smem[];
    const int local = threadIdx.x / G::NT, tf = threadIdx.x % G::NT;
    const int cpc = blockDim.x / G::NT;
    long long col = (long long)blockIdx.x * cpc + local;
    const bool active = col < p.totalCols;
    if (!active) col = p.totalCols - 1;  // keep the thread in the barriers, drop its stores
    const int sp = (int)(col % p.M);
    const long long page = col / p.M;  // r + nAnts*b
    const int s = (sp + p.nSym / 2) % p.nSym;  // ifftshift on the symbol axis (fft2D.m:44)
    const float2* __restrict__ rx = p.rx + (page * p.nSym + s) * (long long)p.nSc;
    const float2* __restrict__ tx = p.tx + (page * p.nSym + s) * (long long)p.nSc;
    const float* __restrict__ w1 = p.win1;
    const int nSc = p.nSc;
    auto load = [&](int n) -> float2 {
        if (n < nSc) {
            float2 a = ld_stream(rx + n), b = ld_stream(tx + n);
            return cscale(cmulc(a, b), __ldg(w1 + n));  // rx .* conj(tx) .* rngWin  (fft2D.m:37,43)
        }
        return make_float2(0.f, 0.f);
    };
    float2 v[16];
    block_fft<R1, R2, +1, true>(v, smem + local * G::kElems, 1, tf, p.twR, load);
    if (active) {
        float2* __restrict__ out = p.inter + (page * p.M + sp) * (long long)p.nIFFT;
        const float sgn = (sp & 1) ? -1.f : 1.f;  // e^{+i pi s'}: Doppler fftshift folded in
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const int n = tf + G::NT * d;
            out[n] = cscale(v[d], __ldg(p.win2 + n) * sgn);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Kernel A' (N = 4096): persistent range IFFT with TMA staging.  One CTA loops over columns; the rx and tx
// columns of the NEXT column are fetched by two cp.async.bulk (TMA 1-D) copies into shared memory, completion
// signalled on an mbarrier, while the current column's FFT runs from registers -> the long-scoreboard stalls of
// kernel A disappear and 2 CTAs/SM keep the copy engine and the FMA pipe busy at the same time.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int R1, int R2>
__global__ void __launch_bounds__(R1* R2, 2)
rdm_range_ifft_tma_kernel(const RdmDev p) {
    using G = FftGeom<R1, R2, true>;
    static_assert(G::NT == R1 * R2 && R1 == 16, "one column per CTA pass, 16 values per thread");
    extern __shared__ __align__(128) unsigned char smraw[];
    float2* fftbuf = reinterpret_cast<float2*>(smraw);
    float2* stageRx = fftbuf + G::kElems + 8;           // keep 16-byte alignment
    float2* stageTx = stageRx + p.nSc;
    __shared__ __align__(8) unsigned long long bar;
    const int tf = threadIdx.x;
    const unsigned colBytes = (unsigned)p.nSc * sizeof(float2);
    auto col_ptrs = [&](long long col, const float2*& rx, const float2*& tx, int& sp, long long& page) {
        sp = (int)(col % p.M);
        page = col / p.M;
        const int s = (sp + p.nSym / 2) % p.nSym;  // ifftshift on the symbol axis (fft2D.m:44)
        rx = p.rx + (page * p.nSym + s) * (long long)p.nSc;
        tx = p.tx + (page * p.nSym + s) * (long long)p.nSc;
    };
    if (tf == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long col = blockIdx.x;
    if (tf == 0 && col < p.totalCols) {
        const float2 *rx, *tx; int sp; long long page;
        col_ptrs(col, rx, tx, sp, page);
        mbar_expect_tx(&bar, 2 * colBytes);
        tma_load_1d(stageRx, rx, colBytes, &bar);
        tma_load_1d(stageTx, tx, colBytes, &bar);
    }
    unsigned parity = 0;
    const float* __restrict__ w1 = p.win1;
    for (; col < p.totalCols; col += gridDim.x) {
        mbar_wait(&bar, parity);
        parity ^= 1;
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = tf + G::NT * j;
            v[j] = (n < p.nSc) ? cscale(cmulc(stageRx[n], stageTx[n]), __ldg(w1 + n)) : make_float2(0.f, 0.f);
        }
        __syncthreads();  // every thread has consumed the stage
        const long long next = col + gridDim.x;
        if (tf == 0 && next < p.totalCols) {
            const float2 *rx, *tx; int sp; long long page;
            col_ptrs(next, rx, tx, sp, page);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before async writes
            mbar_expect_tx(&bar, 2 * colBytes);
            tma_load_1d(stageRx, rx, colBytes, &bar);
            tma_load_1d(stageTx, tx, colBytes, &bar);
        }
        auto noload = [](int) -> float2 { return make_float2(0.f, 0.f); };
        block_fft<R1, R2, +1, true, decltype(noload), true>(v, fftbuf, 1, tf, p.twR, noload);
        const int sp = (int)(col % p.M);
        const long long page = col / p.M;
        float2* __restrict__ out = p.inter + (page * p.M + sp) * (long long)p.nIFFT;
        const float sgn = (sp & 1) ? -1.f : 1.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const int n = tf + G::NT * d;
            out[n] = cscale(v[d], __ldg(p.win2 + n) * sgn);
        }
        __syncthreads();  // fftbuf is reused by the next column's first pass
    }
}

// ------------------------------------------------------------------------------------------
// Kernel B: Doppler FFT + |.|^2
// ------------------------------------------------------------------------------------------
template <int R1, int R2>
__global__ void __launch_bounds__(512)
rdm_doppler_fft_kernel(const RdmDev p, const int RT, const float invF) {
    using G = FftGeom<R1, R2, false>;
    extern __shared__ float2 smem[];
    const int nl = threadIdx.x % RT, tf = threadIdx.x / RT;
    const int tilesPerPage = p.nIFFT / RT;
    const long long page = blockIdx.x / tilesPerPage;
    const int n = (blockIdx.x % tilesPerPage) * RT + nl;
    const float2* __restrict__ in = p.inter + page * (long long)p.M * p.nIFFT + n;
    const int M = p.M, nIFFT = p.nIFFT;
    auto load = [&](int sp) -> float2 {
        if (sp < M) return __ldcg(in + (long long)sp * nIFFT);
        return make_float2(0.f, 0.f);
    };
    float2 v[16];
    block_fft<R1, R2, -1, false>(v, smem + nl, RT, tf, p.twD, load);
    float* __restrict__ out = p.pow + page * (long long)p.nFFT * nIFFT + n;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        // Doppler-axis fftshift (fft2D.m:46) already applied: kernel A modulated symbol s' by (-1)^s'
        const int q = tf + G::NT * d;
        out[(long long)q * nIFFT] = (v[d].x * v[d].x + v[d].y * v[d].y) * invF;  // abs(rdm).^2 (fft2D.m:61)
    }
}

// ------------------------------------------------------------------------------------------
// Kernel C: CA-CFAR decisions
// ------------------------------------------------------------------------------------------
struct CfarDev {
    const float* pow;
    uint8_t* flags;
    uint32_t* rowmask;
    int nIFFT, nFFT, nAnts;
    int row0, col0;  // 0-based first CUT row / col
    int nCutRows, nCut;
    int gr, gc, hr, hc;  // guard half-sizes, guard+training half-sizes
    int rowWords;
    double alpha, nTrain;
    long long total;  // nCut * nAnts * batch
};

__global__ void __launch_bounds__(256) cfar2d_flags_kernel(const CfarDev p) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.total) return;
    const int i = (int)(gid % p.nCut);
    const long long page = gid / p.nCut;  // r + nAnts*b
    const int row = p.row0 + i % p.nCutRows;
    const int col = p.col0 + i / p.nCutRows;
    const float* __restrict__ P = p.pow + page * (long long)p.nFFT * p.nIFFT;
    double acc = 0.0;
    for (int dc = -p.hc; dc <= p.hc; ++dc) {
        const float* __restrict__ c = P + (long long)(col + dc) * p.nIFFT + row;
        const bool colInGuard = (dc >= -p.gc && dc <= p.gc);
        for (int dr = -p.hr; dr <= p.hr; ++dr) {
            if (colInGuard && dr >= -p.gr && dr <= p.gr) continue;
            acc = acc + (double)__ldcg(c + dr);
        }
    }
    const double thr = p.alpha * (acc / p.nTrain);
    const double x = (double)__ldcg(P + (long long)col * p.nIFFT + row);
    const bool det = x > thr;  // strict (CFARDetector2D)
    p.flags[gid] = det ? 1 : 0;
    if (det) {
        const long long b = page / p.nAnts;
        atomicOr(p.rowmask + b * p.rowWords + (row >> 5), 1u << (row & 31));
    }
}

// Reference configuration (cfar2D.m:32-33): guard [2 2], training [1 1] -> 7x7 window minus the 5x5 guard block.
// Fully unrolled: the 24 training loads are independent and issue back to back (same summation order as above).
__global__ void __launch_bounds__(256) cfar2d_flags_7x7_kernel(const CfarDev p) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.total) return;
    const int i = (int)(gid % p.nCut);
    const long long page = gid / p.nCut;
    const int row = p.row0 + i % p.nCutRows;
    const int col = p.col0 + i / p.nCutRows;
    const float* __restrict__ P = p.pow + page * (long long)p.nFFT * p.nIFFT + (long long)col * p.nIFFT + row;
    float t[24];
    int q = 0;
#pragma unroll
    for (int dc = -3; dc <= 3; ++dc)
#pragma unroll
        for (int dr = -3; dr <= 3; ++dr) {
            if (dc >= -2 && dc <= 2 && dr >= -2 && dr <= 2) continue;
            t[q++] = __ldcg(P + (long long)dc * p.nIFFT + dr);
        }
    const float x = __ldcg(P);
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 24; ++k) acc = acc + (double)t[k];
    const bool det = (double)x > p.alpha * (acc / p.nTrain);
    p.flags[gid] = det ? 1 : 0;
    if (det) atomicOr(p.rowmask + (page / p.nAnts) * p.rowWords + (row >> 5), 1u << (row & 31));
}

// ------------------------------------------------------------------------------------------
// Kernel D: ordered compaction -> 'Detection index' [row; col] (1-based) in CUT order
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) cfar2d_compact_kernel(const CfarDev p, int2* det, float* peak, int32_t* detCount) {
    // each thread owns a contiguous run of CUTs (keeps CUT order), one block-wide exclusive scan of the run counts
    __shared__ int warpTot[32];
    const long long page = blockIdx.x;
    const uint8_t* __restrict__ f = p.flags + page * (long long)p.nCut;
    const float* __restrict__ P = p.pow + page * (long long)p.nFFT * p.nIFFT;
    int2* __restrict__ o = det + page * (long long)p.nCut;
    float* __restrict__ pk = peak + page * (long long)p.nCut;
    const int per = (p.nCut + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(lo + per, p.nCut);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += f[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warpTot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warpTot[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += v;
        }
        warpTot[lane] = w;  // inclusive totals of the warps
    }
    __syncthreads();
    int off = incl - cnt + (warp ? warpTot[warp - 1] : 0);
    for (int i = lo; i < hi; ++i)
        if (f[i]) {
            const int row = p.row0 + i % p.nCutRows, col = p.col0 + i / p.nCutRows;
            o[off] = make_int2(row + 1, col + 1);
            pk[off] = P[(long long)col * p.nIFFT + row];
            ++off;
        }
    if (threadIdx.x == blockDim.x - 1) detCount[page] = warpTot[31];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0, q = x * x / 4.0;
    for (int k = 1; k < 200; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}

// kaiser(n, beta)  (Signal Processing Toolbox; fft2D.m:135)
static std::vector<double> kaiser_window(int n, double beta) {
    std::vector<double> w(n, 1.0);
    if (n == 1) return w;
    const double a = (n - 1) / 2.0, den = bessel_i0(beta);
    for (int k = 0; k < n; ++k) {
        double r = (k - a) / a;
        double arg = 1.0 - r * r;
        if (arg < 0) arg = 0;
        w[k] = bessel_i0(beta * std::sqrt(arg)) / den;
    }
    return w;
}

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

int rdm_plan_create(Ctx* ctx, const RdmConfig& c, RdmPlan** out) {
    if (c.nSc < 1 || c.nSym < 1 || c.nAnts < 1 || c.maxBatch < 1) {
        set_error(ctx, "rdm_plan_create: non-positive dimension");
        return kErrInvalidArg;
    }
    if (!is_pow2(c.nIFFT) || c.nIFFT < 256 || c.nIFFT > 4096 || c.nIFFT < c.nSc) {
        set_error(ctx, "rdm_plan_create: nIFFT must be a power of two in [256,4096] and >= nSc");
        return kErrUnsupported;
    }
    if (!is_pow2(c.nFFT) || c.nFFT < 16 || c.nFFT > 4096) {
        set_error(ctx, "rdm_plan_create: nFFT must be a power of two in [16,4096]");
        return kErrUnsupported;
    }
    if (c.guardRows < 0 || c.guardCols < 0 || c.trainRows < 0 || c.trainCols < 0 ||
        (c.trainRows == 0 && c.trainCols == 0)) {
        set_error(ctx, "rdm_plan_create: invalid CFAR band sizes");
        return kErrInvalidArg;
    }
    if (!(c.pfa > 0.0 && c.pfa < 1.0)) {
        set_error(ctx, "rdm_plan_create: Pfa must be in (0,1)");
        return kErrInvalidArg;
    }
    const int hr = c.guardRows + c.trainRows, hc = c.guardCols + c.trainCols;
    if (c.cutRow0 > c.cutRow1 || c.cutCol0 > c.cutCol1) {
        set_error(ctx, "rdm_plan_create: empty CUT rectangle");
        return kErrInvalidArg;
    }
    // phased.CFARDetector2D errors when a CUT's training window leaves the matrix
    if (c.cutRow0 - 1 - hr < 0 || c.cutRow1 - 1 + hr >= c.nIFFT || c.cutCol0 - 1 - hc < 0 ||
        c.cutCol1 - 1 + hc >= c.nFFT) {
        set_error(ctx, "rdm_plan_create: CUT training window exceeds the range-Doppler map");
        return kErrCfarWindow;
    }
    RdmPlan* p = new RdmPlan();
    p->ctx = ctx;
    p->cfg = c;
    p->M = c.nSym < c.nFFT ? c.nSym : c.nFFT;
    p->nCutRows = c.cutRow1 - c.cutRow0 + 1;
    p->nCutCols = c.cutCol1 - c.cutCol0 + 1;
    p->nCut = p->nCutRows * p->nCutCols;
    p->nTrain = (2 * hr + 1) * (2 * hc + 1) - (2 * c.guardRows + 1) * (2 * c.guardCols + 1);
    p->alpha = (double)p->nTrain * (std::pow(c.pfa, -1.0 / (double)p->nTrain) - 1.0);
    p->rowWords = (c.nIFFT + 31) / 32;

    std::vector<double> w1 = kaiser_window(c.nSc, c.kaiserBeta);
    std::vector<double> w2 = kaiser_window(c.nIFFT, c.kaiserBeta);
    std::vector<float> f1(c.nSc), f2(c.nIFFT);
    for (int k = 0; k < c.nSc; ++k) f1[k] = (float)w1[k];
    const double inv = 1.0 / std::sqrt((double)c.nIFFT);
    for (int n = 0; n < c.nIFFT; ++n) {
        int idx = ((n - c.nIFFT / 2) % c.nIFFT + c.nIFFT) % c.nIFFT;
        f2[n] = (float)(w2[idx] * inv);
    }
    const size_t B = (size_t)c.maxBatch, A = (size_t)c.nAnts;
#define ALLOC(ptr, bytes)                                                    \
    do {                                                                     \
        cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                \
        if (e_ != cudaSuccess) {                                             \
            set_error(ctx, std::string("cudaMalloc: ") + cudaGetErrorString(e_)); \
            rdm_plan_destroy(p);                                             \
            return kErrCuda;                                                 \
        }                                                                    \
    } while (0)
    ALLOC(p->d_win1, sizeof(float) * c.nSc);
    ALLOC(p->d_win2, sizeof(float) * c.nIFFT);
    ALLOC(p->d_inter, sizeof(float2) * (size_t)c.nIFFT * p->M * A);  // one map-set, reused (L2 resident)
    ALLOC(p->d_pow, sizeof(float) * (size_t)c.nIFFT * c.nFFT * A * B);
    ALLOC(p->d_flags, (size_t)p->nCut * A * B);
    ALLOC(p->d_rowmask, sizeof(uint32_t) * (size_t)p->rowWords * B);
    ALLOC(p->d_detCount, sizeof(int32_t) * A * B);
    ALLOC(p->d_det, sizeof(int2) * (size_t)p->nCut * A * B);
    ALLOC(p->d_peak, sizeof(float) * (size_t)p->nCut * A * B);
#undef ALLOC
    cudaMemcpy(p->d_win1, f1.data(), sizeof(float) * c.nSc, cudaMemcpyHostToDevice);
    cudaMemcpy(p->d_win2, f2.data(), sizeof(float) * c.nIFFT, cudaMemcpyHostToDevice);
    *out = p;
    return kOk;
}

void rdm_plan_destroy(RdmPlan* p) {
    if (!p) return;
    cudaFree(p->d_win1);
    cudaFree(p->d_win2);
    cudaFree(p->d_inter);
    cudaFree(p->d_pow);
    cudaFree(p->d_flags);
    cudaFree(p->d_rowmask);
    cudaFree(p->d_detCount);
    cudaFree(p->d_det);
    cudaFree(p->d_peak);
    delete p;
}

template <int R1, int R2>
static cudaError_t launch_range(const RdmDev& d, cudaStream_t st) {
    using G = FftGeom<R1, R2, true>;
    const int cpc = G::NT >= 256 ? 1 : 256 / G::NT;
    const int threads = G::NT * cpc;
    const size_t smem = (size_t)cpc * G::kElems * sizeof(float2);
    auto k = rdm_range_ifft_kernel<R1, R2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long blocks = (d.totalCols + cpc - 1) / cpc;
    k<<<(unsigned)blocks, threads, smem, st>>>(d);
    return cudaGetLastError();
}

static cudaError_t launch_range_tma(const RdmDev& d, int numSMs, cudaStream_t st) {
    using G = FftGeom<16, 16, true>;
    const size_t smem = sizeof(float2) * ((size_t)G::kElems + 8 + 2 * (size_t)d.nSc) + 128;
    auto k = rdm_range_ifft_tma_kernel<16, 16>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long blocks = 2LL * numSMs;
    if (blocks > d.totalCols) blocks = d.totalCols;
    k<<<(unsigned)blocks, G::NT, smem, st>>>(d);
    return cudaGetLastError();
}

template <int R1, int R2>
static cudaError_t launch_doppler(const RdmDev& d, long long pages, cudaStream_t st) {
    using G = FftGeom<R1, R2, false>;
    int RT = 8192 / G::N;
    if (RT > 32) RT = 32;
    if (RT < 1) RT = 1;
    const int threads = RT * G::NT;
    const size_t smem = (size_t)RT * G::kElems * sizeof(float2);
    auto k = rdm_doppler_fft_kernel<R1, R2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const long long blocks = pages * (d.nIFFT / RT);
    k<<<(unsigned)blocks, threads, smem, st>>>(d, RT, 1.0f / (float)G::N);
    return cudaGetLastError();
}

static CfarDev make_cfar_dev(RdmPlan* p, const float* pow, int batch) {
    const RdmConfig& c = p->cfg;
    CfarDev d{};
    d.pow = pow;
    d.flags = p->d_flags;
    d.rowmask = p->d_rowmask;
    d.nIFFT = c.nIFFT;
    d.nFFT = c.nFFT;
    d.nAnts = c.nAnts;
    d.row0 = c.cutRow0 - 1;
    d.col0 = c.cutCol0 - 1;
    d.nCutRows = p->nCutRows;
    d.nCut = p->nCut;
    d.gr = c.guardRows;
    d.gc = c.guardCols;
    d.hr = c.guardRows + c.trainRows;
    d.hc = c.guardCols + c.trainCols;
    d.rowWords = p->rowWords;
    d.alpha = p->alpha;
    d.nTrain = (double)p->nTrain;
    d.total = (long long)p->nCut * c.nAnts * batch;
    return d;
}

int rdm_cfar_only(RdmPlan* p, const float* pow, int batch, cudaStream_t st) {
    Ctx* ctx = p->ctx;
    if (batch < 1 || batch > p->cfg.maxBatch) {
        set_error(ctx, "rdm: batch out of range");
        return kErrInvalidArg;
    }
    CfarDev d = make_cfar_dev(p, pow, batch);
    ISAC_CUDA_CHECK(ctx, cudaMemsetAsync(p->d_rowmask, 0, sizeof(uint32_t) * (size_t)p->rowWords * batch, st));
    const long long blocks = (d.total + 255) / 256;
    const int pr = prof_begin(ctx, kProfCfar, st);
    if (d.gr == 2 && d.gc == 2 && d.hr == 3 && d.hc == 3) cfar2d_flags_7x7_kernel<<<(unsigned)blocks, 256, 0, st>>>(d);
    else cfar2d_flags_kernel<<<(unsigned)blocks, 256, 0, st>>>(d);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    cfar2d_compact_kernel<<<(unsigned)(p->cfg.nAnts * batch), 1024, 0, st>>>(d, p->d_det, p->d_peak, p->d_detCount);
    prof_end(ctx, pr, st);
    count_launches(ctx, 2);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    p->lastPow = pow;
    p->lastBatch = batch;
    return kOk;
}

int rdm_run(RdmPlan* p, const float2* rx, const float2* tx, int batch, float* powOut, cudaStream_t st) {
    Ctx* ctx = p->ctx;
    const RdmConfig& c = p->cfg;
    if (batch < 1 || batch > c.maxBatch) {
        set_error(ctx, "rdm: batch out of range");
        return kErrInvalidArg;
    }
    if (!rx || !tx) {
        set_error(ctx, "rdm: null grid pointer");
        return kErrInvalidArg;
    }
    float* pow = powOut ? powOut : p->d_pow;
    // One map-set at a time through the SAME range-profile buffer: the 8*nIFFT*M*nAnts-byte
    // intermediate (44 MB at 273 PRB / 8 antennas) then lives in the 126 MB L2 between the range and
    // Doppler kernels instead of making a round trip through HBM.
    const size_t gridElems = (size_t)c.nSc * c.nSym * c.nAnts;
    const size_t powElems = (size_t)c.nIFFT * c.nFFT * c.nAnts;
    const int prR = prof_begin(ctx, kProfRdmRange, st);  // range+Doppler pairs are timed together
    for (int b = 0; b < batch; ++b) {
        RdmDev d{};
        d.rx = rx + b * gridElems;
        d.tx = tx + b * gridElems;
        d.win1 = p->d_win1;
        d.win2 = p->d_win2;
        ctx_fft_tw(ctx, c.nIFFT, &d.twR.tw1, &d.twR.tw2);
        ctx_fft_tw(ctx, c.nFFT, &d.twD.tw1, &d.twD.tw2);
        d.inter = p->d_inter;
        d.pow = pow + b * powElems;
        d.nSc = c.nSc;
        d.nSym = c.nSym;
        d.nAnts = c.nAnts;
        d.nIFFT = c.nIFFT;
        d.nFFT = c.nFFT;
        d.M = p->M;
        d.totalCols = (long long)p->M * c.nAnts;
        cudaError_t e = cudaSuccess;
        switch (c.nIFFT) {
            case 256: e = launch_range<1, 16>(d, st); break;
            case 512: e = launch_range<2, 16>(d, st); break;
            case 1024: e = launch_range<4, 16>(d, st); break;
            case 2048: e = launch_range<8, 16>(d, st); break;
            case 4096:
                if ((((uintptr_t)d.rx | (uintptr_t)d.tx) & 15) == 0 && (c.nSc % 2) == 0 && !p->noTma) e = launch_range_tma(d, ctx_num_sms(ctx), st);
                else e = launch_range<16, 16>(d, st);
                break;
            default: set_error(ctx, "rdm: unsupported nIFFT"); return kErrUnsupported;
        }
        ISAC_CUDA_CHECK(ctx, e);
        const long long pages = (long long)c.nAnts;
        switch (c.nFFT) {
            case 16: e = launch_doppler<1, 1>(d, pages, st); break;
            case 32: e = launch_doppler<1, 2>(d, pages, st); break;
            case 64: e = launch_doppler<1, 4>(d, pages, st); break;
            case 128: e = launch_doppler<1, 8>(d, pages, st); break;
            case 256: e = launch_doppler<1, 16>(d, pages, st); break;
            case 512: e = launch_doppler<2, 16>(d, pages, st); break;
            case 1024: e = launch_doppler<4, 16>(d, pages, st); break;
            case 2048: e = launch_doppler<8, 16>(d, pages, st); break;
            case 4096: e = launch_doppler<16, 16>(d, pages, st); break;
            default: set_error(ctx, "rdm: unsupported nFFT"); return kErrUnsupported;
        }
        ISAC_CUDA_CHECK(ctx, e);
    }
    prof_end(ctx, prR, st);
    count_launches(ctx, 2 * batch);
    return rdm_cfar_only(p, pow, batch, st);
}

}  // namespace isac
