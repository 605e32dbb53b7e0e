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
#include <cstdlib>
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
    int* ticketR;         // work counters of the persistent kernels (zeroed per run): columns / Doppler tiles are handed
    int* ticketD;         // out in order, so every SM stays busy until the last item
    int hints;            // L2 eviction priorities, 2 bits each (0 normal, 1 evict_first, 2 evict_last): [1:0] rx/tx bulk loads,
                          // [3:2] range-profile stores, [5:4] range-profile tile loads, [7:6] power-map stores; bit 8: discard the
                          // range-profile lines of a tile from L2 once the Doppler kernel has staged them (no write-back)
};

// L2 eviction-priority policy for ld/st/bulk-copy cache hints
__device__ __forceinline__ unsigned long long l2_policy(int kind) {
    unsigned long long pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_hint(float2* p, float2 v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(float* p, float v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

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

__device__ __forceinline__ void tma_load_1d_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                                 unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
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
// Kernel A'' (N = 4096): lean persistent range IFFT.  Same TMA staging as kernel A', but everything that is constant
// across the columns a CTA walks lives in registers, shared memory or folds into immediates: the range window w1
// (NJ registers per thread), the second-pass twiddles (shared memory), every shared-memory / twiddle / output offset
// (one base register each), no per-column predicate (the stage tail [nSc, NJ*256) is zeroed once, w1 = 0 there; the
// register rows j >= NJ are compile-time zeros), no per-column 64-bit division (out = inter + col * N), and one barrier
// less per column (the "stage consumed" barrier of the next column also orders its first-pass stores after this
// column's third-pass loads).  The output is the RAW unnormalised IFFT: the shifted Doppler-axis window w2, the
// 1/sqrt(N) scale and the (-1)^s' Doppler centring all move into the Doppler kernel, where they cost one multiply per
// output (a per-row power scale) and an output-index rotation by F/2.
// NJ = ceil(nSc / 256): 13 for 273 PRB (nSc = 3276), 16 in general.
// ------------------------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(256, 2)
rdm_range4096_lean_kernel(const RdmDev p) {
    constexpr int N = 4096, NT = 256, S1 = 257;  // FftGeom<16,16,true>: addr(k1, mid, lo) = k1*257 + mid*16 + lo
    constexpr int NS = NJ * NT;                  // staged entries per column
    extern __shared__ __align__(128) unsigned char smraw[];
    float2* stageRx = reinterpret_cast<float2*>(smraw);
    float2* stageTx = stageRx + NS;
    float2* fftbuf = stageTx + NS;
    float2* tw2s = fftbuf + 16 * S1;                        // [15][16]
    float* w1s = reinterpret_cast<float*>(tw2s + 240);     // [NS], zero beyond nSc
    __shared__ __align__(8) unsigned long long bar;
    const int tf = threadIdx.x;
    const int nSc = p.nSc;
    const unsigned colBytes = (unsigned)nSc * sizeof(float2);
    for (int n = nSc + tf; n < NS; n += NT) {
        stageRx[n] = make_float2(0.f, 0.f);
        stageTx[n] = make_float2(0.f, 0.f);
    }
    if (tf < 240) tw2s[tf] = __ldg(p.twR.tw2 + tf);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int n = tf + NT * j;
        w1s[n] = (n < nSc) ? __ldg(p.win1 + n) : 0.f;
    }
    float2 tw1r[15];  // first-pass twiddles exp(+2 pi i tf k1 / N): constant across the columns this thread works on
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) tw1r[k1 - 1] = __ldg(p.twR.tw1 + (k1 - 1) * NT + tf);
    const float* const w1 = w1s + tf;                            // + 256*j
    const float2* const tw2 = tw2s + (tf & 15);                  // + (c-1)*16
    float2* const f1 = fftbuf + tf;                                     // pass-1 stores: + k1*257
    float2* const f2 = fftbuf + (tf >> 4) * S1 + (tf & 15);             // pass-2 loads/stores: + a*16
    const float2* const f3 = fftbuf + (tf & 15) * S1 + (tf >> 4) * 16;  // pass-3 loads: + b
    const float2* const sRx = stageRx + tf;
    const float2* const sTx = stageTx + tf;
    const int M = p.M, nSym = p.nSym, half = p.nSym / 2;
    const unsigned long long polIn = l2_policy(p.hints & 3), polOut = l2_policy((p.hints >> 2) & 3);
    auto issue = [&](int col) {  // thread 0: fetch the rx and tx columns of `col`
        const int sp = col % M, page = col / M;
        int s = sp + half;
        if (s >= nSym) s -= nSym;  // ifftshift on the symbol axis (fft2D.m:44)
        const size_t off = ((size_t)page * nSym + s) * (size_t)nSc;
        mbar_expect_tx(&bar, 2 * colBytes);
        tma_load_1d_hint(stageRx, p.rx + off, colBytes, &bar, polIn);
        tma_load_1d_hint(stageTx, p.tx + off, colBytes, &bar, polIn);
    };
    if (tf == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ int sNext;
    const int total = (int)p.totalCols;
    if (tf == 0) {
        const int c0 = atomicAdd(p.ticketR, 1);
        sNext = c0;
        if (c0 < total) issue(c0);
    }
    __syncthreads();
    // Programmatic dependent launch: the Doppler kernel of this map-set may be scheduled as soon as every CTA is here
    // (it waits for this grid's completion before it reads the range profiles).
    asm volatile("griddepcontrol.launch_dependents;");
    int col = sNext;
    unsigned parity = 0;
    bool mustWait = true;  // the previous map-set's Doppler kernel may still be reading the range-profile buffer
    while (col < total) {
        int ticket = 0;
        if (tf == 0) ticket = atomicAdd(p.ticketR, 1);  // next column of this CTA; the round trip hides behind the prologue
        mbar_wait(&bar, parity);
        parity ^= 1;
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            v[j] = (j < NJ) ? pk_scale(pk_cmulc(sRx[NT * j], sTx[NT * j]), w1[NT * j])  // rx.*conj(tx).*rngWin (fft2D.m:37,43)
                            : make_float2(0.f, 0.f);
        if (tf == 0) sNext = ticket;
        __syncthreads();  // stage consumed by every thread (and the previous column's pass-3 loads are done)
        const int nxt = sNext;
        if (tf == 0 && nxt < total) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before async writes
            issue(nxt);
        }
        dft16<+1>(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) v[k1] = pk_cmul(v[k1], tw1r[k1 - 1]);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) f1[k1 * S1] = v[k1];
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 16; ++a) v[a] = f2[a * 16];
        dft16<+1>(v);
#pragma unroll
        for (int c = 1; c < 16; ++c) v[c] = pk_cmul(v[c], tw2[(c - 1) * 16]);
#pragma unroll
        for (int c = 0; c < 16; ++c) f2[c * 16] = v[c];
        __syncthreads();
#pragma unroll
        for (int b = 0; b < 16; ++b) v[b] = f3[b];
        dft16<+1>(v);
        if (mustWait) {  // first column only: everything above overlapped the tail of the preceding grid
            asm volatile("griddepcontrol.wait;" ::: "memory");
            mustWait = false;
        }
        float2* __restrict__ out = p.inter + (size_t)col * N + tf;
#pragma unroll
        for (int d = 0; d < 16; ++d) st_hint(out + NT * d, v[d], polOut);
        col = nxt;
    }
}

// ------------------------------------------------------------------------------------------
// Kernel A-direct (N = 4096): direct-load persistent range IFFT.  As the lean kernel but without the shared-memory stage:
// every thread fetches its NJ rx / tx samples of the column with coalesced streaming loads straight into registers (thread tf
// owns n = tf + 256 j, so a warp reads 256 contiguous bytes per load).  That removes the stage's write + read (42 % of the lean
// kernel's shared-memory traffic) and shrinks the CTA to 48 KB of shared memory; the load latency is hidden by the other CTAs
// of the SM instead of by a prefetch.  MINB = CTAs per SM the register budget is sized for.
// ------------------------------------------------------------------------------------------
template <int NJ, int MINB>
__global__ void __launch_bounds__(256, MINB)
rdm_range4096_direct_kernel(const RdmDev p) {
    constexpr int N = 4096, NT = 256, S1 = 257;
    extern __shared__ __align__(128) unsigned char smraw[];
    float2* fftbuf = reinterpret_cast<float2*>(smraw);
    float2* tw2s = fftbuf + 16 * S1;                       // [15][16]
    float* w1s = reinterpret_cast<float*>(tw2s + 240);    // [NJ*256], zero beyond nSc
    __shared__ int sNext, sNext2;
    const int tf = threadIdx.x;
    const int nSc = p.nSc;
    if (tf < 240) tw2s[tf] = __ldg(p.twR.tw2 + tf);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int n = tf + NT * j;
        w1s[n] = (n < nSc) ? __ldg(p.win1 + n) : 0.f;
    }
    float2 tw1r[15];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) tw1r[k1 - 1] = __ldg(p.twR.tw1 + (k1 - 1) * NT + tf);
    const float* const w1 = w1s + tf;
    const float2* const tw2 = tw2s + (tf & 15);
    float2* const f1 = fftbuf + tf;
    float2* const f2 = fftbuf + (tf >> 4) * S1 + (tf & 15);
    const float2* const f3 = fftbuf + (tf & 15) * S1 + (tf >> 4) * 16;
    const int M = p.M, nSym = p.nSym, half = p.nSym / 2;
    const unsigned long long polOut = l2_policy((p.hints >> 2) & 3);
    const int total = (int)p.totalCols;
    const bool lastRow = tf + NT * (NJ - 1) < nSc;   // the last register row is only partly inside the grid
    const bool pf = (p.hints >> 9) & 1;   // bit 9: bulk-prefetch the rx / tx columns of the ticket after next into L2
    auto prefetch = [&](int c) {          // thread 0
        const int sp = c % M, page = c / M;
        int sy = sp + half;
        if (sy >= nSym) sy -= nSym;
        const size_t off = ((size_t)page * nSym + sy) * (size_t)nSc;
        const unsigned bytes = (unsigned)nSc * (unsigned)sizeof(float2);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.rx + off), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.tx + off), "r"(bytes) : "memory");
    };
    if (tf == 0) {
        sNext = atomicAdd(p.ticketR, 1);
        sNext2 = atomicAdd(p.ticketR, 1);
        if (pf && sNext2 < total) prefetch(sNext2);
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;");
    int col = sNext;
    bool mustWait = true;
    while (col < total) {
        const int sp = col % M, page = col / M;
        int sy = sp + half;
        if (sy >= nSym) sy -= nSym;  // ifftshift on the symbol axis (fft2D.m:44)
        const size_t off = ((size_t)page * nSym + sy) * (size_t)nSc + tf;
        const float2* __restrict__ rx = p.rx + off;
        const float2* __restrict__ tx = p.tx + off;
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < NJ) {
                const bool in = j < NJ - 1 || lastRow;
                const float2 a = in ? ld_stream(rx + NT * j) : make_float2(0.f, 0.f);
                const float2 b = in ? ld_stream(tx + NT * j) : make_float2(0.f, 0.f);
                v[j] = pk_scale(pk_cmulc(a, b), w1[NT * j]);  // rx.*conj(tx).*rngWin (fft2D.m:37,43)
            } else {
                v[j] = make_float2(0.f, 0.f);
            }
        }
        __syncthreads();   // the previous column's pass-3 loads are done (fftbuf reuse) and its sNext has been read
        if (tf == 0) {     // this CTA's columns: col (running), sNext2 (next, already prefetched), a fresh ticket (prefetched now)
            sNext = sNext2;
            const int t2 = atomicAdd(p.ticketR, 1);
            sNext2 = t2;
            if (pf && t2 < total) prefetch(t2);
        }
        dft16<+1>(v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) v[k1] = pk_cmul(v[k1], tw1r[k1 - 1]);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) f1[k1 * S1] = v[k1];
        __syncthreads();
#pragma unroll
        for (int a2 = 0; a2 < 16; ++a2) v[a2] = f2[a2 * 16];
        dft16<+1>(v);
#pragma unroll
        for (int c = 1; c < 16; ++c) v[c] = pk_cmul(v[c], tw2[(c - 1) * 16]);
#pragma unroll
        for (int c = 0; c < 16; ++c) f2[c * 16] = v[c];
        __syncthreads();
        const int nxt = sNext;   // written behind the first barrier of this column
#pragma unroll
        for (int b2 = 0; b2 < 16; ++b2) v[b2] = f3[b2];
        dft16<+1>(v);
        if (mustWait) {  // first column only: everything above overlapped the tail of the preceding grid
            asm volatile("griddepcontrol.wait;" ::: "memory");
            mustWait = false;
        }
        float2* __restrict__ out = p.inter + (size_t)col * N + tf;
#pragma unroll
        for (int d = 0; d < 16; ++d) st_hint(out + NT * d, v[d], polOut);
        col = nxt;
    }
}

// ------------------------------------------------------------------------------------------
// Kernel B' (F = 256): persistent Doppler FFT with bulk-copy staging.  One 256-thread CTA walks tiles of 16
// consecutive range rows of one antenna page; the [M symbols x 16 rows] input tile of the NEXT tile is fetched from
// the (L2-resident) range profiles by ONE 2-D tensor copy (cp.async.bulk.tensor, one mbarrier per buffer) while
// the current tile's two radix-16 passes run -> no thread ever waits on a global load.  The exchange between the two
// passes happens in place in the staged tile (each thread overwrites exactly the 16 entries it read), so two 32 KB
// buffers per CTA suffice and 3 CTAs fit per SM.  Twiddles sit in shared memory.  Epilogue: |X|^2 times the per-row
// scale w2[(n - N/2) mod N]^2 / (N F) (fft2D.m:44-46, :61), written at the Doppler index rotated by F/2 (fftshift).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar,
                                            unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
        "[%4], %5;" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

template <int NI>  // range size when known at compile time (every output offset becomes an immediate), else 0
__global__ void __launch_bounds__(256, 3)
rdm_doppler256_tma_kernel(const RdmDev p, const int nPages, const __grid_constant__ CUtensorMap interMap) {
    constexpr int F = 256, RT = 16, TILE = F * RT;
    extern __shared__ __align__(128) unsigned char smraw[];
    float2* buf0 = reinterpret_cast<float2*>(smraw);  // 2 x [256 symbols][16 rows]
    float2* tws = buf0 + 2 * TILE;                    // [15][16]
    __shared__ __align__(8) unsigned long long bar[2];
    const int tid = threadIdx.x, nl = tid & 15, tf = tid >> 4;
    const int M = p.M, nIFFT = NI > 0 ? NI : p.nIFFT;
    const int tilesPerPage = nIFFT / RT;
    const int total = nPages * tilesPerPage;
    if (tid < 240) {
        float2 w = __ldg(p.twD.tw2 + tid);
        tws[tid] = make_float2(w.x, -w.y);  // forward transform: conjugate table
    }
    const float2* const tw = tws + tf;  // + (c-1)*16
    const int nA = (M - tf + 15) >> 4;  // valid first-pass inputs: a*16 + tf < M
    const unsigned long long polIn = l2_policy((p.hints >> 4) & 3), polOut = l2_policy((p.hints >> 6) & 3);
    const bool discard = (p.hints >> 8) & 1;
    auto issue = [&](int tile, int s) {  // thread 0: one 2-D tensor copy [M symbols x 16 rows] -> buffer s
        const int page = tile / tilesPerPage, row0 = (tile - page * tilesPerPage) * RT;
        mbar_expect_tx(&bar[s], (unsigned)M * RT * (unsigned)sizeof(float2));
        tma_load_2d(buf0 + s * TILE, &interMap, row0, page * M, &bar[s], polIn);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Programmatic dependent launch: the next map-set's range kernel may start its prologue / first column now (it waits
    // for this grid before its first store); this grid waits for the range kernel that produced the profiles.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ int sNext;
    if (tid == 0) {
        const int t0 = atomicAdd(p.ticketD, 1);
        sNext = t0;
        if (t0 < total) issue(t0, 0);
    }
    __syncthreads();
    int tile = sNext;
    unsigned phase = 0;  // bit s = parity to wait for on bar[s]
    for (int i = 0; tile < total; ++i) {
        const int s = i & 1;
        int ticket = 0;
        if (tid == 0) ticket = atomicAdd(p.ticketD, 1);
        float2* const buf = buf0 + s * TILE + tf * RT + nl;        // pass A: + a*256 (symbol a*16 + tf)
        const float2* const bufB = buf0 + s * TILE + tf * TILE / 16 + nl;  // pass B: + b*16   (entry tf*16 + b)
        const int page = tile / tilesPerPage, n = (tile - page * tilesPerPage) * RT + nl;
        const float sc = __ldg(p.win2 + n);  // per-row power scale (see rdm_plan_create)
        mbar_wait(&bar[s], (phase >> s) & 1u);
        phase ^= 1u << s;
        if (discard && tid < M) {  // the staged tile is the only consumer of these 128-byte lines: drop them without write-back
            const float2* line = p.inter + ((size_t)page * M + tid) * nIFFT + (n - nl);
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(line) : "memory");
        }
        float2 v[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) v[a] = (a < nA) ? buf[a * TILE / 16] : make_float2(0.f, 0.f);
        dft16<-1>(v);
#pragma unroll
        for (int c = 1; c < 16; ++c) v[c] = pk_cmul(v[c], tw[(c - 1) * 16]);
#pragma unroll
        for (int c = 0; c < 16; ++c) buf[c * TILE / 16] = v[c];
        if (tid == 0) sNext = ticket;
        __syncthreads();  // exchange complete; every thread is also past its pass-B loads of the previous tile
        const int nxt = sNext;
        if (tid == 0 && nxt < total) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic accesses before async writes
            issue(nxt, s ^ 1);
        }
#pragma unroll
        for (int b = 0; b < 16; ++b) v[b] = bufB[b * RT];
        dft16<-1>(v);
        float* __restrict__ out = p.pow + (size_t)page * F * nIFFT + n;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const int q = tf + 16 * ((d + 8) & 15);  // Doppler-axis fftshift (fft2D.m:46) as an output rotation
            st_hint(out + (size_t)q * nIFFT, (v[d].x * v[d].x + v[d].y * v[d].y) * sc, polOut);  // abs(rdm).^2 (fft2D.m:61)
        }
        tile = nxt;
    }
}

// ------------------------------------------------------------------------------------------
// Kernel B: Doppler FFT + |.|^2
// RT (rows interleaved per CTA) and, for the hot shapes, the range size NI are compile-time so that every
// shared-memory and global offset is an immediate.  RAW: the input is the raw IFFT of the lean range kernel -> the
// Doppler-axis fftshift (fft2D.m:46) is an output-index rotation by F/2 and the power is scaled per range row by
// p.win2[n] = w2[(n - N/2) mod N]^2 / (N F); !RAW: kernel A / A' already windowed, scaled and modulated by (-1)^s'.
// ------------------------------------------------------------------------------------------
template <int R1, int R2, int RT, int NI, bool RAW>
__global__ void __launch_bounds__(RT * R1 * R2)
rdm_doppler_fft_kernel(const RdmDev p, const float invF) {
    using G = FftGeom<R1, R2, false>;
    extern __shared__ float2 smem[];
    const int nl = threadIdx.x % RT, tf = threadIdx.x / RT;
    const int nIFFT = NI > 0 ? NI : p.nIFFT;
    const int tilesPerPage = nIFFT / RT;
    const int page = blockIdx.x / tilesPerPage;
    const int n = (blockIdx.x % tilesPerPage) * RT + nl;
    const int M = p.M;
    const float2* __restrict__ in = p.inter + (size_t)page * M * nIFFT + n;
    auto load = [&](int sp) -> float2 {
        if (sp < M) return __ldcg(in + (size_t)sp * nIFFT);
        return make_float2(0.f, 0.f);
    };
    float2 v[16];
    block_fft<R1, R2, -1, false>(v, smem + nl, RT, tf, p.twD, load);
    float* __restrict__ out = p.pow + (size_t)page * G::N * nIFFT + n;
    const float sc = RAW ? __ldg(p.win2 + n) : invF;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        const int q = tf + G::NT * (RAW ? ((d + 8) & 15) : d);
        out[(size_t)q * nIFFT] = (v[d].x * v[d].x + v[d].y * v[d].y) * sc;  // abs(rdm).^2 (fft2D.m:61)
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
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // no-op unless launched as a programmatic dependent
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
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // no-op unless launched as a programmatic dependent
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
    asm volatile("griddepcontrol.wait;" ::: "memory");  // flags of the preceding grid (programmatic dependent launch)
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
// Kernel C+D fused: one CTA per (antenna, map) page.  The CUT rectangle plus its training halo ((nCutRows + 2 hr) x
// (nCutCols + 2 hc) cells: 376 x 29 at 273 PRB) is staged once in shared memory with coalesced loads along the range axis,
// every thread decides a contiguous run of CUTs from the staged tile (the training cells are added in float64 in exactly
// the order of the stand-alone decision kernels, so the decisions are bit-identical), the run counts are scanned across the
// CTA with warp shuffles and the detections are written in CUT order.  One launch instead of two, no flag array, and the
// 24 training loads of a CUT come from shared memory instead of L2.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) cfar2d_fused_kernel(const CfarDev p, int2* det, float* peak, int32_t* detCount) {
    extern __shared__ float tile[];   // [tileCols][tileRows], range fastest
    __shared__ int warpTot[32];
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the power map of the preceding grid (programmatic dependent launch)
    const long long page = blockIdx.x;
    const int tr = p.nCutRows + 2 * p.hr, nCutCols = p.nCut / p.nCutRows, tc = nCutCols + 2 * p.hc;
    const float* __restrict__ P = p.pow + page * (long long)p.nFFT * p.nIFFT + (long long)(p.col0 - p.hc) * p.nIFFT + (p.row0 - p.hr);
    for (int i = threadIdx.x; i < tr * tc; i += blockDim.x) {
        const int r = i % tr, c = i / tr;
        tile[i] = __ldcg(P + (long long)c * p.nIFFT + r);
    }
    __syncthreads();
    const int per = (p.nCut + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(lo + per, p.nCut);
    unsigned long long mask = 0ull;   // per <= 64 (checked by the launcher)
    int cnt = 0;
    for (int i = lo; i < hi; ++i) {
        const int r = i % p.nCutRows + p.hr, c = i / p.nCutRows + p.hc;   // tile coordinates of the CUT
        double acc = 0.0;
        for (int dc = -p.hc; dc <= p.hc; ++dc) {
            const float* __restrict__ col = tile + (c + dc) * tr + r;
            const bool colInGuard = (dc >= -p.gc && dc <= p.gc);
            for (int dr = -p.hr; dr <= p.hr; ++dr) {
                if (colInGuard && dr >= -p.gr && dr <= p.gr) continue;
                acc = acc + (double)col[dr];
            }
        }
        const bool d = (double)tile[c * tr + r] > p.alpha * (acc / p.nTrain);   // strict (CFARDetector2D)
        if (d) {
            mask |= 1ull << (i - lo);
            ++cnt;
        }
    }
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
    int2* __restrict__ o = det + page * (long long)p.nCut;
    float* __restrict__ pk = peak + page * (long long)p.nCut;
    for (int i = lo; i < hi; ++i)
        if ((mask >> (i - lo)) & 1ull) {
            const int rr = i % p.nCutRows, cc = i / p.nCutRows;
            const int row = p.row0 + rr, col = p.col0 + cc;
            o[off] = make_int2(row + 1, col + 1);
            pk[off] = tile[(cc + p.hc) * tr + rr + p.hr];
            atomicOr(p.rowmask + (page / p.nAnts) * p.rowWords + (row >> 5), 1u << (row & 31));
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

static bool make_inter_tensor_map(CUtensorMap* map, void* inter, int nIFFT, int M, int nAnts);

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
    ALLOC(p->d_rowScale, sizeof(float) * c.nIFFT);
    ALLOC(p->d_tickets, sizeof(int) * 2 * B);
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
    for (int n = 0; n < c.nIFFT; ++n) {  // power scale of range row n when the range kernel emits the raw IFFT
        int idx = ((n - c.nIFFT / 2) % c.nIFFT + c.nIFFT) % c.nIFFT;
        f2[n] = (float)(w2[idx] * w2[idx] / ((double)c.nIFFT * (double)c.nFFT));
    }
    cudaMemcpy(p->d_rowScale, f2.data(), sizeof(float) * c.nIFFT, cudaMemcpyHostToDevice);
    if (const char* e = getenv("ISAC_RDM_PDL")) p->pdl = atoi(e) != 0;
    if (const char* e = getenv("ISAC_RDM_HINTS")) p->hints = (int)strtol(e, nullptr, 0);
    p->hasInterMap = c.nFFT == 256 && c.nIFFT % 16 == 0 && p->M <= 256 &&
                     make_inter_tensor_map(&p->interMap, p->d_inter, c.nIFFT, p->M, c.nAnts);
    *out = p;
    return kOk;
}

void rdm_plan_destroy(RdmPlan* p) {
    if (!p) return;
    cudaFree(p->d_win1);
    cudaFree(p->d_win2);
    cudaFree(p->d_rowScale);
    cudaFree(p->d_tickets);
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

// Launch with (pdl) or without the programmatic-stream-serialization attribute: with it the grid may start while its
// predecessor in the stream is still running and orders itself with griddepcontrol.wait.
template <class... KArgs, class... Args>
static cudaError_t launch_ex(void (*k)(KArgs...), unsigned blocks, unsigned threads, size_t smem, cudaStream_t st, bool pdl,
                             Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, args...);
}

template <int NJ>
static cudaError_t launch_range_lean_nj(const RdmDev& d, int numSMs, bool pdl, cudaStream_t st) {
    const size_t smem = sizeof(float2) * (2 * NJ * 256 + 16 * 257 + 240) + sizeof(float) * NJ * 256;
    auto k = rdm_range4096_lean_kernel<NJ>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    long long blocks = 2LL * numSMs;
    if (blocks > d.totalCols) blocks = d.totalCols;
    return launch_ex(k, (unsigned)blocks, 256, smem, st, pdl, d);
}
template <int NJ, int MINB>
static cudaError_t launch_range_direct_nj(const RdmDev& d, int numSMs, bool pdl, cudaStream_t st) {
    const size_t smem = sizeof(float2) * (16 * 257 + 240) + sizeof(float) * NJ * 256;
    auto k = rdm_range4096_direct_kernel<NJ, MINB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    long long blocks = (long long)MINB * numSMs;
    if (blocks > d.totalCols) blocks = d.totalCols;
    return launch_ex(k, (unsigned)blocks, 256, smem, st, pdl, d);
}
static cudaError_t launch_range_direct(const RdmDev& d, int numSMs, int minb, bool pdl, cudaStream_t st) {
    if (d.nSc <= 13 * 256 && d.nSc > 12 * 256)
        return minb >= 4 ? launch_range_direct_nj<13, 4>(d, numSMs, pdl, st)
                         : minb == 3 ? launch_range_direct_nj<13, 3>(d, numSMs, pdl, st) : launch_range_direct_nj<13, 2>(d, numSMs, pdl, st);
    return minb >= 3 ? launch_range_direct_nj<16, 3>(d, numSMs, pdl, st) : launch_range_direct_nj<16, 2>(d, numSMs, pdl, st);
}
static cudaError_t launch_range_lean(const RdmDev& d, int numSMs, bool pdl, cudaStream_t st) {
    return d.nSc <= 13 * 256 ? launch_range_lean_nj<13>(d, numSMs, pdl, st) : launch_range_lean_nj<16>(d, numSMs, pdl, st);
}

// F = 256, raw range profiles, nIFFT a multiple of 16
static cudaError_t launch_doppler256_tma(const RdmDev& d, int pages, int numSMs, const CUtensorMap& map, bool pdl,
                                         cudaStream_t st) {
    const size_t smem = sizeof(float2) * (2 * 4096 + 240);
    long long blocks = 3LL * numSMs;
    const long long tiles = (long long)pages * (d.nIFFT / 16);
    if (blocks > tiles) blocks = tiles;
    auto k = d.nIFFT == 4096 ? rdm_doppler256_tma_kernel<4096> : rdm_doppler256_tma_kernel<0>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return launch_ex(k, (unsigned)blocks, 256, smem, st, pdl, d, pages, map);
}

// 2-D view of the range-profile buffer for the bulk-staged Doppler kernel: dim0 = range row (nIFFT, contiguous),
// dim1 = symbol of every page (M * nAnts), 8-byte elements (one complex64), box = [16 rows x M symbols], no swizzle.
static bool make_inter_tensor_map(CUtensorMap* map, void* inter, int nIFFT, int M, int nAnts) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
        return false;
    const cuuint64_t dims[2] = {(cuuint64_t)nIFFT, (cuuint64_t)M * nAnts};
    const cuuint64_t strides[1] = {(cuuint64_t)nIFFT * sizeof(float2)};
    const cuuint32_t box[2] = {16u, (cuuint32_t)M};
    const cuuint32_t estr[2] = {1u, 1u};
    return ((EncodeTiled)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, inter, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int R1, int R2, int RT, int NI>
static cudaError_t launch_doppler_rt(const RdmDev& d, long long pages, bool raw, cudaStream_t st) {
    using G = FftGeom<R1, R2, false>;
    const int threads = RT * G::NT;
    const size_t smem = (size_t)RT * G::kElems * sizeof(float2);
    const long long blocks = pages * (d.nIFFT / RT);
    const float invF = 1.0f / (float)G::N;
    if (raw) {
        auto k = rdm_doppler_fft_kernel<R1, R2, RT, NI, true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<(unsigned)blocks, threads, smem, st>>>(d, invF);
    } else {
        auto k = rdm_doppler_fft_kernel<R1, R2, RT, NI, false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<(unsigned)blocks, threads, smem, st>>>(d, invF);
    }
    return cudaGetLastError();
}

// RT = rows interleaved per CTA: 8192 / N capped to [1, 32] (64 KB of shared memory, <= 512 threads)
template <int R1, int R2>
static cudaError_t launch_doppler(const RdmDev& d, long long pages, bool raw, cudaStream_t st) {
    constexpr int N = R1 * R2 * 16;
    constexpr int RT = (8192 / N > 32) ? 32 : (8192 / N < 1 ? 1 : 8192 / N);
    if (N == 256 && d.nIFFT == 4096) return launch_doppler_rt<R1, R2, RT, (N == 256 ? 4096 : 0)>(d, pages, raw, st);
    if (N == 1024 && d.nIFFT == 1024) return launch_doppler_rt<R1, R2, RT, (N == 1024 ? 1024 : 0)>(d, pages, raw, st);
    return launch_doppler_rt<R1, R2, RT, 0>(d, pages, raw, st);
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

// chained: called at the tail of rdm_run -- the row bitmap was cleared before the range kernels, no profiling events are
// recorded in between, and the flags kernel is launched as a programmatic dependent of the last Doppler kernel.
static int rdm_cfar_launch(RdmPlan* p, const float* pow, int batch, bool chained, cudaStream_t st) {
    Ctx* ctx = p->ctx;
    CfarDev d = make_cfar_dev(p, pow, batch);
    int pr = -1;
    if (!chained) {
        ISAC_CUDA_CHECK(ctx, cudaMemsetAsync(p->d_rowmask, 0, sizeof(uint32_t) * (size_t)p->rowWords * batch, st));
        pr = prof_begin(ctx, kProfCfar, st);
    }
    // fused decision + compaction kernel (opt-in, ISAC_CFAR_FUSED=1): measured SLOWER than the two-kernel path inside the chain
    // (+27 us per chain at 1 and 4 map-sets, profiles/r2_rdm_experiments.txt), so the wide one-thread-per-CUT decision kernel
    // followed by the per-page compaction stays the default
    static const int fusedOff = getenv("ISAC_CFAR_FUSED") ? atoi(getenv("ISAC_CFAR_FUSED")) == 0 : 1;
    const size_t tileBytes = sizeof(float) * (size_t)(p->nCutRows + 2 * d.hr) * (size_t)(p->nCutCols + 2 * d.hc);
    if (!fusedOff && tileBytes <= 200 * 1024 && (p->nCut + 1023) / 1024 <= 64) {
        cudaFuncSetAttribute(cfar2d_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tileBytes);
        // same shared-memory carve-out as the range / Doppler kernels of the chain: a different one makes the SMs drain first
        cudaFuncSetAttribute(cfar2d_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        ISAC_CUDA_CHECK(ctx, launch_ex(cfar2d_fused_kernel, (unsigned)(p->cfg.nAnts * batch), 1024, tileBytes, st, chained && p->pdl, d,
                                       p->d_det, p->d_peak, p->d_detCount));
        if (!chained) prof_end(ctx, pr, st);
        count_launches(ctx, 1);
        p->lastPow = pow;
        p->lastBatch = batch;
        return kOk;
    }
    const long long blocks = (d.total + 255) / 256;
    const bool ref7 = d.gr == 2 && d.gc == 2 && d.hr == 3 && d.hc == 3;
    ISAC_CUDA_CHECK(ctx, launch_ex(ref7 ? cfar2d_flags_7x7_kernel : cfar2d_flags_kernel, (unsigned)blocks, 256, 0, st,
                                   chained && p->pdl, d));
    ISAC_CUDA_CHECK(ctx, launch_ex(cfar2d_compact_kernel, (unsigned)(p->cfg.nAnts * batch), 1024, 0, st, p->pdl, d, p->d_det,
                                   p->d_peak, p->d_detCount));
    if (!chained) prof_end(ctx, pr, st);
    count_launches(ctx, 2);
    p->lastPow = pow;
    p->lastBatch = batch;
    return kOk;
}

int rdm_cfar_only(RdmPlan* p, const float* pow, int batch, cudaStream_t st) {
    if (batch < 1 || batch > p->cfg.maxBatch) {
        set_error(p->ctx, "rdm: batch out of range");
        return kErrInvalidArg;
    }
    return rdm_cfar_launch(p, pow, batch, false, st);
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
    ISAC_CUDA_CHECK(ctx, cudaMemsetAsync(p->d_rowmask, 0, sizeof(uint32_t) * (size_t)p->rowWords * batch, st));
    ISAC_CUDA_CHECK(ctx, cudaMemsetAsync(p->d_tickets, 0, sizeof(int) * 2 * (size_t)batch, st));
    const int prR = prof_begin(ctx, kProfRdmRange, st);  // the whole range / Doppler / CFAR chain is timed as one group
    bool prevTmaDoppler = false;
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
        d.hints = p->hints;
        d.ticketR = p->d_tickets + 2 * b;
        d.ticketD = p->d_tickets + 2 * b + 1;
        cudaError_t e = cudaSuccess;
        bool raw = false;
        switch (c.nIFFT) {
            case 256: e = launch_range<1, 16>(d, st); break;
            case 512: e = launch_range<2, 16>(d, st); break;
            case 1024: e = launch_range<4, 16>(d, st); break;
            case 2048: e = launch_range<8, 16>(d, st); break;
            case 4096:
                if ((((uintptr_t)d.rx | (uintptr_t)d.tx) & 15) == 0 && (c.nSc % 2) == 0 && p->variant != 2) {
                    if (p->variant == 0 || p->variant == 3) {
                        d.win2 = p->d_rowScale;  // raw range profiles: window / scale / centring move to the Doppler kernel
                        // PDL edge Doppler(b-1) -> range(b): only behind this plan's own bulk-staged Doppler kernel
                        static const int directMinb = getenv("ISAC_RDM_DIRECT") ? atoi(getenv("ISAC_RDM_DIRECT")) : 0;
                        if (directMinb >= 2 && c.nSc > 15 * 256)   // experimental direct-load range kernel (NJ = 16 row guard)
                            e = launch_range_direct(d, ctx_num_sms(ctx), directMinb, p->pdl && b > 0 && prevTmaDoppler, st);
                        else if (directMinb >= 2 && c.nSc > 12 * 256 && c.nSc <= 13 * 256)
                            e = launch_range_direct(d, ctx_num_sms(ctx), directMinb, p->pdl && b > 0 && prevTmaDoppler, st);
                        else
                            e = launch_range_lean(d, ctx_num_sms(ctx), p->pdl && b > 0 && prevTmaDoppler, st);
                        raw = true;
                    } else {
                        e = launch_range_tma(d, ctx_num_sms(ctx), st);
                    }
                } else {
                    e = launch_range<16, 16>(d, st);
                }
                break;
            default: set_error(ctx, "rdm: unsupported nIFFT"); return kErrUnsupported;
        }
        ISAC_CUDA_CHECK(ctx, e);
        const long long pages = (long long)c.nAnts;
        if (raw && c.nFFT == 256 && p->variant == 0 && p->hasInterMap) {
            ISAC_CUDA_CHECK(ctx, launch_doppler256_tma(d, (int)pages, ctx_num_sms(ctx), p->interMap, p->pdl, st));
            prevTmaDoppler = true;
            continue;
        }
        switch (c.nFFT) {
            case 16: e = launch_doppler<1, 1>(d, pages, raw, st); break;
            case 32: e = launch_doppler<1, 2>(d, pages, raw, st); break;
            case 64: e = launch_doppler<1, 4>(d, pages, raw, st); break;
            case 128: e = launch_doppler<1, 8>(d, pages, raw, st); break;
            case 256: e = launch_doppler<1, 16>(d, pages, raw, st); break;
            case 512: e = launch_doppler<2, 16>(d, pages, raw, st); break;
            case 1024: e = launch_doppler<4, 16>(d, pages, raw, st); break;
            case 2048: e = launch_doppler<8, 16>(d, pages, raw, st); break;
            case 4096: e = launch_doppler<16, 16>(d, pages, raw, st); break;
            default: set_error(ctx, "rdm: unsupported nFFT"); return kErrUnsupported;
        }
        ISAC_CUDA_CHECK(ctx, e);
    }
    count_launches(ctx, 2 * batch);
    const int rc = rdm_cfar_launch(p, pow, batch, true, st);
    prof_end(ctx, prR, st);
    return rc;
}

}  // namespace isac
