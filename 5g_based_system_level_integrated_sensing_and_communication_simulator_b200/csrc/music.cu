// K5 + K6: antenna covariance, float64 one-sided Jacobi eigen-solver, MUSIC scans, findpeaks.
//
// Replaces the arithmetic of sensing.estimation.doaEstimation.music
// (+sensing/+estimation/+doaEstimation/music.m:11-125), the covariance at fft2D.m:106-107 and the
// range/velocity MUSIC of sensing.estimation.music2D (+sensing/+estimation/music2D.m:57-123).
//
// Everything after the covariance runs in float64: the pseudo-spectrum 1/(a' Un Un' a + eps) is
// evaluated near nulls of the denominator, where float32 cannot hold the 1e-5 bar.
// Noise-subspace quadratic forms are evaluated either directly (sum over the noise eigenvectors,
// no cancellation; small arrays) or in complement form ||a||^2 - sum_{k<L} |u_k' a|^2 (large
// arrays, where only the L leading eigenvectors are meaningful because the matrix is rank
// deficient; Un Un' = I - Us Us' exactly for an orthonormal eigenbasis).
#include "music.cuh"
#include "ctx.cuh"
#include <cmath>
#include <vector>

namespace isac {

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MATLAB sind/cosd: exact at multiples of 90 degrees (argument reduction in degrees)
__device__ __forceinline__ double sind_dev(double x) {
    double r = fmod(x, 360.0);
    if (r > 180.0) r -= 360.0;
    if (r < -180.0) r += 360.0;
    if (r > 90.0) r = 180.0 - r;
    if (r < -90.0) r = -180.0 - r;
    const double k = 0.017453292519943295769;  // pi/180
    if (fabs(r) <= 45.0) return sin(r * k);
    const double c = cos((90.0 - fabs(r)) * k);
    return r > 0 ? c : -c;
}
__device__ __forceinline__ double cosd_dev(double x) { return sind_dev(x + 90.0); }

__device__ __forceinline__ double2 cis2pi(double t) {  // exp(2*pi*j*t)
    double s, c;
    sincospi(2.0 * t, &s, &c);
    return make_double2(c, s);
}

// ------------------------------------------------------------------------------------------
// K5: antenna covariance, deterministic two-stage reduction
// ------------------------------------------------------------------------------------------
constexpr int kCovBI = 4, kCovBJ = 4, kCovThreads = 256;  // 4x4 blocks of the upper triangle: 16 float64 accumulators per thread

__global__ void __launch_bounds__(kCovThreads, 2)
cov_partial_kernel(const float2* __restrict__ rx, long long N, int nAnts, int jBlocks, int chunks, double2* __restrict__ part) {
    const int chunk = blockIdx.x, pair = blockIdx.y, b = blockIdx.z;
    const int ib = pair / jBlocks, jb = pair % jBlocks;
    const int i0 = ib * kCovBI, j0 = jb * kCovBJ;
    double2 acc[kCovBI][kCovBJ];
#pragma unroll
    for (int i = 0; i < kCovBI; ++i)
#pragma unroll
        for (int j = 0; j < kCovBJ; ++j) acc[i][j] = make_double2(0.0, 0.0);
    const bool needed = (j0 + kCovBJ - 1 >= i0);  // block touches the upper triangle
    if (needed) {
        const float2* __restrict__ base = rx + (long long)b * nAnts * N;
        const long long per = (N + chunks - 1) / chunks;
        const long long t0 = (long long)chunk * per;
        const long long t1 = (t0 + per < N) ? t0 + per : N;
#pragma unroll 2
        for (long long t = t0 + threadIdx.x; t < t1; t += kCovThreads) {
            double2 xi[kCovBI], xj[kCovBJ];
#pragma unroll
            for (int i = 0; i < kCovBI; ++i) {
                float2 v = (i0 + i < nAnts) ? __ldg(base + (long long)(i0 + i) * N + t) : make_float2(0.f, 0.f);
                xi[i] = make_double2((double)v.x, (double)v.y);
            }
#pragma unroll
            for (int j = 0; j < kCovBJ; ++j) {
                float2 v = (j0 + j < nAnts) ? __ldg(base + (long long)(j0 + j) * N + t) : make_float2(0.f, 0.f);
                xj[j] = make_double2((double)v.x, (double)v.y);
            }
#pragma unroll
            for (int i = 0; i < kCovBI; ++i)
#pragma unroll
                for (int j = 0; j < kCovBJ; ++j) {
                    // acc += conj(xi) * xj as 4 DFMA (the sum-of-products form costs DMUL + DFMA + DADD per part)
                    acc[i][j].x = fma(xi[i].x, xj[j].x, fma(xi[i].y, xj[j].y, acc[i][j].x));
                    acc[i][j].y = fma(xi[i].x, xj[j].y, fma(-xi[i].y, xj[j].x, acc[i][j].y));
                }
        }
    }
    __shared__ double2 red[kCovThreads / 32][kCovBI * kCovBJ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kCovBI; ++i)
#pragma unroll
        for (int j = 0; j < kCovBJ; ++j) {
            double re = warp_sum(acc[i][j].x), im = warp_sum(acc[i][j].y);
            if (lane == 0) red[warp][i * kCovBJ + j] = make_double2(re, im);
        }
    __syncthreads();
    if (threadIdx.x < kCovBI * kCovBJ) {
        double2 s = make_double2(0.0, 0.0);
        for (int w = 0; w < kCovThreads / 32; ++w) s = zadd(s, red[w][threadIdx.x]);
        part[(((long long)b * gridDim.y + pair) * chunks + chunk) * (kCovBI * kCovBJ) + threadIdx.x] = s;
    }
}

// one warp per matrix entry: the lanes stride over the chunk partials, then a fixed-order shuffle reduction (deterministic)
__global__ void __launch_bounds__(256) cov_final_kernel(const double2* __restrict__ part, int nAnts, int jBlocks, int nPairs,
                                                        int chunks, double invN, double2* __restrict__ Ra) {
    const int b = blockIdx.y;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (idx >= nAnts * nAnts) return;
    const int i = idx % nAnts, j = idx / nAnts;
    const int ii = i <= j ? i : j, jj = i <= j ? j : i;  // take (ii,jj) from the upper triangle
    // find a block pair containing (ii,jj): ib = ii/BI, jb = jj/BJ always satisfies the 'needed' test
    const int ib = ii / kCovBI, jb = jj / kCovBJ;
    const int pair = ib * jBlocks + jb;
    const int e = (ii - ib * kCovBI) * kCovBJ + (jj - jb * kCovBJ);
    double2 s = make_double2(0.0, 0.0);
    for (int c = lane; c < chunks; c += 32)
        s = zadd(s, part[(((long long)b * nPairs + pair) * chunks + c) * (kCovBI * kCovBJ) + e]);
    s.x = warp_sum(s.x);
    s.y = warp_sum(s.y);
    if (lane) return;
    s.x *= invN;
    s.y *= invN;
    if (i > j) s.y = -s.y;
    if (i == j) s.y = 0.0;
    Ra[(long long)b * nAnts * nAnts + idx] = s;
}

// Arrays of up to 8 elements (the sensing arrays of the shipped scenarios): ONE pass over the grid.  Every thread keeps the whole
// upper triangle (NA(NA+1)/2 complex float64 accumulators) and walks its share of the samples two at a time with 128-bit
// streaming loads -- 2 NA loads in flight per thread and iteration, each antenna stream read exactly once, coalesced -- instead
// of one CTA set per 4x4 block pair re-reading the streams of its rows and columns.  One CTA per SM (a single wave); the chunk
// partials are added in chunk order by the CTA that finishes last (ticket counter), so the result does not depend on the
// scheduling and no second launch is needed.
template <int NA>
__global__ void __launch_bounds__(kCovThreads, 1)
cov_full_kernel(const float2* __restrict__ rx, long long N, int nAnts, int chunks, double2* __restrict__ part,
                unsigned* __restrict__ tickets, double invN, double2* __restrict__ Ra) {
    constexpr int NT = NA * (NA + 1) / 2;
    const int chunk = blockIdx.x, b = blockIdx.y;
    double2 acc[NT];
#pragma unroll
    for (int e = 0; e < NT; ++e) acc[e] = make_double2(0.0, 0.0);
    const float2* __restrict__ base = rx + (long long)b * nAnts * N;
    const long long nP = N >> 1;   // sample pairs (N even: checked on the host)
    const long long per = (nP + chunks - 1) / chunks;
    const long long u0 = (long long)chunk * per, u1 = (u0 + per < nP) ? u0 + per : nP;
    // software pipeline: the loads of the next sample pair are issued right after the current pair has been widened to float64,
    // so their latency runs behind the 8 NT fused multiply-adds of the current pair
    auto fetch = [&](long long u, float4 (&v)[NA]) {
#pragma unroll
        for (int a = 0; a < NA; ++a)
            v[a] = (a < nAnts && u < u1) ? __ldcs(reinterpret_cast<const float4*>(base + (long long)a * N) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 v[NA];
    fetch(u0 + threadIdx.x, v);
#pragma unroll 1
    for (long long u = u0 + threadIdx.x; u < u1; u += kCovThreads) {
        double2 xl[NA], xh[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            xl[a] = make_double2((double)v[a].x, (double)v[a].y);   // (widening on the integer pipe instead of F2F: measured no gain)
            xh[a] = make_double2((double)v[a].z, (double)v[a].w);
        }
        fetch(u + kCovThreads, v);
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int j = i; j < NA; ++j) {   // acc += conj(x_i) x_j, products of float32 values are exact in float64
                double2& c = acc[i * NA - i * (i - 1) / 2 + (j - i)];
                c.x = fma(xl[i].x, xl[j].x, fma(xl[i].y, xl[j].y, c.x));
                c.y = fma(xl[i].x, xl[j].y, fma(-xl[i].y, xl[j].x, c.y));
                c.x = fma(xh[i].x, xh[j].x, fma(xh[i].y, xh[j].y, c.x));
                c.y = fma(xh[i].x, xh[j].y, fma(-xh[i].y, xh[j].x, c.y));
            }
    }
    __shared__ double2 red[kCovThreads / 32][NT];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < NT; ++e) {
        const double re = warp_sum(acc[e].x), im = warp_sum(acc[e].y);
        if (lane == 0) red[warp][e] = make_double2(re, im);
    }
    __syncthreads();
    if (threadIdx.x < NT) {
        double2 s = make_double2(0.0, 0.0);
        for (int w = 0; w < kCovThreads / 32; ++w) s = zadd(s, red[w][threadIdx.x]);
        part[((long long)b * chunks + chunk) * NT + threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(tickets + b, 1u) == (unsigned)chunks - 1u;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // one warp per entry, the lanes stride over the chunk partials (independent loads), then a fixed-order shuffle tree: deterministic
    for (int e = warp; e < NT; e += kCovThreads / 32) {
        double2 s = make_double2(0.0, 0.0);
        const double2* __restrict__ pb = part + (long long)b * chunks * NT + e;
        for (int c = lane; c < chunks; c += 32) s = zadd(s, __ldcg(pb + (long long)c * NT));
        s.x = warp_sum(s.x);
        s.y = warp_sum(s.y);
        int i = 0, r = e;
        while (r >= NA - i) { r -= NA - i; ++i; }
        const int j = i + r;
        if (lane == 0 && i < nAnts && j < nAnts) {
            s.x *= invN;
            s.y *= invN;
            if (i == j) s.y = 0.0;
            Ra[(long long)b * nAnts * nAnts + i + (long long)j * nAnts] = s;
            if (i != j) Ra[(long long)b * nAnts * nAnts + j + (long long)i * nAnts] = make_double2(s.x, -s.y);
        }
    }
    if (threadIdx.x == 0) tickets[b] = 0u;   // ready for the next launch on this stream
}

int cov_antenna(Ctx* ctx, const float2* rx, long long N, int nAnts, int batch, double2* Ra, cudaStream_t st) {
    if (!rx || !Ra || N < 1 || nAnts < 1 || batch < 1) {
        set_error(ctx, "cov_antenna: invalid argument");
        return kErrInvalidArg;
    }
    if (nAnts <= 8 && (N & 1) == 0 && batch <= 1024 && (reinterpret_cast<uintptr_t>(rx) & 15) == 0) {   // single-pass kernel
        const int NA = nAnts <= 2 ? 2 : (nAnts <= 4 ? 4 : 8), NT = NA * (NA + 1) / 2;
        int chunks = ctx->numSMs / batch;
        if (chunks < 1) chunks = 1;
        const long long maxChunks = ((N >> 1) + kCovThreads - 1) / kCovThreads;
        if (chunks > maxChunks) chunks = (int)maxChunks;
        void* part = nullptr;
        void* tick = nullptr;
        int s = ctx_scratch(ctx, 8, sizeof(double2) * (size_t)batch * chunks * NT, &part);
        if (s) return s;
        static_assert(sizeof(unsigned) * 1024 <= 4096, "ticket slot");
        const bool fresh = ctx->scratchBytes[24] < 4096;
        if ((s = ctx_scratch(ctx, 24, 4096, &tick))) return s;
        if (fresh) ISAC_CUDA_CHECK(ctx, cudaMemsetAsync(tick, 0, 4096, st));   // the kernels leave the counters at zero
        dim3 grid(chunks, batch);
        const int pr = prof_begin(ctx, kProfCov, st);
        const double invN = 1.0 / (double)N;
        if (NA == 2) cov_full_kernel<2><<<grid, kCovThreads, 0, st>>>(rx, N, nAnts, chunks, (double2*)part, (unsigned*)tick, invN, Ra);
        else if (NA == 4) cov_full_kernel<4><<<grid, kCovThreads, 0, st>>>(rx, N, nAnts, chunks, (double2*)part, (unsigned*)tick, invN, Ra);
        else cov_full_kernel<8><<<grid, kCovThreads, 0, st>>>(rx, N, nAnts, chunks, (double2*)part, (unsigned*)tick, invN, Ra);
        prof_end(ctx, pr, st);
        count_launches(ctx, 1);
        ISAC_CUDA_CHECK(ctx, cudaGetLastError());
        return kOk;
    }
    const int iBlocks = (nAnts + kCovBI - 1) / kCovBI, jBlocks = (nAnts + kCovBJ - 1) / kCovBJ;
    const int nPairs = iBlocks * jBlocks;
    // one full wave of CTAs that do work (2 per SM, launch bounds): only the block pairs touching the upper triangle count
    int needed = 0;
    for (int ib = 0; ib < iBlocks; ++ib)
        for (int jb = 0; jb < jBlocks; ++jb) needed += (jb * kCovBJ + kCovBJ - 1 >= ib * kCovBI);
    int chunks = (2 * ctx->numSMs) / (needed * batch);
    if (chunks < 1) chunks = 1;
    const long long maxChunks = (N + kCovThreads - 1) / kCovThreads;
    if (chunks > maxChunks) chunks = (int)maxChunks;
    void* part = nullptr;
    int s = ctx_scratch(ctx, 8, sizeof(double2) * (size_t)batch * nPairs * chunks * kCovBI * kCovBJ, &part);
    if (s) return s;
    dim3 grid(chunks, nPairs, batch);
    const int pr = prof_begin(ctx, kProfCov, st);
    cov_partial_kernel<<<grid, kCovThreads, 0, st>>>(rx, N, nAnts, jBlocks, chunks, (double2*)part);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    dim3 g2((nAnts * nAnts + 7) / 8, batch);   // 8 warps = 8 entries per CTA
    cov_final_kernel<<<g2, 256, 0, st>>>((const double2*)part, nAnts, jBlocks, nPairs, chunks, 1.0 / (double)N, Ra);
    prof_end(ctx, pr, st);
    count_launches(ctx, 2);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// ------------------------------------------------------------------------------------------
// round-robin (circle method) pairing: np even players, round `step` in [0, np-1), slot k in [0, np/2)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void rr_pair(int np, int step, int k, int& p, int& q) {
    const int m = np - 1;
    if (k == 0) {
        p = m;
        q = step % m;
    } else {
        p = (step + k) % m;
        q = (step + m - k) % m;
    }
    if (p > q) {
        int t = p;
        p = q;
        q = t;
    }
}

// Rotation that makes columns (gp, gq) orthogonal given alpha=|gp|^2, beta=|gq|^2, gamma=gp'gq.
// Returns false when already orthogonal to working precision.
__device__ __forceinline__ bool jacobi_rotation(double alpha, double beta, double2 gamma, double tol, double& c,
                                                double& s, double2& ph) {
    const double g2 = gamma.x * gamma.x + gamma.y * gamma.y;
    if (!(g2 > tol * tol * alpha * beta) || g2 == 0.0) return false;
    // reciprocal square roots instead of sqrt + division: the rotation parameters sit on the serial path of every Jacobi step
    // (one warp per column pair waits for them), and a rotation that is a few ulp off is corrected by the next sweep
    const double ig = rsqrt(g2);
    ph = make_double2(gamma.x * ig, -gamma.y * ig);  // e^{-i phi}
    const double zeta = 0.5 * (beta - alpha) * ig;
    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
    c = rsqrt(fma(t, t, 1.0));
    s = c * t;
    return true;
}

// ------------------------------------------------------------------------------------------
// small Hermitian PSD eigen-solver: one CTA per matrix, matrix + eigenvectors in shared memory
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
eig_small_kernel(const double2* __restrict__ A, int n, double* __restrict__ w, double2* __restrict__ Vout, int maxSweeps) {
    extern __shared__ double2 smd[];
    double2* G = smd;
    double2* V = smd + n * n;
    __shared__ int rotated;
    __shared__ double lam[kSmallEigMax];
    const int b = blockIdx.x;
    const double2* __restrict__ Ab = A + (long long)b * n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        const double2 a = Ab[i + j * n], at = Ab[j + i * n];
        G[idx] = make_double2(0.5 * (a.x + at.x), 0.5 * (a.y - at.y));  // (A + A')/2
        V[idx] = make_double2(i == j ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    const int np = n + (n & 1), nPairs = np / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const double tol = 1e-15;
    for (int sweep = 0; sweep < maxSweeps; ++sweep) {
        if (threadIdx.x == 0) rotated = 0;
        __syncthreads();
        for (int step = 0; step < np - 1; ++step) {
            for (int k = warp; k < nPairs; k += nWarps) {
                int p, q;
                rr_pair(np, step, k, p, q);
                if (q >= n) continue;
                double al = 0.0, be = 0.0, gr = 0.0, gi = 0.0;
                for (int r = lane; r < n; r += 32) {
                    const double2 a = G[r + p * n], c2 = G[r + q * n];
                    al += a.x * a.x + a.y * a.y;
                    be += c2.x * c2.x + c2.y * c2.y;
                    gr += a.x * c2.x + a.y * c2.y;  // conj(a)*c
                    gi += a.x * c2.y - a.y * c2.x;
                }
                al = warp_sum(al);
                be = warp_sum(be);
                gr = warp_sum(gr);
                gi = warp_sum(gi);
                double c, s;
                double2 ph;
                if (jacobi_rotation(al, be, make_double2(gr, gi), tol, c, s, ph)) {
                    if (lane == 0) rotated = 1;
                    for (int r = lane; r < n; r += 32) {
                        double2 a = G[r + p * n], qt = zmul(ph, G[r + q * n]);
                        G[r + p * n] = make_double2(c * a.x - s * qt.x, c * a.y - s * qt.y);
                        G[r + q * n] = make_double2(s * a.x + c * qt.x, s * a.y + c * qt.y);
                        a = V[r + p * n];
                        qt = zmul(ph, V[r + q * n]);
                        V[r + p * n] = make_double2(c * a.x - s * qt.x, c * a.y - s * qt.y);
                        V[r + q * n] = make_double2(s * a.x + c * qt.x, s * a.y + c * qt.y);
                    }
                }
            }
            __syncthreads();
        }
        const int again = rotated;
        __syncthreads();
        if (!again) break;
    }
    // eigenvalue = Rayleigh quotient v' A v = Re(sum conj(V[:,i]) .* G[:,i])   (G = A V)
    for (int i = warp; i < n; i += nWarps) {
        double acc = 0.0;
        for (int r = lane; r < n; r += 32) {
            const double2 v = V[r + i * n], g = G[r + i * n];
            acc += v.x * g.x + v.y * g.y;
        }
        acc = warp_sum(acc);
        if (lane == 0) lam[i] = acc;
    }
    __syncthreads();
    // rank (descending, stable)
    for (int i = warp; i < n; i += nWarps) {
        int rank = 0;
        const double li = lam[i];
        for (int j = 0; j < n; ++j) rank += (lam[j] > li) || (lam[j] == li && j < i);
        if (lane == 0) w[(long long)b * n + rank] = li;
        for (int r = lane; r < n; r += 32) Vout[(long long)b * n * n + r + (long long)rank * n] = V[r + i * n];
    }
}

int eig_psd_small(Ctx* ctx, const double2* A, int n, int batch, double* w, double2* V, cudaStream_t st) {
    if (n < 1 || n > kSmallEigMax || batch < 1) {
        set_error(ctx, "eig_psd_small: n out of range");
        return kErrInvalidArg;
    }
    const int np = n + (n & 1);
    int threads = 32 * (np / 2);
    if (threads < 32) threads = 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = sizeof(double2) * 2 * (size_t)n * n;
    cudaFuncSetAttribute(eig_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int pr = prof_begin(ctx, kProfMusic, st);
    eig_small_kernel<<<batch, threads, smem, st>>>(A, n, w, V, 40);
    prof_end(ctx, pr, st);
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// ------------------------------------------------------------------------------------------
// large one-sided Jacobi: one CTA per column pair and round, matrix in global memory (L2 resident)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
jacobi_round_kernel(double2* __restrict__ G, int m, int n, double2* __restrict__ V, int np, int step, double tol,
                    int* __restrict__ rotated) {
    int p, q;
    rr_pair(np, step, blockIdx.x, p, q);
    if (q >= n) return;
    double2* __restrict__ gp = G + (long long)p * m;
    double2* __restrict__ gq = G + (long long)q * m;
    double al = 0.0, be = 0.0, gr = 0.0, gi = 0.0;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const double2 a = gp[r], c2 = gq[r];
        al += a.x * a.x + a.y * a.y;
        be += c2.x * c2.x + c2.y * c2.y;
        gr += a.x * c2.x + a.y * c2.y;
        gi += a.x * c2.y - a.y * c2.x;
    }
    __shared__ double red[4][8];
    __shared__ double rot[4];
    __shared__ int doRot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    al = warp_sum(al);
    be = warp_sum(be);
    gr = warp_sum(gr);
    gi = warp_sum(gi);
    if (lane == 0) {
        red[0][warp] = al;
        red[1][warp] = be;
        red[2][warp] = gr;
        red[3][warp] = gi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b2 = 0, x = 0, y = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) {
            a += red[0][wv];
            b2 += red[1][wv];
            x += red[2][wv];
            y += red[3][wv];
        }
        double c, s;
        double2 ph;
        const bool r = jacobi_rotation(a, b2, make_double2(x, y), tol, c, s, ph);
        doRot = r ? 1 : 0;
        if (r) {
            rot[0] = c;
            rot[1] = s;
            rot[2] = ph.x;
            rot[3] = ph.y;
            *rotated = 1;
        }
    }
    __syncthreads();
    if (!doRot) return;
    const double c = rot[0], s = rot[1];
    const double2 ph = make_double2(rot[2], rot[3]);
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const double2 a = gp[r], qt = zmul(ph, gq[r]);
        gp[r] = make_double2(c * a.x - s * qt.x, c * a.y - s * qt.y);
        gq[r] = make_double2(s * a.x + c * qt.x, s * a.y + c * qt.y);
    }
    if (V) {
        double2* __restrict__ vp = V + (long long)p * n;
        double2* __restrict__ vq = V + (long long)q * n;
        for (int r = threadIdx.x; r < n; r += blockDim.x) {
            const double2 a = vp[r], qt = zmul(ph, vq[r]);
            vp[r] = make_double2(c * a.x - s * qt.x, c * a.y - s * qt.y);
            vq[r] = make_double2(s * a.x + c * qt.x, s * a.y + c * qt.y);
        }
    }
}

__global__ void __launch_bounds__(256)
colnorm_kernel(const double2* __restrict__ G, int m, int n, double* __restrict__ sigma) {
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const double2 a = G[(long long)j * m + r];
        acc += a.x * a.x + a.y * a.y;
    }
    __shared__ double red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += red[wv];
        sigma[j] = sqrt(s);
    }
}

// order[rank] = column index, descending sigma (stable)
__global__ void rank_desc_kernel(const double* __restrict__ sigma, int n, int* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double li = sigma[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        const double lj = sigma[j];
        rank += (lj > li) || (lj == li && j < i);
    }
    order[rank] = i;
}

__global__ void identity_kernel(double2* V, int n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * n) return;
    V[idx] = make_double2((idx % n) == (idx / n) ? 1.0 : 0.0, 0.0);
}

int set_identity(Ctx* ctx, double2* V, int n, cudaStream_t st) {
    const long long tot = (long long)n * n;
    identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(V, n);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

int svd_onesided_jacobi(Ctx* ctx, double2* G, int m, int n, double2* V, double* sigma, int* order, int* sweepsOut,
                        cudaStream_t st) {
    if (!G || m < 1 || n < 1 || !sigma || !order) {
        set_error(ctx, "svd_onesided_jacobi: invalid argument");
        return kErrInvalidArg;
    }
    void* flag = nullptr;
    int s = ctx_scratch(ctx, 9, sizeof(int) * 4, &flag);
    if (s) return s;
    int* dFlag = (int*)flag;
    const int np = n + (n & 1);
    const double tol = 2.2204460492503131e-16 * 4.0 * std::sqrt((double)(m > 64 ? m : 64));
    int sweeps = 0;
    if (n > 1) {
        for (; sweeps < 40; ++sweeps) {
            ISAC_CUDA_CHECK(ctx, cudaMemsetAsync(dFlag, 0, sizeof(int), st));
            for (int step = 0; step < np - 1; ++step)
                jacobi_round_kernel<<<np / 2, 256, 0, st>>>(G, m, n, V, np, step, tol, dFlag);
            ISAC_CUDA_CHECK(ctx, cudaGetLastError());
            int h = 0;
            ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(&h, dFlag, sizeof(int), cudaMemcpyDeviceToHost, st));
            ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
            if (!h) {
                ++sweeps;
                break;
            }
        }
    }
    if (sweepsOut) *sweepsOut = sweeps;
    colnorm_kernel<<<n, 256, 0, st>>>(G, m, n, sigma);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    rank_desc_kernel<<<(n + 255) / 256, 256, 0, st>>>(sigma, n, order);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// ------------------------------------------------------------------------------------------
// number of targets: eigen-gap rule (music.m:109-125) on eig()'s ascending order
// ------------------------------------------------------------------------------------------
__device__ int num_targets_rule(const double* wDesc, int n) {
    // ascending V(i) = wDesc[n-1-i]; deltaV = -diff(V), length n-1
    const int nd = n - 1;
    if (nd < 1) return 0;
    const int start = (nd + 1 + 1) / 2;  // ceil((nd+1)/2), 1-based
    double sum = 0.0;
    int cnt = 0;
    for (int i = start - 1; i < nd; ++i) {
        sum += -(wDesc[n - 1 - (i + 1)] - wDesc[n - 1 - i]);
        ++cnt;
    }
    const double halfMean = sum / (double)cnt;
    int best = 0;
    double bestV = -INFINITY;
    for (int i = 0; i < nd; ++i) {
        const double dv = -(wDesc[n - 1 - (i + 1)] - wDesc[n - 1 - i]) - 2.0 * halfMean;
        if (dv > bestV) {
            bestV = dv;
            best = i;
        }
    }
    return best + 1;
}

__global__ void num_targets_kernel(const double* wDesc, int n, int* Lout) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *Lout = num_targets_rule(wDesc, n);
}

int music_num_targets(Ctx* ctx, const double* wDesc, int n, int* Lout, cudaStream_t st) {
    num_targets_kernel<<<1, 32, 0, st>>>(wDesc, n, Lout);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// ------------------------------------------------------------------------------------------
// findpeaks(y,'NPeaks',L,'SortStr','descend') on one CTA; y in global/shared, S samples
// ------------------------------------------------------------------------------------------
__device__ void findpeaks_block(const double* __restrict__ y, int S, int L, unsigned char* cand /*[S] scratch*/,
                                int* peakLoc, int* nPeaksOut) {
    __shared__ double bestV[32];
    __shared__ int bestI[32];
    __shared__ int chosen;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        unsigned char c = 0;
        if (i >= 1 && i < S - 1 && y[i] > y[i - 1]) {
            int j = i;
            while (j < S - 1 && y[j + 1] == y[i]) ++j;
            if (j < S - 1 && y[j + 1] < y[i]) c = 1;
        }
        cand[i] = c;
    }
    __syncthreads();
    if (L > kMaxPeaks) L = kMaxPeaks;
    int found = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    for (int it = 0; it < L; ++it) {
        double bv = -INFINITY;
        int bi = -1;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
            if (cand[i] && (bi < 0 || y[i] > bv)) {  // strided ascending i: keeps the lowest index on ties
                bv = y[i];
                bi = i;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            bestV[warp] = bv;
            bestI[warp] = bi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = -INFINITY;
            int ix = -1;
            for (int wv = 0; wv < nWarps; ++wv)
                if (bestI[wv] >= 0 && (ix < 0 || bestV[wv] > v || (bestV[wv] == v && bestI[wv] < ix))) {
                    v = bestV[wv];
                    ix = bestI[wv];
                }
            chosen = ix;
            if (ix >= 0) {
                peakLoc[it] = ix + 1;  // 1-based
                cand[ix] = 0;
            }
        }
        __syncthreads();
        if (chosen < 0) break;
        ++found;
        __syncthreads();
    }
    if (threadIdx.x == 0) *nPeaksOut = found;
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// ULA MUSIC (music.m:73-104), direct noise-subspace form, one CTA per batch item
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
music_ula_kernel(const double* __restrict__ w, const double2* __restrict__ V, int n, DoaConfig cfg, LSource ls, int method,
                 int aSteps, int* __restrict__ Lout, double* __restrict__ P, double* __restrict__ PdB,
                 int* __restrict__ peakLoc, int* __restrict__ nPeaks, int* __restrict__ status) {
    extern __shared__ unsigned char smraw[];
    double2* Vs = (double2*)smraw;                       // n*n
    double* Pb = (double*)(Vs + n * n);                  // aSteps
    unsigned char* cand = (unsigned char*)(Pb + aSteps); // aSteps
    __shared__ int Lsh;
    __shared__ double red[16];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) Vs[i] = V[(long long)b * n * n + i];
    if (threadIdx.x == 0) {
        int L;
        if (ls.givenL) L = ls.givenL[b];
        else if (ls.rowmask) {
            L = 0;
            for (int k = 0; k < ls.rowWords; ++k) L += __popc(ls.rowmask[(long long)b * ls.rowWords + k]);
        } else if (ls.fixedL > 0) L = ls.fixedL;
        else L = num_targets_rule(w + (long long)b * n, n);
        Lsh = L;
        Lout[b] = L;
    }
    __syncthreads();
    const int L = Lsh;
    if (L < 1) {  // findpeaks(...,'NPeaks',0) errors in the reference
        if (threadIdx.x == 0) {
            status[b] = kErrNumDetsZero;
            nPeaks[b] = 0;
        }
        return;
    }
    const double eps1 = 2.220446049250313e-16;
    // All three scanners are sum_k g_k |u_k' a|^2 over the eigenpairs (w_k, u_k) of Ra:
    //   MUSIC  g_k = [k >= L]   (a' Un Un' a,  music.m:28-29,90)
    //   MVDR   g_k = 1 / w_k    (a' Ra^-1 a,   mvdrBF.m:73)
    //   DBF    g_k = w_k        (a' Ra a,      digitalBF.m:73)
    const double* __restrict__ wb = w + (long long)b * n;
    for (int a = threadIdx.x; a < aSteps; a += blockDim.x) {
        const double ang = a * cfg.aGran - cfg.aMax / 2.0;            // music.m:88
        const double sd = sind_dev(ang);
        double q = 0.0;
        for (int k = (method == kDoaMusic ? L : 0); k < n; ++k) {      // noise eigenvectors (music.m:28-29)
            double re = 0.0, im = 0.0;
            for (int m = 0; m < n; ++m) {
                const double2 av = cis2pi(-(double)m * cfg.d * sd);    // aULA (music.m:82)
                const double2 u = Vs[m + k * n];
                re += u.x * av.x + u.y * av.y;                         // conj(u)*a
                im += u.x * av.y - u.y * av.x;
            }
            const double g = method == kDoaMusic ? 1.0 : (method == kDoaMvdr ? 1.0 / wb[k] : wb[k]);
            q += g * (re * re + im * im);
        }
        Pb[a] = method == kDoaDbf ? fabs(q) : fabs(1.0 / (q + eps1));  // music.m:90,94 / mvdrBF.m:73,77 / digitalBF.m:73,77
    }
    __syncthreads();
    double mx = 0.0;
    for (int a = threadIdx.x; a < aSteps; a += blockDim.x) mx = fmax(mx, Pb[a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) mx = fmax(mx, red[wv]);
    for (int a = threadIdx.x; a < aSteps; a += blockDim.x) {
        const double pv = Pb[a];
        P[(long long)b * aSteps + a] = pv;
        const double db = 20.0 * log10(pv / mx);                       // music.m:95-96
        PdB[(long long)b * aSteps + a] = db;
        Pb[a] = db;
    }
    __syncthreads();
    findpeaks_block(Pb, aSteps, L, cand, peakLoc + (long long)b * kMaxPeaks, nPeaks + b);
    if (threadIdx.x == 0) status[b] = kOk;
}

int music_doa_ula(Ctx* ctx, const double* w, const double2* V, int n, int batch, const DoaConfig& cfg,
                  const LSource& ls, int* Lout, double* P, double* PdB, int* peakLoc, int* nPeaks, int* status,
                  cudaStream_t st, int method) {
    if (n < 2 || n > kSmallEigMax) {
        set_error(ctx, "music_doa_ula: array size must be in [2,64]");
        return kErrUnsupported;
    }
    const int aSteps = (int)std::floor((cfg.aMax + 1.0) / cfg.aGran);  // music.m:79
    const size_t smem = sizeof(double2) * (size_t)n * n + sizeof(double) * aSteps + aSteps + 16;
    cudaFuncSetAttribute(music_ula_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int pr = prof_begin(ctx, kProfMusic, st);
    music_ula_kernel<<<batch, 512, smem, st>>>(w, V, n, cfg, ls, method, aSteps, Lout, P, PdB, peakLoc, nPeaks, status);
    prof_end(ctx, pr, st);
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// ------------------------------------------------------------------------------------------
// complement-form scans (large arrays)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scan1d_kernel(const double2* __restrict__ vecs, long long ld, const int* __restrict__ order,
              const double* __restrict__ colInvNorm, int len, int nVecs, int conjVec, const int* __restrict__ dL,
              double coef, double x0, double dx, int steps, double* __restrict__ q) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * (blockDim.x >> 5) + warp;
    if (i >= steps) return;
    int L = *dL;
    if (L > nVecs) L = nVecs;
    const double x = x0 + i * dx;
    double acc = 0.0;
    for (int k = 0; k < L; ++k) {
        const int col = order ? order[k] : k;
        const double2* __restrict__ u = vecs + (long long)col * ld;
        double re = 0.0, im = 0.0;
        for (int n = lane; n < len; n += 32) {
            const double2 a = cis2pi(coef * x * (double)n);
            const double2 uv = u[n];
            if (conjVec) {  // eigenvector is conj(u): <conj(u), a> = sum u*a
                re += uv.x * a.x - uv.y * a.y;
                im += uv.x * a.y + uv.y * a.x;
            } else {        // <u, a> = sum conj(u)*a
                re += uv.x * a.x + uv.y * a.y;
                im += uv.x * a.y - uv.y * a.x;
            }
        }
        re = warp_sum(re);
        im = warp_sum(im);
        const double sc = colInvNorm ? colInvNorm[col] : 1.0;
        acc += (re * re + im * im) * sc * sc;
    }
    if (lane == 0) q[i] = (double)len - acc;
}

int music_scan_1d(Ctx* ctx, const double2* vecs, long long ld, const int* order, const double* colInvNorm, int len,
                  int nVecs, int conjVec, const int* dL, double coef, double x0, double dx, int steps, double* q,
                  cudaStream_t st) {
    const int wpb = 8;
    scan1d_kernel<<<(steps + wpb - 1) / wpb, wpb * 32, 0, st>>>(vecs, ld, order, colInvNorm, len, nVecs, conjVec, dL,
                                                               coef, x0, dx, steps, q);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

__global__ void __launch_bounds__(1024)
finish1d_kernel(const double* __restrict__ q, int steps, const int* __restrict__ dL, double* __restrict__ P,
                double* __restrict__ PdB, int* __restrict__ peakLoc, int* __restrict__ nPeaks) {
    extern __shared__ unsigned char smraw[];
    double* yb = (double*)smraw;
    unsigned char* cand = (unsigned char*)(yb + steps);
    __shared__ double red[32];
    double mx = 0.0;
    for (int i = threadIdx.x; i < steps; i += blockDim.x) {
        const double pv = fabs(1.0 / q[i]);  // music2D.m:101,111
        yb[i] = pv;
        P[i] = pv;
        mx = fmax(mx, pv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) mx = fmax(mx, red[wv]);
    __syncthreads();
    for (int i = threadIdx.x; i < steps; i += blockDim.x) {
        const double db = 20.0 * log10(yb[i] / mx);  // music2D.m:112-113
        PdB[i] = db;
        yb[i] = db;
    }
    __syncthreads();
    int L = *dL;
    if (L < 1) {
        if (threadIdx.x == 0) *nPeaks = 0;
        return;
    }
    findpeaks_block(yb, steps, L, cand, peakLoc, nPeaks);
}

int music_finish_1d(Ctx* ctx, const double* q, int steps, const int* dL, double* P, double* PdB, int* peakLoc,
                    int* nPeaks, cudaStream_t st) {
    const size_t smem = sizeof(double) * steps + steps + 16;
    cudaFuncSetAttribute(finish1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    finish1d_kernel<<<1, 1024, smem, st>>>(q, steps, dL, P, PdB, peakLoc, nPeaks);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// UPA (music.m:31-63): one warp per (elevation, azimuth) point
__global__ void __launch_bounds__(256)
upa_scan_kernel(const double2* __restrict__ vecs, long long ld, const int* __restrict__ order, int n, DoaConfig cfg,
                const int* __restrict__ dL, int aSteps, int eSteps, double* __restrict__ P, int method,
                const double* __restrict__ wDesc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (pt >= (long long)aSteps * eSteps) return;
    const int e = (int)(pt % eSteps), a = (int)(pt / eSteps);  // Pmusic(e,a), column-major
    const double el = e * cfg.eGran - cfg.eMax / 2.0;          // music.m:52
    const double az = a * cfg.aGran - cfg.aMax / 2.0;          // music.m:53
    const double st = sind_dev(el), ca = cosd_dev(az), sa = sind_dev(az);
    int L = method == kDoaMusic ? *dL : n;   // MVDR / DBF: weighted sum over ALL eigenpairs (see music_ula_kernel)
    if (L > n) L = n;
    double acc = 0.0;
    for (int k = 0; k < L; ++k) {
        const double2* __restrict__ u = vecs + (long long)(order ? order[k] : k) * ld;
        double re = 0.0, im = 0.0;
        for (int idx = lane; idx < n; idx += 32) {
            const int ny = idx % cfg.nY, mx = idx / cfg.nY;    // reshape of [nY x nX] (music.m:55)
            const double2 av = cis2pi(-st * ((double)mx * cfg.d * ca + (double)ny * cfg.d * sa));  // aUPA (music.m:44)
            const double2 uv = u[idx];
            re += uv.x * av.x + uv.y * av.y;
            im += uv.x * av.y - uv.y * av.x;
        }
        re = warp_sum(re);
        im = warp_sum(im);
        const double g = method == kDoaMusic ? 1.0 : (method == kDoaMvdr ? 1.0 / wDesc[k] : wDesc[k]);
        acc += g * (re * re + im * im);
    }
    if (lane == 0) {
        if (method == kDoaMusic) {
            double q = (double)n - acc;
            if (L >= n) q = 0.0;
            P[pt] = fabs(1.0 / (q + 2.220446049250313e-16));   // music.m:56 then abs (:61)
        } else if (method == kDoaMvdr) {
            P[pt] = fabs(1.0 / (acc + 2.220446049250313e-16)); // mvdrBF.m:40,43
        } else {
            P[pt] = fabs(acc);                                  // digitalBF.m:40,43
        }
    }
}

// Pmusic = -abs(P); PmusicNorm = Pmusic./max(Pmusic); mag2db (music.m:61-63, mvdrBF.m:43-45, digitalBF.m:43-45).
// max() of the [eSteps x aSteps] matrix is MATLAB's column-wise maximum (along the single row when eSteps == 1) and
// implicit expansion divides every column by its own maximum = minus the column's smallest magnitude.
// One warp per azimuth column (eSteps > 1) / one warp for the whole row (eSteps == 1).
__global__ void __launch_bounds__(256)
upa_norm_db_kernel(const double* __restrict__ P, int eSteps, int aSteps, double* __restrict__ PdB) {
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nCols = eSteps > 1 ? aSteps : 1, len = eSteps > 1 ? eSteps : aSteps;
    if (col >= nCols) return;
    const double* __restrict__ c = P + (long long)col * len;
    double mn = INFINITY;
    for (int i = lane; i < len; i += 32) mn = fmin(mn, c[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    for (int i = lane; i < len; i += 32) PdB[(long long)col * len + i] = 20.0 * log10(c[i] / mn);
}

int music_doa_upa(Ctx* ctx, const double2* vecs, long long ld, const int* order, int n, const DoaConfig& cfg,
                  const int* dL, double* P, double* PdB, cudaStream_t st, int method, const double* wDesc) {
    const int aSteps = (int)std::floor((cfg.aMax + 1.0) / cfg.aGran);  // music.m:40
    const int eSteps = (int)std::floor((cfg.eMax + 1.0) / cfg.eGran);  // music.m:41
    if (cfg.nX * cfg.nY != n) {
        set_error(ctx, "music_doa_upa: nX*nY must equal the covariance size");
        return kErrInvalidArg;
    }
    const long long tot = (long long)aSteps * eSteps;
    if (method != kDoaMusic && !wDesc) {
        set_error(ctx, "music_doa_upa: MVDR / beamscan need the eigenvalues");
        return kErrInvalidArg;
    }
    upa_scan_kernel<<<(unsigned)((tot + 7) / 8), 256, 0, st>>>(vecs, ld, order, n, cfg, dL, aSteps, eSteps, P, method, wDesc);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    const int nCols = eSteps > 1 ? aSteps : 1;
    upa_norm_db_kernel<<<(unsigned)((nCols + 7) / 8), 256, 0, st>>>(P, eSteps, aSteps, PdB);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// ------------------------------------------------------------------------------------------
// music2D channel matrix (music2D.m:67-68): H = rx(:,:,1).*conj(tx(:,:,1)), optionally H'
// ------------------------------------------------------------------------------------------
__global__ void channel_kernel(const float2* __restrict__ rx, const float2* __restrict__ tx, int nSc, int nSym,
                               int transpose, double2* __restrict__ H) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nSc * nSym) return;
    const int k = (int)(idx % nSc), l = (int)(idx / nSc);
    const float2 a = rx[idx], b = tx[idx];
    const double re = (double)a.x * b.x + (double)a.y * b.y;
    const double im = (double)a.y * b.x - (double)a.x * b.y;
    if (!transpose) H[idx] = make_double2(re, im);
    else H[(long long)l + (long long)k * nSym] = make_double2(re, -im);  // H' [nSym x nSc]
}

int music2d_channel(Ctx* ctx, const float2* rx, const float2* tx, int nSc, int nSym, int transpose, double2* H,
                    cudaStream_t st) {
    const long long tot = (long long)nSc * nSym;
    channel_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(rx, tx, nSc, nSym, transpose, H);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

}  // namespace isac
