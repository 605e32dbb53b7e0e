// K9'' : Gram-pair SINR search with the subband accumulation fused in (the RI / CQI report path).
//
// The report path (riSelect.m:254-285 -> dlPMISelect.m:385-501, cqiSelect.m:507-632) never looks at SINRPerRE itself, only at
// (a) the per-candidate sum over every RE and layer (dlPMISelect.m:444), (b) the per-subband means at the chosen i1
// (dlPMISelect.m:471-501) and (c) the per-CQI-subband means at the reported PMI (cqiSelect.m:586-614).  All three are sums
// over runs of consecutive CSI-RS REs, so this kernel never writes SINRPerRE (450 MB per 32-UE report at 273 PRB / 8 ports /
// ranks 1-8): one CTA takes a CHUNK of up to G consecutive REs of one subband and one UE, builds the column-pair tables of
// all G REs side by side in shared memory (element e of entry q at [q*G + e], so the G lanes that work on the same
// candidate read one contiguous 16*G-byte segment), evaluates every (candidate, RE) item with the G REs of a candidate on
// adjacent lanes, adds them with warp shuffles in a fixed order and writes ONE partial sum per (chunk, layer, candidate).
// pmi_select_fused_kernel then adds the chunk partials in a fixed order (subband-major) and takes the decisions of
// dlPMISelect.m:449-501.  Item-level parallelism also removes the idle lanes of the per-RE kernel on the high ranks
// (64 candidates of rank 7 / 8 against 128 threads) and amortises the index decoding of the table-building stages over G REs.
//
// Preconditions checked on the host (else the per-RE kernel + pmi_subband_kernel path of comm.cu runs): the mean-of-means
// weights (dlPMISelect.m:481) are uniform inside every PMI and CQI subband (true whenever the CSI-RS occupies one symbol),
// and the Gram-pair dictionary of the report fits in shared memory.
#include "comm.cuh"
#include "ctx.cuh"
#include "sinr_core.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace isac {

struct FusedRank {
    const uint16_t* ent;     // [nCand][ntPad]: table slot (of RE 0) of each packed lower-triangle entry, id * G + swizzle(id) (bit 15: conjugate)
    const uint8_t* valid;
    const double* invScale2; // per candidate 1/scale^2 (explicit codebooks) or nullptr
    double* part;            // [batch][nChunks][nu][nCand] partial sums (NaN: nothing to add)
    double invS2;
    int nCand, nu, ntPad;
};
struct FusedDev {
    const float2* H;
    const double2* beams;
    const uint32_t* pairs;     // [nPairs] atom a | atom a' << 16
    const double2* pal;        // [nPal]
    const uint32_t* cpTerms;   // [cpT/4][nCP][4]: pair index (bit 15: conjugate) | palette index << 16
    const int* reK;
    const int* reL;
    const int* chunkRe0;       // [nChunks] first RE of the chunk
    const int* chunkN;         // [nChunks] REs in the chunk (1..G)
    int K, L, R, P, NB, Pb, nBeams, nChunks, nPairs, nPal, nCP, cpT, nRanks;
    FusedRank rk[kMaxLayers];
    double nVar[kMaxPmiBatch];
};

// Slot of RE e inside the G-wide group of table entry `id`: XOR-swizzled with the entry's index so that lanes which read the
// SAME e of DIFFERENT entries (the table-building stages) spread over all eight 16-byte bank groups, while the G lanes that read
// one entry (the candidate stage) still touch one contiguous 16*G-byte segment.
template <int G>
__device__ __forceinline__ int swz(int id, int e) {
    if (G == 4) return e ^ ((id >> 1) & 3);
    if (G == 2) return e ^ ((id >> 2) & 1);
    return 0;
}

// Low ranks: E of the G REs of a candidate inside one thread (E = G for ranks 1-3, E = 2 for ranks 4-5).  For these ranks the
// factorisation is a handful of flops and the lane-per-RE form spends its time on index decoding, validity loads and the shuffle
// sums; here the index words are decoded once for E REs, the E independent factorisations interleave (instruction-level
// parallelism where the lane-per-RE form had only 3 warps per scheduler to hide the float64 latency) and the chunk sum needs
// fewer (or no) shuffles.  The sum keeps the order of the shuffle tree, (e0 + e1) + (e2 + e3), so all forms give identical bits.
template <int NU, int G, int E>
__device__ __noinline__ void fused_rank_eval_multi(const FusedRank rk, const double2* __restrict__ Gt, double nVar,
                                                   double* __restrict__ part, int nValid) {
    static_assert(G % E == 0 && (E == 1 || E == 2 || E == 4), "REs per thread");
    constexpr int NT = NU * (NU + 1) / 2;
    constexpr int NW = (NT + 7) / 8;
    constexpr int LPC = G / E;   // lanes per candidate
    const int nItems = rk.nCand * LPC;
    for (int base = 0; base < nItems; base += blockDim.x) {   // warp-uniform trip count: every lane reaches the shuffles
        const int item = base + threadIdx.x;
        const bool act = item < nItems;
        const int c = act ? item / LPC : 0, e0 = (item % LPC) * E;
        if (item + (int)blockDim.x < nItems)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(rk.ent + (size_t)((item + blockDim.x) / LPC) * rk.ntPad)));
        double s[E][NU];
        bool ok[E];
#pragma unroll
        for (int j = 0; j < E; ++j) ok[j] = false;
        if (act) {
            const uint4* __restrict__ ep = reinterpret_cast<const uint4*>(rk.ent + (size_t)c * rk.ntPad);
            uint4 ev[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) ev[w] = __ldg(ep + w);
            if (rk.valid[c]) {   // else: restricted precoder, contributes nothing (dlPMISelect.m:418)
                double2 A[E][NT];
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const uint32_t q[4] = {ev[w].x, ev[w].y, ev[w].z, ev[w].w};
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int t = w * 8 + u;
                        if (t < NT) {
                            const uint32_t qq = q[u >> 1];                                   // two entries: slot of RE 0 | conjugate flag << 15
                            const uint32_t ix = ((u & 1) ? (qq >> 16) : qq) & 0x7fffu;
                            const uint32_t sgn = ((u & 1) ? qq : (qq << 16)) & 0x80000000u;  // the flag moved onto the sign bit
#pragma unroll
                            for (int j = 0; j < E; ++j) {
                                const double2 g = Gt[ix ^ (uint32_t)(e0 + j)];            // the swizzle is an XOR of the low bits
                                A[j][t] = make_double2(g.x, __hiloint2double(__double2hiint(g.y) ^ (int)sgn, __double2loint(g.y)));
                            }
                        }
                    }
                }
                const double nv = nVar * (rk.invScale2 ? rk.invScale2[c] : rk.invS2);
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    chol_sinr<NU>(A[j], nv, s[j], 1);
                    ok[j] = e0 + j < nValid;
#pragma unroll
                    for (int l = 0; l < NU; ++l) ok[j] = ok[j] && (s[j][l] == s[j][l]);   // sum(...,'omitnan'): a NaN RE is skipped
                }
            }
        }
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            cnt += ok[j] ? 1 : 0;
            if (!ok[j]) {
#pragma unroll
                for (int l = 0; l < NU; ++l) s[j][l] = 0.0;
            }
        }
        double tot[NU];
#pragma unroll
        for (int l = 0; l < NU; ++l) {
            if (E == 4) tot[l] = (s[0][l] + s[1][l]) + (s[2][l] + s[3][l]);
            else if (E == 2) tot[l] = s[0][l] + s[1][l];
            else tot[l] = s[0][l];
        }
#pragma unroll
        for (int o = 1; o < LPC; o <<= 1) {
#pragma unroll
            for (int l = 0; l < NU; ++l) tot[l] += __shfl_xor_sync(0xffffffffu, tot[l], o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (act && e0 == 0) {
#pragma unroll
            for (int l = 0; l < NU; ++l) part[(size_t)l * rk.nCand + c] = cnt ? tot[l] : NAN;
        }
    }
}

template <int NU, int G>
__device__ __noinline__ void fused_rank_eval(const FusedRank rk, const double2* __restrict__ Gt, double nVar, double* __restrict__ part,
                                             int nValid) {
    constexpr int NT = NU * (NU + 1) / 2;
    constexpr int NW = (NT + 7) / 8;   // 16-byte words of indices per candidate
    const int nItems = rk.nCand * G;
    for (int base = 0; base < nItems; base += blockDim.x) {   // warp-uniform trip count: every lane reaches the shuffles
        const int item = base + threadIdx.x;
        const bool act = item < nItems;
        const int c = act ? item / G : 0, e = item % G;
        if (item + (int)blockDim.x < nItems) {   // pull the next round's index words into L1 behind this round's arithmetic
            const char* nx = reinterpret_cast<const char*>(rk.ent + (size_t)((item + blockDim.x) / G) * rk.ntPad);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
            if (NW > 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + 64));
        }
        double s[NU];
        bool ok = false;
        if (act && e < nValid) {
            const uint4* __restrict__ ep = reinterpret_cast<const uint4*>(rk.ent + (size_t)c * rk.ntPad);
            uint4 ev[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) ev[w] = __ldg(ep + w);   // issued together with the validity flag
            if (rk.valid[c]) {   // else: restricted precoder, contributes nothing (dlPMISelect.m:418)
                double2 A[NT];
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const uint4 v = ev[w];
                    const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int t = w * 8 + u;
                        if (t < NT) {   // bit 15: the pair is stored as (j,i) -> conjugate
                            const uint32_t qq = q[u >> 1];                                   // two entries: slot of RE 0 | conjugate flag << 15
                            const uint32_t slot = ((u & 1) ? (qq >> 16) : qq) & 0x7fffu;
                            const uint32_t sgn = ((u & 1) ? qq : (qq << 16)) & 0x80000000u;  // the flag moved onto the sign bit
                            const double2 g = Gt[slot ^ (uint32_t)e];                        // slot of RE e: the swizzle is an XOR of the low bits
                            A[t] = make_double2(g.x, __hiloint2double(__double2hiint(g.y) ^ (int)sgn, __double2loint(g.y)));
                        }
                    }
                }
                const double nv = nVar * (rk.invScale2 ? rk.invScale2[c] : rk.invS2);
                chol_sinr<NU>(A, nv, s, 1);
                ok = true;
#pragma unroll
                for (int l = 0; l < NU; ++l) ok = ok && (s[l] == s[l]);   // sum(...,'omitnan'): a NaN RE is skipped
            }
        }
        if (!ok) {
#pragma unroll
            for (int l = 0; l < NU; ++l) s[l] = 0.0;
        }
        int cnt = ok ? 1 : 0;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {   // the G REs of a candidate sit on adjacent lanes: fixed-order tree sum
#pragma unroll
            for (int l = 0; l < NU; ++l) s[l] += __shfl_xor_sync(0xffffffffu, s[l], o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (act && e == 0) {
#pragma unroll
            for (int l = 0; l < NU; ++l) part[(size_t)l * rk.nCand + c] = cnt ? s[l] : NAN;
        }
    }
}

// grid: x = chunk, y = UE.  T threads; MINB CTAs per SM the register budget is sized for.
template <int G, int T, int MINB>
__global__ void __launch_bounds__(T, MINB)
pmi_pair_fused_kernel(const __grid_constant__ FusedDev p) {
    extern __shared__ double2 sm[];
    const int R = p.R, P = p.P, nAtoms = p.NB * p.nBeams;
    double2* Hs = sm;                           // [R*P][G]
    double2* pal = Hs + (size_t)R * P * G;      // [nPal]
    double2* Gm = pal + p.nPal;                 // [nPairs][G]
    double2* Bf = Gm + (size_t)p.nPairs * G;    // [R][nAtoms][G]; dead after step 2, the column-pair table reuses it
    double2* Gt = Bf;                           // [nCP][G]
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int re0 = p.chunkRe0[chunk], nValid = p.chunkN[chunk];
    const float2* __restrict__ Hb = p.H + (long long)b * p.K * p.L * R * P;
    for (int i = threadIdx.x; i < R * P * G; i += T) {
        const int e = i % G, rp = i / G, r = rp % R, pp = rp / R;
        double2 h = make_double2(0.0, 0.0);
        if (e < nValid) {
            const long long kk = p.reK[re0 + e] - 1, ll = p.reL[re0 + e] - 1;
            const float2 v = __ldg(Hb + kk + p.K * (ll + (long long)p.L * (r + (long long)R * pp)));
            h = make_double2((double)v.x, (double)v.y);
        }
        Hs[(r * P + pp) * G + e] = h;
    }
    for (int i = threadIdx.x; i < p.nPal; i += T) pal[i] = p.pal[i];
    __syncthreads();
    // 1. beam responses of the G REs: Bf[r][a][e] = H_e[r, block(a)] . v_beam(a)
    for (int i = threadIdx.x; i < nAtoms * G; i += T) {
        const int e = i % G, a = i / G, blk = a / p.nBeams, bm = a - blk * p.nBeams;
        const double2* __restrict__ bv = p.beams + (size_t)bm * p.Pb;
        for (int r = 0; r < R; ++r) {
            const double2* __restrict__ hr = Hs + (size_t)(r * P + blk * p.Pb) * G + e;
            double2 acc = make_double2(0.0, 0.0);
            for (int q = 0; q < p.Pb; ++q) acc = zfma(acc, hr[q * G], __ldg(bv + q));
            Bf[((size_t)r * nAtoms + a) * G + swz<G>(a, e)] = acc;
        }
    }
    __syncthreads();
    // 2. atom Gram pairs Gm[pi][e] = <Bf[a], Bf[a']>: one thread per pair, the G REs inside (the pair word is decoded once)
    //    (two threads per pair with half of the REs each -- 3 full rounds instead of 1.5 at 576 pairs -- measured slower)
    for (int pi = threadIdx.x; pi < p.nPairs; pi += T) {
        const uint32_t w = __ldg(p.pairs + pi);
        const int a0 = (int)(w & 0xffffu), a1 = (int)(w >> 16);
        const double2* __restrict__ pa = Bf + (size_t)a0 * G;
        const double2* __restrict__ pb = Bf + (size_t)a1 * G;
        double2 acc[G];
#pragma unroll
        for (int e = 0; e < G; ++e) acc[e] = make_double2(0.0, 0.0);
        for (int r = 0; r < R; ++r, pa += (size_t)nAtoms * G, pb += (size_t)nAtoms * G) {
#pragma unroll
            for (int e = 0; e < G; ++e) acc[e] = zfmac(acc[e], pb[swz<G>(a1, e)], pa[swz<G>(a0, e)]);  // conj(Bf[a]) Bf[a']
        }
#pragma unroll
        for (int e = 0; e < G; ++e) Gm[(size_t)pi * G + swz<G>(pi, e)] = acc[e];
    }
    __syncthreads();
    // 3. column-pair table Gt[q][e] = sum_t pal[.] Gm[.][e]: one thread per column pair, its terms decoded once for the G REs;
    //    the conjugation flag flips the sign bit of the imaginary part
    const uint4* __restrict__ cpt = reinterpret_cast<const uint4*>(p.cpTerms);
    for (int q = threadIdx.x; q < p.nCP; q += T) {
        double2 acc[G];
#pragma unroll
        for (int e = 0; e < G; ++e) acc[e] = make_double2(0.0, 0.0);
        for (int t = 0; t < p.cpT / 4; ++t) {
            const uint4 v = __ldg(cpt + (size_t)t * p.nCP + q);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double2 c = pal[w4[u] >> 16];
                const int pid = (int)(w4[u] & 0x7fffu);
                const double2* __restrict__ gm = Gm + (size_t)pid * G;
                const long long flip = (long long)(w4[u] & 0x8000u) << 48;   // bit 15 -> bit 63
#pragma unroll
                for (int e = 0; e < G; ++e) {
                    const double2 g = gm[swz<G>(pid, e)];
                    acc[e] = zfma(acc[e], make_double2(g.x, __longlong_as_double(__double_as_longlong(g.y) ^ flip)), c);
                }
            }
        }
#pragma unroll
        for (int e = 0; e < G; ++e) Gt[(size_t)q * G + swz<G>(q, e)] = acc[e];
    }
    __syncthreads();
    // 4. every (candidate, RE) item of every rank
    const double nVar = p.nVar[b];
#pragma unroll
    for (int q = 0; q < kMaxLayers; ++q) {   // static index: the descriptors stay in the parameter bank
        if (q >= p.nRanks) break;
        const FusedRank rk = p.rk[q];
        double* __restrict__ part = rk.part + ((size_t)b * p.nChunks + chunk) * rk.nu * (size_t)rk.nCand;
        switch (rk.nu) {
            case 1: fused_rank_eval_multi<1, G, G>(rk, Gt, nVar, part, nValid); break;
            case 2: fused_rank_eval_multi<2, G, G>(rk, Gt, nVar, part, nValid); break;
            case 3: fused_rank_eval_multi<3, G, G>(rk, Gt, nVar, part, nValid); break;
            case 4: fused_rank_eval_multi<4, G, (G > 2 ? 2 : G)>(rk, Gt, nVar, part, nValid); break;
            case 5: fused_rank_eval_multi<5, G, (G > 2 ? 2 : G)>(rk, Gt, nVar, part, nValid); break;
            case 6: fused_rank_eval<6, G>(rk, Gt, nVar, part, nValid); break;
            case 7: fused_rank_eval<7, G>(rk, Gt, nVar, part, nValid); break;
            default: fused_rank_eval<8, G>(rk, Gt, nVar, part, nValid); break;
        }
    }
}

// ---- selection from the chunk partials ----
struct FusedPostRank {
    const double* part;
    int* sel;
    double* sinrSel;
    double* sinrWb;
    int nCand, nu, n2, n11, n12, n13;
};
struct FusedPostDev {
    FusedPostRank rk[kMaxLayers];
    const int* sbChunk;      // [nSB+1] chunk ranges of the PMI subbands
    const int* cqiChunk;     // [nCqiSB+1]
    const double* sbW;       // [nSB] mean-of-means weight of an RE of the subband
    const double* cqiW;      // [nCqiSB]
    int nRanks, nChunks, nSB, nCqiSB;
};

constexpr int kSelThreads = 512;

// grid: x = UE, y = rank slot.  dynamic shared memory: max(groups * nCand, nSB * n2 * nu) doubles
__global__ void __launch_bounds__(kSelThreads) pmi_select_fused_kernel(const __grid_constant__ FusedPostDev pd) {
    extern __shared__ double tots[];   // [groups][nCand], later [nSB][n2][nu]
    __shared__ double bv[kSelThreads / 32];
    __shared__ int bi[kSelThreads / 32];
    __shared__ int best;
    FusedPostRank p = pd.rk[0];
#pragma unroll
    for (int q = 1; q < kMaxLayers; ++q)
        if (q < pd.nRanks && (int)blockIdx.y == q) p = pd.rk[q];
    const int nSB = pd.nSB, nu = p.nu, nCand = p.nCand, b = blockIdx.x;
    const size_t chunkStride = (size_t)nu * nCand;
    const double* __restrict__ part = p.part + (size_t)b * pd.nChunks * chunkStride;
    // totalSINR (dlPMISelect.m:444) = sum over every RE and layer, NaN skipped: the partials of candidate c are the rows
    // (chunk, layer) of a [nChunks*nu][nCand] array; thread group g adds rows g, g + groups, ... (16 independent loads in
    // flight), then the groups are added in order -- a fixed summation order
    int span = 32;
    while (span < nCand && span < kSelThreads) span <<= 1;
    const int groups = kSelThreads / span;
    const int rows = pd.nChunks * nu;
    {
        const int g = threadIdx.x / span;
        for (int c = threadIdx.x % span; c < nCand; c += span) {
            double tot = 0.0;
            int r = g;
            for (; r + 15 * groups < rows; r += 16 * groups) {   // 16 loads in flight: a 384-candidate rank has one row group only
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = __ldcs(part + (size_t)(r + u * groups) * nCand + c);
#pragma unroll
                for (int u = 0; u < 16; ++u) tot += (v[u] == v[u]) ? v[u] : 0.0;
            }
            for (; r + 7 * groups < rows; r += 8 * groups) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldcs(part + (size_t)(r + u * groups) * nCand + c);
#pragma unroll
                for (int u = 0; u < 8; ++u) tot += (v[u] == v[u]) ? v[u] : 0.0;
            }
            for (; r < rows; r += groups) {
                const double v = __ldcs(part + (size_t)r * nCand + c);
                tot += (v == v) ? v : 0.0;
            }
            tots[g * nCand + c] = tot;
        }
    }
    __syncthreads();
    double v = -INFINITY;
    int ix = -1;
    for (int c = threadIdx.x; c < nCand; c += kSelThreads) {
        double tot = 0.0;
        for (int g = 0; g < groups; ++g) tot += tots[g * nCand + c];
        const double t = round4(tot);  // dlPMISelect.m:449
        if (ix < 0 || t > v) {
            v = t;
            ix = c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
        if (oi >= 0 && (ix < 0 || ov > v || (ov == v && oi < ix))) {
            v = ov;
            ix = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        bv[threadIdx.x >> 5] = v;
        bi[threadIdx.x >> 5] = ix;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = -INFINITY;
        int mi = -1;
        for (int w = 0; w < kSelThreads / 32; ++w)
            if (bi[w] >= 0 && (mi < 0 || bv[w] > m || (bv[w] == m && bi[w] < mi))) {
                m = bv[w];
                mi = bi[w];
            }
        best = mi;  // first linear index of the maximum (find(...,1), dlPMISelect.m:453)
    }
    __syncthreads();
    const int lin = best;
    const int i2wb = lin % p.n2, i11 = (lin / p.n2) % p.n11, i12 = (lin / (p.n2 * p.n11)) % p.n12,
              i13 = lin / (p.n2 * p.n11 * p.n12);
    int* __restrict__ sel = p.sel + (size_t)b * (4 + nSB);
    if (threadIdx.x == 0) {
        sel[0] = i2wb;
        sel[1] = i11;
        sel[2] = i12;
        sel[3] = i13;
    }
    const int base1 = p.n2 * (i11 + p.n11 * (i12 + p.n12 * i13));
    // subband means at the chosen i1 (dlPMISelect.m:471-501): mean-of-means = weight * plain sum (uniform weights)
    auto sub_at = [&](const int* range, double w, int sb, int l, int cand) -> double {
        double acc = 0.0;
        bool any = false;
        for (int ch = range[sb]; ch < range[sb + 1]; ++ch) {
            const double x = part[ch * chunkStride + (size_t)l * nCand + cand];
            if (x == x) {
                acc += x;
                any = true;
            }
        }
        return any ? w * acc : NAN;
    };
    double* subv = tots;   // [nSB][n2][nu]; the totals are dead (every thread is past the barrier above)
    for (int i = threadIdx.x; i < nSB * p.n2 * nu; i += kSelThreads) {
        const int l = i % nu, i2 = (i / nu) % p.n2, sb = i / (nu * p.n2);
        subv[i] = sub_at(pd.sbChunk, pd.sbW[sb], sb, l, base1 + i2);
    }
    __syncthreads();
    for (int sb = threadIdx.x; sb < nSB; sb += kSelThreads) {
        int pick = -1;
        if (pd.sbChunk[sb + 1] > pd.sbChunk[sb]) {  // CSI-RS present in the subband
            double bestT = -INFINITY;
            for (int i2 = 0; i2 < p.n2; ++i2) {
                double acc = 0.0;
                for (int l = 0; l < nu; ++l) {
                    const double x = subv[(sb * p.n2 + i2) * nu + l];
                    if (x == x) acc += x;  // sum(...,2,'omitnan')  (dlPMISelect.m:492)
                }
                const double t = round4(acc);
                if (pick < 0 || t > bestT) {  // [~,i2] = max(...) -> first maximum (dlPMISelect.m:496)
                    bestT = t;
                    pick = i2;
                }
            }
        }
        sel[4 + sb] = pick;
        for (int l = 0; l < nu; ++l)
            p.sinrSel[((size_t)b * nSB + sb) * nu + l] = pick >= 0 ? subv[(sb * p.n2 + pick) * nu + l] : NAN;
    }
    __syncthreads();
    // CQI-subband SINR with one wideband i2 (cqiSelect.m:586-596 -> getSubbandSINR :768-800)
    const int i2first = sel[4];
    for (int idx = threadIdx.x; idx < pd.nCqiSB * nu; idx += kSelThreads) {
        const int l = idx % nu, cs = idx / nu;
        double acc = NAN;
        if (i2first >= 0 && pd.cqiChunk[cs + 1] > pd.cqiChunk[cs]) acc = sub_at(pd.cqiChunk, pd.cqiW[cs], cs, l, base1 + i2first);
        p.sinrWb[((size_t)b * pd.nCqiSB + cs) * nu + l] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <class T>
static int upload_vec(Ctx* ctx, T** dst, const std::vector<T>& v) {
    cudaFree(*dst);
    *dst = nullptr;
    const size_t n = v.empty() ? 1 : v.size();
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)dst, sizeof(T) * n));
    if (!v.empty()) ISAC_CUDA_CHECK(ctx, cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return kOk;
}

// chunk list for G REs per chunk: chunks never straddle a PMI- or CQI-subband boundary
int pmi_plan_prepare_fused(PmiPlan* p, int G) {
    if (p->fG == G && p->d_part) return kOk;
    Ctx* ctx = p->ctx;
    ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    const int nRE = (int)p->reK.size();
    std::vector<char> cut(nRE + 1, 0);
    for (int s : p->sbStartH) cut[s] = 1;
    for (int s : p->cqiStartH) cut[s] = 1;
    std::vector<int> re0, cn;
    for (int e = 0; e < nRE;) {
        int n = 1;
        while (n < G && e + n < nRE && !cut[e + n]) ++n;
        re0.push_back(e);
        cn.push_back(n);
        e += n;
    }
    const int nChunks = (int)re0.size();
    auto ranges = [&](const std::vector<int>& start) {
        std::vector<int> r(start.size(), nChunks);
        size_t q = 0;
        for (int ch = 0; ch <= nChunks; ++ch) {
            const int e = ch < nChunks ? re0[ch] : nRE;
            while (q < start.size() && start[q] <= e) r[q++] = ch;
        }
        return r;
    };
    const std::vector<int> sbChunk = ranges(p->sbStartH), cqiChunk = ranges(p->cqiStartH);
    std::vector<double> sbW(p->nSB, 0.0), cqW(p->nCqiSB, 0.0);
    for (int sb = 0; sb < p->nSB; ++sb)
        if (p->sbStartH[sb + 1] > p->sbStartH[sb]) sbW[sb] = p->wH[p->sbStartH[sb]];
    for (int cs = 0; cs < p->nCqiSB; ++cs)
        if (p->cqiStartH[cs + 1] > p->cqiStartH[cs]) cqW[cs] = p->cwH[p->cqiStartH[cs]];
    int s;
    if ((s = upload_vec(ctx, &p->d_chunkRe0, re0)) || (s = upload_vec(ctx, &p->d_chunkN, cn)) ||
        (s = upload_vec(ctx, &p->d_sbChunk, sbChunk)) || (s = upload_vec(ctx, &p->d_cqiChunk, cqiChunk)) ||
        (s = upload_vec(ctx, &p->d_sbW, sbW)) || (s = upload_vec(ctx, &p->d_cqiSbW, cqW)))
        return s;
    cudaFree(p->d_part);
    p->d_part = nullptr;
    const size_t n = (size_t)p->maxBatch * (nChunks ? nChunks : 1) * p->nLayers * p->tab.nCand();
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&p->d_part, sizeof(double) * n));
    {   // index words of the candidate stage as slot numbers: slot(id, e) = (id * G + k(id)) ^ e with k = the swizzle of swz<G>
        std::vector<uint16_t> ef(p->entH.size());
        for (size_t i = 0; i < ef.size(); ++i) {
            const uint32_t id = p->entH[i] & 0x7fffu;
            const uint32_t k = G == 4 ? ((id >> 1) & 3u) : (G == 2 ? ((id >> 2) & 1u) : 0u);
            ef[i] = (uint16_t)((id * (uint32_t)G + k) | (p->entH[i] & 0x8000u));
        }
        if ((s = upload_vec(ctx, &p->d_entF, ef))) return s;
    }
    p->fG = G;
    p->nChunks = nChunks;
    return kOk;
}

static size_t fused_smem_bytes(const PmiShared* sh, int R, int G) {
    const size_t nCP = sh->cpTerms.size() / (sh->cpT ? sh->cpT : 1), nBf = (size_t)R * sh->NB * sh->nBeams;
    return sizeof(double2) * (G * ((size_t)R * sh->P + sh->pairs.size() + (nCP > nBf ? nCP : nBf)) + sh->pal.size());
}

// largest G in {4, 2, 1} whose tables fit in shared memory (0: none); *threads / *minb: the launch shape that goes with it
int pmi_fused_pick(const PmiShared* sh, int R, int* threads, int* minb) {
    static const int forceG = getenv("ISAC_PAIR_G") ? atoi(getenv("ISAC_PAIR_G")) : 0;
    static const int forceT = getenv("ISAC_PAIR_T") ? atoi(getenv("ISAC_PAIR_T")) : 0;
    const size_t cap = 227 * 1024 - 1024;
    const int order[3] = {4, 2, 1};
    for (int G : order) {
        if (forceG && G != forceG) continue;
        const size_t b = fused_smem_bytes(sh, R, G);
        if (b > cap) continue;
        if ((sh->cpTerms.size() / (sh->cpT ? sh->cpT : 1)) * (size_t)G > 0x8000u) continue;   // slot numbers are 15-bit
        int T = G == 4 ? 384 : 128, mb = G == 4 ? 1 : (G == 2 ? 2 : 3);
        if (G == 4 && forceT == 256) T = 256;
        if (G == 2 && forceT == 256) { T = 256; mb = 1; }
        while (mb > 1 && (b + 1024) * mb > 227 * 1024) --mb;
        *threads = T;
        *minb = mb;
        return G;
    }
    return 0;
}

template <int G, int T, int MINB>
static cudaError_t launch_fused(const FusedDev& d, int batch, size_t smem, cudaStream_t st) {
    auto k = pmi_pair_fused_kernel<G, T, MINB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(d.nChunks, batch);
    k<<<grid, T, smem, st>>>(d);
    return cudaGetLastError();
}

// SINR search + selection of the plans `grp` (all share the dictionary `sh`, the RE list and both subband partitions)
int pmi_fused_run(PmiPlan* const* grp, int n, const float2* H, const double* nv, int batch, cudaStream_t st) {
    PmiPlan* p = grp[0];
    Ctx* ctx = p->ctx;
    PmiShared* sh = p->sh;
    int T = 0, minb = 0;
    const int G = pmi_fused_pick(sh, p->cfg.nRx, &T, &minb);
    if (!G) { set_error(ctx, "pmi_fused_run: dictionary does not fit"); return kErrCapacity; }
    int s = pair_sync_dict(ctx, sh, st);
    if (s) return s;
    for (int i = 0; i < n; ++i)
        if ((s = pmi_plan_prepare_fused(grp[i], G))) return s;
    FusedDev d{};
    d.H = H; d.beams = p->d_beams; d.pairs = sh->d_pairs; d.pal = sh->d_pal; d.cpTerms = sh->d_cpTerms;
    d.reK = p->d_reK; d.reL = p->d_reL; d.chunkRe0 = p->d_chunkRe0; d.chunkN = p->d_chunkN;
    d.K = p->cfg.K; d.L = p->cfg.L; d.R = p->cfg.nRx; d.P = sh->P; d.NB = sh->NB; d.Pb = sh->Pb; d.nBeams = sh->nBeams;
    d.nChunks = p->nChunks; d.nPairs = (int)sh->pairs.size(); d.nPal = (int)sh->pal.size();
    d.cpT = sh->cpT; d.nCP = (int)(sh->cpTerms.size() / sh->cpT);
    for (int b = 0; b < batch; ++b) d.nVar[b] = nv[b];
    FusedPostDev pd{};
    pd.sbChunk = p->d_sbChunk; pd.cqiChunk = p->d_cqiChunk; pd.sbW = p->d_sbW; pd.cqiW = p->d_cqiSbW;
    pd.nChunks = p->nChunks; pd.nSB = p->nSB; pd.nCqiSB = p->nCqiSB;
    int maxCand = 1;
    for (int i = 0; i < n; ++i) {
        PmiPlan* q = grp[i];
        FusedRank& rk = d.rk[d.nRanks++];
        rk.ent = q->d_entF; rk.valid = q->d_valid; rk.invScale2 = q->d_invScale2; rk.part = q->d_part; rk.invS2 = q->invS2;
        rk.nCand = q->tab.nCand(); rk.nu = q->nLayers; rk.ntPad = q->ntPad;
        FusedPostRank& pr = pd.rk[pd.nRanks++];
        const CodebookTable& t = q->tab;
        pr.part = q->d_part; pr.sel = q->d_sel; pr.sinrSel = q->d_sinrSel; pr.sinrWb = q->d_sinrWb;
        pr.nCand = t.nCand(); pr.nu = q->nLayers; pr.n2 = t.n2; pr.n11 = t.n11; pr.n12 = t.n12; pr.n13 = t.n13;
        maxCand = std::max(maxCand, pr.nCand);
    }
    static const bool dbg = getenv("ISAC_PAIR_DEBUG") != nullptr;
    if (dbg) {
        static bool once = false;
        if (!once) {
            once = true;
            fprintf(stderr, "[isac] fused report: G=%d T=%d atoms=%d pairs=%d colPairs=%d cpT=%d palette=%d chunks=%d smem=%zu ranks=%d\n", G, T,
                    d.NB * d.nBeams, d.nPairs, d.nCP, d.cpT, d.nPal, d.nChunks, fused_smem_bytes(sh, d.R, G), d.nRanks);
        }
    }
    const size_t smem = fused_smem_bytes(sh, d.R, G);
    if (smem > 227 * 1024 - 1024) { set_error(ctx, "pmi_fused_run: tables exceed shared memory"); return kErrCapacity; }
    cudaError_t e;
    if (G == 4) e = T == 384 ? launch_fused<4, 384, 1>(d, batch, smem, st) : launch_fused<4, 256, 1>(d, batch, smem, st);
    else if (G == 2) e = T == 256 ? launch_fused<2, 256, 1>(d, batch, smem, st) : launch_fused<2, 128, 2>(d, batch, smem, st);
    else e = launch_fused<1, 128, 3>(d, batch, smem, st);
    (void)minb;
    ISAC_CUDA_CHECK(ctx, e);
    size_t selSmem = sizeof(double) * (size_t)std::max(kSelThreads, maxCand);   // groups * nCand <= max(kSelThreads, nCand)
    for (int i = 0; i < n; ++i)
        selSmem = std::max(selSmem, sizeof(double) * (size_t)p->nSB * grp[i]->tab.n2 * grp[i]->nLayers);
    if (selSmem > 48 * 1024) cudaFuncSetAttribute(pmi_select_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)selSmem);
    dim3 g3(batch, pd.nRanks);
    pmi_select_fused_kernel<<<g3, kSelThreads, selSmem, st>>>(pd);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    count_launches(ctx, 2);
    return kOk;
}

}  // namespace isac
