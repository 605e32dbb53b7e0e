// K8 + K9 (+K12, K10): per-RE LMMSE SINR of every Type-I precoder candidate, wideband / subband PMI
// selection, UL TPMI selection and PRG precoding.
//
// Replaces the interpreted loop nest of communication.phyLayer.dlPMISelect (dlPMISelect.m:385-428 calling
// getPrecodedSINR :1825-1834, selection :444-501), pmiSelect (pmiSelect.m:44-59, precodedSINR.m:11-18,
// sinrPerSubband.m:26-34) and prgPrecode (prgPrecode.m:103-144).
//
// K8 (the dense H*W contraction): every Type-I precoder column is [c0 v ; c1 v (; c2 v ; c3 v)] with v a DFT
// beam, so H*W for ALL candidates follows from the beam responses  B[blk][beam] = H[:, block blk] * v_beam,
// computed once per RE into shared memory (R x P/2 x nBeams complex MACs; P/2 <= 16 is far too short a
// contraction to feed tcgen05 — K = 2..16 would idle the tensor pipe — so it runs on the FP64 pipe with the
// rest of the SINR arithmetic).  K9: per candidate A = (HW)'(HW) + nVar I in registers (float64: A squares the
// conditioning of HW, float32 cannot hold 1e-5 here), Cholesky, diag(A^-1), sinr_l = 1/(nVar [A^-1]_ll) - 1.
#include "comm.cuh"
#include "ctx.cuh"
#include "sinr_core.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace isac {

struct PmiDev {
    const float2* H;
    const double2* beams;
    const int* layerBeam;
    const double2* layerCoef;
    const double* candScale;
    const uint8_t* valid;
    const int* reK;
    const int* reL;
    double* S;
    int K, L, R, P, NB, Pb, nBeams, nCand, nRE;
    double scale;
    double nVar[kMaxPmiBatch];  // per UE, already clipped at 1e-10 (dlPMISelect.m:846-848)
};

template <int NU>
__global__ void __launch_bounds__(128)
pmi_sinr_kernel(const PmiDev p) {
    extern __shared__ double2 sm[];
    double2* Hs = sm;                  // [R][P]
    double2* Bf = sm + p.R * p.P;      // [NB][nBeams][R]
    const int e = blockIdx.x, b = blockIdx.y;
    const int R = p.R, P = p.P;
    const long long kk = p.reK[e] - 1, ll = p.reL[e] - 1;
    const float2* __restrict__ Hb = p.H + (long long)b * p.K * p.L * R * P;
    for (int i = threadIdx.x; i < R * P; i += blockDim.x) {
        const int r = i % R, pp = i / R;
        const float2 h = __ldg(Hb + kk + p.K * (ll + (long long)p.L * (r + (long long)R * pp)));
        Hs[r * P + pp] = make_double2((double)h.x, (double)h.y);
    }
    __syncthreads();
    const int nB = p.NB * p.nBeams * R;
    // Bf[(r*NB + blk)*nBeams + beam]: for a fixed (r, blk) the lanes of a warp (consecutive candidates = consecutive
    // beams in the host-chosen storage order) read consecutive 16-byte words -> no shared-memory bank conflicts
    for (int i = threadIdx.x; i < nB; i += blockDim.x) {
        const int bm = i % p.nBeams, blk = (i / p.nBeams) % p.NB, r = i / (p.nBeams * p.NB);
        double2 acc = make_double2(0.0, 0.0);
        for (int q = 0; q < p.Pb; ++q) acc = zfma(acc, Hs[r * P + blk * p.Pb + q], p.beams[bm * p.Pb + q]);
        Bf[i] = acc;
    }
    __syncthreads();
    const double nVar = p.nVar[b];
    double* __restrict__ Sout = p.S + ((long long)b * p.nRE + e) * NU * (long long)p.nCand;
    for (int c = threadIdx.x; c < p.nCand; c += blockDim.x) {
        if (!p.valid[c]) {
#pragma unroll
            for (int j = 0; j < NU; ++j) Sout[(long long)j * p.nCand + c] = NAN;  // restricted precoder (dlPMISelect.m:418)
            continue;
        }
        const double sc = p.candScale ? p.candScale[c] : p.scale;
        constexpr int NT = NU * (NU + 1) / 2;
        double2 A[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) A[i] = make_double2(0.0, 0.0);
        int beamOf[NU];
        double2 cf[NU][2];
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            beamOf[j] = p.layerBeam[c * NU + j];
            cf[j][0] = p.layerCoef[(c * NU + j) * p.NB];
            cf[j][1] = p.NB > 1 ? p.layerCoef[(c * NU + j) * p.NB + 1] : make_double2(0.0, 0.0);
        }
        for (int r = 0; r < R; ++r) {
            double2 g[NU];
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                const double2* __restrict__ Br = Bf + (size_t)r * p.NB * p.nBeams + beamOf[j];
                double2 acc = zmul(cf[j][0], Br[0]);
                if (p.NB > 1) acc = zfma(acc, cf[j][1], Br[p.nBeams]);
                for (int blk = 2; blk < p.NB; ++blk)
                    acc = zfma(acc, p.layerCoef[(c * NU + j) * p.NB + blk], Br[blk * p.nBeams]);
                g[j] = make_double2(acc.x * sc, acc.y * sc);
            }
#pragma unroll
            for (int i = 0; i < NU; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) A[TRI(i, j)] = zfmac(A[TRI(i, j)], g[j], g[i]);  // conj(g_i) g_j
        }
        chol_sinr<NU>(A, nVar, Sout + c, p.nCand);
    }
}

// ------------------------------------------------------------------------------------------
// K9': Gram-pair form.  Every precoder column is g = sum_blk c_blk (e_blk (x) v_beam): a (beam, co-phasing) pair, of
// which a Type-I codebook has only a few hundred distinct ones ("columns") however many candidates and ranks it spans.
// The entries of (HW)'(HW) are inner products <H g_i, H g_j> of those columns, shared massively between candidates
// (2.2 k distinct column pairs against 1920 candidates x up to 36 entries for the 8-port (2,2) panel, ranks 1-8).
// One CTA per (RE, UE):
//   1. beam responses      Bf[a]   = H[:, blk] v_beam            (atoms a = (blk, beam))
//   2. atom Gram pairs     Gam[p]  = <Bf[a], Bf[a']>             (the distinct pairs the column pairs need)
//   3. column-pair table   G[q]    = sum_t pal[.] Gam[.]         (NB^2 terms: conj(c_i,blk) c_j,blk' products)
//   4. per rank, per candidate: A_ij = G[ent_ij] (a table look-up), Cholesky of A + (nVar/scale^2) I, SINR.
// The codebook scale is folded into the noise term: sinr = 1/(nVar' [(A' + nVar' I)^-1]_ll) - 1 with A' = A/scale^2,
// nVar' = nVar/scale^2, so the table is shared by all ranks.
// ------------------------------------------------------------------------------------------
struct PairRank {
    const uint16_t* ent;     // [nCand][ntPad]: column-pair index of each packed lower-triangle entry
    const uint8_t* valid;
    const double* invScale2; // per candidate 1/scale^2 (explicit codebooks) or nullptr
    double* S;
    double invS2;            // 1/scale^2 of the rank
    int nCand, nu, ntPad;
};
struct PairDev {
    const float2* H;
    const double2* beams;
    const uint32_t* pairs;     // [nPairs] atom a | atom a' << 16
    const double2* pal;        // [nPal]
    const uint32_t* cpTerms;   // [cpT][nCP]: pair index | palette index << 16
    const int* reK;
    const int* reL;
    int K, L, R, P, NB, Pb, nBeams, nRE, nPairs, nPal, nCP, cpT, nRanks;
    PairRank rk[kMaxLayers];
    double nVar[kMaxPmiBatch];
};

template <int NU>
__device__ __noinline__ void pair_rank_eval(const PairRank rk, const double2* __restrict__ G, double nVar, double* __restrict__ Sout) {
    constexpr int NT = NU * (NU + 1) / 2;
    constexpr int NW = (NT + 7) / 8;   // 16-byte words of indices per candidate
    for (int c = threadIdx.x; c < rk.nCand; c += blockDim.x) {
        const uint4* __restrict__ ep = reinterpret_cast<const uint4*>(rk.ent + (size_t)c * rk.ntPad);
        uint4 ev[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) ev[w] = __ldg(ep + w);   // issued together with the validity flag
        if (!rk.valid[c]) {
#pragma unroll
            for (int j = 0; j < NU; ++j) Sout[(long long)j * rk.nCand + c] = NAN;  // restricted precoder (dlPMISelect.m:418)
            continue;
        }
        double2 A[NT];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint4 v = ev[w];
            const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = w * 8 + u;
                if (e < NT) {   // bit 15: the pair is stored as (j,i) -> conjugate
                    const uint32_t id = (q[u >> 1] >> ((u & 1) * 16)) & 0xffffu;
                    const double2 g = G[id & 0x7fffu];
                    A[e] = make_double2(g.x, (id & 0x8000u) ? -g.y : g.y);
                }
            }
        }
        const double nv = nVar * (rk.invScale2 ? rk.invScale2[c] : rk.invS2);
        chol_sinr<NU>(A, nv, Sout + c, rk.nCand);
    }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
pmi_pair_kernel(const __grid_constant__ PairDev p) {
    extern __shared__ double2 sm[];
    const int R = p.R, P = p.P, nAtoms = p.NB * p.nBeams;
    double2* Hs = sm;                       // [R][P]
    double2* pal = Hs + R * P;              // [nPal]
    double2* Gm = pal + p.nPal;             // [nPairs]
    double2* Bf = Gm + p.nPairs;            // [R][NB][nBeams]; dead after step 2, the column-pair table G reuses it
    double2* G = Bf;                        // [nCP]
    const int e = blockIdx.x, b = blockIdx.y;
    const long long kk = p.reK[e] - 1, ll = p.reL[e] - 1;
    const float2* __restrict__ Hb = p.H + (long long)b * p.K * p.L * R * P;
    for (int i = threadIdx.x; i < R * P; i += blockDim.x) {
        const int r = i % R, pp = i / R;
        const float2 h = __ldg(Hb + kk + p.K * (ll + (long long)p.L * (r + (long long)R * pp)));
        Hs[r * P + pp] = make_double2((double)h.x, (double)h.y);
    }
    for (int i = threadIdx.x; i < p.nPal; i += blockDim.x) pal[i] = p.pal[i];
    __syncthreads();
    for (int blk = 0; blk < p.NB; ++blk)
        for (int bm = threadIdx.x; bm < p.nBeams; bm += blockDim.x) {
            const double2* __restrict__ bv = p.beams + (size_t)bm * p.Pb;
            for (int r = 0; r < R; ++r) {
                const double2* __restrict__ hr = Hs + r * P + blk * p.Pb;
                double2 acc = make_double2(0.0, 0.0);
                for (int q = 0; q < p.Pb; ++q) acc = zfma(acc, hr[q], __ldg(bv + q));
                Bf[r * nAtoms + blk * p.nBeams + bm] = acc;
            }
        }
    __syncthreads();
    for (int i = threadIdx.x; i < p.nPairs; i += blockDim.x) {
        const uint32_t w = __ldg(p.pairs + i);
        const double2* __restrict__ pa = Bf + (w & 0xffffu);
        const double2* __restrict__ pb = Bf + (w >> 16);
        double2 acc = make_double2(0.0, 0.0);
        for (int r = 0; r < R; ++r, pa += nAtoms, pb += nAtoms) acc = zfmac(acc, *pb, *pa);  // conj(Bf[a]) Bf[a']
        Gm[i] = acc;
    }
    __syncthreads();
    const uint4* __restrict__ cpt = reinterpret_cast<const uint4*>(p.cpTerms);
    for (int i = threadIdx.x; i < p.nCP; i += blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
        for (int t = 0; t < p.cpT / 4; ++t) {   // [cpT/4][nCP] words of 4 terms: pair index (bit 15: conjugate) | palette << 16
            const uint4 v = __ldg(cpt + (size_t)t * p.nCP + i);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double2 g = Gm[w4[u] & 0x7fffu];
                acc = zfma(acc, make_double2(g.x, (w4[u] & 0x8000u) ? -g.y : g.y), pal[w4[u] >> 16]);
            }
        }
        G[i] = acc;
    }
    __syncthreads();
    const double nVar = p.nVar[b];
#pragma unroll
    for (int q = 0; q < kMaxLayers; ++q) {   // static index: the descriptors stay in the parameter bank
        if (q >= p.nRanks) break;
        const PairRank rk = p.rk[q];
        double* __restrict__ Sout = rk.S + ((long long)b * p.nRE + e) * rk.nu * (long long)rk.nCand;
        switch (rk.nu) {
            case 1: pair_rank_eval<1>(rk, G, nVar, Sout); break;
            case 2: pair_rank_eval<2>(rk, G, nVar, Sout); break;
            case 3: pair_rank_eval<3>(rk, G, nVar, Sout); break;
            case 4: pair_rank_eval<4>(rk, G, nVar, Sout); break;
            case 5: pair_rank_eval<5>(rk, G, nVar, Sout); break;
            case 6: pair_rank_eval<6>(rk, G, nVar, Sout); break;
            case 7: pair_rank_eval<7>(rk, G, nVar, Sout); break;
            default: pair_rank_eval<8>(rk, G, nVar, Sout); break;
        }
    }
}

// ---- selection kernels, all ranks of a report in one launch each ----
struct PostRank {
    const double* S;
    double* sub;
    double* psum;
    int* sel;
    double* sinrSel;
    double* sinrWb;
    int nCand, nu, n2, n11, n12, n13, chunk0;
};
struct PostDev {
    PostRank rk[kMaxLayers];
    const int* sbStart;
    const int* cqiStart;
    const double* reW;
    const double* cqiW;
    int nRanks, nRE, nSB, nCqiSB;
};

__device__ __forceinline__ PostRank pick_rank(const PostDev& p, int key, bool byChunk) {
    PostRank rk = p.rk[0];
#pragma unroll
    for (int q = 1; q < kMaxLayers; ++q)   // static indices: the descriptors stay in the parameter bank
        if (q < p.nRanks && (byChunk ? key >= p.rk[q].chunk0 : key == q)) rk = p.rk[q];
    return rk;
}

// subband means (dlPMISelect.m:481): REs are sorted by subband; weight = mean-of-means weight.
// grid: x = 128-candidate chunks of all ranks, y = subband, z = UE
__global__ void __launch_bounds__(128) pmi_subband_kernel(const __grid_constant__ PostDev p) {
    const PostRank rk = pick_rank(p, blockIdx.x, true);
    const int c = (blockIdx.x - rk.chunk0) * blockDim.x + threadIdx.x;
    const int sb = blockIdx.y, b = blockIdx.z;
    if (c >= rk.nCand) return;
    const int nCand = rk.nCand, nu = rk.nu;
    const int e0 = p.sbStart[sb], e1 = p.sbStart[sb + 1];
    for (int l = 0; l < nu; ++l) {
        const double* __restrict__ s = rk.S + ((long long)b * p.nRE * nu + l) * nCand + c;
        double acc = 0.0, plain = 0.0;
        bool any = false;
        // 16 independent loads in flight per thread (one 16-PRB subband = 16 CSI-RS REs); the sums keep the RE order
        for (int eb = e0; eb < e1; eb += 16) {
            double v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (eb + j < e1) ? __ldcs(s + (long long)(eb + j) * nu * nCand) : NAN;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (!isnan(v[j])) {
                    acc += p.reW[min(eb + j, e1 - 1)] * v[j];
                    plain += v[j];   // un-weighted partial of sum(SINRPerRE,[1 2 3],'omitnan') (dlPMISelect.m:444)
                    any = true;
                }
        }
        const long long o = (((long long)b * p.nSB + sb) * nu + l) * nCand + c;
        rk.sub[o] = any ? acc : NAN;
        rk.psum[o] = plain;
    }
}

// grid: x = UE, y = rank slot
__global__ void __launch_bounds__(256) pmi_select_kernel(const __grid_constant__ PostDev pd) {
    __shared__ double bv[8];
    __shared__ int bi[8];
    __shared__ int best;
    const PostRank p = pick_rank(pd, blockIdx.y, false);
    const int nSB = pd.nSB;
    const int b = blockIdx.x;
    const double* __restrict__ ps = p.psum + (long long)b * nSB * p.nu * p.nCand;
    double v = -INFINITY;
    int ix = -1;
    for (int c = threadIdx.x; c < p.nCand; c += blockDim.x) {
        double tot = 0.0;  // totalSINR: fixed summation order (subband-major, then layer)
        for (int q = 0; q < nSB * p.nu; ++q) tot += ps[(long long)q * p.nCand + c];
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
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
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
    int* __restrict__ sel = p.sel + (long long)b * (4 + nSB);
    if (threadIdx.x == 0) {
        sel[0] = i2wb;
        sel[1] = i11;
        sel[2] = i12;
        sel[3] = i13;
    }
    const long long base1 = (long long)p.n2 * (i11 + (long long)p.n11 * (i12 + (long long)p.n12 * i13));
    for (int sb = threadIdx.x; sb < nSB; sb += blockDim.x) {
        int pick = -1;
        if (pd.sbStart[sb + 1] > pd.sbStart[sb]) {  // CSI-RS present in the subband
            double bestT = -INFINITY;
            for (int i2 = 0; i2 < p.n2; ++i2) {
                double acc = 0.0;
                for (int l = 0; l < p.nu; ++l) {
                    const double x = p.sub[(((long long)b * nSB + sb) * p.nu + l) * p.nCand + base1 + i2];
                    if (!isnan(x)) acc += x;  // sum(...,2,'omitnan')  (dlPMISelect.m:492)
                }
                const double t = round4(acc);
                if (pick < 0 || t > bestT) {  // [~,i2] = max(...) -> first maximum (dlPMISelect.m:496)
                    bestT = t;
                    pick = i2;
                }
            }
        }
        sel[4 + sb] = pick;
        for (int l = 0; l < p.nu; ++l)
            p.sinrSel[((long long)b * nSB + sb) * p.nu + l] =
                pick >= 0 ? p.sub[(((long long)b * nSB + sb) * p.nu + l) * p.nCand + base1 + pick] : NAN;
    }
    __syncthreads();
    // CQI-subband SINR with one wideband i2 (cqiSelect.m:586-596 -> getSubbandSINR :768-800)
    const int i2first = sel[4];
    for (int idx = threadIdx.x; idx < pd.nCqiSB * p.nu; idx += blockDim.x) {
        const int l = idx % p.nu, cs = idx / p.nu;
        double acc = NAN;
        if (i2first >= 0 && pd.cqiStart[cs + 1] > pd.cqiStart[cs]) {
            acc = 0.0;
            for (int e = pd.cqiStart[cs]; e < pd.cqiStart[cs + 1]; ++e)
                acc += pd.cqiW[e] * p.S[(((long long)b * pd.nRE + e) * p.nu + l) * p.nCand + base1 + i2first];
        }
        p.sinrWb[((long long)b * pd.nCqiSB + cs) * p.nu + l] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
template <class T>
static int upload(Ctx* ctx, T** dst, const std::vector<T>& v) {
    const size_t n = v.empty() ? 1 : v.size();
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)dst, sizeof(T) * n));
    if (!v.empty()) ISAC_CUDA_CHECK(ctx, cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return kOk;
}

// sort REs by (k,l), assign subbands and mean-of-means weights (mean over k per symbol, then over symbols)
static void partition_res(const std::vector<int>& k, const std::vector<int>& l, const std::vector<int>& sizes,
                          std::vector<int>& start, std::vector<double>& w) {
    const int nSB = (int)sizes.size(), n = (int)k.size();
    start.assign(nSB + 1, 0);
    w.assign(n, 0.0);
    int prb0 = 0, e = 0;
    for (int sb = 0; sb < nSB; ++sb) {
        const int lo = prb0 * 12 + 1, hi = (prb0 + sizes[sb]) * 12;
        start[sb] = e;
        while (e < n && k[e] >= lo && k[e] <= hi) ++e;
        prb0 += sizes[sb];
        // weights
        std::vector<int> syms;
        for (int i = start[sb]; i < e; ++i)
            if (std::find(syms.begin(), syms.end(), l[i]) == syms.end()) syms.push_back(l[i]);
        for (int i = start[sb]; i < e; ++i) {
            int cnt = 0;
            for (int q = start[sb]; q < e; ++q) cnt += (l[q] == l[i]);
            w[i] = 1.0 / ((double)cnt * (double)syms.size());
        }
    }
    start[nSB] = e;
}

int pmi_plan_create(Ctx* ctx, const CsiConfig& cin, int nLayers, int maxBatch, PmiPlan** out, PmiShared* share, bool fused) {
    if (maxBatch < 1 || cin.nRx < 1 || cin.nPorts < 1 || cin.K < 12 || cin.L < 1 || cin.nRE < 0) {
        set_error(ctx, "pmi_plan_create: invalid configuration");
        return kErrInvalidArg;
    }
    if (nLayers > cin.nRx || nLayers > cin.nPorts) {
        set_error(ctx, "nr5g:hDLPMISelect:InvalidNumLayers");
        return kErrInvalidArg;
    }
    PmiPlan* p = new PmiPlan();
    p->ctx = ctx;
    p->cfg = cin;
    p->nLayers = nLayers;
    p->maxBatch = maxBatch;
    p->fused = fused;
    if (cin.subsetRestriction) {
        const int n = cin.nPorts > 2 ? cin.N1 * cin.O1 * cin.N2 * cin.O2 : 6;
        p->csr.assign(cin.subsetRestriction, cin.subsetRestriction + n);
        p->cfg.subsetRestriction = p->csr.data();
    }
    if (cin.i2Restriction) {
        p->i2r.assign(cin.i2Restriction, cin.i2Restriction + 16);
        p->cfg.i2Restriction = p->i2r.data();
    }
    // RE list sorted by (k, l) (only the summation order depends on it)
    std::vector<int> order(cin.nRE);
    for (int i = 0; i < cin.nRE; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        return cin.reK[a] != cin.reK[b] ? cin.reK[a] < cin.reK[b] : cin.reL[a] < cin.reL[b];
    });
    for (int i : order) {
        if (cin.reK[i] < 1 || cin.reK[i] > cin.nSizeBWP * 12 || cin.reL[i] < 1 || cin.reL[i] > cin.L) continue;  // :352-353
        p->reK.push_back(cin.reK[i]);
        p->reL.push_back(cin.reL[i]);
    }
    p->cfg.nRE = (int)p->reK.size();
    p->cfg.reK = p->reK.data();
    p->cfg.reL = p->reL.data();
    const bool multiPanel = p->cfg.nPanels >= 2;
    int s = multiPanel ? build_type1mp_table(ctx, p->cfg, nLayers, p->tab) : build_type1sp_table(ctx, p->cfg, nLayers, kVariantUE, p->tab);
    if (s) { delete p; return s; }
    if (multiPanel) p->direct = true;   // the Gram-pair dictionary is sized for the single-panel column structure
    subband_info(p->cfg.pmiSubband != 0, p->cfg.nStartBWP, p->cfg.nSizeBWP, p->cfg.subbandSize, p->sbSizes);
    subband_info(p->cfg.cqiSubband != 0, p->cfg.nStartBWP, p->cfg.nSizeBWP, p->cfg.subbandSize, p->cqiSbSizes);
    p->nSB = (int)p->sbSizes.size();
    p->nCqiSB = (int)p->cqiSbSizes.size();
    std::vector<int> sbStart, cqiStart;
    std::vector<double> w, cw;
    partition_res(p->reK, p->reL, p->sbSizes, sbStart, w);
    partition_res(p->reK, p->reL, p->cqiSbSizes, cqiStart, cw);
    p->sbStartH = sbStart; p->cqiStartH = cqiStart; p->wH = w; p->cwH = cw;
    p->uniformW = true;
    for (int i = 0; i < p->nSB; ++i)
        for (int e = sbStart[i]; e < sbStart[i + 1]; ++e) p->uniformW = p->uniformW && w[e] == w[sbStart[i]];
    for (int i = 0; i < p->nCqiSB; ++i)
        for (int e = cqiStart[i]; e < cqiStart[i + 1]; ++e) p->uniformW = p->uniformW && cw[e] == cw[cqiStart[i]];
    p->sbHasRE.resize(p->nSB);
    for (int i = 0; i < p->nSB; ++i) p->sbHasRE[i] = sbStart[i + 1] > sbStart[i];
    p->cqiSbHasRE.resize(p->nCqiSB);
    for (int i = 0; i < p->nCqiSB; ++i) p->cqiSbHasRE[i] = cqiStart[i + 1] > cqiStart[i];
    const CodebookTable& t = p->tab;
    const int nCand = t.nCand(), nu = nLayers;
    // storage order of the beams in shared memory: beam (l, m) = l*L2 + m is stored at m*L1 + l so that candidates that
    // are consecutive in the index order (i2 fastest, then i11 = l) touch consecutive shared-memory words
    const int L2 = t.nBeams > 1 ? p->cfg.N2 * p->cfg.O2 : 1, L1 = t.nBeams / L2;
    auto store_idx = [&](int beam) { return t.nBeams > 1 ? (beam % L2) * L1 + beam / L2 : 0; };
    std::vector<double2> beams(t.beams.size()), coef((size_t)nCand * nu * t.NB);
    for (int bm = 0; bm < t.nBeams; ++bm)
        for (int q = 0; q < t.Pb; ++q) {
            const std::complex<double> v = t.beams[(size_t)bm * t.Pb + q];
            beams[(size_t)store_idx(bm) * t.Pb + q] = make_double2(v.real(), v.imag());
        }
    std::vector<int> lb((size_t)nCand * nu);
    for (int c = 0; c < nCand; ++c)
        for (int j = 0; j < nu; ++j) {
            const LayerDesc& d = t.layers[(size_t)c * nu + j];
            lb[(size_t)c * nu + j] = store_idx(d.beam);
            for (int b = 0; b < t.NB; ++b) coef[((size_t)c * nu + j) * t.NB + b] = make_double2(d.coef[b].real(), d.coef[b].imag());
        }
    // Gram-pair form (K9'): register this rank's columns / column pairs in the (shared) dictionary
    {
        PmiShared* sh = share;
        auto compatible = [&](const PmiShared* q) {
            if (q->beams.empty()) return true;
            if (q->NB != t.NB || q->Pb != t.Pb || q->nBeams != t.nBeams || q->P != t.P || q->beams.size() != beams.size()) return false;
            for (size_t i = 0; i < beams.size(); ++i)
                if (q->beams[i].x != beams[i].x || q->beams[i].y != beams[i].y) return false;
            return true;
        };
        if (!sh || !compatible(sh)) sh = new PmiShared();
        if (sh->beams.empty()) {
            sh->NB = t.NB; sh->Pb = t.Pb; sh->nBeams = t.nBeams; sh->P = t.P;
            sh->beams = beams;
            sh->pal.push_back(make_double2(0.0, 0.0));   // palette 0 = 0, pair 0 = (0,0): the "absent term"
            sh->pairs.push_back(0u);
            sh->pairIdx[0u] = 0;
            sh->cpT = (t.NB * t.NB + 3) / 4 * 4;
        }
        ++sh->refs;
        p->sh = sh;
        auto qkey = [](std::complex<double> q) {
            return std::pair<long long, long long>(std::llround(q.real() * 1099511627776.0), std::llround(q.imag() * 1099511627776.0));
        };
        auto column_of = [&](const LayerDesc& d) {   // distinct (beam, co-phasing coefficients)
            std::vector<long long> key{(long long)store_idx(d.beam)};
            for (int b = 0; b < t.NB; ++b) { auto k = qkey(d.coef[b]); key.push_back(k.first); key.push_back(k.second); }
            auto it = sh->colIdx.find(key);
            if (it != sh->colIdx.end()) return it->second;
            const int id = (int)sh->cols.size();
            PmiShared::Column col;
            col.beam = store_idx(d.beam);
            for (int b = 0; b < t.NB; ++b) col.coef[b] = d.coef[b];
            sh->cols.push_back(col);
            sh->colIdx[key] = id;
            return id;
        };
        // G = <g_ci, g_cj> = sum conj(c_i,blk) c_j,blk' Gamma[(blk,b_i),(blk',b_j)], stored once per unordered pair
        auto colpair_impl = [&](int ci, int cj) -> int {
            const unsigned long long key = ((unsigned long long)ci << 32) | (unsigned)cj;
            auto it = sh->cpIdx.find(key);
            if (it != sh->cpIdx.end()) return it->second;
            const int id = (int)sh->cpTerms.size() / sh->cpT;
            const PmiShared::Column &a = sh->cols[ci], &bcol = sh->cols[cj];
            std::vector<uint32_t> v;
            for (int bi = 0; bi < t.NB; ++bi)
                for (int bj = 0; bj < t.NB; ++bj) {
                    const std::complex<double> q = std::conj(a.coef[bi]) * bcol.coef[bj];
                    if (std::abs(q) < 1e-300) continue;
                    uint32_t a1 = (uint32_t)(bi * t.nBeams + a.beam), a2 = (uint32_t)(bj * t.nBeams + bcol.beam);
                    uint32_t cj = 0;                       // Gamma[a2,a1] = conj(Gamma[a1,a2]): keep one orientation
                    if (a1 > a2) { std::swap(a1, a2); cj = 0x8000u; }
                    const uint32_t pk = a1 | (a2 << 16);
                    auto pit = sh->pairIdx.find(pk);
                    int pi;
                    if (pit == sh->pairIdx.end()) {
                        pi = (int)sh->pairs.size();
                        sh->pairs.push_back(pk);
                        sh->pairIdx[pk] = pi;
                    } else pi = pit->second;
                    auto qit = sh->palIdx.find(qkey(q));
                    int qi;
                    if (qit == sh->palIdx.end()) {
                        qi = (int)sh->pal.size();
                        sh->pal.push_back(make_double2(q.real(), q.imag()));
                        sh->palIdx[qkey(q)] = qi;
                    } else qi = qit->second;
                    if (pi > 0x7fff || qi > 0xffff) { sh->ok = false; pi = qi = 0; }
                    v.push_back((uint32_t)pi | cj | ((uint32_t)qi << 16));
                }
            v.resize(sh->cpT, 0u);
            sh->cpTerms.insert(sh->cpTerms.end(), v.begin(), v.end());
            sh->cpIdx[key] = id;
            if (id > 0x7fff || t.NB * t.nBeams > 0xffff) sh->ok = false;
            return id;
        };
        auto colpair_of = [&](int ci, int cj) -> int {    // <g_cj, g_ci> = conj(<g_ci, g_cj>): bit 15 of the entry
            return ci > cj ? (colpair_impl(cj, ci) | 0x8000) : colpair_impl(ci, cj);
        };
        const int NT = nu * (nu + 1) / 2, ntPad = (NT + 7) / 8 * 8;
        p->ntPad = ntPad;
        std::vector<uint16_t> ent((size_t)nCand * ntPad, 0);
        std::vector<double> is2;
        if (!t.candScale.empty()) is2.assign(nCand, 1.0);
        p->invS2 = 1.0 / (t.scale * t.scale);
        for (int c = 0; c < nCand; ++c) {
            if (!t.valid[c]) continue;
            if (!t.candScale.empty()) is2[c] = 1.0 / (t.candScale[c] * t.candScale[c]);
            int colId[kMaxLayers];
            for (int i = 0; i < nu; ++i) colId[i] = column_of(t.layers[(size_t)c * nu + i]);
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j <= i; ++j)
                    ent[(size_t)c * ntPad + i * (i + 1) / 2 + j] = (uint16_t)(colpair_of(colId[i], colId[j]) & 0xffff);
        }
        p->entH = ent;
        if ((s = upload(ctx, &p->d_ent, ent)) || (!is2.empty() && (s = upload(ctx, &p->d_invScale2, is2)))) {
            pmi_plan_destroy(p);
            return s;
        }
    }
    PmiPlan* ex = p;
#define UP(dst, vec)                                   \
    if ((s = upload(ctx, &(dst), vec))) {              \
        pmi_plan_destroy(p);                           \
        return s;                                      \
    }
    UP(p->d_beams, beams);
    UP(p->d_layerBeam, lb);
    UP(p->d_layerCoef, coef);
    UP(p->d_valid, t.valid);
    UP(p->d_reK, p->reK);
    UP(p->d_reL, p->reL);
    UP(p->d_reW, w);
    UP(p->d_reCqiW, cw);
    UP(ex->d_sbStart, sbStart);
    UP(ex->d_cqiStart, cqiStart);
#undef UP
    const size_t B = maxBatch, nRE = p->reK.size() ? p->reK.size() : 1;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 8); };
    if (!fused) {   // fused report plans allocate these on the first launch that needs them (pmi_plan_legacy_buffers)
        A((void**)&p->d_S, sizeof(double) * nCand * nu * nRE * B);
        A((void**)&p->d_total, sizeof(double) * nCand * nu * p->nSB * B);  // plain per-subband sums
        A((void**)&p->d_sub, sizeof(double) * nCand * nu * p->nSB * B);
    }
    // selection results of the plan: one arena (sel | sinrSel | sinrWb, laid out for maxBatch) -> one D2H copy
    p->resOff[0] = 0;
    p->resOff[1] = (sizeof(int) * (4 + p->nSB) * B + 15) / 16 * 16;
    p->resOff[2] = p->resOff[1] + (sizeof(double) * nu * p->nSB * B + 15) / 16 * 16;
    p->resBytes = p->resOff[2] + (sizeof(double) * nu * p->nCqiSB * B + 15) / 16 * 16;
    A((void**)&p->d_res, p->resBytes);
    p->ownsRes = true;
    if (e == cudaSuccess) pmi_plan_use_arena(p, p->d_res, nullptr);
    A((void**)&ex->d_nVar, sizeof(double) * B);
    if (e != cudaSuccess) {
        set_error(ctx, std::string("pmi_plan_create: cudaMalloc: ") + cudaGetErrorString(e));
        pmi_plan_destroy(p);
        return kErrCuda;
    }
    *out = p;
    return kOk;
}

// SINRPerRE / per-subband arrays of the per-RE kernels, for plans that were created without them
static int pmi_plan_legacy_buffers(PmiPlan* p) {
    if (p->d_S) return kOk;
    Ctx* ctx = p->ctx;
    const size_t nCand = p->tab.nCand(), nu = p->nLayers, nRE = p->reK.size() ? p->reK.size() : 1, B = p->maxBatch;
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&p->d_S, sizeof(double) * nCand * nu * nRE * B));
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&p->d_total, sizeof(double) * nCand * nu * p->nSB * B));
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&p->d_sub, sizeof(double) * nCand * nu * p->nSB * B));
    return kOk;
}

// point the plan's result arrays into `dev` (resBytes bytes) and its host staging area to `host` (may be null)
void pmi_plan_use_arena(PmiPlan* p, char* dev, char* host) {
    if (p->ownsRes && dev != p->d_res) {
        cudaFree(p->d_res);
        p->ownsRes = false;
    }
    p->d_res = dev;
    p->d_sel = (int*)(dev + p->resOff[0]);
    p->d_sinrSel = (double*)(dev + p->resOff[1]);
    p->d_sinrWb = (double*)(dev + p->resOff[2]);
    p->hostRes = host;
}

void pmi_plan_destroy(PmiPlan* p) {
    if (!p) return;
    cudaFree(p->d_sbStart); cudaFree(p->d_cqiStart); cudaFree(p->d_nVar);
    if (p->pin) cudaFreeHost(p->pin);
    cudaFree(p->d_beams); cudaFree(p->d_layerBeam); cudaFree(p->d_layerCoef); cudaFree(p->d_candScale);
    cudaFree(p->d_valid); cudaFree(p->d_reK); cudaFree(p->d_reL); cudaFree(p->d_reSb); cudaFree(p->d_reW);
    cudaFree(p->d_reCqiSb); cudaFree(p->d_reCqiW); cudaFree(p->d_S); cudaFree(p->d_total); cudaFree(p->d_sub);
    if (p->ownsRes) cudaFree(p->d_res);
    cudaFree(p->d_ent); cudaFree(p->d_entF); cudaFree(p->d_invScale2);
    cudaFree(p->d_chunkRe0); cudaFree(p->d_chunkN); cudaFree(p->d_sbChunk); cudaFree(p->d_cqiChunk);
    cudaFree(p->d_sbW); cudaFree(p->d_cqiSbW); cudaFree(p->d_part);
    if (p->sh && --p->sh->refs == 0) {
        cudaFree(p->sh->d_pairs);
        cudaFree(p->sh->d_pal);
        cudaFree(p->sh->d_cpTerms);
        delete p->sh;
    }
    delete p;
}

template <int NU>
static cudaError_t launch_sinr(const PmiDev& d, int batch, cudaStream_t st) {
    const size_t smem = sizeof(double2) * ((size_t)d.R * d.P + (size_t)d.NB * d.nBeams * d.R);
    auto k = pmi_sinr_kernel<NU>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(d.nRE, batch);
    k<<<grid, 128, smem, st>>>(d);
    return cudaGetLastError();
}

static int pmi_direct_launch(PmiPlan* p, const float2* H, const double* nv, int batch, cudaStream_t st) {
    const CodebookTable& t = p->tab;
    PmiDev d{};
    d.H = H; d.beams = p->d_beams; d.layerBeam = p->d_layerBeam; d.layerCoef = p->d_layerCoef;
    d.candScale = nullptr; d.valid = p->d_valid; d.reK = p->d_reK; d.reL = p->d_reL; d.S = p->d_S;
    for (int b = 0; b < batch; ++b) d.nVar[b] = nv[b];
    d.K = p->cfg.K; d.L = p->cfg.L; d.R = p->cfg.nRx; d.P = t.P; d.NB = t.NB; d.Pb = t.Pb; d.nBeams = t.nBeams;
    d.nCand = t.nCand(); d.nRE = (int)p->reK.size(); d.scale = t.scale;
    cudaError_t e;
    switch (p->nLayers) {
        case 1: e = launch_sinr<1>(d, batch, st); break;
        case 2: e = launch_sinr<2>(d, batch, st); break;
        case 3: e = launch_sinr<3>(d, batch, st); break;
        case 4: e = launch_sinr<4>(d, batch, st); break;
        case 5: e = launch_sinr<5>(d, batch, st); break;
        case 6: e = launch_sinr<6>(d, batch, st); break;
        case 7: e = launch_sinr<7>(d, batch, st); break;
        default: e = launch_sinr<8>(d, batch, st); break;
    }
    ISAC_CUDA_CHECK(p->ctx, e);
    count_launches(p->ctx, 1);
    return kOk;
}

static size_t pair_smem_bytes(const PmiShared* sh, int R) {
    const size_t nCP = sh->cpTerms.size() / (sh->cpT ? sh->cpT : 1), nBf = (size_t)R * sh->NB * sh->nBeams;
    return sizeof(double2) * ((size_t)R * sh->P + sh->pal.size() + sh->pairs.size() + (nCP > nBf ? nCP : nBf));
}

// (re-)upload the dictionary when ranks were added since the last launch
int pair_sync_dict(Ctx* ctx, PmiShared* sh, cudaStream_t st) {
    if (sh->upPairs == sh->pairs.size() && sh->upPal == sh->pal.size() && sh->upCp == sh->cpTerms.size()) return kOk;
    ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    cudaFree(sh->d_pairs);
    cudaFree(sh->d_pal);
    cudaFree(sh->d_cpTerms);
    sh->d_pairs = nullptr; sh->d_pal = nullptr; sh->d_cpTerms = nullptr;
    const size_t nCP = sh->cpTerms.size() / sh->cpT;
    std::vector<uint32_t> tt(sh->cpTerms.size());   // [cpT/4][nCP][4]: one 16-byte word of 4 terms per column pair, coalesced
    for (size_t q = 0; q < nCP; ++q)
        for (int t = 0; t < sh->cpT; ++t) tt[((size_t)(t / 4) * nCP + q) * 4 + (t % 4)] = sh->cpTerms[q * sh->cpT + t];
    int s;
    if ((s = upload(ctx, &sh->d_pairs, sh->pairs))) return s;
    if ((s = upload(ctx, &sh->d_pal, sh->pal))) return s;
    if ((s = upload(ctx, &sh->d_cpTerms, tt))) return s;
    sh->upPairs = sh->pairs.size();
    sh->upPal = sh->pal.size();
    sh->upCp = sh->cpTerms.size();
    return kOk;
}

int pmi_select_run_multi(PmiPlan* const* plans, int n, const float2* H, const double* nVar, int batch, cudaStream_t st) {
    if (n < 1 || !plans || !plans[0]) return kErrInvalidArg;
    Ctx* ctx = plans[0]->ctx;
    if (batch < 1 || !H || !nVar) {
        set_error(ctx, "pmi_select_run: invalid argument");
        return kErrInvalidArg;
    }
    if (batch > kMaxPmiBatch) {
        set_error(ctx, "pmi_select_run: batch exceeds kMaxPmiBatch");
        return kErrCapacity;
    }
    double nv[kMaxPmiBatch];
    for (int b = 0; b < batch; ++b) {
        if (!(nVar[b] >= 0.0) || !std::isfinite(nVar[b])) {
            set_error(ctx, "dlPMISelect: NVAR must be real, nonnegative and finite");
            return kErrInvalidArg;
        }
        nv[b] = nVar[b] < 1e-10 ? 1e-10 : nVar[b];  // dlPMISelect.m:846-848
    }
    std::vector<PmiPlan*> live;
    for (int i = 0; i < n; ++i) {
        PmiPlan* p = plans[i];
        if (batch > p->maxBatch) {
            set_error(ctx, "pmi_select_run: invalid argument");
            return kErrInvalidArg;
        }
        if (!p->reK.empty()) live.push_back(p);  // else everything NaN (dlPMISelect.m:362-379), decided at collect time
    }
    if (live.empty()) return kOk;
    const int pr = prof_begin(ctx, kProfPmi, st);
    std::vector<char> done(live.size(), 0);
    // report plans: SINR search with the subband accumulation fused in + selection from the chunk partials (comm_fused.cu)
    for (size_t i = 0; i < live.size(); ++i) {
        PmiPlan* p = live[i];
        PmiShared* sh = p->sh;
        int T, mb;
        if (done[i] || !p->fused || !p->uniformW || p->direct || !sh || !sh->ok || !pmi_fused_pick(sh, p->cfg.nRx, &T, &mb)) continue;
        std::vector<PmiPlan*> grp;
        for (size_t j = i; j < live.size(); ++j) {
            PmiPlan* q = live[j];
            if (done[j] || !q->fused || !q->uniformW || q->direct || q->sh != sh || q->reK != p->reK || q->reL != p->reL ||
                q->sbSizes != p->sbSizes || q->cqiSbSizes != p->cqiSbSizes)
                continue;
            grp.push_back(q);
            done[j] = 2;
        }
        int s = pmi_fused_run(grp.data(), (int)grp.size(), H, nv, batch, st);
        if (s) return s;
    }
    for (size_t i = 0; i < live.size(); ++i)
        if (done[i] != 2) {
            int s = pmi_plan_legacy_buffers(live[i]);
            if (s) return s;
        }
    // SINR of every candidate: one fused launch per dictionary, the direct kernel for the rest
    for (size_t i = 0; i < live.size(); ++i) {
        if (done[i]) continue;
        PmiPlan* p = live[i];
        PmiShared* sh = p->sh;
        const bool pairOk = sh && sh->ok && !p->direct && pair_smem_bytes(sh, p->cfg.nRx) <= 200 * 1024;
        if (!pairOk) {
            int s = pmi_direct_launch(p, H, nv, batch, st);
            if (s) return s;
            done[i] = 1;
            continue;
        }
        int s = pair_sync_dict(ctx, sh, st);
        if (s) return s;
        PairDev d{};
        d.H = H; d.beams = p->d_beams; d.pairs = sh->d_pairs; d.pal = sh->d_pal; d.reK = p->d_reK; d.reL = p->d_reL;
        d.K = p->cfg.K; d.L = p->cfg.L; d.R = p->cfg.nRx; d.P = sh->P; d.NB = sh->NB; d.Pb = sh->Pb; d.nBeams = sh->nBeams;
        d.nRE = (int)p->reK.size(); d.nPairs = (int)sh->pairs.size(); d.nPal = (int)sh->pal.size();
        d.cpTerms = sh->d_cpTerms; d.cpT = sh->cpT; d.nCP = (int)(sh->cpTerms.size() / sh->cpT);
        for (int b = 0; b < batch; ++b) d.nVar[b] = nv[b];
        for (size_t j = i; j < live.size(); ++j) {
            PmiPlan* q = live[j];
            if (done[j] || q->sh != sh || q->direct || q->reK != p->reK) continue;
            PairRank& rk = d.rk[d.nRanks++];
            rk.ent = q->d_ent; rk.valid = q->d_valid; rk.invScale2 = q->d_invScale2; rk.S = q->d_S; rk.invS2 = q->invS2;
            rk.nCand = q->tab.nCand(); rk.nu = q->nLayers; rk.ntPad = q->ntPad;
            done[j] = 1;
        }
        const size_t smem = pair_smem_bytes(sh, d.R);
        static const int minb = getenv("ISAC_PAIR_MINB") ? atoi(getenv("ISAC_PAIR_MINB")) : 3;
        dim3 grid(d.nRE, batch);
        if (minb == 4) {
            cudaFuncSetAttribute(pmi_pair_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            pmi_pair_kernel<4><<<grid, 128, smem, st>>>(d);
        } else {
            cudaFuncSetAttribute(pmi_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            pmi_pair_kernel<3><<<grid, 128, smem, st>>>(d);
        }
        ISAC_CUDA_CHECK(ctx, cudaGetLastError());
        count_launches(ctx, 1);
    }
    // subband means + selection: one launch each for the plans that share the RE partition (all ranks of a CSI plan)
    for (size_t i = 0; i < live.size(); ++i) done[i] = done[i] == 2 ? 2 : 0;
    for (size_t i = 0; i < live.size(); ++i) {
        if (done[i]) continue;
        PmiPlan* p = live[i];
        PostDev pd{};
        pd.sbStart = p->d_sbStart; pd.cqiStart = p->d_cqiStart; pd.reW = p->d_reW; pd.cqiW = p->d_reCqiW;
        pd.nRE = (int)p->reK.size(); pd.nSB = p->nSB; pd.nCqiSB = p->nCqiSB;
        int chunks = 0;
        for (size_t j = i; j < live.size(); ++j) {
            PmiPlan* q = live[j];
            if (done[j] || q->reK != p->reK || q->reL != p->reL || q->nSB != p->nSB || q->nCqiSB != p->nCqiSB || q->sbSizes != p->sbSizes ||
                q->cqiSbSizes != p->cqiSbSizes)
                continue;
            const CodebookTable& t = q->tab;
            PostRank& rk = pd.rk[pd.nRanks++];
            rk.S = q->d_S; rk.sub = q->d_sub; rk.psum = q->d_total; rk.sel = q->d_sel; rk.sinrSel = q->d_sinrSel; rk.sinrWb = q->d_sinrWb;
            rk.nCand = t.nCand(); rk.nu = q->nLayers; rk.n2 = t.n2; rk.n11 = t.n11; rk.n12 = t.n12; rk.n13 = t.n13;
            rk.chunk0 = chunks;
            chunks += (rk.nCand + 127) / 128;
            done[j] = 1;
        }
        dim3 g2(chunks, p->nSB, batch);
        pmi_subband_kernel<<<g2, 128, 0, st>>>(pd);
        dim3 g3(batch, pd.nRanks);
        pmi_select_kernel<<<g3, 256, 0, st>>>(pd);
        count_launches(ctx, 2);
    }
    prof_end(ctx, pr, st);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

int pmi_select_run(PmiPlan* p, const float2* H, const double* nVar, int batch, cudaStream_t st) {
    return pmi_select_run_multi(&p, 1, H, nVar, batch, st);
}

static bool plan_all_nan(const PmiPlan* p) {
    bool anyValid = false;
    for (uint8_t v : p->tab.valid) anyValid |= (v != 0);
    return p->reK.empty() || !anyValid;  // dlPMISelect.m:362-379
}

// enqueue the D2H copy of the selection results into pinned host memory (no synchronisation)
int pmi_select_collect_enqueue(PmiPlan* p, int batch, cudaStream_t st) {
    Ctx* ctx = p->ctx;
    (void)batch;
    if (plan_all_nan(p)) return kOk;
    if (!p->ownsRes) {   // slice of a shared arena: copy just this plan's slice to its place in the owner's pinned buffer
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(p->hostRes, p->d_res, p->resBytes, cudaMemcpyDeviceToHost, st));
        return kOk;
    }
    if (p->pinBytes < p->resBytes) {
        if (p->pin) cudaFreeHost(p->pin);
        p->pin = nullptr;
        ISAC_CUDA_CHECK(ctx, cudaMallocHost(&p->pin, p->resBytes));
        p->pinBytes = p->resBytes;
    }
    p->hostRes = (char*)p->pin;
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(p->pin, p->d_res, p->resBytes, cudaMemcpyDeviceToHost, st));
    return kOk;
}

// parse the staged results; the caller has synchronised the stream after pmi_select_collect_enqueue
int pmi_select_collect_finish(PmiPlan* p, int batch, std::vector<PmiResult>& out) {
    const int nu = p->nLayers, nSB = p->nSB, nC = p->nCqiSB;
    out.resize(batch);   // elements (and their vectors' capacity) survive when the caller keeps `out` between reports
    for (auto& r : out) { r.allNaN = false; r.i1[0] = r.i1[1] = r.i1[2] = 0; }
    if (plan_all_nan(p)) {
        for (auto& r : out) {
            r.allNaN = true;
            r.i2.assign(nSB, NAN);
            r.sinrSel.assign((size_t)nSB * nu, NAN);
            r.sinrWbSel.assign((size_t)nC * nu, NAN);
        }
        return kOk;
    }
    const char* h = p->hostRes;
    const int* sel = (const int*)(h + p->resOff[0]);
    const double* ss = (const double*)(h + p->resOff[1]);
    const double* sw = (const double*)(h + p->resOff[2]);
    for (int b = 0; b < batch; ++b) {
        PmiResult& r = out[b];
        const int* s = sel + (size_t)b * (4 + nSB);
        r.i1[0] = s[1] + 1; r.i1[1] = s[2] + 1; r.i1[2] = s[3] + 1;
        r.i2.resize(nSB);
        for (int sb = 0; sb < nSB; ++sb) r.i2[sb] = s[4 + sb] >= 0 ? (double)(s[4 + sb] + 1) : NAN;
        r.sinrSel.resize((size_t)nSB * nu);
        for (int sb = 0; sb < nSB; ++sb)
            for (int l = 0; l < nu; ++l) r.sinrSel[(size_t)l * nSB + sb] = ss[((size_t)b * nSB + sb) * nu + l];  // [nSB x nu] col-major
        r.sinrWbSel.resize((size_t)nC * nu);
        for (int cs = 0; cs < nC; ++cs)
            for (int l = 0; l < nu; ++l) r.sinrWbSel[(size_t)l * nC + cs] = sw[((size_t)b * nC + cs) * nu + l];
    }
    return kOk;
}

int pmi_select_collect(PmiPlan* p, int batch, std::vector<PmiResult>& out) {
    Ctx* ctx = p->ctx;
    int s = pmi_select_collect_enqueue(p, batch, ctx->stream);
    if (s) return s;
    ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return pmi_select_collect_finish(p, batch, out);
}

int pmi_get_sinr_arrays(PmiPlan* p, int batch, double* sinrPerRE, double* sinrPerSubband) {
    Ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    const size_t nCand = p->tab.nCand(), nu = p->nLayers, nRE = p->reK.size();
    if (!p->d_S && nRE) {
        set_error(ctx, "SINRPerRE / SINRPerSubband are kept by dlPMISelect plans only (report plans never store them)");
        return kErrUnsupported;
    }
    // device layout [cand][layer][RE|SB][batch] -> host MATLAB layout [RE|SB x layer x cand x batch]
    auto fetch = [&](const double* dsrc, size_t n3, double* dst) -> int {
        std::vector<double> tmp(nCand * nu * n3 * batch);
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(tmp.data(), dsrc, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        for (size_t b = 0; b < (size_t)batch; ++b)
            for (size_t e = 0; e < n3; ++e)
                for (size_t l = 0; l < nu; ++l)
                    for (size_t c = 0; c < nCand; ++c)
                        dst[e + n3 * (l + nu * (c + nCand * b))] = tmp[c + nCand * (l + nu * (e + n3 * b))];
        return kOk;
    };
    int s = kOk;
    if (sinrPerRE && nRE) s = fetch(p->d_S, nRE, sinrPerRE);
    if (!s && sinrPerSubband && nRE) s = fetch(p->d_sub, (size_t)p->nSB, sinrPerSubband);
    return s;
}

// ------------------------------------------------------------------------------------------
// host tails: riSelect / cqiSelect
// ------------------------------------------------------------------------------------------
double ri_total_sinr(const PmiResult& r, int nSB, int rank) {
    if (r.allNaN) return NAN;
    double total = 0.0;
    for (int l = 0; l < rank; ++l) {
        double acc = 0.0;
        int cnt = 0;
        for (int sb = 0; sb < nSB; ++sb) {
            if (std::isnan(r.i2[sb])) continue;                               // riSelect.m:263
            const double v = r.sinrSel[(size_t)l * nSB + sb] * rank;         // :265
            if (!std::isnan(v)) { acc += v; ++cnt; }
        }
        const double mean = cnt ? acc / cnt : NAN;                            // :278
        if (mean >= 1.0) total += mean;                                       // :282
    }
    return total;
}

static double get_cqi(double lin, const double* table, int n) {  // cqiSelect.m:697-722
    if (std::isnan(lin)) return NAN;
    const double db = 10.0 * std::log10(lin);
    int last = -1;
    for (int i = 0; i < n; ++i)
        if (table[i] <= db) last = i;
    return last < 0 ? 0.0 : (double)(last + 1);
}

void cqi_from_pmi(const CsiConfig& cfg, int nu, const PmiResult& r, int nSB, int nCqiSB, const double* table, int tableLen,
                  CsiReport& rep) {
    const int nCW = (nu + 3) / 4;
    rep.nCW = nCW;
    std::vector<double> sbl((size_t)nCqiSB * nu, NAN);  // SINRperSubband [nCqiSB x nu]
    if (!r.allNaN) {
        if (!cfg.pmiSubband || nSB == 1) sbl = r.sinrWbSel;                   // cqiSelect.m:586-596
        else sbl = r.sinrSel;                                                 // :604-614 (PMI subbands == CQI subbands)
    }
    std::vector<double> cw((size_t)nCqiSB * nCW);
    for (int s = 0; s < nCqiSB; ++s) {                                        // :610-627
        bool nan = false;
        for (int l = 0; l < nu; ++l) nan |= std::isnan(sbl[(size_t)l * nCqiSB + s]);
        for (int c = 0; c < nCW; ++c) {
            if (nan) { cw[(size_t)c * nCqiSB + s] = NAN; continue; }
            int lo = 0, hi = nu;
            if (nCW == 2) { lo = c == 0 ? 0 : nu / 2; hi = c == 0 ? nu / 2 : nu; }  // nrLayerDemap (TS 38.211 Table 7.3.1.3-1)
            double a = 0.0;
            for (int l = lo; l < hi; ++l) a += sbl[(size_t)l * nCqiSB + s];
            cw[(size_t)c * nCqiSB + s] = a;
        }
    }
    std::vector<double> full;  // [rows x nCW]
    int rows = nCqiSB;
    if (nCqiSB > 1) {                                                         // :631-633 wideband row = omitnan mean
        rows = nCqiSB + 1;
        full.assign((size_t)rows * nCW, NAN);
        for (int c = 0; c < nCW; ++c) {
            double a = 0.0; int n = 0;
            for (int s = 0; s < nCqiSB; ++s) { const double v = cw[(size_t)c * nCqiSB + s]; if (!std::isnan(v)) { a += v; ++n; } }
            full[(size_t)c * rows] = n ? a / n : NAN;
            for (int s = 0; s < nCqiSB; ++s) full[(size_t)c * rows + 1 + s] = cw[(size_t)c * nCqiSB + s];
        }
    } else full = cw;
    if (r.allNaN) {                                                           // :636-650
        const int ns = nCqiSB == 1 ? 0 : nCqiSB;
        rep.nCqiRows = ns + 1;
        rep.cqi.assign((size_t)(ns + 1) * nCW, NAN);
        rep.sinrPerSubbandPerCW.assign((size_t)(ns + 1) * nCW, NAN);
        return;
    }
    std::vector<double> cqAll(full.size());
    for (size_t i = 0; i < full.size(); ++i) cqAll[i] = get_cqi(full[i], table, tableLen);   // :653
    rep.sinrPerSubbandPerCW = full;
    if (cfg.cqiSubband) {                                                     // :656-677
        rep.nCqiRows = rows;
        rep.cqi.assign((size_t)rows * nCW, NAN);
        for (int c = 0; c < nCW; ++c) {
            rep.cqi[(size_t)c * rows] = cqAll[(size_t)c * rows];
            for (int s = 1; s < rows; ++s) {
                const double d = cqAll[(size_t)c * rows + s] - cqAll[(size_t)c * rows];
                double off = NAN;
                if (d == 0) off = 0; else if (d == 1) off = 1; else if (d >= 2) off = 2; else if (d <= -1) off = 3;
                rep.cqi[(size_t)c * rows + s] = off;
            }
        }
    } else {
        rep.nCqiRows = 1;
        rep.cqi.resize(nCW);
        for (int c = 0; c < nCW; ++c) rep.cqi[c] = cqAll[(size_t)c * rows];
    }
}

// ------------------------------------------------------------------------------------------
// UL TPMI selection (pmiSelect.m:28-66)
// ------------------------------------------------------------------------------------------
template <int NU>
__global__ void __launch_bounds__(128)
ul_sinr_kernel(const float2* __restrict__ hestAll, int K, int nSym, int R, int P, const double2* __restrict__ W /*[P][NU][nT]*/,
               int nT, double nVar, double* __restrict__ sinrAll /*[K*nSym][nT][batch]*/) {
    const long long re = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (re >= (long long)K * nSym) return;
    const float2* __restrict__ hest = hestAll + (long long)blockIdx.y * K * nSym * R * P;
    double* __restrict__ sinr = sinrAll + (long long)blockIdx.y * K * nSym * nT;
    double2 h[16 * 4];  // R <= 16, P <= 4
    double sr = 0.0, si = 0.0;
    for (int r = 0; r < R; ++r)
        for (int p = 0; p < P; ++p) {
            const float2 v = __ldg(hest + re + (long long)K * nSym * (r + (long long)R * p));
            h[r * 4 + p] = make_double2((double)v.x, (double)v.y);
            sr += v.x;
            si += v.y;
        }
    const bool has = (sr != 0.0 || si != 0.0);  // find(sum(hest,3:4) ~= 0)  (pmiSelect.m:35)
    for (int t = 0; t < nT; ++t) {
        double out = 0.0;
        if (has) {
            double2 A[NU][NU];
#pragma unroll
            for (int i = 0; i < NU; ++i)
#pragma unroll
                for (int j = 0; j < NU; ++j) A[i][j] = make_double2(0.0, 0.0);
            for (int r = 0; r < R; ++r) {
                double2 g[NU];
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    double2 acc = make_double2(0.0, 0.0);
                    for (int p = 0; p < P; ++p) acc = zadd(acc, zmul(h[r * 4 + p], W[(t * NU + j) * P + p]));
                    g[j] = acc;
                }
#pragma unroll
                for (int i = 0; i < NU; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) A[i][j] = zadd(A[i][j], zmulc(g[j], g[i]));
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) A[i][i].x += nVar;
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                double d = A[j][j].x;
#pragma unroll
                for (int k = 0; k < j; ++k) d -= A[j][k].x * A[j][k].x + A[j][k].y * A[j][k].y;
                const double ljj = sqrt(d);
                A[j][j] = make_double2(ljj, 0.0);
                const double inv = 1.0 / ljj;
#pragma unroll
                for (int i = j + 1; i < NU; ++i) {
                    double2 s = A[i][j];
#pragma unroll
                    for (int k = 0; k < j; ++k) s = zsub(s, zmulc(A[i][k], A[j][k]));
                    A[i][j] = make_double2(s.x * inv, s.y * inv);
                }
            }
#pragma unroll
            for (int cc = 0; cc < NU; ++cc) {
                double2 x[NU];
                x[cc] = make_double2(1.0 / A[cc][cc].x, 0.0);
                double nrm = x[cc].x * x[cc].x;
#pragma unroll
                for (int i = cc + 1; i < NU; ++i) {
                    double2 s = make_double2(0.0, 0.0);
#pragma unroll
                    for (int k = cc; k < i; ++k) s = zadd(s, zmul(A[i][k], x[k]));
                    const double inv = -1.0 / A[i][i].x;
                    x[i] = make_double2(s.x * inv, s.y * inv);
                    nrm = fma(x[i].x, x[i].x, fma(x[i].y, x[i].y, nrm));
                }
                out += 1.0 / (nVar * nrm) - 1.0;  // sum over layers (precodedSINR.m:16)
            }
        }
        sinr[re * nT + t] = out;
    }
}

// sinrPerSubband (sinrPerSubband.m:26-34): one CTA per (band, tpmi)
__global__ void __launch_bounds__(256)
ul_band_kernel(const double* __restrict__ sinrAll, int K, int nSym, int nT, const int* __restrict__ bandLo,
               const int* __restrict__ bandHi, double* __restrict__ outAll /*[nSB][nT][batch]*/) {
    const int sb = blockIdx.x, t = blockIdx.y;
    const double* __restrict__ sinr = sinrAll + (long long)blockIdx.z * K * nSym * nT;
    double* __restrict__ out = outAll + (long long)blockIdx.z * gridDim.x * nT;
    double acc = 0.0;
    long long cnt = 0;
    const int lo = bandLo[sb] - 1, hi = bandHi[sb];  // 0-based [lo, hi)
    for (long long i = threadIdx.x; i < (long long)(hi - lo) * nSym; i += blockDim.x) {
        const int k = lo + (int)(i % (hi - lo)), l = (int)(i / (hi - lo));
        const double* __restrict__ row = sinr + ((long long)k + (long long)K * l) * nT;
        acc += row[t];
        double s = 0.0;
        for (int q = 0; q < nT; ++q) s += row[q];
        cnt += (s != 0.0);
    }
    __shared__ double ra[8];
    __shared__ long long rc[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = acc; rc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0; long long c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += ra[w]; c += rc[w]; }
        out[(long long)sb * nT + t] = a / (double)c;  // 0/0 -> NaN like MATLAB
    }
}

int ul_pmi_select_run(Ctx* ctx, int nu, const float2* hest, int K, int nSym, int R, int P, double noiseEst, int bandSize,
                      UlPmiResult& out, cudaStream_t st) {
    std::vector<UlPmiResult> v;
    int s = ul_pmi_select_batch(ctx, nu, hest, K, nSym, R, P, noiseEst, bandSize, 1, v, st);
    if (!s) out = v[0];
    return s;
}

// UL state of one context: the report whose kernels and result copy are enqueued but whose host tail has not run yet,
// and the uploaded PUSCH codebooks / band limits.  Owned by the context (Ctx::ulState) and freed with it, so that a
// context created later at the same address (MEX mexAtExit + re-create, another device) never inherits stale buffers.
struct UlCache { int nu, P, K, band; double2* dW; int* dIdx; };
struct UlPending {
    bool active = false, launched = false;
    int batch = 0, nSB = 0, nT = 0;
    UlPmiResult proto;
    double* hBands = nullptr;   // pinned [nSB x nT x batch]
    size_t hBytes = 0;
    cudaEvent_t ready = nullptr;
    std::vector<UlCache> cache;
};
static void ul_state_free(void* s) {
    UlPending* u = static_cast<UlPending*>(s);
    if (!u) return;
    if (u->hBands) cudaFreeHost(u->hBands);
    if (u->ready) cudaEventDestroy(u->ready);
    for (UlCache& e : u->cache) { cudaFree(e.dW); cudaFree(e.dIdx); }
    delete u;
}
static UlPending* ul_pending(Ctx* ctx) {
    if (!ctx->ulState) {
        ctx->ulState = new UlPending();
        ctx->ulFree = ul_state_free;
    }
    return static_cast<UlPending*>(ctx->ulState);
}

int ul_pmi_select_batch(Ctx* ctx, int nu, const float2* hest, int K, int nSym, int R, int P, double noiseEst, int bandSize,
                        int batch, std::vector<UlPmiResult>& outs, cudaStream_t st) {
    const int s = ul_pmi_select_batch_enqueue(ctx, nu, hest, K, nSym, R, P, noiseEst, bandSize, batch, st);
    return s ? s : ul_pmi_select_batch_finish(ctx, outs);
}

int ul_pmi_select_batch_enqueue(Ctx* ctx, int nu, const float2* hest, int K, int nSym, int R, int P, double noiseEst, int bandSize,
                                int batch, cudaStream_t st) {
    if (batch < 1) { set_error(ctx, "pmiSelect: batch < 1"); return kErrInvalidArg; }
    UlPending* pend = ul_pending(ctx);
    if (pend->active) { set_error(ctx, "pmiSelect: the previous enqueued report has not been finished"); return kErrInvalidArg; }
    UlPmiResult out;
    if (!hest || K < 12 || nSym < 1 || R < 1 || R > 16 || bandSize < 1) {
        set_error(ctx, "pmiSelect: invalid argument");
        return kErrInvalidArg;
    }
    CodebookTable t;
    int s = build_pusch_table(ctx, nu, P, t);
    if (s) return s;
    const int nT = t.n2;
    out = UlPmiResult();
    out.nTPMI = nT;
    const double nrb = K / 12.0, r = nrb / bandSize;
    const int nSB = (int)std::ceil(r);
    out.nSB = nSB;
    std::vector<int> lo(nSB), hi(nSB);
    for (int i = 0; i < nSB; ++i) {  // sinrPerSubband.m:20-21
        lo[i] = 12 * bandSize * i + 1;
        hi[i] = i < (int)std::floor(r) ? 12 * bandSize * (i + 1) : (int)(12 * bandSize * r);
    }
    out.subbandIndices.resize((size_t)nSB * 2);
    for (int i = 0; i < nSB; ++i) { out.subbandIndices[i] = lo[i]; out.subbandIndices[nSB + i] = hi[i]; }
    pend->batch = batch; pend->nSB = nSB; pend->nT = nT; pend->launched = false;
    if (noiseEst == 0.0) {  // pmiSelect.m:39
        out.none = true;
        pend->proto = out;
        pend->active = true;
        return kOk;
    }
    // the PUSCH codebook of (nu, P) and the band limits are uploaded once per context and cached
    std::vector<UlCache>& cache = pend->cache;
    UlCache* uc = nullptr;
    for (auto& e : cache)
        if (e.nu == nu && e.P == P && e.K == K && e.band == bandSize) uc = &e;
    if (!uc) {
        std::vector<std::complex<double>> Wc;
        materialize_codebook(t, Wc);  // [P][nu][nT]
        std::vector<double2> Wd(Wc.size());
        for (size_t i = 0; i < Wc.size(); ++i) Wd[i] = make_double2(Wc[i].real(), Wc[i].imag());
        UlCache e{nu, P, K, bandSize, nullptr, nullptr};
        ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&e.dW, sizeof(double2) * Wd.size()));
        ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&e.dIdx, sizeof(int) * 2 * nSB));
        ISAC_CUDA_CHECK(ctx, cudaMemcpy(e.dW, Wd.data(), sizeof(double2) * Wd.size(), cudaMemcpyHostToDevice));
        std::vector<int> lohi(lo);
        lohi.insert(lohi.end(), hi.begin(), hi.end());
        ISAC_CUDA_CHECK(ctx, cudaMemcpy(e.dIdx, lohi.data(), sizeof(int) * 2 * nSB, cudaMemcpyHostToDevice));
        cache.push_back(e);
        uc = &cache.back();
    }
    void *dW = uc->dW, *dIdx = uc->dIdx, *dS = nullptr, *dB = nullptr;
    if ((s = ctx_scratch(ctx, 3, sizeof(double) * (size_t)K * nSym * nT * batch, &dS))) return s;
    if ((s = ctx_scratch(ctx, 4, sizeof(double) * (size_t)nSB * nT * batch, &dB))) return s;
    const long long nre = (long long)K * nSym;
    const dim3 blocks((unsigned)((nre + 127) / 128), batch);
    const int pr = prof_begin(ctx, kProfUlPmi, st);
    switch (nu) {
        case 1: ul_sinr_kernel<1><<<blocks, 128, 0, st>>>(hest, K, nSym, R, P, (const double2*)dW, nT, noiseEst, (double*)dS); break;
        case 2: ul_sinr_kernel<2><<<blocks, 128, 0, st>>>(hest, K, nSym, R, P, (const double2*)dW, nT, noiseEst, (double*)dS); break;
        case 3: ul_sinr_kernel<3><<<blocks, 128, 0, st>>>(hest, K, nSym, R, P, (const double2*)dW, nT, noiseEst, (double*)dS); break;
        default: ul_sinr_kernel<4><<<blocks, 128, 0, st>>>(hest, K, nSym, R, P, (const double2*)dW, nT, noiseEst, (double*)dS); break;
    }
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    dim3 g(nSB, nT, batch);
    ul_band_kernel<<<g, 256, 0, st>>>((const double*)dS, K, nSym, nT, (const int*)dIdx, (const int*)dIdx + nSB, (double*)dB);
    prof_end(ctx, pr, st);
    count_launches(ctx, 2);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    const size_t bytes = sizeof(double) * (size_t)nSB * nT * batch;
    if (pend->hBytes < bytes) {
        if (pend->hBands) cudaFreeHost(pend->hBands);
        pend->hBands = nullptr;
        pend->hBytes = 0;
        ISAC_CUDA_CHECK(ctx, cudaMallocHost((void**)&pend->hBands, bytes));
        pend->hBytes = bytes;
    }
    if (!pend->ready) ISAC_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&pend->ready, cudaEventDisableTiming));
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(pend->hBands, dB, bytes, cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaEventRecord(pend->ready, st));
    pend->proto = out;
    pend->launched = true;
    pend->active = true;
    return kOk;
}

// Waits for the band SINRs of the enqueued report only (not for work enqueued behind it) and runs the host tail.
int ul_pmi_select_batch_finish(Ctx* ctx, std::vector<UlPmiResult>& outs) {
    UlPending* pend = ul_pending(ctx);
    if (!pend->active) { set_error(ctx, "pmiSelect: nothing enqueued"); return kErrInvalidArg; }
    pend->active = false;
    const int batch = pend->batch, nSB = pend->nSB, nT = pend->nT;
    outs.assign(batch, pend->proto);
    if (!pend->launched) return kOk;   // zero noise estimate: every report is "none" (pmiSelect.m:39,60-64)
    ISAC_CUDA_CHECK(ctx, cudaEventSynchronize(pend->ready));
    for (int b = 0; b < batch; ++b) {
        UlPmiResult& o = outs[b];
        const double* bands = pend->hBands + (size_t)b * nSB * nT;
        // "no channel estimates" <=> every band is 0/0 (pmiSelect.m:39,60-64)
        bool any = false;
        for (size_t i = 0; i < (size_t)nSB * nT; ++i) any |= !std::isnan(bands[i]);
        if (!any) { o.none = true; continue; }
        o.pmi.resize(nSB);
        o.sinr.resize((size_t)nSB * nT);
        for (int sb = 0; sb < nSB; ++sb) {
            int best = 0;
            for (int tt = 0; tt < nT; ++tt) {
                o.sinr[(size_t)tt * nSB + sb] = bands[(size_t)sb * nT + tt];
                if (bands[(size_t)sb * nT + tt] > bands[(size_t)sb * nT + best]) best = tt;  // first max (pmiSelect.m:56)
            }
            o.pmi[sb] = std::isnan(bands[(size_t)sb * nT]) ? NAN : (double)best;           // :57-58 (0-based)
        }
    }
    return kOk;
}

// ------------------------------------------------------------------------------------------
// PRG precoding (prgPrecode.m:53-144)
// ------------------------------------------------------------------------------------------
// One thread per RE, looping over the antenna ports.  The reference writes the layer symbols into a K x L x nLayers grid by
// linear index (portgrid(indin) = symin, prgPrecode.m:131), multiplies every RE by F(:,:,prg) (:134) and reads the result back at
// the RE positions of the first layer (:139-144).  nrPDSCHIndices-style input has the same RE positions in every layer
// column, so the symbol that lands at (position of row i, layer l) is symin(i,l): the kernel checks exactly that and
// otherwise falls back to scanning the index list for the target linear index (last write wins; absent -> 0), which
// reproduces the grid semantics for arbitrary indices without a scratch grid or a second pass.  The layer symbols, the index
// arithmetic (a 64-bit modulo) and the PRG number are evaluated once per RE and reused for all P ports; every store of a warp
// is one contiguous run along the RE axis.
template <int NU>   // layers, compile time: a run-time bound on the unrolled layer loops issues every predicated-off iteration
__global__ void __launch_bounds__(256)
prg_precode_kernel(const float2* __restrict__ sym, const int* __restrict__ ind, int NRE, long long plane, int K,
                   const float2* __restrict__ F, int P, int NPRG, int nStartGrid, int Pd, float2* __restrict__ antsym,
                   int* __restrict__ antind) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NRE) return;
    constexpr int nu = NU;
    // blockIdx.y: independent allocations of a batch (e.g. the cells of one slot), every array stacked along its last dim
    sym += (long long)blockIdx.y * NRE * nu;
    ind += (long long)blockIdx.y * NRE * nu;
    F += (long long)blockIdx.y * nu * P * NPRG;
    antsym += (long long)blockIdx.y * NRE * P;
    antind += (long long)blockIdx.y * NRE * P;
    const long long i0 = (long long)ind[i] - 1;
    const long long pos = (plane <= 0x7fffffffLL && i0 >= 0 && i0 <= 0xffffffffLL) ? (long long)((unsigned)i0 % (unsigned)plane)
                                                                                  : i0 % plane;  // RE position of the first layer's index
    const int k = (int)(pos % K);
    const int prg = (nStartGrid + k / 12) / Pd;             // getPRGSet (prgPrecode.m:94-100), 0-based
    float2 x[NU];
#pragma unroll
    for (int l = 0; l < NU; ++l) {
        x[l] = make_float2(0.f, 0.f);
        const long long target = pos + plane * l + 1;
        if ((long long)ind[i + (long long)NRE * l] == target) x[l] = sym[i + (long long)NRE * l];
        else
            for (long long q = 0; q < (long long)NRE * nu; ++q)
                if ((long long)ind[q] == target) x[l] = sym[q];
    }
    const float2* __restrict__ Fp = F + (long long)nu * P * prg;
    for (int p = 0; p < P; ++p) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = 0; l < NU; ++l) {
            const float2 f = __ldg(Fp + l + nu * p);
            acc.x += x[l].x * f.x - x[l].y * f.y;           // portgrid * F(:,:,prg) (prgPrecode.m:134)
            acc.y += x[l].x * f.y + x[l].y * f.x;
        }
        antsym[i + (long long)NRE * p] = acc;
        antind[i + (long long)NRE * p] = (int)(pos + 1 + plane * p);
    }
}

int prg_precode_run(Ctx* ctx, int K, int Lsym, int nStartGrid, const float2* portsym, const int* portind, int NRE, int nu,
                    const float2* F, int P, int NPRG, float2* antsym, int* antind, int batch, cudaStream_t st) {
    if (!portsym || !portind || !F || !antsym || !antind || K < 12 || K % 12 || Lsym < 1 || nu < 1 || P < 1 || NPRG < 1 || NRE < 0 ||
        batch < 1 || batch > 65535) {
        set_error(ctx, "prgPrecode: invalid argument");
        return kErrInvalidArg;
    }
    if (NRE == 0) return kOk;
    const long long plane = (long long)K * Lsym;
    const int nrb = K / 12;
    const int Pd = (nrb + nStartGrid + NPRG - 1) / NPRG;  // Pd_BWP = ceil((NRB+nstartgrid)/NPRG)
    const int pr = prof_begin(ctx, kProfPrecode, st);
    dim3 grid((unsigned)((NRE + 255) / 256), batch);
#define ISAC_PRG(N_) case N_: prg_precode_kernel<N_><<<grid, 256, 0, st>>>(portsym, portind, NRE, plane, K, F, P, NPRG, nStartGrid, Pd, antsym, antind); break;
    switch (nu) {
        ISAC_PRG(1) ISAC_PRG(2) ISAC_PRG(3) ISAC_PRG(4) ISAC_PRG(5) ISAC_PRG(6) ISAC_PRG(7) ISAC_PRG(8)
        default: set_error(ctx, "prgPrecode: more than 8 layers"); return kErrInvalidArg;
    }
#undef ISAC_PRG
    prof_end(ctx, pr, st);
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

// ------------------------------------------------------------------------------------------
// precodedSINR (precodedSINR.m:11-18): sum over the layers of the LMMSE SINR of one RE, for a batch of REs that
// share the precoder W.  float64 in and out (the reference passes doubles); one thread per RE.
// ------------------------------------------------------------------------------------------
template <int NU>
__global__ void __launch_bounds__(128)
precoded_sinr_kernel(const double2* __restrict__ H, int R, int P, const double2* __restrict__ W, double nVar, int batch,
                     double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    constexpr int NT = NU * (NU + 1) / 2;
    double2 A[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) A[i] = make_double2(0.0, 0.0);
    const double2* __restrict__ Hb = H + (size_t)b * R * P;
    for (int r = 0; r < R; ++r) {
        double2 g[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) g[j] = make_double2(0.0, 0.0);
        for (int q = 0; q < P; ++q) {
            const double2 h = Hb[r + (size_t)R * q];
#pragma unroll
            for (int j = 0; j < NU; ++j) g[j] = zfma(g[j], h, __ldg(W + q + (size_t)P * j));
        }
#pragma unroll
        for (int i = 0; i < NU; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) A[TRI(i, j)] = zfmac(A[TRI(i, j)], g[j], g[i]);
    }
    double s[NU];
    chol_sinr<NU>(A, nVar, s, 1);
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < NU; ++j) tot += s[j];
    out[b] = tot;
}

int precoded_sinr_run(Ctx* ctx, const double2* H, int R, int P, double sigma, const double2* W, int nLayers, int batch,
                      double* out, cudaStream_t st) {
    if (!H || !W || !out || R < 1 || P < 1 || nLayers < 1 || nLayers > kMaxLayers || batch < 1 || !(sigma > 0.0)) {
        set_error(ctx, "precodedSINR: invalid argument (1 <= nLayers <= 8, sigma > 0)");
        return kErrInvalidArg;
    }
    const double nVar = sigma * sigma;
    const int grid = (batch + 127) / 128;
    switch (nLayers) {
        case 1: precoded_sinr_kernel<1><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
        case 2: precoded_sinr_kernel<2><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
        case 3: precoded_sinr_kernel<3><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
        case 4: precoded_sinr_kernel<4><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
        case 5: precoded_sinr_kernel<5><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
        case 6: precoded_sinr_kernel<6><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
        case 7: precoded_sinr_kernel<7><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
        default: precoded_sinr_kernel<8><<<grid, 128, 0, st>>>(H, R, P, W, nVar, batch, out); break;
    }
    count_launches(ctx, 1);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

}  // namespace isac
