// K11: CDL frequency-response generator (TR 38.901 7.7.1, steps of 7.5 for fixed angles).
//
// The reference filters time-domain waveforms through the toolbox's nrCDLChannel (uePhy.m:731, gNBPhy.m:840,
// objects built at +parameters/+channelModels/+communication/cdl.m:48-88) and recovers H with
// nrChannelEstimate (uePhy.m:897).  This kernel produces the channel matrix H[K x L x nRx x nTx] directly in
// the frequency domain.  PARITY: statistical only — the toolbox's source and its mt19937 ray-coupling /
// initial-phase draws are not reproducible here; tables are TR 38.901 Tables 7.7.1-1/-3/-4 and 7.5-3.
//
// Host (float64, tiny): per-ray static coefficients g_m[u,s] (field patterns, XPR matrix, array phases, power)
// and Dopplers.  Device: (1) C_n[l,u,s] = sum_{m in n} g_m[u,s] e^{2 pi j nu_m t_l}; (2) the write-bound
// contraction H[k,(l,u,s)] = sum_n e^{-2 pi j f_k tau_n} C_n[(l,u,s)] (contraction length = #clusters <= 24).
#include "cdl.cuh"
#include <algorithm>
#include "ctx.cuh"
#include <cmath>
#include <cstring>

namespace isac {

namespace {
struct Row { double delay, powerdB, aod, aoa, zod, zoa; };
struct Profile { std::vector<Row> rows; double cASD, cASA, cZSD, cZSA, xprdB; bool los; double losPowerdB; };

const Profile& profile_table(int id) {
    static const Profile A{{{0.0000,-13.4,-178.1,51.3,50.2,125.4},{0.3819,0,-4.2,-152.7,93.2,91.3},{0.4025,-2.2,-4.2,-152.7,93.2,91.3},
        {0.5868,-4,-4.2,-152.7,93.2,91.3},{0.4610,-6,90.2,76.6,122,94},{0.5375,-8.2,90.2,76.6,122,94},{0.6708,-9.9,90.2,76.6,122,94},
        {0.5750,-10.5,121.5,-1.8,150.2,47.1},{0.7618,-7.5,-81.7,-41.9,55.2,56},{1.5375,-15.9,158.4,94.2,26.4,30.1},
        {1.8978,-6.6,-83,51.9,126.4,58.8},{2.2242,-16.7,134.8,-115.9,171.6,26},{2.1718,-12.4,-153,26.6,151.4,49.2},
        {2.4942,-15.2,-172,76.6,157.2,143.1},{2.5119,-10.8,-129.9,-7,47.2,117.4},{3.0582,-11.3,-136,-23,40.4,122.7},
        {4.0810,-12.7,165.4,-47.2,43.3,123.2},{4.4579,-16.2,148.4,110.4,161.8,32.6},{4.5695,-18.3,132.7,144.5,10.8,27.2},
        {4.7966,-18.9,-118.6,155.3,16.7,15.2},{5.0066,-16.6,-154.1,102,171.7,146},{5.3043,-19.9,126.5,-151.8,22.7,150.7},
        {9.6586,-29.7,-56.2,55.2,144.9,156.1}}, 5, 11, 3, 3, 10, false, 0};
    static const Profile Cc{{{0,-4.4,-46.6,-101,97.2,87.6},{0.2099,-1.2,-22.8,120,98.6,72.1},{0.2219,-3.5,-22.8,120,98.6,72.1},
        {0.2329,-5.2,-22.8,120,98.6,72.1},{0.2176,-2.5,-40.7,-127.5,100.6,70.1},{0.6366,0,0.3,170.4,99.2,75.3},
        {0.6448,-2.2,0.3,170.4,99.2,75.3},{0.6560,-3.9,0.3,170.4,99.2,75.3},{0.6584,-7.4,73.1,55.4,105.2,67.4},
        {0.7935,-7.1,-64.5,66.5,95.3,63.8},{0.8213,-10.7,80.2,-48.1,106.1,71.4},{0.9336,-11.1,-97.1,46.9,93.5,60.5},
        {1.2285,-5.1,-55.3,68.1,103.7,90.6},{1.3083,-6.8,-64.3,-68.7,104.2,60.1},{2.1704,-8.7,-78.5,81.5,93.0,61.0},
        {2.7105,-13.2,102.7,30.7,104.2,100.7},{4.2589,-13.9,99.2,-16.4,94.9,62.3},{4.6003,-13.9,88.8,3.8,93.1,66.7},
        {5.4902,-15.8,-101.9,-13.7,92.2,52.9},{5.6077,-17.1,92.2,9.7,106.7,61.8},{6.3065,-16,93.3,5.6,93.0,51.9},
        {6.6374,-15.7,106.6,0.7,92.9,61.7},{7.0427,-21.6,119.5,-21.9,105.2,58},{8.6523,-22.8,-123.8,33.6,107.8,57}},
        2, 15, 3, 7, 7, false, 0};
    static const Profile D{{{0,-13.5,0,-180,98.5,81.5},{0.035,-18.8,89.2,89.2,85.5,86.9},{0.612,-21,89.2,89.2,85.5,86.9},
        {1.363,-22.8,89.2,89.2,85.5,86.9},{1.405,-17.9,13,163,97.5,79.4},{1.804,-20.1,13,163,97.5,79.4},{2.596,-21.9,13,163,97.5,79.4},
        {1.775,-22.9,34.6,-137,98.5,78.2},{4.042,-27.8,-64.5,74.5,88.4,73.6},{7.937,-23.6,-32.9,127.7,91.3,78.3},
        {9.424,-24.8,52.6,-119.6,103.8,87},{9.708,-30.0,-132.1,-9.1,80.3,70.6},{12.525,-27.7,77.2,-83.8,86.5,72.9}},
        5, 8, 3, 3, 11, true, -0.2};
    // TR 38.901 Table 7.7.1-2 (CDL-B, NLOS) and Table 7.7.1-5 (CDL-E, LOS: K = 22 dB; the first row is the Rayleigh part of the
    // LOS cluster, the specular path carries -0.03 dB)
    static const Profile B{{{0.0000,0,9.3,-173.3,105.8,78.9},{0.1072,-2.2,9.3,-173.3,105.8,78.9},{0.2155,-4,9.3,-173.3,105.8,78.9},
        {0.2095,-3.2,-34.1,125.5,115.3,63.3},{0.2870,-9.8,-65.4,-88.0,119.3,59.9},{0.2986,-1.2,-11.4,155.1,103.2,67.5},
        {0.3752,-3.4,-11.4,155.1,103.2,67.5},{0.5055,-5.2,-11.4,155.1,103.2,67.5},{0.3681,-7.6,-67.2,-89.8,118.2,82.6},
        {0.3697,-3,52.5,132.1,102.0,66.3},{0.5700,-8.9,-72,-83.6,100.4,61.6},{0.5283,-9,74.3,95.3,98.3,58.0},
        {1.1021,-4.8,-52.2,103.7,103.4,78.2},{1.2756,-5.7,-50.5,-87.8,102.5,82.0},{1.5474,-7.5,61.4,-92.5,101.4,62.4},
        {1.7842,-1.9,30.6,-139.1,103.0,78.0},{2.0169,-7.6,-72.5,-90.6,100.0,60.9},{2.8294,-12.2,-90.6,58.6,115.2,82.9},
        {3.0219,-9.8,-77.6,-79.0,100.5,60.8},{3.6187,-11.4,-82.6,65.8,119.6,57.3},{4.1067,-14.9,-103.6,52.7,118.7,59.9},
        {4.2790,-9.2,75.6,88.7,117.8,60.1},{4.7834,-11.3,-77.6,-60.4,115.7,62.3}}, 10, 22, 3, 7, 8, false, 0};
    static const Profile E{{{0,-22.03,0,-180,99.6,80.4},{0.5133,-15.8,57.5,18.2,104.2,80.4},{0.5440,-18.1,57.5,18.2,104.2,80.4},
        {0.5630,-19.8,57.5,18.2,104.2,80.4},{0.5440,-22.9,-20.1,101.8,99.4,80.8},{0.7112,-22.4,16.2,112.9,100.8,86.3},
        {1.9092,-18.6,9.3,-155.5,98.8,82.7},{1.9293,-20.8,9.3,-155.5,98.8,82.7},{1.9589,-22.6,9.3,-155.5,98.8,82.7},
        {2.6426,-22.3,19,-143.3,100.8,82.9},{3.7136,-25.6,32.7,-94.7,96.4,88},{5.4524,-20.2,0.5,147,98.9,81},
        {12.0034,-29.8,55.9,-36.2,95.6,88.6},{20.6519,-29.2,57.6,-26,104.6,78.3}}, 5, 11, 3, 7, 8, true, -0.03};
    switch (id) {
        case 0: return A;
        case 1: return B;
        case 2: return Cc;
        case 3: return D;
        default: return E;
    }
}

// splitmix64: the documented generator of the ray coupling / initial phases (restated in oracle/cdl.py)
struct SplitMix {
    unsigned long long s;
    unsigned long long next() {
        unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
};

// TR 38.901 Table 7.3-1 element power pattern (linear), theta = zenith, phi = azimuth in degrees
double pattern38901(double theta, double phi) {
    auto wrap = [](double a) { a = std::fmod(a + 180.0, 360.0); if (a < 0) a += 360.0; return a - 180.0; };
    const double ph = wrap(phi);
    const double av = -std::min(12.0 * ((theta - 90.0) / 65.0) * ((theta - 90.0) / 65.0), 30.0);
    const double ah = -std::min(12.0 * (ph / 65.0) * (ph / 65.0), 30.0);
    const double a = -std::min(-(av + ah), 30.0) + 8.0;  // 8 dBi maximum gain
    return std::pow(10.0, a / 10.0);
}
}  // namespace

int cdl_build_rays(Ctx* ctx, const CdlConfig& c, CdlRays& r) {
    if (c.profile < 0 || c.profile > 4) {
        set_error(ctx, "CDL: only CDL-A, CDL-C and CDL-D are tabulated");
        return kErrUnsupported;
    }
    for (int i = 0; i < 3; ++i)
        if (c.txSize[i] < 1 || c.rxSize[i] < 1 || (i == 2 && (c.txSize[i] > 2 || c.rxSize[i] > 2))) {
            set_error(ctx, "CDL: invalid antenna array size");
            return kErrInvalidArg;
        }
    const Profile& p = profile_table(c.profile);
    static const double alpha[20] = {0.0447, -0.0447, 0.1413, -0.1413, 0.2492, -0.2492, 0.3715, -0.3715, 0.5129, -0.5129,
                                     0.6797, -0.6797, 0.8844, -0.8844, 1.1481, -1.1481, 1.5195, -1.5195, 2.1551, -2.1551};
    const int nCl = (int)p.rows.size(), M = 20;
    const int nTx = c.txSize[0] * c.txSize[1] * c.txSize[2], nRx = c.rxSize[0] * c.rxSize[1] * c.rxSize[2];
    r = CdlRays();
    r.nCl = nCl; r.nRay = M; r.nRx = nRx; r.nTx = nTx; r.los = p.los;
    // powers, normalised so that all path gains sum to 1 (NormalizePathGains)
    std::vector<double> pw(nCl);
    double tot = 0.0, plos = p.los ? std::pow(10.0, p.losPowerdB / 10.0) : 0.0;
    for (int n = 0; n < nCl; ++n) { pw[n] = std::pow(10.0, p.rows[n].powerdB / 10.0); tot += pw[n]; }
    tot += plos;
    for (int n = 0; n < nCl; ++n) pw[n] /= tot;
    plos /= tot;
    r.power = pw;
    r.tau.resize(nCl);
    for (int n = 0; n < nCl; ++n) r.tau[n] = p.rows[n].delay * c.delaySpread;
    const double kappa = std::pow(10.0, p.xprdB / 10.0);
    const double deg = M_PI / 180.0;
    auto elemPos = [](const int size[3], int e, double& y, double& z, int& pol) {
        const int m = e % size[0], n = (e / size[0]) % size[1];
        pol = e / (size[0] * size[1]);
        y = 0.5 * n;  // columns along y, rows along z, half-wavelength spacing (toolbox default [0.5 0.5 1 1])
        z = 0.5 * m;
    };
    auto field = [&](bool pat, int npol, int pol, double theta, double phi, double& Ft, double& Fp) {
        const double a = pat ? std::sqrt(pattern38901(theta, phi)) : 1.0;
        const double zeta = npol == 2 ? (pol == 0 ? 45.0 : -45.0) : 0.0;  // PolarizationAngles [45 -45]; single pol: vertical
        Ft = a * std::cos(zeta * deg);  // polarisation model 2 (TR 38.901 7.3.2)
        Fp = a * std::sin(zeta * deg);
    };
    SplitMix rng{c.seed};
    const int nRaysTot = nCl * M + (p.los ? 1 : 0);
    r.g.assign((size_t)nRaysTot * nRx * nTx, 0.0);
    r.nu.assign(nRaysTot, 0.0);
    r.cluster.assign(nRaysTot, 0);
    auto ray = [&](int idx, int cl, double amp, double aod, double aoa, double zod, double zoa, const std::complex<double> X[4]) {
        r.cluster[idx] = cl;
        r.nu[idx] = c.maxDoppler * std::sin(zoa * deg) * std::cos(aoa * deg);  // UT travels along +x (UTDirectionOfTravel [0;90])
        const double rxv[3] = {std::sin(zoa * deg) * std::cos(aoa * deg), std::sin(zoa * deg) * std::sin(aoa * deg), std::cos(zoa * deg)};
        const double txv[3] = {std::sin(zod * deg) * std::cos(aod * deg), std::sin(zod * deg) * std::sin(aod * deg), std::cos(zod * deg)};
        for (int u = 0; u < nRx; ++u) {
            double yu, zu; int pu;
            elemPos(c.rxSize, u, yu, zu, pu);
            double Frt, Frp;
            field(c.rxPattern38901 != 0, c.rxSize[2], pu, zoa, aoa, Frt, Frp);
            const double phr = 2.0 * M_PI * (rxv[1] * yu + rxv[2] * zu);
            for (int s = 0; s < nTx; ++s) {
                double ys, zs; int ps;
                elemPos(c.txSize, s, ys, zs, ps);
                double Ftt, Ftp;
                field(c.txPattern38901 != 0, c.txSize[2], ps, zod, aod, Ftt, Ftp);
                const double pht = 2.0 * M_PI * (txv[1] * ys + txv[2] * zs);
                const std::complex<double> pol = Frt * (X[0] * Ftt + X[1] * Ftp) + Frp * (X[2] * Ftt + X[3] * Ftp);
                r.g[((size_t)idx * nRx + u) * nTx + s] = amp * pol * std::polar(1.0, phr + pht) / std::sqrt((double)nRx);  // NormalizeChannelOutputs
            }
        }
    };
    for (int n = 0; n < nCl; ++n) {
        // random coupling of the rays within the cluster (TR 38.901 7.5 step 8): permutations for AOA, ZOD, ZOA
        int perm[3][20];
        for (int q = 0; q < 3; ++q) {
            for (int m = 0; m < M; ++m) perm[q][m] = m;
            for (int m = M - 1; m > 0; --m) {
                const int j = (int)(rng.uniform() * (m + 1));
                std::swap(perm[q][m], perm[q][j]);
            }
        }
        for (int m = 0; m < M; ++m) {
            std::complex<double> X[4];
            for (int q = 0; q < 4; ++q) X[q] = std::polar(1.0, (2.0 * rng.uniform() - 1.0) * M_PI);  // initial phases (step 10)
            X[1] *= std::sqrt(1.0 / kappa);
            X[2] *= std::sqrt(1.0 / kappa);
            const Row& row = p.rows[n];
            ray(n * M + m, n, std::sqrt(pw[n] / M), row.aod + p.cASD * alpha[m], row.aoa + p.cASA * alpha[perm[0][m]],
                row.zod + p.cZSD * alpha[perm[1][m]], row.zoa + p.cZSA * alpha[perm[2][m]], X);
        }
    }
    if (p.los) {  // specular ray of cluster 1: XPR matrix diag(1,-1) (TR 38.901 eq. 7.5-29)
        const std::complex<double> X[4] = {1.0, 0.0, 0.0, -1.0};
        const Row& row = p.rows[0];
        ray(nCl * M, 0, std::sqrt(plos), row.aod, row.aoa, row.zod, row.zoa, X);
    }
    return kOk;
}

constexpr int kCdlMaxSym = 16;
struct CdlTimes { double t[kCdlMaxSym]; };

// C_n[l,u,s] = sum_{m in n} g_m[u,s] exp(2 pi j nu_m t_l); rays of cluster n are [n*nRay, (n+1)*nRay) (+ the LOS ray).
// One CTA per (cluster, symbol): the <= 21 ray phasors are computed once, then threads sweep the antenna pairs.
// blockIdx.z = channel of the batch (all channels of a launch share nCl / nRay / array sizes; ray tables differ).
constexpr int kCdlMaxBatch = 32;
struct CdlBatch {
    const double2* g[kCdlMaxBatch];
    const double* nu[kCdlMaxBatch];
    const double* tau[kCdlMaxBatch];
    double t0[kCdlMaxBatch];
};

// LPER = symbols per work item (compile time: a run-time bound on fully unrolled 16-symbol loops issued every predicated-off
// iteration -- 9 of 16 at L = 14, 15 of 16 for the one-symbol SRS channels)
template <int LPER>
__global__ void __launch_bounds__(128)
cdl_cluster_kernel(const CdlBatch bt, int nRay, int losRay, int nCl, int nRx, int nTx, int L, const CdlTimes tl,
                   float2* __restrict__ Call /*[batch][nCl][J]*/) {
    // one CTA per (cluster, channel): the (<= 21 rays) x (L symbols) phasors once, then each thread keeps one antenna
    // pair's ray coefficients in flight while it sweeps its share of the symbols
    __shared__ double2 ph[kCdlMaxSym][24];
    __shared__ int rayIdx[24];
    const int n = blockIdx.x, RT = nRx * nTx;
    const double2* __restrict__ g = bt.g[blockIdx.y];
    const double* __restrict__ nu = bt.nu[blockIdx.y];
    float2* __restrict__ C = Call + (size_t)blockIdx.y * nCl * L * RT;
    const int cnt = nRay + ((n == 0 && losRay >= 0) ? 1 : 0);
    for (int i = threadIdx.x; i < cnt * L; i += blockDim.x) {
        const int q = i % cnt, l = i / cnt;
        const int m = q < nRay ? n * nRay + q : losRay;
        double s, c;
        sincospi(2.0 * nu[m] * (bt.t0[blockIdx.y] + tl.t[l]), &s, &c);
        ph[l][q] = make_double2(c, s);
        if (l == 0) rayIdx[q] = m;
    }
    __syncthreads();
    // work item = (antenna pair us, symbol group): groups of symbols so that all 128 threads are busy when RT < 128
    const int groups = RT >= (int)blockDim.x ? 1 : (int)blockDim.x / RT;
    const int lper = LPER;   // == ceil(L / groups), chosen by the host
    for (int w = threadIdx.x; w < RT * groups; w += blockDim.x) {
        const int us = w % RT, grp = w / RT;
        const int u = us / nTx, sx = us % nTx;
        const int l0 = grp * lper, l1 = min(L, l0 + lper);
        double re[LPER], im[LPER];
#pragma unroll
        for (int l = 0; l < LPER; ++l) re[l] = im[l] = 0.0;
        // ray coefficients five at a time, double-buffered: the loads of group k + 1 are in flight while group k is accumulated,
        // so one global-memory latency is exposed per thread instead of one per ray (the serial form spent 20 x a dependent
        // shared -> global load chain); same summation order
        constexpr int kGrp = 5;
        double2 gv[kGrp], gn[kGrp];
        auto fetch = [&](int q0, double2 (&dst)[kGrp]) {
#pragma unroll
            for (int u2 = 0; u2 < kGrp; ++u2)
                dst[u2] = q0 + u2 < cnt ? __ldg(g + (size_t)rayIdx[q0 + u2] * RT + us) : make_double2(0.0, 0.0);
        };
        fetch(0, gv);
        for (int q0 = 0; q0 < cnt; q0 += kGrp) {
            fetch(q0 + kGrp, gn);
#pragma unroll
            for (int u2 = 0; u2 < kGrp; ++u2) {
                const int q = q0 + u2;
                if (q < cnt) {
#pragma unroll
                    for (int d = 0; d < LPER; ++d) {
                        const int l = l0 + d;
                        if (l < l1) {
                            const double2 p = ph[l][q];
                            re[d] = fma(gv[u2].x, p.x, fma(-gv[u2].y, p.y, re[d]));
                            im[d] = fma(gv[u2].x, p.y, fma(gv[u2].y, p.x, im[d]));
                        }
                    }
                }
            }
#pragma unroll
            for (int u2 = 0; u2 < kGrp; ++u2) gv[u2] = gn[u2];
        }
        // MATLAB order of H(k,l,u,s): j = l + L*(u + nRx*s)
#pragma unroll
        for (int d = 0; d < LPER; ++d) {
            const int l = l0 + d;
            if (l < l1)
                C[(size_t)n * L * RT + l + (size_t)L * (u + (size_t)nRx * sx)] = make_float2((float)re[d], (float)im[d]);
        }
    }
}

// H[k, j] = sum_n E[k,n] C[n, j],  E[k,n] = exp(-2 pi j f_k tau_n);  j = (l,u,s) flattened, output [K x J].
// The complex contraction is run as two real GEMMs that share B on the tensor pipe:
//   Re H = [Er, -Ei] * [Cr ; Ci],   Im H = [Ei, Er] * [Cr ; Ci]        (M = K rows, N = J cols, inner = 2*nClusters <= 48)
// with mma.sync.m16n8k8 TF32 in the error-compensated 3xTF32 form (a = a_hi + a_lo, b = b_hi + b_lo;
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, ~2^-21 relative) so the result keeps float32-level accuracy.  The kernel is
// write-bound (8*K*J bytes); the inner dimension is far too short for a tcgen05/TMEM pipeline to pay off.
constexpr int kCdlTK = 64, kCdlTJ = 128, kCdlMaxCl = 24;

__device__ __forceinline__ unsigned to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Shared memory holds the operands already split into TF32 hi/lo parts and stored in mma FRAGMENT order (one 16-byte
// word per lane and fragment), so the inner loop is 12 conflict-free LDS.128 + 48 MMAs per k-step and warp; splitting
// every operand again at each use cost 3x the instructions of the MMAs themselves.
//   Af[s][mt][v][lane] = {a0,a1,a2,a3} of the 16x8 A fragment of m-tile mt; v: 0 re-hi, 1 re-lo, 2 im-hi, 3 im-lo
//   Bf[s][nt][lane]    = {b0_hi, b1_hi, b0_lo, b1_lo} of the 8x8 B fragment of n-tile nt
constexpr int kCdlKSteps = (2 * kCdlMaxCl + 7) / 8;                                   // 6
constexpr size_t kCdlAfBytes = sizeof(uint4) * kCdlKSteps * (kCdlTK / 16) * 4 * 32;   // 49152
constexpr size_t kCdlBfBytes = sizeof(uint4) * kCdlKSteps * (kCdlTJ / 8) * 32;        // 49152
constexpr size_t kCdlSmemBytes = kCdlAfBytes + kCdlBfBytes;

__global__ void __launch_bounds__(256, 2)
cdl_response_kernel(const float2* __restrict__ Call, const CdlBatch bt, int nCl, int K, long long J, double scs,
                    float2* __restrict__ Hall) {
    extern __shared__ uint4 cdl_sm[];
    uint4* Af = cdl_sm;
    uint4* Bf = cdl_sm + kCdlAfBytes / sizeof(uint4);
    float2* Es = reinterpret_cast<float2*>(Bf);   // [kCdlMaxCl][kCdlTK] staging of E, dead before Bf is filled
    const float2* __restrict__ C = Call + (size_t)blockIdx.z * nCl * J;
    const double* __restrict__ tau = bt.tau[blockIdx.z];
    float2* __restrict__ H = Hall + (size_t)blockIdx.z * K * J;
    const int k0 = blockIdx.x * kCdlTK;
    const long long j0 = (long long)blockIdx.y * kCdlTJ;
    const int ksteps = (2 * nCl + 7) / 8;
    for (int i = threadIdx.x; i < kCdlMaxCl * kCdlTK; i += blockDim.x) {
        const int n = i / kCdlTK, kk = i % kCdlTK;
        float2 e = make_float2(0.f, 0.f);
        if (n < nCl) {
            const double f = ((double)(k0 + kk) - (double)(K / 2)) * scs;
            double sn, cs;
            sincospi(-2.0 * f * tau[n], &sn, &cs);
            e = make_float2((float)cs, (float)sn);
        }
        Es[i] = e;
    }
    __syncthreads();
    // inner index q = 8s + t (+4): cluster n = 4s + t/2 (+2); part = t & 1: 0 -> (Er | Cr), 1 -> (-Ei | Ci) for Re H,
    // (Ei | Cr), (Er | Ci) for Im H
    for (int i = threadIdx.x; i < ksteps * (kCdlTK / 16) * 32; i += blockDim.x) {
        const int lane = i & 31, mt = (i >> 5) % (kCdlTK / 16), sidx = i / (32 * (kCdlTK / 16));
        const int g = lane >> 2, t = lane & 3, part = t & 1;
        const int nA = 4 * sidx + (t >> 1), nB = nA + 2, r0 = mt * 16 + g;
        const float2 z = make_float2(0.f, 0.f);
        const float2 e[4] = {nA < kCdlMaxCl ? Es[nA * kCdlTK + r0] : z, nA < kCdlMaxCl ? Es[nA * kCdlTK + r0 + 8] : z,
                             nB < kCdlMaxCl ? Es[nB * kCdlTK + r0] : z, nB < kCdlMaxCl ? Es[nB * kCdlTK + r0 + 8] : z};
        unsigned rh[4], rl[4], ih[4], il[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            split_tf32(part ? -e[q].y : e[q].x, rh[q], rl[q]);
            split_tf32(part ? e[q].x : e[q].y, ih[q], il[q]);
        }
        uint4* dst = Af + ((size_t)(sidx * (kCdlTK / 16) + mt) * 4) * 32 + lane;
        dst[0] = make_uint4(rh[0], rh[1], rh[2], rh[3]);
        dst[32] = make_uint4(rl[0], rl[1], rl[2], rl[3]);
        dst[64] = make_uint4(ih[0], ih[1], ih[2], ih[3]);
        dst[96] = make_uint4(il[0], il[1], il[2], il[3]);
    }
    __syncthreads();   // Es is dead: Bf may overwrite it
    for (int i = threadIdx.x; i < ksteps * (kCdlTJ / 8) * 32; i += blockDim.x) {
        const int lane = i & 31, nt = (i >> 5) % (kCdlTJ / 8), sidx = i / (32 * (kCdlTJ / 8));
        const int g = lane >> 2, t = lane & 3, part = t & 1;
        const int nA = 4 * sidx + (t >> 1), nB = nA + 2;
        const long long j = j0 + nt * 8 + g;
        const float2 z = make_float2(0.f, 0.f);
        const float2 c0 = (nA < nCl && j < J) ? __ldg(C + (size_t)nA * J + j) : z;
        const float2 c1 = (nB < nCl && j < J) ? __ldg(C + (size_t)nB * J + j) : z;
        unsigned h0, l0, h1, l1;
        split_tf32(part ? c0.y : c0.x, h0, l0);
        split_tf32(part ? c1.y : c1.x, h1, l1);
        Bf[i] = make_uint4(h0, h1, l0, l1);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wk = (warp & 1) * 32, wj = (warp >> 1) * 32;   // warp tile: 32 (k) x 32 (j)
    float dre[2][4][4], dim[2][4][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int q = 0; q < 4; ++q) dre[mi][ni][q] = dim[mi][ni][q] = 0.f;
    for (int sidx = 0; sidx < ksteps; ++sidx) {
        unsigned bh[4][2], bl[4][2];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const uint4 b = Bf[(size_t)(sidx * (kCdlTJ / 8) + (wj >> 3) + ni) * 32 + lane];
            bh[ni][0] = b.x; bh[ni][1] = b.y; bl[ni][0] = b.z; bl[ni][1] = b.w;
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            const uint4* src = Af + ((size_t)(sidx * (kCdlTK / 16) + (wk >> 4) + mi) * 4) * 32 + lane;
            const uint4 v0 = src[0], v1 = src[32], v2 = src[64], v3 = src[96];
            const unsigned ah[4] = {v0.x, v0.y, v0.z, v0.w}, al[4] = {v1.x, v1.y, v1.z, v1.w};
            const unsigned ch[4] = {v2.x, v2.y, v2.z, v2.w}, cl[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                mma_tf32(dre[mi][ni], al, bh[ni]);
                mma_tf32(dre[mi][ni], ah, bl[ni]);
                mma_tf32(dre[mi][ni], ah, bh[ni]);
                mma_tf32(dim[mi][ni], cl, bh[ni]);
                mma_tf32(dim[mi][ni], ch, bl[ni]);
                mma_tf32(dim[mi][ni], ch, bh[ni]);
            }
        }
    }
    // accumulator (row g / g+8, col 2t / 2t+1) -> H[j*K + k]
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + wk + mi * 16 + g + ((q >> 1) ? 8 : 0);
                const long long j = j0 + wj + ni * 8 + 2 * t + (q & 1);
                if (k < K && j < J) H[j * K + k] = make_float2(dre[mi][ni][q], dim[mi][ni][q]);
            }
}

// ------------------------------------------------------------------------------------------
// K11 on the 5th-generation tensor cores: tcgen05.mma kind::tf32, operands in shared memory, accumulators in TMEM.
//
// One CTA per 128-subcarrier tile of one channel; it walks the 128-column tiles j of H (E is built once).  With E = Er + j Ei [128 x nCl] and
// C = Cr + j Ci [nCl x 128]:   Re H = Er Cr - Ei Ci,   Im H = Ei Cr + Er Ci   -> four real products per k-step of 8
// clusters, each in the 3xTF32 form (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM, ~2^-21 relative), the minus
// sign through the instruction descriptor's a_negate bit.  36 MMAs of 128x128x8 per tile (~1.2 us of tensor time),
// issued by one thread; the other 255 threads only build operands and drain TMEM, so the kernel is bound by the
// 8*K*J bytes it writes.
//
// Shared-memory operand layout (K-major, no swizzle; cute::UMMA "INTERLEAVE" canonical form): core matrix = 8 rows
// x 16 bytes (4 tf32), rows 16 B apart; the two core matrices of a k-step (k = 0..3 | 4..7) are LBO = 128 B apart,
// consecutive 8-row groups SBO = 256 B apart: element (row, k) of k-step s at  s*4096 + (row/8)*256 + (k/4)*128 +
// (row%8)*16 + (k%4)*4  bytes from the start of its tile.
// ------------------------------------------------------------------------------------------
constexpr int kUmmaM = 128, kUmmaN = 128, kUmmaKSteps = kCdlMaxCl / 8;          // 3 k-steps of 8 clusters
constexpr int kUmmaTileBytes = kUmmaKSteps * kUmmaM * 8 * 4;                    // 12288 per operand tile
constexpr size_t kUmmaSmemBytes = 8 * (size_t)kUmmaTileBytes + 1024;            // Er/Ei/Cr/Ci x hi/lo (+ alignment slack)

__device__ __forceinline__ unsigned smem_addr_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long umma_smem_desc(unsigned addr) {
    // start address [0,14) (>>4), LBO [16,30) = 128 B, SBO [32,46) = 256 B, descriptor version 1 at [46,48), no swizzle
    return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)(128 >> 4) << 16) |
           ((unsigned long long)(256 >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(unsigned tmemD, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                          unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmemD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ int umma_off(int row, int k) {   // byte offset inside one k-step block (4096 B)
    return (row >> 3) * 256 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4;
}

__global__ void __launch_bounds__(256, 2)
cdl_response_umma_kernel(const float2* __restrict__ Call, const CdlBatch bt, int nCl, int K, long long J, double scs,
                         float2* __restrict__ Hall) {
    extern __shared__ unsigned char umma_raw[];
    __shared__ unsigned long long mmaBar;
    __shared__ unsigned tmemBase;
    unsigned char* sm = (unsigned char*)(((size_t)umma_raw + 1023) & ~(size_t)1023);
    // tiles: 0 Er_hi, 1 Er_lo, 2 Ei_hi, 3 Ei_lo, 4 Cr_hi, 5 Cr_lo, 6 Ci_hi, 7 Ci_lo
    const float2* __restrict__ C = Call + (size_t)blockIdx.y * nCl * J;
    const double* __restrict__ tau = bt.tau[blockIdx.y];
    float2* __restrict__ H = Hall + (size_t)blockIdx.y * K * J;
    const int k0 = blockIdx.x * kUmmaM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {   // TMEM: 256 columns (Re | Im accumulators, 128 fp32 columns each)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(&tmemBase)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(&mmaBar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // E[k0+row, n] = exp(-2 pi j f tau_n), once per CTA (shared by all column tiles): the phase f*tau is reduced to
    // [0,1) in float64, the sine/cosine of the reduced phase are float32 (as accurate as the fp32 operand they feed)
    for (int i = threadIdx.x; i < (kCdlMaxCl / 4) * kUmmaM; i += blockDim.x) {   // item = (4 consecutive clusters, row): 16-byte stores
        const int g = i / kUmmaM, row = i % kUmmaM;
        unsigned eh[4], el[4], ih[4], il[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int n = g * 4 + kk;
            float er = 0.f, ei = 0.f;
            if (n < nCl) {
                const double c = ((double)(k0 + row) - (double)(K / 2)) * scs * tau[n];
                sincospif(-2.0f * (float)(c - floor(c)), &ei, &er);
            }
            split_tf32(er, eh[kk], el[kk]);
            split_tf32(ei, ih[kk], il[kk]);
        }
        const int o = (g >> 1) * 4096 + umma_off(row, (g & 1) * 4);
        *(uint4*)(sm + 0 * kUmmaTileBytes + o) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
        *(uint4*)(sm + 1 * kUmmaTileBytes + o) = make_uint4(el[0], el[1], el[2], el[3]);
        *(uint4*)(sm + 2 * kUmmaTileBytes + o) = make_uint4(ih[0], ih[1], ih[2], ih[3]);
        *(uint4*)(sm + 3 * kUmmaTileBytes + o) = make_uint4(il[0], il[1], il[2], il[3]);
    }
    // instruction descriptor: D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, K-major both, N>>3 at [17,23), M>>4 at [24,29)
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kUmmaN >> 3) << 17) | ((unsigned)(kUmmaM >> 4) << 24);
    const unsigned idescNegA = idesc | (1u << 13);
    const unsigned base = smem_addr_u32(sm), bar = smem_addr_u32(&mmaBar);
    const int ksteps = (nCl + 7) / 8;
    const int row = (warp & 3) * 32 + lane, k = k0 + row;   // epilogue: warp w drains TMEM lanes 32*(w%4).., column half w/4
    const int colHalf = (warp >> 2) * 64;
    const int nTiles = (int)((J + kUmmaN - 1) / kUmmaN);
    constexpr int kCPer = kCdlMaxCl * kUmmaN / 256;   // C elements per thread and tile (12)
    // Work item of the operand fill = (group of 4 consecutive clusters, column): its four hi (lo) words are one 16-byte row of a
    // core matrix, so each thread issues ONE 128-bit shared-memory store per operand array and item instead of four scattered
    // 32-bit ones (the 32-bit stores of a warp hit 8 banks four times over: 75 % of the kernel's shared-memory wavefronts were
    // conflicts, profiles/r1_cdl_umma_v1_summary.txt).  Item i = threadIdx.x + q*256: group g = i / 128, column i % 128.
    constexpr int kItems = kCPer / 4;                 // 3 items of 4 clusters per thread
    float2 cnext[kCPer];
    auto fetch_c = [&](long long j0) {   // global loads of one column tile of C (zero beyond nCl / J)
#pragma unroll
        for (int q = 0; q < kItems; ++q) {
            const int i = threadIdx.x + q * 256, g = i / kUmmaN, col = i % kUmmaN;
            const long long j = j0 + col;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int n = g * 4 + kk;
                cnext[q * 4 + kk] = (n < nCl && j < J) ? __ldg(C + (size_t)n * J + j) : make_float2(0.f, 0.f);
            }
        }
    };
    fetch_c(0);
    for (int tile = 0; tile < nTiles; ++tile) {
        const long long j0 = (long long)tile * kUmmaN;
        // C tile split into TF32 hi/lo; the previous tile's MMAs have completed (barrier wait below), its loads were
        // issued before that wait so their latency hides behind the tensor work
#pragma unroll
        for (int q = 0; q < kItems; ++q) {
            const int i = threadIdx.x + q * 256, g = i / kUmmaN, col = i % kUmmaN;
            uint4 rh, rl, ih, il;
            split_tf32(cnext[q * 4 + 0].x, rh.x, rl.x); split_tf32(cnext[q * 4 + 1].x, rh.y, rl.y);
            split_tf32(cnext[q * 4 + 2].x, rh.z, rl.z); split_tf32(cnext[q * 4 + 3].x, rh.w, rl.w);
            split_tf32(cnext[q * 4 + 0].y, ih.x, il.x); split_tf32(cnext[q * 4 + 1].y, ih.y, il.y);
            split_tf32(cnext[q * 4 + 2].y, ih.z, il.z); split_tf32(cnext[q * 4 + 3].y, ih.w, il.w);
            const int o = (g >> 1) * 4096 + umma_off(col, (g & 1) * 4);   // clusters 4g..4g+3: k-step g/2, k = 0..3 or 4..7
            *(uint4*)(sm + 4 * kUmmaTileBytes + o) = rh;
            *(uint4*)(sm + 5 * kUmmaTileBytes + o) = rl;
            *(uint4*)(sm + 6 * kUmmaTileBytes + o) = ih;
            *(uint4*)(sm + 7 * kUmmaTileBytes + o) = il;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // also: every warp has drained the previous tile's accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned tmem = tmemBase;
        if (threadIdx.x == 0) {
            for (int s = 0; s < ksteps; ++s) {
                unsigned long long d[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) d[q] = umma_smem_desc(base + q * kUmmaTileBytes + s * 4096);
                const unsigned acc = s > 0;
                // Re (columns 0..127): Er*Cr - Ei*Ci
                umma_tf32(tmem, d[1], d[4], idesc, acc);          // Er_lo * Cr_hi
                umma_tf32(tmem, d[0], d[5], idesc, 1u);           // Er_hi * Cr_lo
                umma_tf32(tmem, d[0], d[4], idesc, 1u);           // Er_hi * Cr_hi
                umma_tf32(tmem, d[3], d[6], idescNegA, 1u);       // -Ei_lo * Ci_hi
                umma_tf32(tmem, d[2], d[7], idescNegA, 1u);
                umma_tf32(tmem, d[2], d[6], idescNegA, 1u);
                // Im (columns 128..255): Ei*Cr + Er*Ci
                umma_tf32(tmem + 128, d[3], d[4], idesc, acc);
                umma_tf32(tmem + 128, d[2], d[5], idesc, 1u);
                umma_tf32(tmem + 128, d[2], d[4], idesc, 1u);
                umma_tf32(tmem + 128, d[1], d[6], idesc, 1u);
                umma_tf32(tmem + 128, d[0], d[7], idesc, 1u);
                umma_tf32(tmem + 128, d[0], d[6], idesc, 1u);
            }
            // arrives on the barrier when every MMA above has completed (implies tcgen05.fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        if (tile + 1 < nTiles) fetch_c(j0 + kUmmaN);
        {   // everyone waits for the accumulators (phase parity alternates per tile)
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "WAIT_MMA:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra DONE_MMA;\n\t"
                "bra WAIT_MMA;\n\t"
                "DONE_MMA:\n\t"
                "}\n" ::"r"(bar), "r"((unsigned)(tile & 1))
                : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            unsigned re[32], im[32];
            const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(colHalf + c0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(re[0]), "=r"(re[1]), "=r"(re[2]), "=r"(re[3]), "=r"(re[4]), "=r"(re[5]), "=r"(re[6]), "=r"(re[7]), "=r"(re[8]),
                  "=r"(re[9]), "=r"(re[10]), "=r"(re[11]), "=r"(re[12]), "=r"(re[13]), "=r"(re[14]), "=r"(re[15]), "=r"(re[16]),
                  "=r"(re[17]), "=r"(re[18]), "=r"(re[19]), "=r"(re[20]), "=r"(re[21]), "=r"(re[22]), "=r"(re[23]), "=r"(re[24]),
                  "=r"(re[25]), "=r"(re[26]), "=r"(re[27]), "=r"(re[28]), "=r"(re[29]), "=r"(re[30]), "=r"(re[31])
                : "r"(taddr));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(im[0]), "=r"(im[1]), "=r"(im[2]), "=r"(im[3]), "=r"(im[4]), "=r"(im[5]), "=r"(im[6]), "=r"(im[7]), "=r"(im[8]),
                  "=r"(im[9]), "=r"(im[10]), "=r"(im[11]), "=r"(im[12]), "=r"(im[13]), "=r"(im[14]), "=r"(im[15]), "=r"(im[16]),
                  "=r"(im[17]), "=r"(im[18]), "=r"(im[19]), "=r"(im[20]), "=r"(im[21]), "=r"(im[22]), "=r"(im[23]), "=r"(im[24]),
                  "=r"(im[25]), "=r"(im[26]), "=r"(im[27]), "=r"(im[28]), "=r"(im[29]), "=r"(im[30]), "=r"(im[31])
                : "r"(taddr + 128u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (k < K) {
                float2* __restrict__ hp = H + (j0 + colHalf + c0) * K + k;   // lanes: consecutive subcarriers
#pragma unroll
                for (int q = 0; q < 32; ++q, hp += K)
                    if (j0 + colHalf + c0 + q < J) *hp = make_float2(__uint_as_float(re[q]), __uint_as_float(im[q]));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(256u) : "memory");
}

// upload the ray tables once (they do not change between calls)
int cdl_upload(Ctx* ctx, CdlRays& rays) {
    if (rays.d_g) return kOk;
    std::vector<double2> g(rays.g.size());
    for (size_t i = 0; i < g.size(); ++i) g[i] = make_double2(rays.g[i].real(), rays.g[i].imag());
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&rays.d_g, sizeof(double2) * g.size()));
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&rays.d_nu, sizeof(double) * rays.nu.size()));
    ISAC_CUDA_CHECK(ctx, cudaMalloc((void**)&rays.d_tau, sizeof(double) * rays.tau.size()));
    ISAC_CUDA_CHECK(ctx, cudaMemcpy(rays.d_g, g.data(), sizeof(double2) * g.size(), cudaMemcpyHostToDevice));
    ISAC_CUDA_CHECK(ctx, cudaMemcpy(rays.d_nu, rays.nu.data(), sizeof(double) * rays.nu.size(), cudaMemcpyHostToDevice));
    ISAC_CUDA_CHECK(ctx, cudaMemcpy(rays.d_tau, rays.tau.data(), sizeof(double) * rays.tau.size(), cudaMemcpyHostToDevice));
    return kOk;
}

void cdl_free(CdlRays& rays) {
    cudaFree(rays.d_g);
    cudaFree(rays.d_nu);
    cudaFree(rays.d_tau);
    rays.d_g = nullptr; rays.d_nu = nullptr; rays.d_tau = nullptr;
}

int cdl_generate_batch(Ctx* ctx, CdlRays* const* rays, int n, int K, double scsHz, int L, const double* symTime,
                       const double* t0, float2* H, cudaStream_t st) {
    const bool legacyMma = rays && rays[0] && rays[0]->legacyMma;
    if (!H || !rays || !t0 || n < 1 || K < 1 || L < 1 || L > kCdlMaxSym || !symTime) {
        set_error(ctx, "cdl_generate: invalid argument (L <= 16 symbols per call)");
        return kErrInvalidArg;
    }
    const CdlRays& r0 = *rays[0];
    if (r0.nCl < 1 || r0.nCl > kCdlMaxCl) {
        set_error(ctx, "cdl_generate: invalid ray table");
        return kErrInvalidArg;
    }
    for (int i = 0; i < n; ++i) {
        const CdlRays& r = *rays[i];
        if (r.nCl != r0.nCl || r.nRay != r0.nRay || r.nRx != r0.nRx || r.nTx != r0.nTx || r.los != r0.los) {
            set_error(ctx, "cdl_generate_batch: channels of one batch must share profile and array sizes");
            return kErrInvalidArg;
        }
        int s = cdl_upload(ctx, *rays[i]);
        if (s) return s;
    }
    const int RT = r0.nRx * r0.nTx;
    const long long J = (long long)L * RT;
    CdlTimes tl{};
    for (int l = 0; l < L; ++l) tl.t[l] = symTime[l];
    const int losRay = r0.los ? r0.nCl * r0.nRay : -1;
    for (int i0 = 0; i0 < n; i0 += kCdlMaxBatch) {
        const int nb = std::min(kCdlMaxBatch, n - i0);
        void* dC = nullptr;
        int s = ctx_scratch(ctx, 10, sizeof(float2) * (size_t)nb * r0.nCl * L * RT, &dC);
        if (s) return s;
        CdlBatch bt{};
        for (int i = 0; i < nb; ++i) {
            bt.g[i] = rays[i0 + i]->d_g;
            bt.nu[i] = rays[i0 + i]->d_nu;
            bt.tau[i] = rays[i0 + i]->d_tau;
            bt.t0[i] = t0[i0 + i];
        }
        const int pr = prof_begin(ctx, kProfCdl, st);
        dim3 g1(r0.nCl, nb);
        {
            const int groups = RT >= 128 ? 1 : 128 / RT, lper = (L + groups - 1) / groups;   // as in the kernel
#define ISAC_CDL_CLUSTER(N_) case N_: cdl_cluster_kernel<N_><<<g1, 128, 0, st>>>(bt, r0.nRay, losRay, r0.nCl, r0.nRx, r0.nTx, L, tl, (float2*)dC); break;
            switch (lper) {
                ISAC_CDL_CLUSTER(1) ISAC_CDL_CLUSTER(2) ISAC_CDL_CLUSTER(3) ISAC_CDL_CLUSTER(4) ISAC_CDL_CLUSTER(5) ISAC_CDL_CLUSTER(6)
                ISAC_CDL_CLUSTER(7) ISAC_CDL_CLUSTER(8) ISAC_CDL_CLUSTER(9) ISAC_CDL_CLUSTER(10) ISAC_CDL_CLUSTER(11) ISAC_CDL_CLUSTER(12)
                ISAC_CDL_CLUSTER(13) ISAC_CDL_CLUSTER(14) ISAC_CDL_CLUSTER(15)
                default: cdl_cluster_kernel<16><<<g1, 128, 0, st>>>(bt, r0.nRay, losRay, r0.nCl, r0.nRx, r0.nTx, L, tl, (float2*)dC); break;
            }
#undef ISAC_CDL_CLUSTER
        }
        dim3 grid((K + kCdlTK - 1) / kCdlTK, (unsigned)((J + kCdlTJ - 1) / kCdlTJ), nb);
        if (!legacyMma) {   // tcgen05 / TMEM path
            dim3 gu((K + kUmmaM - 1) / kUmmaM, nb);   // one CTA per (128-subcarrier tile, channel), looping over the column tiles
            cudaFuncSetAttribute(cdl_response_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmemBytes);
            cdl_response_umma_kernel<<<gu, 256, kUmmaSmemBytes, st>>>((const float2*)dC, bt, r0.nCl, K, J, scsHz, H + (size_t)i0 * K * J);
        } else {
            cudaFuncSetAttribute(cdl_response_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCdlSmemBytes);
            cdl_response_kernel<<<grid, 256, kCdlSmemBytes, st>>>((const float2*)dC, bt, r0.nCl, K, J, scsHz, H + (size_t)i0 * K * J);
        }
        prof_end(ctx, pr, st);
        count_launches(ctx, 2);
    }
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    return kOk;
}

int cdl_generate(Ctx* ctx, CdlRays& rays, int K, double scsHz, int L, const double* symTime, double t0, float2* H,
                 cudaStream_t st) {
    CdlRays* one = &rays;
    return cdl_generate_batch(ctx, &one, 1, K, scsHz, L, symTime, &t0, H, st);
}

}  // namespace isac
