// Channel estimation from reference signals on sm_100a.
//
// Replaces the reference's calls of the closed toolbox function
//   [Hest, nVar] = nrChannelEstimate(rxGrid, refInd, refSym, 'CDMLengths', cdmLen [, 'AveragingWindow', [F T]])
// (+communication/+phyLayer/uePhy.m:897 CSI-RS, gNBPhy.m:1030 SRS, uePhy.m:836 / gNBPhy.m:935 DM-RS) so that the channel
// matrix consumed by riSelect / cqiSelect / pmiSelect never leaves the device.  The estimator (LS at the reference REs,
// CDM despreading by block means, optional moving average, linear interpolation with constant extrapolation in
// frequency then time, noise variance from second differences of the despread estimates) is specified in
// oracle/chest.py; PARITY-UNPINNED against the toolbox.
//
// Kernels (all HBM / latency trivial except the last one):
//   chest_despread_kernel : one thread per (block, rx antenna, port, UE): mean of rx * conj(s)/|s|^2 over FD x TD REs
//   chest_average_kernel  : truncated F x T moving average over blocks (only when an averaging window is set)
//   chest_noise_kernel    : one CTA per UE, fixed-order float64 reduction of |second difference|^2  -> nVar
//   chest_interp_kernel   : one thread per (subcarrier, rx antenna, port, UE) writing the L symbols of H: coalesced
//                           float2 stores along the subcarrier axis; the kernel is bound by the 8*K*L*R*P bytes it writes.
#include "chest.cuh"
#include "ctx.cuh"
#include <algorithm>
#include <cmath>

namespace isac {

struct ChestDev {
    const float2* rx;
    const int32_t* refK;
    const int32_t* refL;
    const float2* inv;
    const int32_t* flo;
    const float* fw;
    const int32_t* tlo;
    const float* tw;
    float2* D;
    float2* A;
    float2* H;
    double* nvar;
    int K, L, R, P, nK, nL, nBf, nBt, FD, TD, avgF, avgT, batch;
};

__global__ void __launch_bounds__(256) chest_despread_kernel(const ChestDev p) {
    const long long total = (long long)p.nBf * p.nBt * p.R * p.P * p.batch;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int bf = (int)(gid % p.nBf);
    long long t = gid / p.nBf;
    const int bt = (int)(t % p.nBt); t /= p.nBt;
    const int r = (int)(t % p.R); t /= p.R;
    const int port = (int)(t % p.P);
    const int b = (int)(t / p.P);
    const float2* __restrict__ rx = p.rx + ((size_t)b * p.R + r) * (size_t)p.K * p.L;
    const int k0 = bf * p.FD, k1 = min(k0 + p.FD, p.nK);
    const int l0 = bt * p.TD, l1 = min(l0 + p.TD, p.nL);
    float2 acc = make_float2(0.f, 0.f);
    for (int li = l0; li < l1; ++li) {
        const int l = __ldg(p.refL + li + p.nL * port);
        for (int ki = k0; ki < k1; ++ki) {
            const int k = __ldg(p.refK + ki + p.nK * port);
            const float2 z = cmul(__ldg(rx + k + (size_t)p.K * l), __ldg(p.inv + ki + p.nK * (li + p.nL * port)));  // LS estimate
            acc = cadd(acc, z);
        }
    }
    p.D[gid] = cscale(acc, 1.0f / (float)((k1 - k0) * (l1 - l0)));
}

__global__ void __launch_bounds__(256) chest_average_kernel(const ChestDev p) {
    const long long total = (long long)p.nBf * p.nBt * p.R * p.P * p.batch;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int bf = (int)(gid % p.nBf);
    const int bt = (int)((gid / p.nBf) % p.nBt);
    const float2* __restrict__ D = p.D + (gid - bf - (long long)p.nBf * bt);
    const int hf = p.avgF > 1 ? p.avgF / 2 : 0, ht = p.avgT > 1 ? p.avgT / 2 : 0;
    // frequency first, then time (separable truncated means, as oracle/chest.py::_moving_average)
    float2 acc = make_float2(0.f, 0.f);
    const int t0 = max(0, bt - ht), t1 = min(p.nBt, bt + ht + 1);
    const int f0 = max(0, bf - hf), f1 = min(p.nBf, bf + hf + 1);
    for (int tt = t0; tt < t1; ++tt) {
        float2 row = make_float2(0.f, 0.f);
        for (int ff = f0; ff < f1; ++ff) row = cadd(row, __ldg(D + ff + p.nBf * tt));
        acc = cadd(acc, cscale(row, 1.0f / (float)(f1 - f0)));
    }
    p.A[gid] = cscale(acc, 1.0f / (float)(t1 - t0));
}

__global__ void __launch_bounds__(1024) chest_noise_kernel(const ChestDev p) {
    __shared__ double red[1024];
    const int b = blockIdx.x;
    const long long rows = (long long)p.nBt * p.R * p.P;  // (bt, r, port) rows of nBf blocks each
    const float2* __restrict__ D = p.D + (size_t)b * rows * p.nBf;
    const long long n = p.nBf >= 3 ? rows * (p.nBf - 2) : 0;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {  // fixed assignment and order -> deterministic
        const long long row = i / (p.nBf - 2);
        const int bf = (int)(i % (p.nBf - 2));
        const float2 a = D[row * p.nBf + bf], c = D[row * p.nBf + bf + 1], e = D[row * p.nBf + bf + 2];
        const double dx = (double)e.x - 2.0 * (double)c.x + (double)a.x, dy = (double)e.y - 2.0 * (double)c.y + (double)a.y;
        acc += dx * dx + dy * dy;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.nvar[b] = n ? ((double)(p.FD * p.TD) / 6.0) * red[0] / (double)n : 0.0;
}

__global__ void __launch_bounds__(256) chest_interp_kernel(const ChestDev p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.K) return;
    const int r = blockIdx.y % p.R, port = blockIdx.y / p.R, b = blockIdx.z;
    const float2* __restrict__ D = (p.avgF > 1 || p.avgT > 1 ? p.A : p.D) + (((size_t)b * p.P + port) * p.R + r) * (size_t)p.nBf * p.nBt;
    const int lo = __ldg(p.flo + k + p.K * port);
    const float w = __ldg(p.fw + k + p.K * port);
    const int hi = min(lo + 1, p.nBf - 1);
    float2* __restrict__ H = p.H + ((((size_t)b * p.P + port) * p.R + r) * p.L) * (size_t)p.K + k;
    int tPrev = -1;
    float2 f0 = make_float2(0.f, 0.f), f1 = f0;
    for (int l = 0; l < p.L; ++l) {
        const int t0 = __ldg(p.tlo + l + p.L * port);
        const float tw = __ldg(p.tw + l + p.L * port);
        if (t0 != tPrev) {  // frequency interpolation of the two bracketing block rows
            const int t1 = min(t0 + 1, p.nBt - 1);
            const float2 a0 = __ldg(D + lo + p.nBf * t0), b0 = __ldg(D + hi + p.nBf * t0);
            const float2 a1 = __ldg(D + lo + p.nBf * t1), b1 = __ldg(D + hi + p.nBf * t1);
            f0 = make_float2(fmaf(w, b0.x - a0.x, a0.x), fmaf(w, b0.y - a0.y, a0.y));
            f1 = make_float2(fmaf(w, b1.x - a1.x, a1.x), fmaf(w, b1.y - a1.y, a1.y));
            tPrev = t0;
        }
        H[(size_t)l * p.K] = make_float2(fmaf(tw, f1.x - f0.x, f0.x), fmaf(tw, f1.y - f0.y, f0.y));
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------
static void interp_table(const std::vector<double>& c, int n, std::vector<int32_t>& lo, std::vector<float>& w) {
    lo.assign(n, 0);
    w.assign(n, 0.f);
    if (c.size() < 2) return;
    for (int x = 0; x < n; ++x) {
        int i = (int)(std::upper_bound(c.begin(), c.end(), (double)x) - c.begin()) - 1;
        i = std::max(0, std::min(i, (int)c.size() - 2));
        double t = ((double)x - c[i]) / (c[i + 1] - c[i]);
        t = std::max(0.0, std::min(1.0, t));  // constant extrapolation outside the span of the block centres
        lo[x] = i;
        w[x] = (float)t;
    }
}

template <class T>
static bool upload(T** d, const std::vector<T>& h) {
    if (cudaMalloc((void**)d, sizeof(T) * std::max<size_t>(h.size(), 1)) != cudaSuccess) return false;
    return h.empty() || cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice) == cudaSuccess;
}

int chest_plan_create(Ctx* ctx, const ChestConfig& c, long long nRef, const int32_t* refInd, const float2* refSym,
                      ChestPlan** out) {
    if (c.K < 1 || c.L < 1 || c.nRx < 1 || c.nPorts < 1 || c.cdmFd < 1 || c.cdmTd < 1 || c.maxBatch < 1 || nRef < 1 || !refInd ||
        !refSym || c.avgF < 0 || c.avgT < 0) {
        set_error(ctx, "channelEstimate: invalid configuration");
        return kErrInvalidArg;
    }
    if ((c.avgF > 1 && c.avgF % 2 == 0) || (c.avgT > 1 && c.avgT % 2 == 0)) {
        set_error(ctx, "channelEstimate: averaging window sizes must be odd");
        return kErrInvalidArg;
    }
    const long long KL = (long long)c.K * c.L;
    struct Ref { int k, l; float2 s; };
    std::vector<std::vector<Ref>> per(c.nPorts);
    for (long long i = 0; i < nRef; ++i) {
        const long long ind = (long long)refInd[i] - 1;
        if (ind < 0 || ind >= KL * c.nPorts) {
            set_error(ctx, "channelEstimate: refInd outside the K x L x P grid");
            return kErrInvalidArg;
        }
        const int port = (int)(ind / KL);
        const long long rem = ind % KL;
        per[port].push_back({(int)(rem % c.K), (int)(rem / c.K), refSym[i]});
    }
    int nK = -1, nL = -1;
    std::vector<int32_t> refK, refL, flo, tlo;
    std::vector<float2> inv;
    std::vector<float> fw, tw;
    for (int p = 0; p < c.nPorts; ++p) {
        auto& v = per[p];
        if (v.empty()) {
            set_error(ctx, "channelEstimate: a port has no reference REs");
            return kErrInvalidArg;
        }
        std::vector<int> ks, ls;
        for (auto& r : v) { ks.push_back(r.k); ls.push_back(r.l); }
        std::sort(ks.begin(), ks.end()); ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
        std::sort(ls.begin(), ls.end()); ls.erase(std::unique(ls.begin(), ls.end()), ls.end());
        if (v.size() != ks.size() * ls.size() || (nK >= 0 && ((int)ks.size() != nK || (int)ls.size() != nL))) {
            set_error(ctx, "channelEstimate: the reference REs of every port must form the same-sized (symbols x subcarriers) grid");
            return kErrUnsupported;
        }
        nK = (int)ks.size();
        nL = (int)ls.size();
        std::vector<float2> invp((size_t)nK * nL, make_float2(0.f, 0.f));
        std::vector<char> seen((size_t)nK * nL, 0);
        for (auto& r : v) {
            const int ki = (int)(std::lower_bound(ks.begin(), ks.end(), r.k) - ks.begin());
            const int li = (int)(std::lower_bound(ls.begin(), ls.end(), r.l) - ls.begin());
            const double m2 = (double)r.s.x * r.s.x + (double)r.s.y * r.s.y;
            if (m2 == 0.0 || seen[ki + (size_t)nK * li]) {
                set_error(ctx, "channelEstimate: zero or duplicated reference symbol");
                return kErrInvalidArg;
            }
            seen[ki + (size_t)nK * li] = 1;
            invp[ki + (size_t)nK * li] = make_float2((float)(r.s.x / m2), (float)(-r.s.y / m2));
        }
        refK.insert(refK.end(), ks.begin(), ks.end());
        refL.insert(refL.end(), ls.begin(), ls.end());
        inv.insert(inv.end(), invp.begin(), invp.end());
        // block centres and interpolation tables of this port
        std::vector<double> kc, lc;
        for (int i = 0; i < nK; i += c.cdmFd) {
            const int j = std::min(i + c.cdmFd, nK);
            double s = 0;
            for (int q = i; q < j; ++q) s += ks[q];
            kc.push_back(s / (j - i));
        }
        for (int i = 0; i < nL; i += c.cdmTd) {
            const int j = std::min(i + c.cdmTd, nL);
            double s = 0;
            for (int q = i; q < j; ++q) s += ls[q];
            lc.push_back(s / (j - i));
        }
        std::vector<int32_t> lo;
        std::vector<float> w;
        interp_table(kc, c.K, lo, w);
        flo.insert(flo.end(), lo.begin(), lo.end());
        fw.insert(fw.end(), w.begin(), w.end());
        interp_table(lc, c.L, lo, w);
        tlo.insert(tlo.end(), lo.begin(), lo.end());
        tw.insert(tw.end(), w.begin(), w.end());
    }
    ChestPlan* p = new ChestPlan();
    p->ctx = ctx;
    p->cfg = c;
    p->nK = nK;
    p->nL = nL;
    p->nBf = (nK + c.cdmFd - 1) / c.cdmFd;
    p->nBt = (nL + c.cdmTd - 1) / c.cdmTd;
    const size_t nD = (size_t)p->nBf * p->nBt * c.nRx * c.nPorts * c.maxBatch;
    bool ok = upload(&p->d_refK, refK) && upload(&p->d_refL, refL) && upload(&p->d_inv, inv) && upload(&p->d_flo, flo) &&
              upload(&p->d_fw, fw) && upload(&p->d_tlo, tlo) && upload(&p->d_tw, tw);
    ok = ok && cudaMalloc((void**)&p->d_D, sizeof(float2) * nD) == cudaSuccess;
    if (c.avgF > 1 || c.avgT > 1) ok = ok && cudaMalloc((void**)&p->d_A, sizeof(float2) * nD) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&p->d_nvar, sizeof(double) * c.maxBatch) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&p->h_nvar, sizeof(double) * c.maxBatch) == cudaSuccess;
    if (!ok) {
        set_error(ctx, std::string("channelEstimate: allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        chest_plan_destroy(p);
        return kErrCuda;
    }
    *out = p;
    return kOk;
}

void chest_plan_destroy(ChestPlan* p) {
    if (!p) return;
    cudaFree(p->d_refK); cudaFree(p->d_refL); cudaFree(p->d_inv); cudaFree(p->d_flo); cudaFree(p->d_fw);
    cudaFree(p->d_tlo); cudaFree(p->d_tw); cudaFree(p->d_D); cudaFree(p->d_A); cudaFree(p->d_nvar);
    if (p->h_nvar) cudaFreeHost(p->h_nvar);
    delete p;
}

int chest_run(ChestPlan* p, const float2* rx, int batch, float2* H, double* nVarHost, cudaStream_t st) {
    Ctx* ctx = p->ctx;
    const ChestConfig& c = p->cfg;
    if (batch < 1 || batch > c.maxBatch || !rx || !H) {
        set_error(ctx, "channelEstimate: batch out of range or null pointer");
        return kErrInvalidArg;
    }
    ChestDev d{};
    d.rx = rx; d.refK = p->d_refK; d.refL = p->d_refL; d.inv = p->d_inv; d.flo = p->d_flo; d.fw = p->d_fw;
    d.tlo = p->d_tlo; d.tw = p->d_tw; d.D = p->d_D; d.A = p->d_A; d.H = H; d.nvar = p->d_nvar;
    d.K = c.K; d.L = c.L; d.R = c.nRx; d.P = c.nPorts; d.nK = p->nK; d.nL = p->nL; d.nBf = p->nBf; d.nBt = p->nBt;
    d.FD = c.cdmFd; d.TD = c.cdmTd; d.avgF = c.avgF; d.avgT = c.avgT; d.batch = batch;
    const long long nD = (long long)p->nBf * p->nBt * c.nRx * c.nPorts * batch;
    const int pr = prof_begin(ctx, kProfChest, st);
    chest_despread_kernel<<<(unsigned)((nD + 255) / 256), 256, 0, st>>>(d);
    int launches = 3;
    if (c.avgF > 1 || c.avgT > 1) {
        chest_average_kernel<<<(unsigned)((nD + 255) / 256), 256, 0, st>>>(d);
        ++launches;
    }
    chest_noise_kernel<<<batch, 1024, 0, st>>>(d);
    chest_interp_kernel<<<dim3((c.K + 255) / 256, c.nRx * c.nPorts, batch), 256, 0, st>>>(d);
    prof_end(ctx, pr, st);
    count_launches(ctx, launches);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    if (nVarHost) {
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(p->h_nvar, p->d_nvar, sizeof(double) * batch, cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        for (int b = 0; b < batch; ++b) nVarHost[b] = p->h_nvar[b];
    }
    return kOk;
}

}  // namespace isac
