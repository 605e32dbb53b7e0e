// Sensing estimators: fft2D (fft2D.m:31-115), doaEstimation.music (music.m:11-104), music2D (music2D.m:33-123).
#include "sense.cuh"
#include "ctx.cuh"
#include <algorithm>
#include <cmath>
#include <numeric>

namespace isac {

__global__ void gather_kernel(const double* __restrict__ v, const int* __restrict__ order, int n, double* __restrict__ out,
                              double* __restrict__ inv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (out) out[i] = v[order[i]];
    if (inv) inv[i] = v[i] > 0.0 ? 1.0 / v[i] : 0.0;
}

__global__ void set_int_kernel(int* p, int v) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *p = v;
}

static int doa_spec_len(const DoaConfig& d, int* aSteps, int* eSteps) {
    *aSteps = (int)std::floor((d.aMax + 1.0) / d.aGran);
    *eSteps = d.isUpa ? (int)std::floor((d.eMax + 1.0) / d.eGran) : 1;
    return *aSteps * *eSteps;
}

#define SALLOC(ctx, ptr, bytes, onfail)                                         \
    do {                                                                        \
        cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                   \
        if (e_ != cudaSuccess) {                                                \
            set_error(ctx, std::string("cudaMalloc: ") + cudaGetErrorString(e_)); \
            onfail;                                                             \
            return kErrCuda;                                                    \
        }                                                                       \
    } while (0)

int sense_plan_create(Ctx* ctx, const RdmConfig& rc, const DoaConfig& doa, double rRes, double vRes, SensePlan** out) {
    RdmPlan* r = nullptr;
    int s = rdm_plan_create(ctx, rc, &r);
    if (s) return s;
    SensePlan* p = new SensePlan();
    p->rdm = r;
    p->doa = doa;
    p->rRes = rRes;
    p->vRes = vRes;
    p->specLen = doa_spec_len(doa, &p->aSteps, &p->eSteps);
    const size_t B = rc.maxBatch, n = rc.nAnts;
    if (!doa.isUpa && (doa.nAnts != rc.nAnts)) {
        set_error(ctx, "sense_plan_create: ULA numElements must equal the grid's antenna count");
        sense_plan_destroy(p);
        return kErrInvalidArg;
    }
    if (doa.isUpa && doa.nX * doa.nY != rc.nAnts) {
        set_error(ctx, "sense_plan_create: UPA nV*nH must equal the grid's antenna count");
        sense_plan_destroy(p);
        return kErrInvalidArg;
    }
    SALLOC(ctx, p->d_Ra, sizeof(double2) * n * n * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_w, sizeof(double) * n * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_V, sizeof(double2) * n * n * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_L, sizeof(int) * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_P, sizeof(double) * p->specLen * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_PdB, sizeof(double) * p->specLen * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_peakLoc, sizeof(int) * kMaxPeaks * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_nPeaks, sizeof(int) * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_status, sizeof(int) * B, sense_plan_destroy(p));
    SALLOC(ctx, p->d_order, sizeof(int) * n, sense_plan_destroy(p));
    cudaMemset(p->d_status, 0, sizeof(int) * B);
    cudaMemset(p->d_nPeaks, 0, sizeof(int) * B);
    p->detCap = r->nCut < 512 ? r->nCut : 512;
    const size_t pages = n * B;
    const size_t stageBytes = sizeof(int) * (pages + 3 * B + (size_t)kMaxPeaks * B) + (sizeof(int2) + sizeof(float)) * (size_t)p->detCap * pages + 16;  // + alignment pad of the int2 section
    if (cudaMallocHost((void**)&p->h_stage, stageBytes) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming) != cudaSuccess) {
        set_error(ctx, "sense_plan_create: pinned staging allocation failed");
        sense_plan_destroy(p);
        return kErrCuda;
    }
    *out = p;
    return kOk;
}

// pinned staging layout (see SensePlan::h_stage)
struct StageView {
    int *cnt, *L, *nPk, *stt, *pk;
    int2* det;
    float* peak;
};
static StageView stage_view(const SensePlan* p) {
    const size_t B = p->rdm->cfg.maxBatch, pages = (size_t)p->rdm->cfg.nAnts * B;
    StageView v;
    v.cnt = reinterpret_cast<int*>(p->h_stage);
    v.L = v.cnt + pages;
    v.nPk = v.L + B;
    v.stt = v.nPk + B;
    v.pk = v.stt + B;
    v.det = reinterpret_cast<int2*>(v.pk + (size_t)kMaxPeaks * B + ((pages + 3 * B + (size_t)kMaxPeaks * B) & 1));  // 8-byte aligned
    v.peak = reinterpret_cast<float*>(v.det + (size_t)p->detCap * pages);
    return v;
}

// D2H copies of everything collect needs, enqueued behind the chain; `ready` marks their completion
static int stage_results(SensePlan* p, int batch, cudaStream_t st) {
    RdmPlan* r = p->rdm;
    Ctx* ctx = r->ctx;
    const int pages = r->cfg.nAnts * batch;
    StageView v = stage_view(p);
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(v.cnt, r->d_detCount, sizeof(int32_t) * pages, cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(v.L, p->d_L, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
    if (!p->doa.isUpa) {
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(v.nPk, p->d_nPeaks, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(v.stt, p->d_status, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(v.pk, p->d_peakLoc, sizeof(int) * kMaxPeaks * batch, cudaMemcpyDeviceToHost, st));
    }
    // first detCap detections of every page: pages are nCut entries apart on the device, detCap apart in the staging buffer
    ISAC_CUDA_CHECK(ctx, cudaMemcpy2DAsync(v.det, sizeof(int2) * p->detCap, r->d_det, sizeof(int2) * r->nCut, sizeof(int2) * p->detCap,
                                           pages, cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaMemcpy2DAsync(v.peak, sizeof(float) * p->detCap, r->d_peak, sizeof(float) * r->nCut,
                                           sizeof(float) * p->detCap, pages, cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaEventRecord(p->ready, st));
    p->stagedBatch = batch;
    return kOk;
}

void sense_plan_destroy(SensePlan* p) {
    if (!p) return;
    if (p->rdm) rdm_plan_destroy(p->rdm);
    if (p->h_stage) cudaFreeHost(p->h_stage);
    if (p->ready) cudaEventDestroy(p->ready);
    cudaFree(p->d_Ra);
    cudaFree(p->d_w);
    cudaFree(p->d_V);
    cudaFree(p->d_L);
    cudaFree(p->d_P);
    cudaFree(p->d_PdB);
    cudaFree(p->d_peakLoc);
    cudaFree(p->d_nPeaks);
    cudaFree(p->d_status);
    cudaFree(p->d_order);
    delete p;
}

__global__ void popcount_L_kernel(const uint32_t* __restrict__ rowmask, int rowWords, int* __restrict__ L) {
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        int c = 0;
        for (int k = 0; k < rowWords; ++k) c += __popc(rowmask[(long long)b * rowWords + k]);
        L[b] = c;
    }
}

int sense_fft2d_run(SensePlan* p, const float2* rx, const float2* tx, int batch, float* powOut, cudaStream_t st) {
    RdmPlan* r = p->rdm;
    Ctx* ctx = r->ctx;
    int s = rdm_run(r, rx, tx, batch, powOut, st);  // fft2D.m:37-63
    if (s) return s;
    const RdmConfig& c = r->cfg;
    const int n = c.nAnts;
    s = cov_antenna(ctx, rx, (long long)c.nSc * c.nSym, n, batch, p->d_Ra, st);  // fft2D.m:106-107
    if (s) return s;
    // numDets = numel(unique(allRngEst)) = number of distinct detected range rows (fft2D.m:99,110)
    if (!p->doa.isUpa && n <= kSmallEigMax) {
        s = eig_psd_small(ctx, p->d_Ra, n, batch, p->d_w, p->d_V, st);  // music.m:19-29
        if (s) return s;
        LSource ls{};
        ls.rowmask = r->d_rowmask;
        ls.rowWords = r->rowWords;
        s = music_doa_ula(ctx, p->d_w, p->d_V, n, batch, p->doa, ls, p->d_L, p->d_P, p->d_PdB, p->d_peakLoc,
                          p->d_nPeaks, p->d_status, st);  // music.m:73-104
        return s ? s : stage_results(p, batch, st);
    }
    // UPA (music.m:31-63): spectrum only, the reference's peak picker (tools.find2DPeaks) does not exist
    popcount_L_kernel<<<batch, 32, 0, st>>>(r->d_rowmask, r->rowWords, p->d_L);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    for (int b = 0; b < batch; ++b) {
        double2* V = p->d_V + (size_t)b * n * n;
        double* w = p->d_w + (size_t)b * n;
        const int* order = nullptr;
        if (n <= kSmallEigMax) {
            s = eig_psd_small(ctx, p->d_Ra + (size_t)b * n * n, n, 1, w, V, st);
            if (s) return s;
        } else {
            // eigenvectors of the Hermitian PSD Ra = right singular vectors of Ra itself
            void* G = nullptr;
            s = ctx_scratch(ctx, 11, sizeof(double2) * (size_t)n * n, &G);
            if (s) return s;
            ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(G, p->d_Ra + (size_t)b * n * n, sizeof(double2) * (size_t)n * n,
                                                 cudaMemcpyDeviceToDevice, st));
            s = set_identity(ctx, V, n, st);
            if (s) return s;
            void* sig = nullptr;
            s = ctx_scratch(ctx, 12, sizeof(double) * n, &sig);
            if (s) return s;
            s = svd_onesided_jacobi(ctx, (double2*)G, n, n, V, (double*)sig, p->d_order, nullptr, st);
            if (s) return s;
            gather_kernel<<<(n + 255) / 256, 256, 0, st>>>((const double*)sig, p->d_order, n, w, nullptr);
            ISAC_CUDA_CHECK(ctx, cudaGetLastError());
            order = p->d_order;
        }
        s = music_doa_upa(ctx, V, n, order, n, p->doa, p->d_L + b, p->d_P + (size_t)b * p->specLen,
                          p->d_PdB + (size_t)b * p->specLen, st);
        if (s) return s;
    }
    return stage_results(p, batch, st);
}

int sense_fft2d_collect(SensePlan* p, int batch, std::vector<Fft2dResult>& out) {
    RdmPlan* r = p->rdm;
    Ctx* ctx = r->ctx;
    const RdmConfig& c = r->cfg;
    cudaStream_t st = ctx->stream;
    const int nA = c.nAnts;
    if (batch < 1 || batch > p->stagedBatch) {
        set_error(ctx, "fft2D collect: no staged run covers this batch (call isac_fft2d_dev first)");
        return kErrInvalidArg;
    }
    ISAC_CUDA_CHECK(ctx, cudaEventSynchronize(p->ready));  // the staged copies only, not work enqueued behind them
    const StageView v = stage_view(p);
    out.assign(batch, Fft2dResult());
    std::vector<int2> detBig;
    std::vector<float> peakBig;
    for (int b = 0; b < batch; ++b) {
        Fft2dResult& res = out[b];
        std::vector<double> allR, allV;
        for (int a = 0; a < nA; ++a) {
            const int pg = b * nA + a, n = v.cnt[pg];
            if (n <= 0) continue;
            const int2* det = v.det + (size_t)p->detCap * pg;
            const float* peak = v.peak + (size_t)p->detCap * pg;
            if (n > p->detCap) {  // more detections than the staged prefix: fetch this page directly
                detBig.resize(n);
                peakBig.resize(n);
                ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(detBig.data(), r->d_det + (size_t)r->nCut * pg, sizeof(int2) * n,
                                                     cudaMemcpyDeviceToHost, st));
                ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(peakBig.data(), r->d_peak + (size_t)r->nCut * pg, sizeof(float) * n,
                                                     cudaMemcpyDeviceToHost, st));
                ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
                det = detBig.data();
                peak = peakBig.data();
            }
            std::vector<int> idx(n);
            std::iota(idx.begin(), idx.end(), 0);
            std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return peak[x] > peak[y]; });  // fft2D.m:89
            for (int i : idx) {
                allR.push_back((double)(det[i].x - 1) * p->rRes);                          // fft2D.m:77,81
                allV.push_back(((double)det[i].y - (double)c.nFFT / 2.0 - 1.0) * p->vRes);  // fft2D.m:78,82
            }
        }
        auto uniq = [](const std::vector<double>& v) {  // unique(.,'stable')  fft2D.m:99
            std::vector<double> o;
            for (double x : v)
                if (std::find(o.begin(), o.end(), x) == o.end()) o.push_back(x);
            return o;
        };
        res.rngEst = uniq(allR);
        res.velEst = uniq(allV);
        res.L = v.L[b];
        res.status = p->doa.isUpa ? 0 : v.stt[b];
        if (!p->doa.isUpa && res.status == 0)
            for (int i = 0; i < v.nPk[b]; ++i)
                res.aziEst.push_back((double)(v.pk[(size_t)b * kMaxPeaks + i] - 1) * p->doa.aGran - p->doa.aMax / 2.0);  // music.m:103
    }
    return kOk;
}

// ------------------------------------------------------------------------------------------
// doaEstimation.music on a given covariance
// ------------------------------------------------------------------------------------------
int music_doa_run(Ctx* ctx, const DoaConfig& doa, const double2* dRa, int numDets, int* Lout, std::vector<double>& aziEst,
                  std::vector<double>& PdB, std::vector<double>& P, cudaStream_t st, int method) {
    if (method != kDoaMusic && numDets < 1) {  // mvdrBF / digitalBF have no source-count rule: findpeaks needs NPeaks >= 1
        set_error(ctx, "mvdrBF/digitalBF: numDets must be a positive integer (findpeaks 'NPeaks')");
        return kErrNumDetsZero;
    }
    const int n = doa.isUpa ? doa.nX * doa.nY : doa.nAnts;
    int aSteps, eSteps;
    const int spec = doa_spec_len(doa, &aSteps, &eSteps);
    void *w = nullptr, *V = nullptr, *misc = nullptr, *dP = nullptr, *dPdB = nullptr;
    int s;
    if ((s = ctx_scratch(ctx, 2, sizeof(double) * n, &w))) return s;
    if ((s = ctx_scratch(ctx, 3, sizeof(double2) * (size_t)n * n, &V))) return s;
    if ((s = ctx_scratch(ctx, 4, sizeof(int) * (kMaxPeaks + 8 + n), &misc))) return s;
    if ((s = ctx_scratch(ctx, 5, sizeof(double) * spec, &dP))) return s;
    if ((s = ctx_scratch(ctx, 6, sizeof(double) * spec, &dPdB))) return s;
    int* dL = (int*)misc;
    int* dNp = dL + 1;
    int* dSt = dL + 2;
    int* dPk = dL + 8;
    int* dOrder = dPk + kMaxPeaks;
    aziEst.clear();
    if (!doa.isUpa && n <= kSmallEigMax) {
        if ((s = eig_psd_small(ctx, dRa, n, 1, (double*)w, (double2*)V, st))) return s;
        LSource ls{};
        ls.fixedL = numDets > 0 ? numDets : 0;
        if (numDets == 0) {  // explicit zero detections: findpeaks errors in the reference
            set_error(ctx, "music: numDets == 0 (findpeaks 'NPeaks' must be positive)");
            return kErrNumDetsZero;
        }
        if ((s = music_doa_ula(ctx, (double*)w, (double2*)V, n, 1, doa, ls, dL, (double*)dP, (double*)dPdB, dPk, dNp,
                               dSt, st, method)))
            return s;
        int h[8 + kMaxPeaks];
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(h, dL, sizeof(int) * (8 + kMaxPeaks), cudaMemcpyDeviceToHost, st));
        PdB.resize(spec);
        P.resize(spec);
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(PdB.data(), dPdB, sizeof(double) * spec, cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(P.data(), dP, sizeof(double) * spec, cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        *Lout = h[0];
        if (h[2] != 0) {
            set_error(ctx, "music: zero sources");
            return h[2];
        }
        for (int i = 0; i < h[1]; ++i) aziEst.push_back((double)(h[8 + i] - 1) * doa.aGran - doa.aMax / 2.0);
        return kOk;
    }
    // UPA or large array
    const int* order = nullptr;
    if (n <= kSmallEigMax) {
        if ((s = eig_psd_small(ctx, dRa, n, 1, (double*)w, (double2*)V, st))) return s;
    } else {
        void *G = nullptr, *sig = nullptr;
        if ((s = ctx_scratch(ctx, 11, sizeof(double2) * (size_t)n * n, &G))) return s;
        if ((s = ctx_scratch(ctx, 12, sizeof(double) * n, &sig))) return s;
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(G, dRa, sizeof(double2) * (size_t)n * n, cudaMemcpyDeviceToDevice, st));
        if ((s = set_identity(ctx, (double2*)V, n, st))) return s;
        if ((s = svd_onesided_jacobi(ctx, (double2*)G, n, n, (double2*)V, (double*)sig, dOrder, nullptr, st))) return s;
        gather_kernel<<<(n + 255) / 256, 256, 0, st>>>((const double*)sig, dOrder, n, (double*)w, nullptr);
        ISAC_CUDA_CHECK(ctx, cudaGetLastError());
        order = dOrder;
    }
    if (numDets > 0) {
        set_int_kernel<<<1, 32, 0, st>>>(dL, numDets);
    } else if (numDets == 0) {
        set_error(ctx, "music: numDets == 0");
        return kErrNumDetsZero;
    } else {
        if ((s = music_num_targets(ctx, (const double*)w, n, dL, st))) return s;
    }
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    if (!doa.isUpa) {
        set_error(ctx, "music: ULA arrays larger than 64 elements are not supported");
        return kErrUnsupported;
    }
    if ((s = music_doa_upa(ctx, (double2*)V, n, order, n, doa, dL, (double*)dP, (double*)dPdB, st, method, (const double*)w)))
        return s;
    PdB.resize(spec);
    P.resize(spec);
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(PdB.data(), dPdB, sizeof(double) * spec, cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(P.data(), dP, sizeof(double) * spec, cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(Lout, dL, sizeof(int), cudaMemcpyDeviceToHost, st));
    ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    return kOk;
}

// ------------------------------------------------------------------------------------------
// music2D
// ------------------------------------------------------------------------------------------
int music2d_run(Ctx* ctx, const Music2dConfig& c, const float2* rx, const float2* tx, Music2dResult& out,
                cudaStream_t st) {
    const double kC = 299792458.0;
    const int nSc = c.nSc, nSym = c.nSym, nA = c.nAnts;
    const double lambda = kC / c.fc;
    const double rGran = 0.5, vGran = 0.5;
    const double vMax = c.vZone * 2.0;                                  // music2D.m:43
    const int rSteps = (int)std::floor((c.rMax + 1.0) / rGran);        // music2D.m:45
    const int vSteps = (int)std::floor((vMax + 1.0) / vGran);          // music2D.m:46
    int s;
    // --- DoA (music2D.m:57-63) ---
    void* dRa = nullptr;
    if ((s = ctx_scratch(ctx, 7, sizeof(double2) * (size_t)nA * nA, &dRa))) return s;
    if ((s = cov_antenna(ctx, rx, (long long)nSc * nSym, nA, 1, (double2*)dRa, st))) return s;
    std::vector<double> P;
    int L = 0;
    if ((s = music_doa_run(ctx, c.doa, (const double2*)dRa, c.numDetsOverride > 0 ? c.numDetsOverride : -1, &L,
                           out.aziEst, out.PdoadB, P, st)))
        return s;
    out.L = L;
    // --- H and its SVD by one-sided Jacobi on the thin orientation (music2D.m:67-89) ---
    const bool tall = nSc >= nSym;
    const int m = tall ? nSc : nSym, n = tall ? nSym : nSc;
    void *G = nullptr, *V = nullptr, *sig = nullptr, *inv = nullptr, *misc = nullptr, *q = nullptr, *dP = nullptr, *dPdB = nullptr;
    if ((s = ctx_scratch(ctx, 11, sizeof(double2) * (size_t)m * n, &G))) return s;
    if ((s = ctx_scratch(ctx, 3, sizeof(double2) * (size_t)n * n, &V))) return s;
    if ((s = ctx_scratch(ctx, 12, sizeof(double) * n, &sig))) return s;
    if ((s = ctx_scratch(ctx, 13, sizeof(double) * n, &inv))) return s;
    if ((s = ctx_scratch(ctx, 4, sizeof(int) * (kMaxPeaks + 8 + n), &misc))) return s;
    const int maxSteps = rSteps > vSteps ? rSteps : vSteps;
    if ((s = ctx_scratch(ctx, 14, sizeof(double) * maxSteps, &q))) return s;
    if ((s = ctx_scratch(ctx, 5, sizeof(double) * maxSteps, &dP))) return s;
    if ((s = ctx_scratch(ctx, 6, sizeof(double) * maxSteps, &dPdB))) return s;
    int* dL = (int*)misc;
    int* dNp = dL + 1;
    int* dPk = dL + 8;
    int* dOrder = dPk + kMaxPeaks;
    if ((s = music2d_channel(ctx, rx, tx, nSc, nSym, tall ? 0 : 1, (double2*)G, st))) return s;
    if ((s = set_identity(ctx, (double2*)V, n, st))) return s;
    if ((s = svd_onesided_jacobi(ctx, (double2*)G, m, n, (double2*)V, (double*)sig, dOrder, &out.sweeps, st))) return s;
    gather_kernel<<<(n + 255) / 256, 256, 0, st>>>((const double*)sig, dOrder, n, nullptr, (double*)inv);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    set_int_kernel<<<1, 32, 0, st>>>(dL, L);
    ISAC_CUDA_CHECK(ctx, cudaGetLastError());
    // left singular vectors of H span Rr's signal space; conj(right singular vectors) span Rv's
    const double2* leftV = tall ? (const double2*)G : (const double2*)V;
    const double* leftScale = tall ? (const double*)inv : nullptr;
    const long long leftLd = tall ? m : n;
    const double2* rightV = tall ? (const double2*)V : (const double2*)G;
    const double* rightScale = tall ? nullptr : (const double*)inv;
    const long long rightLd = tall ? n : m;
    auto finish = [&](int steps, std::vector<double>& Pv, std::vector<double>& PdBv, std::vector<double>& est, double x0,
                      double dx) -> int {
        int s2 = music_finish_1d(ctx, (const double*)q, steps, dL, (double*)dP, (double*)dPdB, dPk, dNp, st);
        if (s2) return s2;
        int h[8 + kMaxPeaks];
        Pv.resize(steps);
        PdBv.resize(steps);
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(h, dL, sizeof(int) * (8 + kMaxPeaks), cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(Pv.data(), dP, sizeof(double) * steps, cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaMemcpyAsync(PdBv.data(), dPdB, sizeof(double) * steps, cudaMemcpyDeviceToHost, st));
        ISAC_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        est.clear();
        for (int i = 0; i < h[1]; ++i) est.push_back((double)(h[8 + i] - 1) * dx + x0);
        return kOk;
    };
    // range scan (music2D.m:92,98-102): a_r[n] = exp(-2j*pi*scs*2*r*n/c)
    if ((s = music_scan_1d(ctx, leftV, leftLd, dOrder, leftScale, nSc, n, 0, dL, -2.0 * c.scsHz / kC, 0.0, rGran, rSteps,
                           (double*)q, st)))
        return s;
    if ((s = finish(rSteps, out.Pr, out.PrdB, out.rngEst, 0.0, rGran))) return s;                 // music2D.m:122
    // velocity scan (music2D.m:93,104-108): a_v[m] = exp(2j*pi*T*2*v*m/lambda)
    if ((s = music_scan_1d(ctx, rightV, rightLd, dOrder, rightScale, nSym, n, 1, dL, 2.0 * c.Tsri / lambda, -vMax / 2.0,
                           vGran, vSteps, (double*)q, st)))
        return s;
    if ((s = finish(vSteps, out.Pv, out.PvdB, out.velEst, -vMax / 2.0, vGran))) return s;        // music2D.m:123
    return kOk;
}

}  // namespace isac
