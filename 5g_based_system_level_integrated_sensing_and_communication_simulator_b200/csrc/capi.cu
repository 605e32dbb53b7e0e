// C ABI (include/isac_b200.h) over the internal C++/CUDA implementation.
#include "../../include/isac_b200.h"
#include "isac_common.cuh"
#include "ctx.cuh"
#include "rdm.cuh"
#include "sense.cuh"
#include "echo.cuh"
#include "ofdm.cuh"
#include "chest.cuh"
#include "los.cuh"
#include "comm.cuh"
#include "cdl.cuh"
#include "link.cuh"
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace isac {

void set_error(Ctx* ctx, const std::string& msg) {
    if (ctx) ctx->err = msg;
}
void ctx_fft_tw(Ctx* ctx, int N, const float2** tw1, const float2** tw2) {
    int i = 0;
    while ((16 << i) < N) ++i;
    *tw1 = ctx->d_twiddle + ctx->tw1Off[i];
    *tw2 = ctx->d_twiddle + ctx->tw2Off[i];
}

static cudaEvent_t take_event(Ctx* ctx) {
    if (!ctx->eventPool.empty()) {
        cudaEvent_t e = ctx->eventPool.back();
        ctx->eventPool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
int prof_begin(Ctx* ctx, int slot, cudaStream_t st) {
    if (!ctx->profiling) return -1;
    Ctx::ProfRec r{slot, take_event(ctx), take_event(ctx)};
    cudaEventRecord(r.a, st);
    ctx->prof.push_back(r);
    return (int)ctx->prof.size() - 1;
}
void prof_end(Ctx* ctx, int rec, cudaStream_t st) {
    if (rec >= 0 && rec < (int)ctx->prof.size()) cudaEventRecord(ctx->prof[rec].b, st);
}
void count_launches(Ctx* ctx, int n) { ctx->launches += n; }
int ctx_num_sms(Ctx* ctx) { return ctx->numSMs; }

int ctx_pinned(Ctx* ctx, int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= Ctx::kPinnedSlots) return kErrInvalidArg;
    if (ctx->pinnedBytes[slot] < bytes) {
        if (ctx->pinned[slot]) cudaFreeHost(ctx->pinned[slot]);
        ctx->pinned[slot] = nullptr;
        ctx->pinnedBytes[slot] = 0;
        ISAC_CUDA_CHECK(ctx, cudaMallocHost(&ctx->pinned[slot], bytes));
        ctx->pinnedBytes[slot] = bytes;
    }
    *out = ctx->pinned[slot];
    return kOk;
}

int ctx_scratch(Ctx* ctx, int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= Ctx::kScratchSlots) return kErrInvalidArg;
    if (ctx->scratchBytes[slot] < bytes) {
        if (ctx->scratch[slot]) cudaFree(ctx->scratch[slot]);
        ctx->scratch[slot] = nullptr;
        ctx->scratchBytes[slot] = 0;
        ISAC_CUDA_CHECK(ctx, cudaMalloc(&ctx->scratch[slot], bytes));
        ctx->scratchBytes[slot] = bytes;
    }
    *out = ctx->scratch[slot];
    return kOk;
}

}  // namespace isac

using namespace isac;

struct isac_ctx {
    Ctx c;
};
struct isac_rdm_plan {
    RdmPlan* p;
};
struct isac_city {
    CityPlan* p;
};
struct isac_sense_plan {
    SensePlan* p;
    isac_rdm_plan rdmView;  // non-owning view handed out by isac_sense_plan_rdm
};

static RdmConfig to_rdm_config(const isac_rdm_config* cfg) {
    RdmConfig c{};
    c.nSc = cfg->nSc; c.nSym = cfg->nSym; c.nAnts = cfg->nAnts; c.nIFFT = cfg->nIFFT; c.nFFT = cfg->nFFT;
    c.cutRow0 = cfg->cutRow0; c.cutRow1 = cfg->cutRow1; c.cutCol0 = cfg->cutCol0; c.cutCol1 = cfg->cutCol1;
    c.guardRows = cfg->guardRows; c.guardCols = cfg->guardCols;
    c.trainRows = cfg->trainRows; c.trainCols = cfg->trainCols;
    c.maxBatch = cfg->maxBatch; c.pfa = cfg->pfa; c.kaiserBeta = cfg->kaiserBeta;
    return c;
}
static DoaConfig to_doa_config(const isac_doa_config* d) {
    DoaConfig c{};
    c.isUpa = d->isUpa; c.nAnts = d->nAnts; c.nX = d->nX; c.nY = d->nY; c.d = d->d;
    c.aGran = d->aGran; c.aMax = d->aMax; c.eGran = d->eGran; c.eMax = d->eMax;
    return c;
}

static thread_local std::string g_createError;

extern "C" {

const char* isac_version(void) { return "isac_b200 0.1.0 (sm_100a)"; }

int isac_create(isac_ctx** out, int device) {
    if (!out) return ISAC_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_createError = "isac_create: no CUDA device (this library has no CPU fallback)";
        return ISAC_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        g_createError = "isac_create: device index out of range";
        return ISAC_ERR_INVALID_ARG;
    }
    isac_ctx* h = new (std::nothrow) isac_ctx();
    if (!h) return ISAC_ERR_CUDA;
    Ctx* c = &h->c;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        g_createError = "isac_create: cudaSetDevice failed";
        delete h;
        return ISAC_ERR_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->numSMs = prop.multiProcessorCount;
    c->ccMajor = prop.major;
    if (prop.major < 10) {
        g_createError = "isac_create: this build targets sm_100a (Blackwell B200) only";
        delete h;
        return ISAC_ERR_UNSUPPORTED;
    }
    if (cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking) != cudaSuccess) {
        g_createError = "isac_create: cudaStreamCreate failed";
        delete h;
        return ISAC_ERR_CUDA;
    }
    c->stream = c->ownStream;
    // twiddle tables for N = 16..4096 (N = R1*R2*16: N <= 256 -> R1 = 1, R2 = N/16; else R1 = N/256, R2 = 16)
    std::vector<float2> tw;
    for (int i = 0; i < 9; ++i) {
        const int N = 16 << i;
        const int R1 = N <= 256 ? 1 : N / 256, R2 = N <= 256 ? N / 16 : 16, N2 = N / R1;
        c->tw1Off[i] = tw.size();
        for (int k1 = 1; k1 < R1; ++k1)
            for (int m = 0; m < N2; ++m) {
                const double a = 2.0 * M_PI * (double)m * (double)k1 / (double)N;
                tw.push_back(make_float2((float)std::cos(a), (float)std::sin(a)));
            }
        c->tw2Off[i] = tw.size();
        for (int cc = 1; cc < R2; ++cc)
            for (int b = 0; b < 16; ++b) {
                const double a = 2.0 * M_PI * (double)b * (double)cc / (double)N2;
                tw.push_back(make_float2((float)std::cos(a), (float)std::sin(a)));
            }
        tw.push_back(make_float2(1.f, 0.f));  // keep the table pointers valid when a pass has no twiddles
    }
    if (cudaMalloc((void**)&c->d_twiddle, sizeof(float2) * tw.size()) != cudaSuccess ||
        cudaMemcpy(c->d_twiddle, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        g_createError = "isac_create: twiddle table upload failed";
        delete h;
        return ISAC_ERR_CUDA;
    }
    *out = h;
    return ISAC_OK;
}

int isac_destroy(isac_ctx* h) {
    if (!h) return ISAC_OK;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < Ctx::kPinnedSlots; ++i)
        if (c->pinned[i]) cudaFreeHost(c->pinned[i]);
    for (int i = 0; i < Ctx::kScratchSlots; ++i)
        if (c->scratch[i]) cudaFree(c->scratch[i]);
    if (c->d_twiddle) cudaFree(c->d_twiddle);
    if (c->ulFree) c->ulFree(c->ulState);
    for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : c->eventPool) cudaEventDestroy(e);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete h;
    return ISAC_OK;
}

const char* isac_last_error(const isac_ctx* h) {
    if (!h) return g_createError.c_str();
    return h->c.err.c_str();
}

int isac_set_stream(isac_ctx* h, void* s) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    h->c.stream = (cudaStream_t)s;  // NULL == the CUDA legacy default stream
    return ISAC_OK;
}

int isac_use_own_stream(isac_ctx* h) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    h->c.stream = h->c.ownStream;
    return ISAC_OK;
}

int isac_profile_enable(isac_ctx* h, int32_t on) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    h->c.profiling = on != 0;
    return ISAC_OK;
}

int isac_profile_collect(isac_ctx* h, double* msPerSlot, int32_t* countPerSlot, int64_t* launches) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < kProfSlots; ++i) {
        if (msPerSlot) msPerSlot[i] = 0.0;
        if (countPerSlot) countPerSlot[i] = 0;
    }
    for (auto& r : c->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.slot >= 0 && r.slot < kProfSlots) {
            if (msPerSlot) msPerSlot[r.slot] += (double)ms;
            if (countPerSlot) countPerSlot[r.slot] += 1;
        }
        c->eventPool.push_back(r.a);
        c->eventPool.push_back(r.b);
    }
    c->prof.clear();
    if (launches) *launches = c->launches;
    c->launches = 0;
    return ISAC_OK;
}

int isac_profile_timeline(isac_ctx* h, void* baseEvent, int32_t maxRec, int32_t* slots, double* beginMs, double* endMs, int32_t* nRec) {
    if (!h || !baseEvent || !slots || !beginMs || !endMs || !nRec || maxRec < 0) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    int n = 0;
    for (auto& r : c->prof) {
        if (n >= maxRec) break;
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, (cudaEvent_t)baseEvent, r.a) != cudaSuccess ||
            cudaEventElapsedTime(&b, (cudaEvent_t)baseEvent, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
        slots[n] = r.slot;
        beginMs[n] = (double)a;
        endMs[n] = (double)b;
        ++n;
    }
    *nRec = n;
    return ISAC_OK;
}

int isac_synchronize(isac_ctx* h) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    ISAC_CUDA_CHECK(&h->c, cudaStreamSynchronize(h->c.stream));
    return ISAC_OK;
}

// ---- RDM + CFAR ------------------------------------------------------------------------------
int isac_rdm_plan_create(isac_ctx* h, const isac_rdm_config* cfg, isac_rdm_plan** out) {
    if (!h || !cfg || !out) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    RdmConfig c = to_rdm_config(cfg);
    RdmPlan* p = nullptr;
    int st = rdm_plan_create(&h->c, c, &p);
    if (st != kOk) return st;
    *out = new isac_rdm_plan{p};
    return ISAC_OK;
}

int isac_rdm_plan_destroy(isac_rdm_plan* pl) {
    if (!pl) return ISAC_OK;
    if (pl->p) {
        cudaSetDevice(pl->p->ctx->device);
        cudaStreamSynchronize(pl->p->ctx->stream);
        rdm_plan_destroy(pl->p);
    }
    delete pl;
    return ISAC_OK;
}

int isac_rdm_plan_info(const isac_rdm_plan* pl, double* alpha, int32_t* nTrain, int32_t* nCut) {
    if (!pl || !pl->p) return ISAC_ERR_INVALID_ARG;
    if (alpha) *alpha = pl->p->alpha;
    if (nTrain) *nTrain = pl->p->nTrain;
    if (nCut) *nCut = pl->p->nCut;
    return ISAC_OK;
}

int isac_rdm_plan_set_variant(isac_rdm_plan* pl, int32_t variant) {
    if (!pl || !pl->p || variant < 0 || variant > 3) return ISAC_ERR_INVALID_ARG;
    pl->p->variant = variant;
    return ISAC_OK;
}

int isac_rdm_cfar_dev(isac_rdm_plan* pl, const void* rx, const void* tx, int32_t batch, float* rdPower) {
    if (!pl || !pl->p) return ISAC_ERR_INVALID_ARG;
    Ctx* c = pl->p->ctx;
    cudaSetDevice(c->device);
    return rdm_run(pl->p, (const float2*)rx, (const float2*)tx, batch, rdPower, c->stream);
}

int isac_cfar2d_dev(isac_rdm_plan* pl, const float* rdPower, int32_t batch) {
    if (!pl || !pl->p || !rdPower) return ISAC_ERR_INVALID_ARG;
    Ctx* c = pl->p->ctx;
    cudaSetDevice(c->device);
    return rdm_cfar_only(pl->p, rdPower, batch, c->stream);
}

int isac_rdm_get_detections(isac_rdm_plan* pl, int32_t batch, int32_t maxDet, int32_t* detCount,
                            int32_t* detRowCol, float* peaks) {
    if (!pl || !pl->p || !detCount) return ISAC_ERR_INVALID_ARG;
    RdmPlan* p = pl->p;
    Ctx* c = p->ctx;
    cudaSetDevice(c->device);
    if (batch < 1 || batch > p->lastBatch) {
        set_error(c, "isac_rdm_get_detections: batch exceeds the last run");
        return ISAC_ERR_INVALID_ARG;
    }
    const int pages = p->cfg.nAnts * batch;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(detCount, p->d_detCount, sizeof(int32_t) * pages, cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    if (!detRowCol && !peaks) return ISAC_OK;
    int status = ISAC_OK;
    for (int pg = 0; pg < pages; ++pg) {
        int n = detCount[pg];
        if (n > maxDet) {
            n = maxDet;
            status = ISAC_ERR_CAPACITY;
            set_error(c, "isac_rdm_get_detections: maxDet smaller than the number of detections");
        }
        if (n <= 0) continue;
        if (detRowCol)
            ISAC_CUDA_CHECK(c, cudaMemcpyAsync(detRowCol + (size_t)2 * maxDet * pg, p->d_det + (size_t)p->nCut * pg,
                                               sizeof(int2) * n, cudaMemcpyDeviceToHost, c->stream));
        if (peaks)
            ISAC_CUDA_CHECK(c, cudaMemcpyAsync(peaks + (size_t)maxDet * pg, p->d_peak + (size_t)p->nCut * pg,
                                               sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return status;
}

int isac_rdm_get_power(isac_rdm_plan* pl, int32_t batch, float* host) {
    if (!pl || !pl->p || !host) return ISAC_ERR_INVALID_ARG;
    RdmPlan* p = pl->p;
    Ctx* c = p->ctx;
    cudaSetDevice(c->device);
    if (batch < 1 || batch > p->lastBatch || !p->lastPow) {
        set_error(c, "isac_rdm_get_power: no power map for that batch");
        return ISAC_ERR_INVALID_ARG;
    }
    const size_t n = (size_t)p->cfg.nIFFT * p->cfg.nFFT * p->cfg.nAnts * batch;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(host, p->lastPow, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return ISAC_OK;
}

int isac_rdm_cfar_host(isac_rdm_plan* pl, const void* rxHost, const void* txHost, int32_t batch, int32_t maxDet,
                       int32_t* detCount, int32_t* detRowCol, float* peaks, float* rdPowerHost) {
    if (!pl || !pl->p || !rxHost || !txHost || !detCount) return ISAC_ERR_INVALID_ARG;
    RdmPlan* p = pl->p;
    Ctx* c = p->ctx;
    cudaSetDevice(c->device);
    if (batch < 1 || batch > p->cfg.maxBatch) {
        set_error(c, "isac_rdm_cfar_host: batch out of range");
        return ISAC_ERR_INVALID_ARG;
    }
    const size_t gridBytes = sizeof(float2) * (size_t)p->cfg.nSc * p->cfg.nSym * p->cfg.nAnts * batch;
    void *dRx = nullptr, *dTx = nullptr;
    int st = ctx_scratch(c, 0, gridBytes, &dRx);
    if (st) return st;
    st = ctx_scratch(c, 1, gridBytes, &dTx);
    if (st) return st;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dRx, rxHost, gridBytes, cudaMemcpyHostToDevice, c->stream));
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dTx, txHost, gridBytes, cudaMemcpyHostToDevice, c->stream));
    st = rdm_run(p, (const float2*)dRx, (const float2*)dTx, batch, nullptr, c->stream);
    if (st) return st;
    st = isac_rdm_get_detections(pl, batch, maxDet, detCount, detRowCol, peaks);
    if (st) return st;
    if (rdPowerHost) return isac_rdm_get_power(pl, batch, rdPowerHost);
    return ISAC_OK;
}

// ---- MUSIC / fft2D / music2D -------------------------------------------------------------------
static bool doa_valid(Ctx* c, const isac_doa_config* d) {
    if (!d || d->aGran <= 0 || d->aMax <= 0 || (d->isUpa && (d->eGran <= 0 || d->eMax <= 0 || d->nX < 1 || d->nY < 1)) ||
        (!d->isUpa && d->nAnts < 2)) {
        set_error(c, "invalid isac_doa_config");
        return false;
    }
    return true;
}

int isac_music_doa_host(isac_ctx* h, const isac_doa_config* doa, const double* Ra, int32_t numDets, int32_t* L,
                        double* aziEst, int32_t* nAzi, double* PmusicdB, double* Pmusic) {
    return isac_doa_scan_host(h, doa, ISAC_DOA_MUSIC, Ra, numDets, L, aziEst, nAzi, PmusicdB, Pmusic);
}

int isac_doa_scan_host(isac_ctx* h, const isac_doa_config* doa, int32_t method, const double* Ra, int32_t numDets, int32_t* L,
                       double* aziEst, int32_t* nAzi, double* PmusicdB, double* Pmusic) {
    if (!h || !Ra || !L || !nAzi || method < ISAC_DOA_MUSIC || method > ISAC_DOA_DBF) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    if (!doa_valid(c, doa)) return ISAC_ERR_INVALID_ARG;
    DoaConfig d = to_doa_config(doa);
    const int n = d.isUpa ? d.nX * d.nY : d.nAnts;
    void* dRa = nullptr;
    int st = ctx_scratch(c, 7, sizeof(double2) * (size_t)n * n, &dRa);
    if (st) return st;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dRa, Ra, sizeof(double2) * (size_t)n * n, cudaMemcpyHostToDevice, c->stream));
    std::vector<double> azi, PdB, P;
    int Lh = 0;
    st = music_doa_run(c, d, (const double2*)dRa, numDets, &Lh, azi, PdB, P, c->stream, method);
    *L = Lh;
    if (st) return st;
    *nAzi = (int32_t)azi.size();
    if (aziEst) std::memcpy(aziEst, azi.data(), sizeof(double) * azi.size());
    if (PmusicdB) std::memcpy(PmusicdB, PdB.data(), sizeof(double) * PdB.size());
    if (Pmusic) std::memcpy(Pmusic, P.data(), sizeof(double) * P.size());
    return ISAC_OK;
}

int isac_sense_plan_create(isac_ctx* h, const isac_rdm_config* rdm, const isac_doa_config* doa, double rRes,
                           double vRes, isac_sense_plan** out) {
    if (!h || !rdm || !out) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    if (!doa_valid(&h->c, doa)) return ISAC_ERR_INVALID_ARG;
    SensePlan* p = nullptr;
    int st = sense_plan_create(&h->c, to_rdm_config(rdm), to_doa_config(doa), rRes, vRes, &p);
    if (st) return st;
    isac_sense_plan* sp = new isac_sense_plan{p, {p->rdm}};
    *out = sp;
    return ISAC_OK;
}

int isac_sense_plan_destroy(isac_sense_plan* sp) {
    if (!sp) return ISAC_OK;
    if (sp->p) {
        Ctx* c = sp->p->rdm->ctx;
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        sense_plan_destroy(sp->p);
    }
    delete sp;
    return ISAC_OK;
}

isac_rdm_plan* isac_sense_plan_rdm(isac_sense_plan* sp) { return sp ? &sp->rdmView : nullptr; }

int isac_fft2d_dev(isac_sense_plan* sp, const void* rx, const void* tx, int32_t batch, float* rdPower) {
    if (!sp || !sp->p) return ISAC_ERR_INVALID_ARG;
    Ctx* c = sp->p->rdm->ctx;
    cudaSetDevice(c->device);
    return sense_fft2d_run(sp->p, (const float2*)rx, (const float2*)tx, batch, rdPower, c->stream);
}

int isac_fft2d_collect(isac_sense_plan* sp, int32_t batch, int32_t maxOut, double* rngEst, int32_t* nRng, double* velEst,
                       int32_t* nVel, double* aziEst, int32_t* nAzi, int32_t* L, int32_t* status) {
    if (!sp || !sp->p || !nRng || !nVel) return ISAC_ERR_INVALID_ARG;
    Ctx* c = sp->p->rdm->ctx;
    cudaSetDevice(c->device);
    if (batch < 1 || batch > sp->p->rdm->lastBatch) {
        set_error(c, "isac_fft2d_collect: batch exceeds the last run");
        return ISAC_ERR_INVALID_ARG;
    }
    std::vector<Fft2dResult> res;
    int st = sense_fft2d_collect(sp->p, batch, res);
    if (st) return st;
    int ret = ISAC_OK;
    for (int b = 0; b < batch; ++b) {
        const Fft2dResult& r = res[b];
        int nr = (int)r.rngEst.size(), nv = (int)r.velEst.size();
        if (nr > maxOut || nv > maxOut) {
            ret = ISAC_ERR_CAPACITY;
            set_error(c, "isac_fft2d_collect: maxOut too small");
            if (nr > maxOut) nr = maxOut;
            if (nv > maxOut) nv = maxOut;
        }
        nRng[b] = nr;
        nVel[b] = nv;
        if (rngEst) std::memcpy(rngEst + (size_t)b * maxOut, r.rngEst.data(), sizeof(double) * nr);
        if (velEst) std::memcpy(velEst + (size_t)b * maxOut, r.velEst.data(), sizeof(double) * nv);
        if (nAzi) nAzi[b] = (int32_t)r.aziEst.size();
        if (aziEst) std::memcpy(aziEst + (size_t)b * ISAC_MAX_PEAKS, r.aziEst.data(), sizeof(double) * r.aziEst.size());
        if (L) L[b] = r.L;
        if (status) status[b] = r.status;
    }
    return ret;
}

int isac_fft2d_get_spectrum(isac_sense_plan* sp, int32_t batch, double* PdB) {
    if (!sp || !sp->p || !PdB) return ISAC_ERR_INVALID_ARG;
    Ctx* c = sp->p->rdm->ctx;
    cudaSetDevice(c->device);
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(PdB, sp->p->d_PdB, sizeof(double) * (size_t)sp->p->specLen * batch,
                                       cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return ISAC_OK;
}

int isac_fft2d_host(isac_sense_plan* sp, const void* rxHost, const void* txHost, int32_t batch, int32_t maxOut,
                    double* rngEst, int32_t* nRng, double* velEst, int32_t* nVel, double* aziEst, int32_t* nAzi,
                    int32_t* L, int32_t* status) {
    if (!sp || !sp->p || !rxHost || !txHost) return ISAC_ERR_INVALID_ARG;
    RdmPlan* p = sp->p->rdm;
    Ctx* c = p->ctx;
    cudaSetDevice(c->device);
    if (batch < 1 || batch > p->cfg.maxBatch) {
        set_error(c, "isac_fft2d_host: batch out of range");
        return ISAC_ERR_INVALID_ARG;
    }
    const size_t gridBytes = sizeof(float2) * (size_t)p->cfg.nSc * p->cfg.nSym * p->cfg.nAnts * batch;
    void *dRx = nullptr, *dTx = nullptr;
    int st = ctx_scratch(c, 0, gridBytes, &dRx);
    if (st) return st;
    st = ctx_scratch(c, 1, gridBytes, &dTx);
    if (st) return st;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dRx, rxHost, gridBytes, cudaMemcpyHostToDevice, c->stream));
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dTx, txHost, gridBytes, cudaMemcpyHostToDevice, c->stream));
    st = sense_fft2d_run(sp->p, (const float2*)dRx, (const float2*)dTx, batch, nullptr, c->stream);
    if (st) return st;
    return isac_fft2d_collect(sp, batch, maxOut, rngEst, nRng, velEst, nVel, aziEst, nAzi, L, status);
}

int isac_music2d_dev(isac_ctx* h, const isac_music2d_config* cfg, const void* rx, const void* tx, int32_t* L,
                     double* aziEst, int32_t* nAzi, double* rngEst, int32_t* nRng, double* velEst, int32_t* nVel,
                     double* PrdB, double* PvdB, int32_t* sweeps) {
    if (!h || !cfg || !rx || !tx || !L) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    if (!doa_valid(c, &cfg->doa)) return ISAC_ERR_INVALID_ARG;
    if (cfg->nSc < 2 || cfg->nSym < 2 || cfg->nAnts < 1) {
        set_error(c, "isac_music2d_dev: invalid grid size");
        return ISAC_ERR_INVALID_ARG;
    }
    Music2dConfig m{};
    m.nSc = cfg->nSc; m.nSym = cfg->nSym; m.nAnts = cfg->nAnts;
    m.scsHz = cfg->scsHz; m.fc = cfg->fc; m.Tsri = cfg->Tsri; m.rMax = cfg->rMax; m.vZone = cfg->vZone;
    m.doa = to_doa_config(&cfg->doa);
    m.numDetsOverride = cfg->numDetsOverride;
    Music2dResult r;
    int st = music2d_run(c, m, (const float2*)rx, (const float2*)tx, r, c->stream);
    *L = r.L;
    if (st) return st;
    auto put = [](const std::vector<double>& v, double* dst, int32_t* n) {
        if (n) *n = (int32_t)v.size();
        if (dst) std::memcpy(dst, v.data(), sizeof(double) * v.size());
    };
    put(r.aziEst, aziEst, nAzi);
    put(r.rngEst, rngEst, nRng);
    put(r.velEst, velEst, nVel);
    if (PrdB) std::memcpy(PrdB, r.PrdB.data(), sizeof(double) * r.PrdB.size());
    if (PvdB) std::memcpy(PvdB, r.PvdB.data(), sizeof(double) * r.PvdB.size());
    if (sweeps) *sweeps = r.sweeps;
    return ISAC_OK;
}

int isac_antenna_covariance_dev(isac_ctx* h, const void* rx, int64_t nScSym, int32_t nAnts, double* RaHost) {
    if (!h || !rx || !RaHost) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    void* dRa = nullptr;
    int st = ctx_scratch(c, 7, sizeof(double2) * (size_t)nAnts * nAnts, &dRa);
    if (st) return st;
    st = cov_antenna(c, (const float2*)rx, nScSym, nAnts, 1, (double2*)dRa, c->stream);
    if (st) return st;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(RaHost, dRa, sizeof(double2) * (size_t)nAnts * nAnts, cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return ISAC_OK;
}

// ---- echo synthesis + OFDM demodulation --------------------------------------------------------
static bool echo_cfg(Ctx* c, const isac_echo_config* cfg, EchoConfig& e) {
    if (!cfg || cfg->T < 1 || cfg->nTx < 1 || cfg->nTargets < 0 || !cfg->range || !cfg->velocity ||
        !cfg->largeScaleFading || !cfg->steeringVec) {
        set_error(c, "invalid isac_echo_config");
        return false;
    }
    e = EchoConfig{};
    e.T = cfg->T; e.nTx = cfg->nTx; e.nTargets = cfg->nTargets; e.fc = cfg->fc; e.fs = cfg->fs; e.N0 = cfg->N0;
    e.range = cfg->range; e.velocity = cfg->velocity; e.largeScaleFading = cfg->largeScaleFading;
    e.steeringVec = cfg->steeringVec; e.los = cfg->los; e.nfft = cfg->nfft; e.nSc = cfg->nSc; e.nSymTx = cfg->nSymTx;
    e.symbolsPerSubframe = cfg->symbolsPerSubframe; e.cpLengths = cfg->cpLengths;
    return true;
}

int isac_radar_channel_dev(isac_ctx* h, const isac_echo_config* cfg, const void* tx, const void* noise, int32_t noiseMode,
                           uint64_t seed, void* rxWave) {
    if (!h || !tx || !rxWave) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    EchoConfig e;
    if (!echo_cfg(c, cfg, e)) return ISAC_ERR_INVALID_ARG;
    return radar_channel_run(c, e, (const float2*)tx, (const float2*)noise, noiseMode, seed, (float2*)rxWave, c->stream);
}

int isac_mono_static_sensing_dev(isac_ctx* h, const isac_echo_config* cfg, const void* tx, const void* noise,
                                 int32_t noiseMode, uint64_t seed, void* echoGrid, int32_t* nSymOut) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    EchoConfig e;
    if (!echo_cfg(c, cfg, e) || !cfg->cpLengths) return ISAC_ERR_INVALID_ARG;
    int n = 0;
    int st = mono_static_sensing_run(c, e, (const float2*)tx, (const float2*)noise, noiseMode, seed, (float2*)echoGrid, &n,
                                     c->stream);
    if (nSymOut) *nSymOut = n;
    return st;
}

// ---- device-memory helpers for gateways that drive `_dev` entry points (MEX: no CUDA runtime in the gateway) --------
int isac_dev_malloc(isac_ctx* h, uint64_t bytes, void** dptr) {
    if (!h || !dptr) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    ISAC_CUDA_CHECK(&h->c, cudaMalloc(dptr, bytes ? (size_t)bytes : 1));
    return ISAC_OK;
}

int isac_dev_free(isac_ctx* h, void* dptr) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    ISAC_CUDA_CHECK(&h->c, cudaStreamSynchronize(h->c.stream));
    ISAC_CUDA_CHECK(&h->c, cudaFree(dptr));
    return ISAC_OK;
}

int isac_memcpy_h2d(isac_ctx* h, void* dst, const void* src, uint64_t bytes) {
    if (!h || (!dst && bytes) || (!src && bytes)) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    ISAC_CUDA_CHECK(&h->c, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, h->c.stream));
    ISAC_CUDA_CHECK(&h->c, cudaStreamSynchronize(h->c.stream));   // the host buffer may be pageable / reused by the caller
    return ISAC_OK;
}

int isac_memcpy_d2h(isac_ctx* h, void* dst, const void* src, uint64_t bytes) {
    if (!h || (!dst && bytes) || (!src && bytes)) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    ISAC_CUDA_CHECK(&h->c, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, h->c.stream));
    ISAC_CUDA_CHECK(&h->c, cudaStreamSynchronize(h->c.stream));
    return ISAC_OK;
}

// ---- city layout: LoS / blockage ---------------------------------------------------------------
int isac_city_create(isac_ctx* h, int32_t nWalls, const int32_t* wallOffsets, const double* corners, isac_city** out) {
    if (!h || !out) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    CityPlan* p = nullptr;
    int st = city_plan_create(&h->c, nWalls, wallOffsets, corners, &p);
    if (st) return st;
    *out = new isac_city{p};
    return ISAC_OK;
}

int isac_city_destroy(isac_city* c) {
    if (!c) return ISAC_OK;
    if (c->p) {
        cudaSetDevice(c->p->ctx->device);
        cudaStreamSynchronize(c->p->ctx->stream);
        city_plan_destroy(c->p);
    }
    delete c;
    return ISAC_OK;
}

int isac_city_check_los_dev(isac_city* c, int32_t nLinks, const double* uePos, const double* antPos, int32_t nAnt, int32_t* los) {
    if (!c || !c->p || (nAnt != 1 && nAnt != nLinks)) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(c->p->ctx->device);
    return city_check_los(c->p, nLinks, uePos, antPos, nAnt == 1 ? 0 : 3, los, c->p->ctx->stream);
}

int isac_city_check_los_host(isac_city* c, int32_t nLinks, const double* uePos, const double* antPos, int32_t nAnt, int32_t* los) {
    if (!c || !c->p || !uePos || !antPos || !los || nLinks < 1 || (nAnt != 1 && nAnt != nLinks)) return ISAC_ERR_INVALID_ARG;
    Ctx* x = c->p->ctx;
    cudaSetDevice(x->device);
    const size_t bU = sizeof(double) * 3 * (size_t)nLinks, bA = sizeof(double) * 3 * (size_t)nAnt, bL = sizeof(int32_t) * (size_t)nLinks;
    void* d = nullptr;
    int st = ctx_scratch(x, 18, bU + bA + bL, &d);
    if (st) return st;
    double* dU = (double*)d;
    double* dA = (double*)((char*)d + bU);
    int* dL = (int*)((char*)d + bU + bA);
    ISAC_CUDA_CHECK(x, cudaMemcpyAsync(dU, uePos, bU, cudaMemcpyHostToDevice, x->stream));
    ISAC_CUDA_CHECK(x, cudaMemcpyAsync(dA, antPos, bA, cudaMemcpyHostToDevice, x->stream));
    st = city_check_los(c->p, nLinks, dU, dA, nAnt == 1 ? 0 : 3, dL, x->stream);
    if (st) return st;
    ISAC_CUDA_CHECK(x, cudaMemcpyAsync(los, dL, bL, cudaMemcpyDeviceToHost, x->stream));
    ISAC_CUDA_CHECK(x, cudaStreamSynchronize(x->stream));
    return ISAC_OK;
}

int isac_ofdm_modulate_dev(isac_ctx* h, const void* txGrid, int32_t nSc, int32_t nSym, int32_t nAnts, int32_t nfft,
                           int32_t symbolsPerSubframe, const int32_t* cpLengths, double scale, void* txWaveform, int64_t* T) {
    if (!h || !cpLengths || symbolsPerSubframe < 1 || nSym < 1) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    OfdmConfig o{};
    o.nSc = nSc; o.nSym = nSym; o.nAnts = nAnts; o.nfft = nfft;
    o.symbolsPerSubframe = symbolsPerSubframe; o.cpLengths = cpLengths; o.scale = scale;
    if (T) *T = ofdm_waveform_length(o);
    if (!txWaveform) return ISAC_OK;  // size query
    return ofdm_modulate_run(c, o, (const float2*)txGrid, (float2*)txWaveform, c->stream);
}

int isac_ofdm_modulate_ex_dev(isac_ctx* h, const void* txGrid, int32_t nSc, int32_t nSym, int32_t nAnts, int64_t gridStride, int32_t nfft,
                              int32_t symbolsPerSubframe, const int32_t* cpLengths, double scale, int32_t windowing, int32_t symPhase,
                              void* txWaveform, int64_t waveStride, int64_t sampleOffset, int64_t* T) {
    if (!h || !cpLengths || symbolsPerSubframe < 1 || nSym < 1) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    OfdmConfig o{};
    o.nSc = nSc; o.nSym = nSym; o.nAnts = nAnts; o.nfft = nfft;
    o.symbolsPerSubframe = symbolsPerSubframe; o.cpLengths = cpLengths; o.scale = scale;
    o.windowing = windowing; o.symPhase = symPhase; o.gridStride = gridStride; o.waveStride = waveStride; o.sampleOffset = sampleOffset;
    const long long len = ofdm_waveform_length(o);
    if (T) *T = len;
    if (!txWaveform) return ISAC_OK;  // size query
    if (waveStride > 0 && sampleOffset + len > waveStride) {
        set_error(c, "ofdmModulate: the block does not fit in the destination buffer");
        return ISAC_ERR_CAPACITY;
    }
    return ofdm_modulate_run(c, o, (const float2*)txGrid, (float2*)txWaveform, c->stream);
}

int isac_mono_static_sensing_host(isac_ctx* h, const isac_echo_config* cfg, const void* txHost, const void* noiseHost,
                                  int32_t noiseMode, uint64_t seed, void* echoHost, int32_t* nSymOut) {
    if (!h || (!txHost && echoHost) || (!echoHost && !nSymOut)) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    EchoConfig e;
    if (!echo_cfg(c, cfg, e) || !cfg->cpLengths) return ISAC_ERR_INVALID_ARG;
    int n = 0;
    int st = mono_static_sensing_run(c, e, nullptr, nullptr, 0, 0, nullptr, &n, c->stream);  // size query
    if (st) return st;
    if (!echoHost) {  // echoGridHost == NULL: query nSymOut only (as the _dev entry point)
        *nSymOut = n;
        return ISAC_OK;
    }
    const size_t wb = sizeof(float2) * (size_t)e.T * e.nTx, gb = sizeof(float2) * (size_t)e.nSc * n * e.nTx;
    void *dTx = nullptr, *dNz = nullptr, *dOut = nullptr;
    if ((st = ctx_scratch(c, 0, wb, &dTx))) return st;
    if ((st = ctx_scratch(c, 2, gb, &dOut))) return st;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dTx, txHost, wb, cudaMemcpyHostToDevice, c->stream));
    if (noiseMode == ISAC_NOISE_TENSOR) {
        if (!noiseHost) return ISAC_ERR_INVALID_ARG;
        if ((st = ctx_scratch(c, 1, wb, &dNz))) return st;
        ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dNz, noiseHost, wb, cudaMemcpyHostToDevice, c->stream));
    }
    st = mono_static_sensing_run(c, e, (const float2*)dTx, (const float2*)dNz, noiseMode, seed, (float2*)dOut, &n, c->stream);
    if (st) return st;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(echoHost, dOut, gb, cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    if (nSymOut) *nSymOut = n;
    return ISAC_OK;
}

// ---- link budget -------------------------------------------------------------------------------
int isac_pathloss_host(isac_ctx* h, int32_t scenario, double fcHz, int32_t nLinks, const double* bsPos, const double* uePos,
                       const int32_t* los, double* plDb) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    return pathloss_run(&h->c, scenario, fcHz, nLinks, bsPos, uePos, los, plDb, h->c.stream);
}

int isac_link_budget_dev(isac_ctx* h, void* H, int64_t elemsPerLink, int32_t nLinks, const double* plDb, double rxGainDb) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    return link_scale_run(&h->c, (float2*)H, elemsPerLink, nLinks, plDb, rxGainDb, h->c.stream);
}

int isac_thermal_noise_power(double noiseFigureDb, double temperatureK, double sampleRate, double* Nt) {
    if (!Nt || !(sampleRate > 0.0)) return ISAC_ERR_INVALID_ARG;
    const double nf = std::pow(10.0, noiseFigureDb / 10.0);
    *Nt = 1.380649e-23 * (temperatureK + 290.0 * (nf - 1.0)) * sampleRate;   // uePhy.m:945-947
    return ISAC_OK;
}

int isac_dft_channel_matrix(int32_t nTx, int32_t nRx, double* H) {
    if (nTx < 1 || nRx < 1 || !H) return ISAC_ERR_INVALID_ARG;
    const int m = nTx > nRx ? nTx : nRx;
    // the rows (or columns) of the truncated DFT matrix stay orthogonal with norm sqrt(m): its 2-norm is sqrt(m)
    const double inv = 1.0 / std::sqrt((double)m), w = -2.0 * 3.14159265358979323846 / m;
    for (int r = 0; r < nRx; ++r)
        for (int t = 0; t < nTx; ++t) {
            const double ph = w * (double)((long long)t * r % m);
            H[2 * (t + (size_t)nTx * r)] = std::cos(ph) * inv;
            H[2 * (t + (size_t)nTx * r) + 1] = std::sin(ph) * inv;
        }
    return ISAC_OK;
}

// ---- codebooks / PMI / RI / CQI / UL TPMI / PRG precoding --------------------------------------
struct isac_pmi_plan {
    PmiPlan* p;
};
struct isac_csi_plan {
    Ctx* ctx;
    CsiConfig cfg;
    std::vector<uint8_t> csr, i2r;
    std::vector<int> reK, reL;
    int maxBatch;
    PmiPlan* byRank[kMaxLayers];
    char* d_arena = nullptr;    // selection results of all ranks, contiguous -> one D2H copy per report
    char* h_arena = nullptr;    // pinned
    size_t arenaBytes = 0;
    // a report enqueued by ri_enqueue and not yet finished (isac_csi_report_enqueue_dev / _finish)
    cudaEvent_t ready = nullptr;   // recorded behind the D2H copy of the arena
    // host-side result records of the last report: kept between reports so that the per-UE vectors are reused, not reallocated
    std::vector<double> riBuf;
    std::vector<PmiResult> chosenBuf;
    std::vector<std::vector<PmiResult>> allBuf;
    bool scored[kMaxLayers] = {};   // ranks whose records in the `all` list of the last ri_finish are current
    const float2* pendH = nullptr;
    std::vector<double> pendNVar;
    int pendBatch = 0;             // 0 = nothing pending
    bool pendLaunched = false;     // false: nothing reportable (riSelect.m:235-245), no kernels were enqueued
};

static CsiConfig to_csi_config(const isac_csi_config* c) {
    CsiConfig o{};
    o.nPorts = c->nPorts; o.N1 = c->N1; o.N2 = c->N2; o.O1 = c->O1; o.O2 = c->O2; o.codebookMode = c->codebookMode;
    o.nSizeBWP = c->nSizeBWP; o.nStartBWP = c->nStartBWP; o.subbandSize = c->subbandSize;
    o.pmiSubband = c->pmiSubband; o.cqiSubband = c->cqiSubband; o.K = c->K; o.L = c->L; o.nRx = c->nRx;
    o.subsetRestriction = c->subsetRestriction; o.i2Restriction = c->i2Restriction;
    std::memcpy(o.riRestriction, c->riRestriction, 8);
    o.nRE = c->nRE; o.reK = c->reK; o.reL = c->reL;
    o.nPanels = c->nPanels;
    return o;
}

int isac_type1sp_codebook(const isac_csi_config* cfg, int32_t nLayers, int32_t variant, int32_t dims[4], double* W) {
    if (!cfg || !dims) return ISAC_ERR_INVALID_ARG;
    CodebookTable t;
    int st = build_type1sp_table(nullptr, to_csi_config(cfg), nLayers, variant, t);
    if (st) return st;
    dims[0] = t.n2; dims[1] = t.n11; dims[2] = t.n12; dims[3] = t.n13;
    if (W) {
        std::vector<std::complex<double>> w;
        materialize_codebook(t, w);
        std::memcpy(W, w.data(), sizeof(std::complex<double>) * w.size());
    }
    return ISAC_OK;
}

int isac_type1mp_codebook(const isac_csi_config* cfg, int32_t nPanels, int32_t nLayers, int32_t dims[9], double* W) {
    if (!cfg || !dims) return ISAC_ERR_INVALID_ARG;
    CsiConfig c = to_csi_config(cfg);
    int d[9];
    std::vector<std::complex<double>> w;
    const int st = type1mp_codebook(nullptr, c, nPanels, nLayers, d, W ? &w : nullptr);
    if (st) return st;
    for (int i = 0; i < 9; ++i) dims[i] = d[i];
    if (W) std::memcpy(W, w.data(), sizeof(std::complex<double>) * w.size());
    return ISAC_OK;
}

int isac_type1mp_codebook_from_table(const isac_csi_config* cfg, int32_t nPanels, int32_t nLayers, int32_t dims[9], double* W) {
    if (!cfg || !dims) return ISAC_ERR_INVALID_ARG;
    CsiConfig c = to_csi_config(cfg);
    c.nPanels = nPanels;
    c.nPorts = 2 * nPanels * c.N1 * c.N2;
    CodebookTable t;
    const int st = build_type1mp_table(nullptr, c, nLayers, t);
    if (st) return st;
    const int d[9] = {t.mp[0], t.mp[1], t.mp[2], t.n11, t.n12, t.mp[3], t.mp[4], t.mp[5], t.mp[6]};
    for (int i = 0; i < 9; ++i) dims[i] = d[i];
    if (W) {
        std::vector<std::complex<double>> w;
        materialize_codebook(t, w);
        std::memcpy(W, w.data(), sizeof(std::complex<double>) * w.size());
    }
    return ISAC_OK;
}

int isac_pusch_codebook(int32_t nLayers, int32_t nPorts, int32_t* nTPMI, double* W) {
    if (!nTPMI) return ISAC_ERR_INVALID_ARG;
    CodebookTable t;
    int st = build_pusch_table(nullptr, nLayers, nPorts, t);
    if (st) return st;
    *nTPMI = t.n2;
    if (W) {
        std::vector<std::complex<double>> w;
        materialize_codebook(t, w);
        std::memcpy(W, w.data(), sizeof(std::complex<double>) * w.size());
    }
    return ISAC_OK;
}

int isac_pmi_plan_create(isac_ctx* h, const isac_csi_config* cfg, int32_t nLayers, int32_t maxBatch, isac_pmi_plan** out) {
    if (!h || !cfg || !out) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    PmiPlan* p = nullptr;
    int st = pmi_plan_create(&h->c, to_csi_config(cfg), nLayers, maxBatch, &p);
    if (st) return st;
    *out = new isac_pmi_plan{p};
    return ISAC_OK;
}

int isac_pmi_plan_destroy(isac_pmi_plan* pl) {
    if (!pl) return ISAC_OK;
    if (pl->p) {
        cudaSetDevice(pl->p->ctx->device);
        cudaStreamSynchronize(pl->p->ctx->stream);
        pmi_plan_destroy(pl->p);
    }
    delete pl;
    return ISAC_OK;
}

int isac_pmi_plan_set_kernel(isac_pmi_plan* pl, int32_t direct) {
    if (!pl || !pl->p) return ISAC_ERR_INVALID_ARG;
    pl->p->direct = direct != 0 || pl->p->cfg.nPanels >= 2;   // Type1MultiPanel plans always run the direct kernel
    return ISAC_OK;
}

int isac_pmi_plan_mp_dims(const isac_pmi_plan* pl, int32_t mpDims[7]) {
    if (!pl || !pl->p || !mpDims) return ISAC_ERR_INVALID_ARG;
    for (int i = 0; i < 7; ++i) mpDims[i] = pl->p->tab.mp[i];
    return ISAC_OK;
}

int isac_pmi_plan_info(const isac_pmi_plan* pl, int32_t dims[4], int32_t* nSB, int32_t* nCqiSB, int32_t* nRE, int32_t* reKs,
                       int32_t* reLs) {
    if (!pl || !pl->p) return ISAC_ERR_INVALID_ARG;
    const PmiPlan* p = pl->p;
    if (dims) { dims[0] = p->tab.n2; dims[1] = p->tab.n11; dims[2] = p->tab.n12; dims[3] = p->tab.n13; }
    if (nSB) *nSB = p->nSB;
    if (nCqiSB) *nCqiSB = p->nCqiSB;
    if (nRE) *nRE = (int32_t)p->reK.size();
    if (reKs) std::memcpy(reKs, p->reK.data(), sizeof(int) * p->reK.size());
    if (reLs) std::memcpy(reLs, p->reL.data(), sizeof(int) * p->reL.size());
    return ISAC_OK;
}

int isac_dl_pmi_select_dev(isac_pmi_plan* pl, const void* H, const double* nVar, int32_t batch) {
    if (!pl || !pl->p) return ISAC_ERR_INVALID_ARG;
    Ctx* c = pl->p->ctx;
    cudaSetDevice(c->device);
    return pmi_select_run(pl->p, (const float2*)H, nVar, batch, c->stream);
}

static void put_pmi(const PmiResult& r, int nSB, int nu, double* i1, double* i2, double* sinrAt) {
    if (i1) for (int q = 0; q < 3; ++q) i1[q] = r.allNaN ? NAN : (double)r.i1[q];
    if (i2) for (int sb = 0; sb < nSB; ++sb) i2[sb] = r.i2[sb];
    if (sinrAt) std::memcpy(sinrAt, r.sinrSel.data(), sizeof(double) * (size_t)nSB * nu);
}

int isac_dl_pmi_collect(isac_pmi_plan* pl, int32_t batch, double* i1, double* i2, double* sinrAt) {
    if (!pl || !pl->p) return ISAC_ERR_INVALID_ARG;
    PmiPlan* p = pl->p;
    cudaSetDevice(p->ctx->device);
    std::vector<PmiResult> res;
    int st = pmi_select_collect(p, batch, res);
    if (st) return st;
    for (int b = 0; b < batch; ++b)
        put_pmi(res[b], p->nSB, p->nLayers, i1 ? i1 + 3 * b : nullptr, i2 ? i2 + (size_t)p->nSB * b : nullptr,
                sinrAt ? sinrAt + (size_t)p->nSB * p->nLayers * b : nullptr);
    return ISAC_OK;
}

int isac_dl_pmi_get_info(isac_pmi_plan* pl, int32_t batch, double* sinrPerRE, double* sinrPerSubband) {
    if (!pl || !pl->p) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(pl->p->ctx->device);
    return pmi_get_sinr_arrays(pl->p, batch, sinrPerRE, sinrPerSubband);
}

// riSelect.m:222-231: ranks 1..min(nRx, nPorts), at most 4 with a Type1MultiPanel codebook
static int csi_max_rank(const CsiConfig& c) {
    int m = c.nRx < c.nPorts ? c.nRx : c.nPorts;
    if (c.nPanels >= 2 && m > 4) m = 4;
    return m < kMaxLayers ? m : kMaxLayers;
}

int isac_csi_plan_create(isac_ctx* h, const isac_csi_config* cfg, int32_t maxBatch, isac_csi_plan** out) {
    if (!h || !cfg || !out) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    isac_csi_plan* pl = new isac_csi_plan();
    pl->ctx = &h->c;
    pl->cfg = to_csi_config(cfg);
    pl->maxBatch = maxBatch;
    for (int r = 0; r < kMaxLayers; ++r) pl->byRank[r] = nullptr;
    const int maxRank = csi_max_rank(pl->cfg);  // riSelect.m:222
    PmiShared* share = nullptr;  // ranks built from the same beams share one Gram-pair dictionary -> one fused SINR launch
    // report plans never return SINRPerRE: subband sums are accumulated inside the SINR kernel (ISAC_PMI_FUSED=0: per-RE path)
    const bool fused = !(getenv("ISAC_PMI_FUSED") && atoi(getenv("ISAC_PMI_FUSED")) == 0);
    for (int r = 1; r <= maxRank && r <= kMaxLayers; ++r) {
        int st = pmi_plan_create(&h->c, pl->cfg, r, maxBatch, &pl->byRank[r - 1], share, fused);
        if (!st && !share) share = pl->byRank[r - 1]->sh;
        if (st) {
            isac_csi_plan_destroy(pl);
            return st;
        }
    }
    for (int r = 0; r < kMaxLayers; ++r)
        if (pl->byRank[r]) pl->arenaBytes += pl->byRank[r]->resBytes;
    if (cudaMalloc((void**)&pl->d_arena, pl->arenaBytes ? pl->arenaBytes : 16) != cudaSuccess ||
        cudaMallocHost((void**)&pl->h_arena, pl->arenaBytes ? pl->arenaBytes : 16) != cudaSuccess) {
        set_error(pl->ctx, "csi_plan_create: result arena allocation failed");
        isac_csi_plan_destroy(pl);
        return ISAC_ERR_CUDA;
    }
    size_t off = 0;
    for (int r = 0; r < kMaxLayers; ++r)
        if (pl->byRank[r]) {
            pmi_plan_use_arena(pl->byRank[r], pl->d_arena + off, pl->h_arena + off);
            off += pl->byRank[r]->resBytes;
        }
    *out = pl;
    return ISAC_OK;
}

int isac_csi_plan_destroy(isac_csi_plan* pl) {
    if (!pl) return ISAC_OK;
    cudaSetDevice(pl->ctx->device);
    cudaStreamSynchronize(pl->ctx->stream);
    cudaFree(pl->d_arena);
    if (pl->h_arena) cudaFreeHost(pl->h_arena);
    if (pl->ready) cudaEventDestroy(pl->ready);
    for (int r = 0; r < kMaxLayers; ++r)
        if (pl->byRank[r]) pmi_plan_destroy(pl->byRank[r]);
    delete pl;
    return ISAC_OK;
}

int isac_csi_plan_set_kernel(isac_csi_plan* pl, int32_t direct) {
    if (!pl) return ISAC_ERR_INVALID_ARG;
    for (int r = 0; r < kMaxLayers; ++r)
        if (pl->byRank[r]) pl->byRank[r]->direct = direct != 0 || pl->cfg.nPanels >= 2;   // multi-panel: direct kernel only
    return ISAC_OK;
}

int isac_csi_plan_mp_dims(const isac_csi_plan* pl, int32_t nLayers, int32_t mpDims[7]) {
    if (!pl || !mpDims || nLayers < 1 || nLayers > kMaxLayers || !pl->byRank[nLayers - 1]) return ISAC_ERR_INVALID_ARG;
    for (int i = 0; i < 7; ++i) mpDims[i] = pl->byRank[nLayers - 1]->tab.mp[i];
    return ISAC_OK;
}

// riSelect.m:254-294 for a batch; keeps every evaluated rank's results for the fused report
// Kernels of every valid rank + the asynchronous D2H copy of the selection arena; no synchronisation.
static int ri_enqueue(isac_csi_plan* pl, const float2* H, const double* nVar, int batch) {
    Ctx* c = pl->ctx;
    const int maxRank = csi_max_rank(pl->cfg);
    pl->pendH = H;
    pl->pendNVar.assign(nVar, nVar + batch);
    pl->pendBatch = batch;
    pl->pendLaunched = false;
    std::vector<PmiPlan*> plans;
    for (int r = 1; r <= maxRank && r <= kMaxLayers; ++r)
        if (pl->cfg.riRestriction[r - 1]) plans.push_back(pl->byRank[r - 1]);
    if (plans.empty() || (pl->byRank[0] && pl->byRank[0]->reK.empty())) return kOk;  // riSelect.m:235-245
    int st = pmi_select_run_multi(plans.data(), (int)plans.size(), H, nVar, batch, c->stream);
    if (st) { pl->pendBatch = 0; return st; }
    if (!pl->ready) ISAC_CUDA_CHECK(c, cudaEventCreateWithFlags(&pl->ready, cudaEventDisableTiming));
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(pl->h_arena, pl->d_arena, pl->arenaBytes, cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaEventRecord(pl->ready, c->stream));  // one copy + one wait for all ranks
    pl->pendLaunched = true;
    return kOk;
}

// Waits for the arena of the pending report (not for work enqueued behind it) and runs the host-side rank selection.
static int ri_finish(isac_csi_plan* pl, std::vector<double>& RI, std::vector<PmiResult>& chosen,
                     std::vector<std::vector<PmiResult>>& all) {
    Ctx* c = pl->ctx;
    const int batch = pl->pendBatch;
    if (batch < 1) { set_error(c, "csi report: nothing enqueued"); return kErrInvalidArg; }
    pl->pendBatch = 0;
    const int maxRank = csi_max_rank(pl->cfg);
    all.resize(kMaxLayers);   // the per-UE records of an earlier report stay allocated (their vectors are overwritten in place)
    for (int r = 0; r < kMaxLayers; ++r) pl->scored[r] = false;
    std::vector<int> valid;
    for (int r = 1; r <= maxRank && r <= kMaxLayers; ++r)
        if (pl->cfg.riRestriction[r - 1]) valid.push_back(r);
    RI.assign(batch, NAN);
    chosen.resize(batch);
    for (auto& r : chosen) { r.allNaN = false; r.i1[0] = r.i1[1] = r.i1[2] = 0; r.i2.clear(); r.sinrSel.clear(); r.sinrWbSel.clear(); }
    const int nSB = pl->byRank[0] ? pl->byRank[0]->nSB : 1;
    if (!pl->pendLaunched) {  // riSelect.m:235-245
        for (auto& r : chosen) { r.allNaN = true; r.i2.assign(nSB, NAN); }
        return kOk;
    }
    ISAC_CUDA_CHECK(c, cudaEventSynchronize(pl->ready));
    for (int r : valid) {
        int st = pmi_select_collect_finish(pl->byRank[r - 1], batch, all[r - 1]);
        if (st) return st;
        pl->scored[r - 1] = true;
    }
    for (int b = 0; b < batch; ++b) {
        double best = -INFINITY;
        bool allNaNTotals = true;
        for (int r : valid) {
            const PmiResult& pr = all[r - 1][b];
            const double total = ri_total_sinr(pr, nSB, r);
            if (!std::isnan(total)) allNaNTotals = false;
            if (total > best + 0.1) {  // riSelect.m:284
                best = total;
                RI[b] = r;
                chosen[b] = pr;
            }
        }
        if (allNaNTotals) {  // riSelect.m:289-292
            RI[b] = NAN;
            chosen[b] = all[valid.back() - 1][b];
        }
    }
    return kOk;
}

static int ri_select_batch(isac_csi_plan* pl, const float2* H, const double* nVar, int batch, std::vector<double>& RI,
                           std::vector<PmiResult>& chosen, std::vector<std::vector<PmiResult>>& all) {
    const int st = ri_enqueue(pl, H, nVar, batch);
    return st ? st : ri_finish(pl, RI, chosen, all);
}

int isac_ri_select_dev(isac_csi_plan* pl, const void* H, const double* nVar, int32_t batch, double* RI, double* i1, double* i2) {
    if (!pl || !H || !nVar || !RI) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(pl->ctx->device);
    if (batch < 1 || batch > pl->maxBatch) { set_error(pl->ctx, "riSelect: batch out of range"); return ISAC_ERR_INVALID_ARG; }
    std::vector<double> ri;
    std::vector<PmiResult> chosen;
    std::vector<std::vector<PmiResult>> all;
    int st = ri_select_batch(pl, (const float2*)H, nVar, batch, ri, chosen, all);
    if (st) return st;
    const int nSB = pl->byRank[0]->nSB;
    for (int b = 0; b < batch; ++b) {
        RI[b] = ri[b];
        put_pmi(chosen[b], nSB, 0, i1 ? i1 + 3 * b : nullptr, i2 ? i2 + (size_t)nSB * b : nullptr, nullptr);
    }
    return ISAC_OK;
}

static void put_cqi(const CsiReport& rep, int rowsOut, double* cqi, double* sbcw, int rowsFull) {
    if (cqi) {
        for (int i = 0; i < rowsOut * 2; ++i) cqi[i] = NAN;
        for (int c = 0; c < rep.nCW; ++c)
            for (int s = 0; s < rep.nCqiRows; ++s) cqi[(size_t)c * rowsOut + s] = rep.cqi[(size_t)c * rep.nCqiRows + s];
    }
    if (sbcw) {
        for (int i = 0; i < rowsFull * 2; ++i) sbcw[i] = NAN;
        const int rows = (int)(rep.sinrPerSubbandPerCW.size() / (rep.nCW ? rep.nCW : 1));
        for (int c = 0; c < rep.nCW; ++c)
            for (int s = 0; s < rows && s < rowsFull; ++s) sbcw[(size_t)c * rowsFull + s] = rep.sinrPerSubbandPerCW[(size_t)c * rows + s];
    }
}

int isac_cqi_select_dev(isac_csi_plan* pl, int32_t nLayers, const void* H, const double* nVar, int32_t batch,
                        const double* table, int32_t tableLen, double* cqi, int32_t* cqiRows, double* i1, double* i2,
                        double* sbcw) {
    if (!pl || !H || !nVar || !table) return ISAC_ERR_INVALID_ARG;
    Ctx* c = pl->ctx;
    cudaSetDevice(c->device);
    if (nLayers < 1 || nLayers > kMaxLayers || !pl->byRank[nLayers - 1]) { set_error(c, "cqiSelect: invalid nLayers"); return ISAC_ERR_INVALID_ARG; }
    if (batch < 1 || batch > pl->maxBatch) { set_error(c, "cqiSelect: batch out of range"); return ISAC_ERR_INVALID_ARG; }
    PmiPlan* p = pl->byRank[nLayers - 1];
    int st = pmi_select_run(p, (const float2*)H, nVar, batch, c->stream);  // cqiSelect.m:507
    if (st) return st;
    std::vector<PmiResult> res;
    if ((st = pmi_select_collect(p, batch, res))) return st;
    const int rowsOut = (pl->cfg.cqiSubband && p->nCqiSB > 1) ? p->nCqiSB + 1 : 1;
    const int rowsFull = p->nCqiSB > 1 ? p->nCqiSB + 1 : 1;
    if (cqiRows) *cqiRows = rowsOut;
    for (int b = 0; b < batch; ++b) {
        CsiReport rep;
        cqi_from_pmi(pl->cfg, nLayers, res[b], p->nSB, p->nCqiSB, table, tableLen, rep);
        put_cqi(rep, rowsOut, cqi ? cqi + (size_t)2 * rowsOut * b : nullptr, sbcw ? sbcw + (size_t)2 * rowsFull * b : nullptr, rowsFull);
        put_pmi(res[b], p->nSB, nLayers, i1 ? i1 + 3 * b : nullptr, i2 ? i2 + (size_t)p->nSB * b : nullptr, nullptr);
    }
    return ISAC_OK;
}

int isac_csi_report_enqueue_dev(isac_csi_plan* pl, const void* H, const double* nVar, int32_t batch) {
    if (!pl || !H || !nVar) return ISAC_ERR_INVALID_ARG;
    Ctx* c = pl->ctx;
    cudaSetDevice(c->device);
    if (batch < 1 || batch > pl->maxBatch) { set_error(c, "csi report: batch out of range"); return ISAC_ERR_INVALID_ARG; }
    if (pl->pendBatch) { set_error(c, "csi report: the previous report of this plan has not been finished"); return ISAC_ERR_INVALID_ARG; }
    return ri_enqueue(pl, (const float2*)H, nVar, batch);
}

int isac_csi_report_finish(isac_csi_plan* pl, const double* table, int32_t tableLen, int32_t rankCap, double* RI, double* i1,
                           double* i2, double* cqi, int32_t* cqiRows) {
    if (!pl || !table || !RI) return ISAC_ERR_INVALID_ARG;
    Ctx* c = pl->ctx;
    cudaSetDevice(c->device);
    const float2* H = pl->pendH;
    const std::vector<double> nVar = pl->pendNVar;
    const int batch = pl->pendBatch;
    std::vector<double>& ri = pl->riBuf;
    std::vector<PmiResult>& chosen = pl->chosenBuf;
    std::vector<std::vector<PmiResult>>& all = pl->allBuf;
    int st = ri_finish(pl, ri, chosen, all);
    if (st) return st;
    PmiPlan* p0 = pl->byRank[0];
    const int rowsOut = (pl->cfg.cqiSubband && p0->nCqiSB > 1) ? p0->nCqiSB + 1 : 1;
    if (cqiRows) *cqiRows = rowsOut;
    const bool launched = pl->pendLaunched;
    for (int b = 0; b < batch; ++b) {
        RI[b] = ri[b];
        if (!launched) {   // no CSI-RS RE in the BWP / every rank restricted: all-NaN report (riSelect.m:235-245, cqiSelect.m:636-650)
            if (i1) for (int q = 0; q < 3; ++q) i1[3 * b + q] = NAN;
            if (i2) for (int sb = 0; sb < p0->nSB; ++sb) i2[(size_t)p0->nSB * b + sb] = NAN;
            if (cqi) for (int q = 0; q < 2 * rowsOut; ++q) cqi[(size_t)2 * rowsOut * b + q] = NAN;
            continue;
        }
        // rank = min(riSelect(..), rankCap) (uePhy.m:901-903): MATLAB's min ignores NaN, so a NaN RI proceeds with the cap
        int rank = std::isnan(ri[b]) ? (rankCap > 0 ? rankCap : 1) : (int)ri[b];
        if (rankCap > 0 && rank > rankCap) rank = rankCap;
        if (!std::isnan(ri[b])) RI[b] = rank;
        if (rank < 1 || rank > kMaxLayers || !pl->byRank[rank - 1]) {
            set_error(c, "nr5g:hDLPMISelect:InvalidNumLayers");
            return ISAC_ERR_INVALID_ARG;
        }
        if (!pl->scored[rank - 1]) {  // rank not scored by the RI loop (restricted): evaluate it now
            if ((st = pmi_select_run(pl->byRank[rank - 1], H, nVar.data(), batch, c->stream))) return st;
            if ((st = pmi_select_collect(pl->byRank[rank - 1], batch, all[rank - 1]))) return st;
            pl->scored[rank - 1] = true;
        }
        const PmiResult& pr = all[rank - 1][b];
        CsiReport rep;
        cqi_from_pmi(pl->cfg, rank, pr, p0->nSB, p0->nCqiSB, table, tableLen, rep);
        put_cqi(rep, rowsOut, cqi ? cqi + (size_t)2 * rowsOut * b : nullptr, nullptr, 0);
        put_pmi(pr, p0->nSB, rank, i1 ? i1 + 3 * b : nullptr, i2 ? i2 + (size_t)p0->nSB * b : nullptr, nullptr);
    }
    return ISAC_OK;
}

int isac_csi_report_dev(isac_csi_plan* pl, const void* H, const double* nVar, int32_t batch, const double* table,
                        int32_t tableLen, int32_t rankCap, double* RI, double* i1, double* i2, double* cqi, int32_t* cqiRows) {
    if (!pl || !H || !nVar || !table || !RI) return ISAC_ERR_INVALID_ARG;
    const int st = isac_csi_report_enqueue_dev(pl, H, nVar, batch);
    return st ? st : isac_csi_report_finish(pl, table, tableLen, rankCap, RI, i1, i2, cqi, cqiRows);
}

int isac_precoded_sinr_host(isac_ctx* h, const void* H, int32_t nRx, int32_t nPorts, double sigma, const void* W, int32_t nLayers,
                            int32_t batch, double* sinr) {
    if (!h || !H || !W || !sinr || batch < 1 || nRx < 1 || nPorts < 1 || nLayers < 1) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    const size_t bH = sizeof(double2) * (size_t)nRx * nPorts * batch, bW = sizeof(double2) * (size_t)nPorts * nLayers;
    void* d = nullptr;
    int st = ctx_scratch(c, 16, bH + bW + sizeof(double) * batch, &d);
    if (st) return st;
    double2* dH = (double2*)d;
    double2* dW = (double2*)((char*)d + bH);
    double* dOut = (double*)((char*)d + bH + bW);
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dH, H, bH, cudaMemcpyHostToDevice, c->stream));
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(dW, W, bW, cudaMemcpyHostToDevice, c->stream));
    st = precoded_sinr_run(c, dH, nRx, nPorts, sigma, dW, nLayers, batch, dOut, c->stream);
    if (st) return st;
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(sinr, dOut, sizeof(double) * batch, cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return ISAC_OK;
}

int isac_ul_pmi_select_dev(isac_ctx* h, int32_t nLayers, const void* hest, int32_t K, int32_t nSym, int32_t nRx, int32_t nPorts,
                           double noiseEst, int32_t bandSize, int32_t maxSB, double* pmi, double* sinr, int32_t* sbIdx,
                           int32_t* nSB, int32_t* nTPMI, int32_t* none) {
    if (!h || !hest || !nSB || !nTPMI || !none) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    UlPmiResult r;
    int st = ul_pmi_select_run(c, nLayers, (const float2*)hest, K, nSym, nRx, nPorts, noiseEst, bandSize, r, c->stream);
    if (st) return st;
    *nSB = r.nSB; *nTPMI = r.nTPMI; *none = r.none ? 1 : 0;
    if (r.nSB > maxSB) { set_error(c, "pmiSelect: maxSB too small"); return ISAC_ERR_CAPACITY; }
    if (sbIdx) std::memcpy(sbIdx, r.subbandIndices.data(), sizeof(int) * r.subbandIndices.size());
    if (!r.none) {
        if (pmi) std::memcpy(pmi, r.pmi.data(), sizeof(double) * r.pmi.size());
        if (sinr) std::memcpy(sinr, r.sinr.data(), sizeof(double) * r.sinr.size());
    }
    return ISAC_OK;
}

static int ul_batch_outputs(Ctx* c, const std::vector<UlPmiResult>& r, int32_t maxSB, double* pmi, double* sinr, int32_t* nSB,
                            int32_t* nTPMI, int32_t* none) {
    const int batch = (int)r.size();
    *nSB = r[0].nSB; *nTPMI = r[0].nTPMI;
    if (r[0].nSB > maxSB) { set_error(c, "pmiSelect: maxSB too small"); return ISAC_ERR_CAPACITY; }
    for (int b = 0; b < batch; ++b) {
        none[b] = r[b].none ? 1 : 0;
        if (r[b].none) continue;
        if (pmi) std::memcpy(pmi + (size_t)maxSB * b, r[b].pmi.data(), sizeof(double) * r[b].pmi.size());
        if (sinr) std::memcpy(sinr + (size_t)maxSB * r[0].nTPMI * b, r[b].sinr.data(), sizeof(double) * r[b].sinr.size());
    }
    return ISAC_OK;
}

int isac_ul_pmi_select_batch_dev(isac_ctx* h, int32_t nLayers, const void* hest, int32_t K, int32_t nSym, int32_t nRx, int32_t nPorts,
                                 double noiseEst, int32_t bandSize, int32_t batch, int32_t maxSB, double* pmi, double* sinr,
                                 int32_t* nSB, int32_t* nTPMI, int32_t* none) {
    if (!h || !hest || !nSB || !nTPMI || !none) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    std::vector<UlPmiResult> r;
    int st = ul_pmi_select_batch(c, nLayers, (const float2*)hest, K, nSym, nRx, nPorts, noiseEst, bandSize, batch, r, c->stream);
    if (st) return st;
    return ul_batch_outputs(c, r, maxSB, pmi, sinr, nSB, nTPMI, none);
}

int isac_ul_pmi_select_batch_enqueue_dev(isac_ctx* h, int32_t nLayers, const void* hest, int32_t K, int32_t nSym, int32_t nRx,
                                         int32_t nPorts, double noiseEst, int32_t bandSize, int32_t batch) {
    if (!h || !hest) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    return ul_pmi_select_batch_enqueue(c, nLayers, (const float2*)hest, K, nSym, nRx, nPorts, noiseEst, bandSize, batch, c->stream);
}

int isac_ul_pmi_select_batch_finish(isac_ctx* h, int32_t maxSB, double* pmi, double* sinr, int32_t* nSB, int32_t* nTPMI,
                                    int32_t* none) {
    if (!h || !nSB || !nTPMI || !none) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    std::vector<UlPmiResult> r;
    int st = ul_pmi_select_batch_finish(c, r);
    if (st) return st;
    return ul_batch_outputs(c, r, maxSB, pmi, sinr, nSB, nTPMI, none);
}

int isac_prg_precode_dev(isac_ctx* h, int32_t K, int32_t Lsym, int32_t nStartGrid, const void* portsym, const int32_t* portind,
                         int32_t NRE, int32_t nLayers, const void* F, int32_t P, int32_t NPRG, void* antsym, int32_t* antind) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    return prg_precode_run(c, K, Lsym, nStartGrid, (const float2*)portsym, portind, NRE, nLayers, (const float2*)F, P, NPRG,
                           (float2*)antsym, antind, 1, c->stream);
}

int isac_prg_precode_batch_dev(isac_ctx* h, int32_t K, int32_t Lsym, int32_t nStartGrid, const void* portsym, const int32_t* portind,
                               int32_t NRE, int32_t nLayers, const void* F, int32_t P, int32_t NPRG, int32_t batch, void* antsym,
                               int32_t* antind) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    return prg_precode_run(c, K, Lsym, nStartGrid, (const float2*)portsym, portind, NRE, nLayers, (const float2*)F, P, NPRG,
                           (float2*)antsym, antind, batch, c->stream);
}

// ---- CDL channel ---------------------------------------------------------------------------------
struct isac_cdl_channel {
    Ctx* ctx;
    CdlRays rays;
};

int isac_cdl_create(isac_ctx* h, const isac_cdl_config* cfg, isac_cdl_channel** out) {
    if (!h || !cfg || !out) return ISAC_ERR_INVALID_ARG;
    CdlConfig c{};
    c.profile = cfg->profile; c.delaySpread = cfg->delaySpread; c.fc = cfg->fc; c.maxDoppler = cfg->maxDoppler;
    for (int i = 0; i < 3; ++i) { c.txSize[i] = cfg->txSize[i]; c.rxSize[i] = cfg->rxSize[i]; }
    c.txPattern38901 = cfg->txPattern38901; c.rxPattern38901 = cfg->rxPattern38901; c.seed = cfg->seed;
    isac_cdl_channel* ch = new isac_cdl_channel();
    ch->ctx = &h->c;
    int st = cdl_build_rays(&h->c, c, ch->rays);
    if (st) { delete ch; return st; }
    *out = ch;
    return ISAC_OK;
}

int isac_cdl_destroy(isac_cdl_channel* ch) {
    if (ch) {
        cudaSetDevice(ch->ctx->device);
        cudaStreamSynchronize(ch->ctx->stream);
        cdl_free(ch->rays);
    }
    delete ch;
    return ISAC_OK;
}

int isac_cdl_get_rays(const isac_cdl_channel* ch, int32_t* nCl, int32_t* nRays, int32_t* nRx, int32_t* nTx, double* tau,
                      double* nu, int32_t* cluster, double* g) {
    if (!ch) return ISAC_ERR_INVALID_ARG;
    const CdlRays& r = ch->rays;
    if (nCl) *nCl = r.nCl;
    if (nRays) *nRays = (int32_t)r.nu.size();
    if (nRx) *nRx = r.nRx;
    if (nTx) *nTx = r.nTx;
    if (tau) std::memcpy(tau, r.tau.data(), sizeof(double) * r.tau.size());
    if (nu) std::memcpy(nu, r.nu.data(), sizeof(double) * r.nu.size());
    if (cluster) std::memcpy(cluster, r.cluster.data(), sizeof(int) * r.cluster.size());
    if (g) std::memcpy(g, r.g.data(), sizeof(std::complex<double>) * r.g.size());
    return ISAC_OK;
}

int isac_cdl_set_kernel(isac_cdl_channel* ch, int32_t legacyMma) {
    if (!ch) return ISAC_ERR_INVALID_ARG;
    ch->rays.legacyMma = legacyMma != 0;
    return ISAC_OK;
}

int isac_cdl_generate_batch_dev(isac_cdl_channel* const* ch, int32_t n, int32_t K, double scsHz, int32_t L, const double* symTime,
                                const double* t0, void* H) {
    if (!ch || n < 1 || !H || !symTime || !t0) return ISAC_ERR_INVALID_ARG;
    std::vector<CdlRays*> rays(n);
    for (int i = 0; i < n; ++i) {
        if (!ch[i] || ch[i]->ctx != ch[0]->ctx) return ISAC_ERR_INVALID_ARG;
        rays[i] = &ch[i]->rays;
    }
    Ctx* c = ch[0]->ctx;
    cudaSetDevice(c->device);
    return cdl_generate_batch(c, rays.data(), n, K, scsHz, L, symTime, t0, (float2*)H, c->stream);
}

int isac_cdl_generate_dev(isac_cdl_channel* ch, int32_t K, double scsHz, int32_t L, const double* symTime, double t0, void* H) {
    if (!ch || !H || !symTime) return ISAC_ERR_INVALID_ARG;
    Ctx* c = ch->ctx;
    cudaSetDevice(c->device);
    return cdl_generate(c, ch->rays, K, scsHz, L, symTime, t0, (float2*)H, c->stream);
}

// ---- channel estimation ---------------------------------------------------------
struct isac_chest_plan {
    ChestPlan* p;
};

int isac_chest_plan_create(isac_ctx* h, int32_t K, int32_t L, int32_t nRx, int32_t nPorts, int64_t nRef, const int32_t* refInd,
                           const void* refSym, int32_t cdmFd, int32_t cdmTd, int32_t avgF, int32_t avgT, int32_t maxBatch,
                           isac_chest_plan** out) {
    if (!h || !out) return ISAC_ERR_INVALID_ARG;
    *out = nullptr;
    cudaSetDevice(h->c.device);
    ChestConfig c{K, L, nRx, nPorts, cdmFd, cdmTd, avgF, avgT, maxBatch};
    ChestPlan* p = nullptr;
    const int st = chest_plan_create(&h->c, c, (long long)nRef, refInd, (const float2*)refSym, &p);
    if (st) return st;
    *out = new isac_chest_plan{p};
    return ISAC_OK;
}

int isac_chest_plan_destroy(isac_chest_plan* plan) {
    if (!plan) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(plan->p->ctx->device);
    cudaStreamSynchronize(plan->p->ctx->stream);
    chest_plan_destroy(plan->p);
    delete plan;
    return ISAC_OK;
}

int isac_channel_estimate_dev(isac_chest_plan* plan, const void* rxGrid, int32_t batch, void* Hest, double* nVar) {
    if (!plan) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(plan->p->ctx->device);
    return chest_run(plan->p, (const float2*)rxGrid, batch, (float2*)Hest, nVar, plan->p->ctx->stream);
}

int isac_chest_get_nvar(isac_chest_plan* plan, int32_t batch, double* nVar) {
    if (!plan || !nVar || batch < 1 || batch > plan->p->cfg.maxBatch) return ISAC_ERR_INVALID_ARG;
    Ctx* c = plan->p->ctx;
    cudaSetDevice(c->device);
    ISAC_CUDA_CHECK(c, cudaMemcpyAsync(plan->p->h_nvar, plan->p->d_nvar, sizeof(double) * batch, cudaMemcpyDeviceToHost, c->stream));
    ISAC_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int b = 0; b < batch; ++b) nVar[b] = plan->p->h_nvar[b];
    return ISAC_OK;
}

}  // extern "C"
