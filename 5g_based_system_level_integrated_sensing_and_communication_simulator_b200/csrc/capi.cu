// C ABI (include/isac_b200.h) over the internal C++/CUDA implementation.
#include "../../include/isac_b200.h"
#include "isac_common.cuh"
#include "ctx.cuh"
#include "rdm.cuh"
#include "sense.cuh"
#include "echo.cuh"
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace isac {

void set_error(Ctx* ctx, const std::string& msg) {
    if (ctx) ctx->err = msg;
}
const float2* ctx_twiddle(Ctx* ctx) { return ctx->d_twiddle; }

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
    std::vector<float2> tw(kTwiddleN);
    for (int m = 0; m < kTwiddleN; ++m) {
        const double a = 2.0 * M_PI * (double)m / (double)kTwiddleN;
        tw[m] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    if (cudaMalloc((void**)&c->d_twiddle, sizeof(float2) * kTwiddleN) != cudaSuccess ||
        cudaMemcpy(c->d_twiddle, tw.data(), sizeof(float2) * kTwiddleN, cudaMemcpyHostToDevice) != cudaSuccess) {
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
    if (!h || !Ra || !L || !nAzi) return ISAC_ERR_INVALID_ARG;
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
    st = music_doa_run(c, d, (const double2*)dRa, numDets, &Lh, azi, PdB, P, c->stream);
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

int isac_mono_static_sensing_host(isac_ctx* h, const isac_echo_config* cfg, const void* txHost, const void* noiseHost,
                                  int32_t noiseMode, uint64_t seed, void* echoHost, int32_t* nSymOut) {
    if (!h || !txHost || !echoHost) return ISAC_ERR_INVALID_ARG;
    Ctx* c = &h->c;
    cudaSetDevice(c->device);
    EchoConfig e;
    if (!echo_cfg(c, cfg, e) || !cfg->cpLengths) return ISAC_ERR_INVALID_ARG;
    int n = 0;
    int st = mono_static_sensing_run(c, e, nullptr, nullptr, 0, 0, nullptr, &n, c->stream);  // size query
    if (st) return st;
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

}  // extern "C"
