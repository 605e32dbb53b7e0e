// C ABI (include/isac_b200.h) over the internal C++/CUDA implementation.
#include "../../include/isac_b200.h"
#include "isac_common.cuh"
#include "ctx.cuh"
#include "rdm.cuh"
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace isac {

void set_error(Ctx* ctx, const std::string& msg) {
    if (ctx) ctx->err = msg;
}
const float2* ctx_twiddle(Ctx* ctx) { return ctx->d_twiddle; }

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

int isac_synchronize(isac_ctx* h) {
    if (!h) return ISAC_ERR_INVALID_ARG;
    ISAC_CUDA_CHECK(&h->c, cudaStreamSynchronize(h->c.stream));
    return ISAC_OK;
}

// ---- RDM + CFAR ------------------------------------------------------------------------------
int isac_rdm_plan_create(isac_ctx* h, const isac_rdm_config* cfg, isac_rdm_plan** out) {
    if (!h || !cfg || !out) return ISAC_ERR_INVALID_ARG;
    cudaSetDevice(h->c.device);
    RdmConfig c{};
    c.nSc = cfg->nSc; c.nSym = cfg->nSym; c.nAnts = cfg->nAnts; c.nIFFT = cfg->nIFFT; c.nFFT = cfg->nFFT;
    c.cutRow0 = cfg->cutRow0; c.cutRow1 = cfg->cutRow1; c.cutCol0 = cfg->cutCol0; c.cutCol1 = cfg->cutCol1;
    c.guardRows = cfg->guardRows; c.guardCols = cfg->guardCols;
    c.trainRows = cfg->trainRows; c.trainCols = cfg->trainCols;
    c.maxBatch = cfg->maxBatch; c.pfa = cfg->pfa; c.kaiserBeta = cfg->kaiserBeta;
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

}  // extern "C"
