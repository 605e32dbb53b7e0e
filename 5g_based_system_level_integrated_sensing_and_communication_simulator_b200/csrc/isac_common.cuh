// Shared device/host helpers for the ISAC B200 hot-path library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace isac {

constexpr int kTwiddleN = 4096;  // master twiddle table: tw[m] = exp(+2*pi*i*m/4096)

// ---- complex helpers (float2 = interleaved complex, MATLAB mxComplexSingle layout) ----
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__host__ __device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

__host__ __device__ __forceinline__ double2 zmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ double2 zmulc(double2 a, double2 b) {  // a*conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// fused complex multiply-accumulate forms (4 DFMA each; acc + a*b written with zadd/zmul costs DMUL+DFMA+DADD per part)
__device__ __forceinline__ double2 zfma(double2 acc, double2 a, double2 b) {    // acc + a*b
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, acc.x)), fma(a.x, b.y, fma(a.y, b.x, acc.y)));
}
__device__ __forceinline__ double2 zfmac(double2 acc, double2 a, double2 b) {   // acc + a*conj(b)
    return make_double2(fma(a.x, b.x, fma(a.y, b.y, acc.x)), fma(a.y, b.x, fma(-a.x, b.y, acc.y)));
}
__device__ __forceinline__ double2 zfms(double2 acc, double2 a, double2 b) {    // acc - a*b
    return make_double2(fma(-a.x, b.x, fma(a.y, b.y, acc.x)), fma(-a.x, b.y, fma(-a.y, b.x, acc.y)));
}
__device__ __forceinline__ double2 zfmsc(double2 acc, double2 a, double2 b) {   // acc - a*conj(b)
    return make_double2(fma(-a.x, b.x, fma(-a.y, b.y, acc.x)), fma(-a.y, b.x, fma(a.x, b.y, acc.y)));
}
// fast float64 reciprocal / reciprocal square root: MUFU seed (rcp/rsqrt.approx.ftz.f64, ~2^-22) + two Newton steps
// (relative error ~1e-16; not correctly rounded, which the 1e-5 parity budget of the SINR path does not need)
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    return y * fma(-hx * y, y, 1.5);
}
__host__ __device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// streaming (read-once) global load: ld.global.cs (evict-first in L1/L2)
__device__ __forceinline__ float2 ld_stream(const float2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

// ---- status codes of the C ABI (mirrored in include/isac_b200.h) ----
enum Status : int {
    kOk = 0,
    kErrInvalidArg = 1,
    kErrCuda = 2,
    kErrNoDevice = 3,
    kErrUnsupported = 4,
    kErrCfarWindow = 5,   // a CUT's training window leaves the RD map (reference: CFARDetector2D errors)
    kErrNoLosTarget = 6,  // every target NLoS (reference: empty waveform, basicRadarChannel.m:59)
    kErrNumDetsZero = 7,  // MUSIC asked for zero sources (reference: findpeaks NPeaks=0 errors)
    kErrCapacity = 8,
};

struct Ctx;  // defined in capi.cu

void set_error(Ctx* ctx, const std::string& msg);
struct FftTw;
// per-size twiddle tables for block_fft (N = 16 << log2N16); see fft_core.cuh
void ctx_fft_tw(Ctx* ctx, int N, const float2** tw1, const float2** tw2);

// Optional CUDA-event profiling of kernel groups on the launching stream (used by bench.py for the
// live roofline figure) and a counter of this library's kernel launches.
enum ProfSlot : int {
    kProfRdmRange = 0, kProfRdmDoppler = 1, kProfCfar = 2, kProfEcho = 3, kProfCov = 4, kProfMusic = 5,
    kProfPmi = 6, kProfCdl = 7, kProfPrecode = 8, kProfUlPmi = 9, kProfOfdmMod = 10, kProfChest = 11, kProfSlots = 16
};
int prof_begin(Ctx* ctx, int slot, cudaStream_t st);  // returns a record index (or -1 when disabled)
void prof_end(Ctx* ctx, int rec, cudaStream_t st);
void count_launches(Ctx* ctx, int n);
int ctx_num_sms(Ctx* ctx);

#define ISAC_CUDA_CHECK(ctx, expr)                                                          \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::isac::set_error((ctx), std::string(#expr) + ": " + cudaGetErrorString(_e));   \
            return ::isac::kErrCuda;                                                        \
        }                                                                                   \
    } while (0)

}  // namespace isac
