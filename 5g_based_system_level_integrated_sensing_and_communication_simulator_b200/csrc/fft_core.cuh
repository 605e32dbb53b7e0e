// In-register radix-2/4/8/16 DFTs and a shared-memory block FFT (N = R1*R2*16, N <= 4096).
//
// Every thread keeps 16 complex values in registers; passes exchange data through shared
// memory.  The decomposition is decimation-in-frequency so that
//   * the first pass reads the input at stride N/R1 with consecutive threads on consecutive
//     elements  (coalesced float2 global loads, natural order), and
//   * the last pass leaves X[tf + (N/16)*d] in register d of thread tf
//     (coalesced natural-order global stores straight from registers).
// `RT` FFT instances can be interleaved in shared memory (element e of instance nl lives at
// smem[pad(e)*RT + nl]); with RT == 1 the layout is padded so that all three passes are
// bank-conflict free for 8-byte accesses.
#pragma once
#include "isac_common.cuh"

namespace isac {

// ---- packed FP32x2 complex arithmetic (Blackwell FADD2 / FMUL2 / FFMA2) ----------------------------------------
// A complex value is one 64-bit register pair (re = low half, im = high half).  sm_100a executes add/mul/fma.f32x2 on
// both halves in ONE instruction, and ptxas folds the half swap (".LO_HI"), a one-sided negation (".NP") and scalar
// broadcasts (".F32") of these patterns into operand modifiers -> complex add/sub = 1 instruction (2 scalar),
// complex multiply = 2 (4 scalar), "+- j*u" = 1 FFMA2 (the j rotation is free).  Per-lane rounding is identical to
// the scalar FADD / FMUL / FFMA forms.
__device__ __forceinline__ unsigned long long pk(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 unpk(unsigned long long r) {
    float2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return unpk(r);
}
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return unpk(r);
}
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) {  // element-wise
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
    return unpk(r);
}
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) {  // element-wise a*b + c
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
    return unpk(r);
}
__device__ __forceinline__ float2 pk_swap(float2 a) { return make_float2(a.y, a.x); }
// a * w  (complex)
__device__ __forceinline__ float2 pk_cmul(float2 a, float2 w) {
    return pk_fma(pk_swap(a), make_float2(-w.y, w.y), pk_mul(a, make_float2(w.x, w.x)));
}
// a * conj(b)
__device__ __forceinline__ float2 pk_cmulc(float2 a, float2 b) {
    return pk_fma(pk_swap(a), make_float2(b.y, -b.y), pk_mul(a, make_float2(b.x, b.x)));
}
__device__ __forceinline__ float2 pk_scale(float2 a, float s) { return pk_mul(a, make_float2(s, s)); }
// t + SIGN*j*u  and  t - SIGN*j*u
template <int SIGN>
__device__ __forceinline__ float2 pk_addj(float2 t, float2 u) {
    return pk_fma(pk_swap(u), make_float2(-(float)SIGN, (float)SIGN), t);
}
template <int SIGN>
__device__ __forceinline__ float2 pk_subj(float2 t, float2 u) {
    return pk_fma(pk_swap(u), make_float2((float)SIGN, -(float)SIGN), t);
}
// multiply by SIGN*j
template <int SIGN>
__device__ __forceinline__ float2 mulj(float2 a) {
    return pk_mul(pk_swap(a), make_float2(-(float)SIGN, (float)SIGN));
}

template <int SIGN>
__device__ __forceinline__ void dft2(float2& a, float2& b) {
    float2 t = pk_sub(a, b);
    a = pk_add(a, b);
    b = t;
}

// y_k = sum_n x_n exp(SIGN*2*pi*i*n*k/4)
template <int SIGN>
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = pk_add(x0, x2), t1 = pk_sub(x0, x2);
    const float2 t2 = pk_add(x1, x3), u = pk_sub(x1, x3);
    x0 = pk_add(t0, t2);
    x2 = pk_sub(t0, t2);
    x1 = pk_addj<SIGN>(t1, u);  // t1 + SIGN*j*(x1 - x3)
    x3 = pk_subj<SIGN>(t1, u);
}

template <int SIGN>
__device__ __forceinline__ void dft8(float2* v) {
    // n = 4*n1 + n2, k = k1 + 2*k2
    constexpr float c = 0.70710678118654752440f;
    float2 A0[4], A1[4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        A0[n2] = pk_add(v[n2], v[4 + n2]);
        A1[n2] = pk_sub(v[n2], v[4 + n2]);
    }
    // twiddles w8^{n2}, w8 = exp(SIGN*2*pi*i/8)
    A1[1] = pk_cmul(A1[1], make_float2(c, SIGN * c));
    A1[2] = mulj<SIGN>(A1[2]);
    A1[3] = pk_cmul(A1[3], make_float2(-c, SIGN * c));
    dft4<SIGN>(A0[0], A0[1], A0[2], A0[3]);
    dft4<SIGN>(A1[0], A1[1], A1[2], A1[3]);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        v[2 * k2] = A0[k2];
        v[2 * k2 + 1] = A1[k2];
    }
}

template <int SIGN>
__device__ __forceinline__ void dft16(float2* v) {
    // n = 4*n1 + n2, k = k1 + 4*k2
    constexpr float c1 = 0.92387953251128675613f;  // cos(pi/8)
    constexpr float s1 = 0.38268343236508977173f;  // sin(pi/8)
    constexpr float c2 = 0.70710678118654752440f;  // cos(pi/4)
    float2 A[4][4];  // A[k1][n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        float2 a = v[n2], b = v[4 + n2], c = v[8 + n2], d = v[12 + n2];
        dft4<SIGN>(a, b, c, d);
        A[0][n2] = a;
        A[1][n2] = b;
        A[2][n2] = c;
        A[3][n2] = d;
    }
    // twiddle w16^{n2*k1}, w16 = exp(SIGN*2*pi*i/16): exponent m -> (cos(pi m/8), SIGN sin(pi m/8))
    A[1][1] = pk_cmul(A[1][1], make_float2(c1, SIGN * s1));    // m=1
    A[1][2] = pk_cmul(A[1][2], make_float2(c2, SIGN * c2));    // m=2
    A[1][3] = pk_cmul(A[1][3], make_float2(s1, SIGN * c1));    // m=3
    A[2][1] = pk_cmul(A[2][1], make_float2(c2, SIGN * c2));    // m=2
    A[2][2] = mulj<SIGN>(A[2][2]);                             // m=4
    A[2][3] = pk_cmul(A[2][3], make_float2(-c2, SIGN * c2));   // m=6
    A[3][1] = pk_cmul(A[3][1], make_float2(s1, SIGN * c1));    // m=3
    A[3][2] = pk_cmul(A[3][2], make_float2(-c2, SIGN * c2));   // m=6
    A[3][3] = pk_cmul(A[3][3], make_float2(-c1, -SIGN * s1));  // m=9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        dft4<SIGN>(A[k1][0], A[k1][1], A[k1][2], A[k1][3]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = A[k1][k2];
    }
}

template <int R, int SIGN>
__device__ __forceinline__ void dftR(float2* v) {
    if constexpr (R == 2) dft2<SIGN>(v[0], v[1]);
    if constexpr (R == 4) dft4<SIGN>(v[0], v[1], v[2], v[3]);
    if constexpr (R == 8) dft8<SIGN>(v);
    if constexpr (R == 16) dft16<SIGN>(v);
}

// Per-size twiddle tables (context-owned, computed in float64 on the host), laid out in the order the passes
// consume them so that every warp-wide load is a contiguous 128/256-byte segment:
//   tw1[(k1-1)*N2 + m] = exp(+2 pi i m k1 / N)     m < N2 = N/R1, k1 = 1..R1-1   (pass 1, lanes = consecutive m)
//   tw2[(c-1)*16 + b]  = exp(+2 pi i b c / N2)     b < 16,        c  = 1..R2-1   (pass 2, lanes = consecutive b)
struct FftTw {
    const float2* tw1;
    const float2* tw2;
};
template <int SIGN>
__device__ __forceinline__ float2 tw_load(const float2* __restrict__ p) {
    float2 w = __ldg(p);
    if (SIGN < 0) w.y = -w.y;
    return w;
}

// Shared-memory geometry of one block FFT.
template <int R1, int R2, bool PAD>
struct FftGeom {
    static constexpr int N = R1 * R2 * 16;
    static constexpr int NT = N / 16;  // threads per FFT instance
    static constexpr int N2 = R2 * 16;
    static constexpr int S2 = (PAD && R1 < 16) ? 16 + R1 : 16;
    static constexpr int S1raw = R2 * S2;
    static constexpr int S1 = (PAD && R1 > 1) ? (S1raw + ((17 - (S1raw % 16)) % 16)) : S1raw;  // == 1 (mod 16)
    static constexpr int kElems = R1 * S1;  // float2 slots per instance
    __device__ __forceinline__ static int addr(int k1, int mid, int lo) { return k1 * S1 + mid * S2 + lo; }
};

// Block FFT.  `v` receives X[tf + NT*d] in v[d].  `load(n)` returns input element n.
// smem points at this instance's slot 0 (already offset by nl); RT = interleave stride.
template <int R1, int R2, int SIGN, bool PAD, class Load, bool PRELOADED = false>
__device__ __forceinline__ void block_fft(float2 (&v)[16], float2* smem, const int RT, const int tf,
                                          const FftTw tw, Load load) {
    using G = FftGeom<R1, R2, PAD>;
    constexpr int NT = G::NT, N2 = G::N2;
    if constexpr (R1 > 1) {
#pragma unroll
        for (int i = 0; i < 16 / R1; ++i) {
            const int m = tf + NT * i;  // 0..N2-1
#pragma unroll
            if constexpr (!PRELOADED) {
#pragma unroll
                for (int j = 0; j < R1; ++j) v[i * R1 + j] = load(N2 * j + m);
            }
            dftR<R1, SIGN>(&v[i * R1]);
#pragma unroll
            for (int k1 = 1; k1 < R1; ++k1)
                v[i * R1 + k1] = pk_cmul(v[i * R1 + k1], tw_load<SIGN>(tw.tw1 + (k1 - 1) * N2 + m));
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) smem[G::addr(k1, m >> 4, m & 15) * RT] = v[i * R1 + k1];
        }
        __syncthreads();
    }
    if constexpr (R2 > 1) {
#pragma unroll
        for (int i = 0; i < 16 / R2; ++i) {
            const int g = tf + NT * i;  // 0..R1*16-1
            const int k1 = g >> 4, b = g & 15;
#pragma unroll
            for (int a = 0; a < R2; ++a) {
                if constexpr (R1 > 1) v[i * R2 + a] = smem[G::addr(k1, a, b) * RT];
                else v[i * R2 + a] = load(a * 16 + b);
            }
            dftR<R2, SIGN>(&v[i * R2]);
#pragma unroll
            for (int c = 1; c < R2; ++c)
                v[i * R2 + c] = pk_cmul(v[i * R2 + c], tw_load<SIGN>(tw.tw2 + (c - 1) * 16 + b));
#pragma unroll
            for (int c = 0; c < R2; ++c) smem[G::addr(k1, c, b) * RT] = v[i * R2 + c];
        }
        __syncthreads();
    }
    {
        const int k1 = tf % R1, c = tf / R1;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            if constexpr (R1 * R2 > 1) v[b] = smem[G::addr(k1, c, b) * RT];
            else v[b] = load(b);
        }
        dft16<SIGN>(v);
    }
}

// log2 helpers for dispatch
inline int ilog2(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

}  // namespace isac
