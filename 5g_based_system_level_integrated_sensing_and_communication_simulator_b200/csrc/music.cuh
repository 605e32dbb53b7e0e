// K5 + K6: covariance matrices, Hermitian eigen-decomposition (one-sided Jacobi, float64) and
// MUSIC pseudo-spectrum scans.  Declarations shared with capi.cu.
#pragma once
#include "isac_common.cuh"

namespace isac {

constexpr int kMaxPeaks = 64;      // NPeaks capacity of the on-device findpeaks
constexpr int kSmallEigMax = 64;   // single-CTA eigen-solver limit

// Ra[i,j] = sum_t conj(x_i[t]) x_j[t] / N  with x_i = page i of the grid (fft2D.m:106-107, music2D.m:57-58)
// rx: [N x nAnts x batch] float2 (N = nSc*nSym);  Ra: [nAnts x nAnts x batch] double2 column-major.
int cov_antenna(Ctx* ctx, const float2* rx, long long N, int nAnts, int batch, double2* Ra, cudaStream_t st);

// Eigen-decomposition of Hermitian PSD matrices A[n x n x batch], n <= kSmallEigMax:
// w[n x batch] descending, V[n x n x batch] matching columns.
int eig_psd_small(Ctx* ctx, const double2* A, int n, int batch, double* w, double2* V, cudaStream_t st);

// One-sided (Hestenes) Jacobi on G[m x n] (in place: G <- G*V), optional V[n x n] accumulation.
// On return: sigma[n] = column norms (descending order given by order[]), order[k] = column index of
// the k-th largest.  Synchronises the stream once per sweep.
int svd_onesided_jacobi(Ctx* ctx, double2* G, int m, int n, double2* V, double* sigma, int* order,
                        int* sweepsOut, cudaStream_t st);

// DoA scanners sharing the eigen-decomposition of Ra: doaEstimation.music (music.m:1), mvdrBF (mvdrBF.m:1), digitalBF (digitalBF.m:1)
enum DoaMethod : int { kDoaMusic = 0, kDoaMvdr = 1, kDoaDbf = 2 };

struct DoaConfig {
    int isUpa;        // 0 = ULA (music.m:73-104), 1 = UPA (music.m:31-71)
    int nAnts;        // ULA: array.numElements
    int nX, nY;       // UPA: array.nV, array.nH
    double d;         // element spacing / lambda (music.m:12)
    double aGran, aMax, eGran, eMax;  // scan granularity / scale in degrees (radarParams.m:120-124)
};

// L selection source for the DoA stage
struct LSource {
    const int* givenL;         // device int per batch, or nullptr
    const uint32_t* rowmask;   // CFAR detected-row bitmap (L = popcount), or nullptr
    int rowWords;
    int fixedL;                // used when > 0 and both pointers are null; <= 0 -> eigen-gap rule (music.m:109-125)
};

// ULA MUSIC from eigenpairs (n <= kSmallEigMax): outputs per batch item
//   Lout[1], P[aSteps] (abs(1/(a'Unn a + eps))), PdB[aSteps], peakLoc[kMaxPeaks] (1-based), nPeaks[1], status[1]
int music_doa_ula(Ctx* ctx, const double* w, const double2* V, int n, int batch, const DoaConfig& cfg,
                  const LSource& ls, int* Lout, double* P, double* PdB, int* peakLoc, int* nPeaks,
                  int* status, cudaStream_t st, int method = kDoaMusic);

// UPA MUSIC (any n = nX*nY): complement form with the L leading eigenvectors.
//   vecs: [n x nVecs] column-major; order: indices of the leading columns (descending eigenvalue);
//   PdB [eSteps x aSteps] column-major = mag2db(Pmusic/max(Pmusic)) with Pmusic = -abs(...) (music.m:61-63)
int music_doa_upa(Ctx* ctx, const double2* vecs, long long ld, const int* order, int n, const DoaConfig& cfg,
                  const int* dL, double* P, double* PdB, cudaStream_t st, int method = kDoaMusic,
                  const double* wDesc = nullptr);

// 1-D complement-form scan  q[i] = len - sum_{k<L} |<u_k, a(x_i)>|^2, a[n] = exp(2*pi*j*coef*x_i*n)
//   vecs columns are scaled by colScale[order[k]] (nullptr -> 1) and conjugated when conjVec != 0.
int music_scan_1d(Ctx* ctx, const double2* vecs, long long ld, const int* order, const double* colInvNorm,
                  int len, int nVecs, int conjVec, const int* dL, double coef, double x0, double dx, int steps,
                  double* q, cudaStream_t st);

// P = abs(1/q) ; PdB = mag2db(P/max(P)) ; findpeaks(PdB,'NPeaks',L,'SortStr','descend')
int music_finish_1d(Ctx* ctx, const double* q, int steps, const int* dL, double* P, double* PdB, int* peakLoc,
                    int* nPeaks, cudaStream_t st);

// eigen-gap rule on descending eigenvalues (device): Lout = determineNumTargets(ascending(w))
int music_num_targets(Ctx* ctx, const double* wDesc, int n, int* Lout, cudaStream_t st);

// H = rx(:,:,1).*conj(tx(:,:,1)) as double2, optionally conjugate-transposed (music2D.m:67-68)
int music2d_channel(Ctx* ctx, const float2* rx, const float2* tx, int nSc, int nSym, int transpose, double2* H,
                    cudaStream_t st);
int set_identity(Ctx* ctx, double2* V, int n, cudaStream_t st);

}  // namespace isac
