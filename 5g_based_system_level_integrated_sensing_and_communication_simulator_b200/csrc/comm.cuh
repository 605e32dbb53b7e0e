// K8-K10, K12: Type-I PMI / RI / CQI selection, UL TPMI selection, PRG precoding.
#pragma once
#include "codebook.cuh"
#include <complex>
#include <map>
#include <unordered_map>
#include <utility>
#include <vector>

namespace isac {

// Dictionary of the beam-response inner products ("Gram pairs") the candidates of one report configuration need, shared
// by all ranks whose codebooks are built from the same beams (K9', comm.cu).  Host side; uploaded lazily.
struct PmiShared {
    struct Column { int beam; std::complex<double> coef[kMaxBlocks]; };
    int NB = 0, Pb = 0, nBeams = 0, P = 0;
    std::vector<double2> beams;                         // storage order [nBeams][Pb]
    std::vector<Column> cols;                           // distinct precoder columns (beam, co-phasing)
    std::map<std::vector<long long>, int> colIdx;
    std::vector<uint32_t> pairs;                        // atom a | atom a' << 16 ; Gamma = <Bf[a], Bf[a']>
    std::unordered_map<uint32_t, int> pairIdx;
    std::vector<double2> pal;                           // palette of coefficient products conj(c_i) c_j; pal[0] = 0
    std::map<std::pair<long long, long long>, int> palIdx;
    int cpT = 4;                                        // terms per column pair (NB^2 rounded up to 4)
    std::vector<uint32_t> cpTerms;                      // [nCP][cpT]: pair index | palette index << 16
    std::unordered_map<unsigned long long, int> cpIdx;
    uint32_t* d_pairs = nullptr;
    double2* d_pal = nullptr;
    uint32_t* d_cpTerms = nullptr;
    size_t upPairs = 0, upPal = 0, upCp = 0;            // sizes of the device copies
    int refs = 0;
    bool ok = true;                                     // false: dictionary outgrew 16-bit indices
};

// results of one dlPMISelect evaluation for one UE (host side)
struct PmiResult {
    bool allNaN = false;             // no CSI-RS RE in the BWP / everything restricted
    int i1[3] = {0, 0, 0};           // 1-based [i11 i12 i13]
    std::vector<double> i2;          // per subband, 1-based, NaN where CSI-RS is absent
    std::vector<double> sinrSel;     // [nSB x nLayers] SINRPerSubband at the reported PMI (NaN rows as above)
    std::vector<double> sinrWbSel;   // [nCqiSB x nLayers] per-CQI-subband SINR with a single (wideband) i2 (cqiSelect.m:586-596)
};

// Device plan for one (report configuration, rank): codebook tables + work buffers for `maxBatch` UEs.
struct PmiPlan {
    Ctx* ctx = nullptr;
    CsiConfig cfg{};
    std::vector<int> reK, reL;
    std::vector<uint8_t> csr, i2r;
    int nLayers = 0, maxBatch = 0;
    CodebookTable tab;
    std::vector<int> sbSizes, cqiSbSizes;  // PMI / CQI subband sizes in PRBs
    int nSB = 0, nCqiSB = 0;
    // device
    double2* d_beams = nullptr;
    int* d_layerBeam = nullptr;     // [nCand*nLayers]
    double2* d_layerCoef = nullptr; // [nCand*nLayers*NB]
    double* d_candScale = nullptr;  // [nCand]
    uint8_t* d_valid = nullptr;     // [nCand]
    int* d_reK = nullptr; int* d_reL = nullptr;
    int* d_reSb = nullptr;          // PMI subband of each RE
    double* d_reW = nullptr;        // mean-of-means weight of each RE within its PMI subband
    int* d_reCqiSb = nullptr;       // CQI subband of each RE
    double* d_reCqiW = nullptr;
    double* d_S = nullptr;          // SINRPerRE compact [nCand][nLayers][nRE][batch]
    double* d_total = nullptr;      // plain per-subband sums [nCand][nLayers][nSB][batch] (wideband totals derive from them)
    double* d_sub = nullptr;        // SINRPerSubband [nCand][nLayers][nSB][batch]
    int* d_sel = nullptr;           // [4 + nSB][batch]: allNaN flag, i11, i12, i13 (0-based), i2 per subband (0-based, -1 = NaN)
    double* d_sinrSel = nullptr;    // [nLayers][nSB][batch]
    double* d_sinrWb = nullptr;     // [nLayers][nCqiSB][batch]
    int* d_sbStart = nullptr;       // [nSB+1] RE ranges of the PMI subbands (REs sorted by subcarrier)
    int* d_cqiStart = nullptr;      // [nCqiSB+1]
    double* d_nVar = nullptr;       // [batch] (unused: nVar travels as a kernel parameter)
    PmiShared* sh = nullptr;        // Gram-pair dictionary (shared between the ranks of a CSI plan)
    uint16_t* d_ent = nullptr;      // [nCand][ntPad]: column-pair index of every packed lower-triangle entry
    std::vector<uint16_t> entH;     // host copy
    uint16_t* d_entF = nullptr;     // fused path: the same entries as table SLOT indices of RE 0, id * G + swizzle(id) (bit 15 kept)
    double* d_invScale2 = nullptr;  // [nCand] 1/scale^2 (explicit codebooks only)
    double invS2 = 1.0;             // 1/scale^2 of the rank
    int ntPad = 0;
    bool direct = false;            // force the direct (H*W) kernel
    char* d_res = nullptr;          // arena holding d_sel | d_sinrSel | d_sinrWb (own allocation or a slice of a CSI plan's)
    size_t resOff[3] = {0, 0, 0}, resBytes = 0;
    bool ownsRes = false;
    char* hostRes = nullptr;        // where the arena lands on the host (pinned)
    void* pin = nullptr;            // pinned staging of the selection results
    size_t pinBytes = 0;
    std::vector<uint8_t> sbHasRE, cqiSbHasRE;
    // fused report path (comm_fused.cu): subband sums accumulated inside the SINR kernel, SINRPerRE never stored
    bool fused = false;             // requested by the owner (RI / CQI report plans); dlPMISelect plans keep SINRPerRE for `info`
    bool uniformW = false;          // the mean-of-means weights are uniform inside every PMI and CQI subband
    std::vector<int> sbStartH, cqiStartH;   // host copies of d_sbStart / d_cqiStart
    std::vector<double> wH, cwH;            // host copies of d_reW / d_reCqiW
    int fG = 0, nChunks = 0;        // REs per chunk the chunk tables were built for, number of chunks
    int* d_chunkRe0 = nullptr; int* d_chunkN = nullptr;
    int* d_sbChunk = nullptr; int* d_cqiChunk = nullptr;    // [nSB+1] / [nCqiSB+1] chunk ranges
    double* d_sbW = nullptr; double* d_cqiSbW = nullptr;   // per-subband RE weight
    double* d_part = nullptr;       // [maxBatch][nChunks][nLayers][nCand] chunk partial sums
};

int pmi_plan_create(Ctx* ctx, const CsiConfig& cfg, int nLayers, int maxBatch, PmiPlan** out, PmiShared* share = nullptr,
                    bool fused = false);
// fused report path (comm_fused.cu)
int pair_sync_dict(Ctx* ctx, PmiShared* sh, cudaStream_t st);
int pmi_fused_pick(const PmiShared* sh, int R, int* threads, int* minb);
int pmi_plan_prepare_fused(PmiPlan* p, int G);
int pmi_fused_run(PmiPlan* const* grp, int n, const float2* H, const double* nVar, int batch, cudaStream_t st);
void pmi_plan_destroy(PmiPlan* p);
void pmi_plan_use_arena(PmiPlan* p, char* dev, char* host);
// H: device complex64 [K x L x nRx x P x batch]; nVar: host [batch]
int pmi_select_run(PmiPlan* p, const float2* H, const double* nVar, int batch, cudaStream_t st);
// several ranks of one report configuration: one fused SINR launch for the plans that share a dictionary
int pmi_select_run_multi(PmiPlan* const* plans, int n, const float2* H, const double* nVar, int batch, cudaStream_t st);
int pmi_select_collect(PmiPlan* p, int batch, std::vector<PmiResult>& out);   // synchronises
// split form: enqueue the D2H copies for several plans, synchronise once, then parse
int pmi_select_collect_enqueue(PmiPlan* p, int batch, cudaStream_t st);
int pmi_select_collect_finish(PmiPlan* p, int batch, std::vector<PmiResult>& out);
// optional big outputs of the last run (host): SINRPerRE [nRE x nLayers x nCand x batch], SINRPerSubband [nSB x nLayers x nCand x batch]
int pmi_get_sinr_arrays(PmiPlan* p, int batch, double* sinrPerRE, double* sinrPerSubband);

// riSelect (riSelect.m:254-294) and cqiSelect tail (cqiSelect.m:610-695) on the host from PmiResults
struct CsiReport {
    double RI = NAN;
    PmiResult pmi;                   // at the reported rank
    std::vector<double> cqi;         // [(nCqiSB+1 or 1) x nCW] column-major
    std::vector<double> sinrPerSubbandPerCW;
    int nCqiRows = 0, nCW = 0;
};
double ri_total_sinr(const PmiResult& r, int nSB, int rank);   // totalSINR(rankIdx) of riSelect.m:266-282
void cqi_from_pmi(const CsiConfig& cfg, int nLayers, const PmiResult& r, int nSB, int nCqiSB, const double* sinrTable,
                  int tableLen, CsiReport& rep);

// UL: pmiSelect (pmiSelect.m:28-66).  hest device complex64 [K x nSym x nRx x P]
struct UlPmiResult {
    bool none = false;               // no estimates / zero noise -> NaN outputs
    std::vector<double> pmi;         // [nSB] 0-based TPMI or NaN
    std::vector<double> sinr;        // [nSB x nTPMI] column-major
    std::vector<int> subbandIndices; // [nSB x 2] column-major, 1-based subcarriers
    int nSB = 0, nTPMI = 0;
};
int ul_pmi_select_run(Ctx* ctx, int nLayers, const float2* hest, int K, int nSym, int nRx, int P, double noiseEst,
                      int bandSize, UlPmiResult& out, cudaStream_t st);

// `batch` independent estimates stacked along a 5th dimension (one synchronisation for all of them)
int ul_pmi_select_batch(Ctx* ctx, int nLayers, const float2* hest, int K, int nSym, int nRx, int P, double noiseEst,
                        int bandSize, int batch, std::vector<UlPmiResult>& out, cudaStream_t st);
// the same in two halves: kernels + async result copy (no synchronisation) / wait for that copy + host tail
int ul_pmi_select_batch_enqueue(Ctx* ctx, int nLayers, const float2* hest, int K, int nSym, int nRx, int P, double noiseEst,
                                int bandSize, int batch, cudaStream_t st);
int ul_pmi_select_batch_finish(Ctx* ctx, std::vector<UlPmiResult>& out);

// precodedSINR (precodedSINR.m:11-18) for `batch` REs sharing W: H [nRx x P x batch], W [P x nLayers] complex128 (device)
int precoded_sinr_run(Ctx* ctx, const double2* H, int R, int P, double sigma, const double2* W, int nLayers, int batch,
                      double* out, cudaStream_t st);

// prgPrecode (prgPrecode.m:53-144): portsym/portind [NRE x nLayers], F [nLayers x P x NPRG] (device);
// antsym [NRE x P] complex64, antind [NRE x P] int32 (1-based).  scratch grid owned by ctx.
int prg_precode_run(Ctx* ctx, int K, int Lsym, int nStartGrid, const float2* portsym, const int* portind, int NRE,
                    int nLayers, const float2* F, int P, int NPRG, float2* antsym, int* antind, int batch, cudaStream_t st);

}  // namespace isac
