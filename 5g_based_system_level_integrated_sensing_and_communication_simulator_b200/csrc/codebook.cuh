// K7: Type-I single-panel codebook (TS 38.214 5.2.2.2.1) as a beam/co-phasing table, and the
// TS 38.211 6.3.1.5 PUSCH codebooks.  Host-side generators; the SINR kernels never see W itself.
#pragma once
#include "isac_common.cuh"
#include <complex>
#include <vector>

namespace isac {

constexpr int kMaxLayers = 8;
constexpr int kMaxPmiBatch = 32;  // UEs per PMI launch (their noise variances travel as kernel parameters)
constexpr int kMaxBlocks = 8;  // column blocks of a precoder: 2 (v_lm ; phi v_lm), 4 (vbar: 1, theta, phi, phi*theta), P (explicit W, P <= 4)
                               // or 2*Ng (multi-panel: both polarisations of every panel)

// Mirrors isac_csi_config of include/isac_b200.h (validated reportConfig of dlPMISelect.m:511-851)
struct CsiConfig {
    int nPorts;                 // csirs.NumCSIRSPorts
    int N1, N2, O1, O2;         // PanelDimensions / OverSamplingFactors (Table 5.2.2.2.1-2)
    int codebookMode;           // 1 or 2
    int nSizeBWP, nStartBWP;
    int subbandSize;            // NSBPRB (0: wideband / BWP < 24 PRB)
    int pmiSubband;             // PMIMode == 'Subband'
    int cqiSubband;             // CQIMode == 'Subband'
    int K, L;                   // carrier.NSizeGrid*12, SymbolsPerSlot
    int nRx;
    const uint8_t* subsetRestriction;  // CodebookSubsetRestriction bits (N1*O1*N2*O2, or 6 for 2 ports), nullptr = all ones
    const uint8_t* i2Restriction;      // 16 bits, nullptr = all ones
    uint8_t riRestriction[8];
    int nRE;
    const int* reK;             // 1-based CSI-RS RE subscripts relative to the BWP (validateInputs :797-833)
    const int* reL;
    int nPanels;                // Ng of a Type1MultiPanel report (PanelDimensions = [Ng N1 N2]); 0 or 1 = Type1SinglePanel
};

// One precoder column = scale * [coef[0]*v ; coef[1]*v ; ...] with v = beams[beam]
struct LayerDesc {
    int beam;
    std::complex<double> coef[kMaxBlocks];
};

struct CodebookTable {
    int P = 0, nLayers = 0, NB = 0, Pb = 0;        // ports, layers, blocks per column, ports per block
    int n2 = 1, n11 = 1, n12 = 1, n13 = 1;         // index-set sizes [i2, i11, i12, i13]
    int nBeams = 0;
    double scale = 1.0;                             // 1/sqrt(nLayers*P) (or the table's own factor)
    std::vector<std::complex<double>> beams;        // [nBeams][Pb]
    std::vector<uint8_t> valid;                     // [nCand] 0 = restricted (all-zero W)
    std::vector<LayerDesc> layers;                  // [nCand][nLayers]
    std::vector<double> candScale;                  // [nCand] per-candidate factor (explicit codebooks), else empty
    // Type1MultiPanel: the reference's 9-D index set [i20 i21 i22 | i11 i12 i13 i141 i142 i143] is walked in MATLAB linear order,
    // so it is stored flattened as n2 = i20*i21*i22 and n13 = i13*i141*i142*i143; mp = {i20,i21,i22,i13,i141,i142,i143} lengths
    int mp[7] = {0, 0, 0, 0, 0, 0, 0};
    int nCand() const { return n2 * n11 * n12 * n13; }
};

enum CodebookVariant { kVariantUE = 0, kVariantGNB = 1 };

// getPMIType1SinglePanelCodebook (dlPMISelect.m:853-1349) / pmiType1SinglePanelCodebook.m:46-554
int build_type1sp_table(Ctx* ctx, const CsiConfig& c, int nLayers, int variant, CodebookTable& t);
// getPMIType1MultiPanelCodebook (dlPMISelect.m:1351-1772) as a beam / co-phasing table (2*Ng blocks per column), flattened
// index set (see CodebookTable::mp); c.nPanels = Ng
int build_type1mp_table(Ctx* ctx, const CsiConfig& c, int nLayers, CodebookTable& t);
// getPMIType1MultiPanelCodebook (dlPMISelect.m:1351-1772): dims = [i20 i21 i22 i11 i12 i13 i141 i142 i143] lengths, W (may be
// nullptr) = explicit [P x nLayers x prod(dims)] array, column-major, restricted precoders zero
int type1mp_codebook(Ctx* ctx, const CsiConfig& c, int nPanels, int nLayers, int dims[9], std::vector<std::complex<double>>* W);
// nrPUSCHCodebook(nlayers,nports,tpmi).' for tpmi = 0..maxTPMI as an explicit table (pmiSelect.m:45)
int build_pusch_table(Ctx* ctx, int nLayers, int nPorts, CodebookTable& t);
// W[P x nLayers x nCand] complex128, column-major (restricted candidates all zero)
void materialize_codebook(const CodebookTable& t, std::vector<std::complex<double>>& W);

// getDownlinkPMISubbandInfo (dlPMISelect.m:1836-1887)
void subband_info(bool subbandMode, int nStartBWP, int nSizeBWP, int nsbprb, std::vector<int>& sizes);

}  // namespace isac
