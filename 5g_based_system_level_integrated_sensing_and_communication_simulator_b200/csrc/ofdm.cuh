// CP-OFDM modulation of the sensing transmit grid (gNBPhy.m:599): declarations shared with capi.cu.
#pragma once
#include "isac_common.cuh"

namespace isac {

constexpr int kOfdmMaxSymPerSubframe = 56;  // symbols per subframe up to 60 kHz subcarrier spacing

struct OfdmConfig {
    int nSc, nSym, nAnts, nfft;
    int symbolsPerSubframe;  // length of cpLengths
    const int* cpLengths;    // nrOFDMInfo.CyclicPrefixLengths of one subframe
    double scale;            // signalAmp (gNBPhy.m:599)
    // extended form (isac_ofdm_modulate_ex_dev): a block of symbols written into resident buffers
    int windowing = 0;       // raised-cosine window / overlap length in samples (0: plain CP-OFDM)
    int symPhase = 0;        // position of the block's first symbol in the subframe's CP pattern
    long long gridStride = 0;   // symbols per antenna page of the SOURCE grid (0: nSym)
    long long waveStride = 0;   // samples per antenna of the DESTINATION buffer (0: the block's own length)
    long long sampleOffset = 0; // first sample of the block in the destination buffer
};

long long ofdm_waveform_length(const OfdmConfig& c);
// grid: device [nSc x nSym x nAnts] float2; wave: device [T x nAnts] float2, T = ofdm_waveform_length(c)
int ofdm_modulate_run(Ctx* ctx, const OfdmConfig& c, const float2* grid, float2* wave, cudaStream_t st);

}  // namespace isac
