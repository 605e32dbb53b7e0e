// K1 + K2 declarations: radar echo synthesis (basicRadarChannel) and fused OFDM demodulation (monoStaticSensing).
#pragma once
#include "isac_common.cuh"
#include <cstring>

namespace isac {

constexpr int kEchoMaxTargets = 16;
constexpr int kEchoMaxSymPerSubframe = 56;  // symbols per subframe up to 60 kHz subcarrier spacing
constexpr int kEchoInlineSteer = 128;        // nAnts*nTargets steering entries carried as kernel parameters  // LoS targets per call (bounded further by shared memory: nTgt*nSc*8 B)

// host-side description (mirrors isac_echo_config of include/isac_b200.h)
struct EchoConfig {
    long long T;            // size(txWaveform,1)
    int nTx;                // size(txWaveform,2) == number of Rx antennas
    int nTargets;
    double fc, fs, N0;      // radarParams.fc / .fs / .N0
    const double* range;    // [nTargets] radarParams.range
    const double* velocity; // [nTargets] radarParams.velocity
    const double* largeScaleFading;  // [nTargets]
    const double* steeringVec;       // complex128 [nTx x nTargets] radarParams.RxSteeringVec
    const int* los;         // [nTargets] targetLoSConditions (nullptr = all LoS)
    int nfft, nSc;          // OFDM numerology (nrOFDMInfo)
    int nSymTx;             // txDimension(2)
    int symbolsPerSubframe; // length of cpLengths
    const int* cpLengths;   // CP length of each symbol of one subframe
};

int radar_channel_run(Ctx* ctx, const EchoConfig& c, const float2* tx, const float2* noise, int noiseMode,
                      unsigned long long seed, float2* rxWave, cudaStream_t st);
// echoGrid == nullptr: only report the number of output symbols
int mono_static_sensing_run(Ctx* ctx, const EchoConfig& c, const float2* tx, const float2* noise, int noiseMode,
                            unsigned long long seed, float2* echoGrid, int* nSymOut, cudaStream_t st);

}  // namespace isac
