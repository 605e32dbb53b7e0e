// Library context: device, stream, error string, twiddle table, reusable staging buffers.
#pragma once
#include "isac_common.cuh"
#include <vector>

namespace isac {

struct Ctx {
    static constexpr int kPinnedSlots = 8;
    static constexpr int kScratchSlots = 26;
    int device = 0;
    int numSMs = 0;
    int ccMajor = 0;
    cudaStream_t ownStream = nullptr;
    cudaStream_t stream = nullptr;  // stream every launch / copy is enqueued on
    float2* d_twiddle = nullptr;    // packed per-size twiddle tables (see fft_core.cuh)
    size_t tw1Off[9] = {}, tw2Off[9] = {};  // offsets for N = 16 << i
    std::string err;
    void* pinned[kPinnedSlots] = {};
    size_t pinnedBytes[kPinnedSlots] = {};
    void* scratch[kScratchSlots] = {};
    size_t scratchBytes[kScratchSlots] = {};
    struct ProfRec { int slot; cudaEvent_t a, b; };
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> eventPool;
    long long launches = 0;
    // per-context state of the UL TPMI path (pending report + codebook / band-limit cache, comm.cu); freed by isac_destroy
    void* ulState = nullptr;
    void (*ulFree)(void*) = nullptr;
};

// grow-only pinned host / device scratch buffers owned by the context
int ctx_pinned(Ctx* ctx, int slot, size_t bytes, void** out);
int ctx_scratch(Ctx* ctx, int slot, size_t bytes, void** out);

}  // namespace isac
