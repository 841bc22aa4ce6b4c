// NVTX ranges around the C-ABI entry points (header-only NVTX v3: a no-op unless a profiler is attached).
// They play the role of the reference's callgrind markers around its hot calls
// (trep/_trep/midpointvi.c:2674-2738) for nsys / ncu timelines: one range per library call, named after it.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace trepb {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace trepb
#define TREPB_NVTX(name) ::trepb::NvtxRange trepb_nvtx_range_(name)
