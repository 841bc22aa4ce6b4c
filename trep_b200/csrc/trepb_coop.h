// Launch interface of the team-cooperative kernels (trepb_coop.cu).
#pragma once
#include <cuda_runtime.h>
#include "trepb_kernels.cuh"
#include "trepb_coop_math.cuh"

namespace trepb {

struct CoopLaunch {
    int grid, warps;        // CTAs, warps (= instances in flight) per CTA
    size_t smem;            // table blob + warps x workspace
    cudaStream_t stream;
    CoopSys sys;            // view whose base is the DEVICE copy of the blob
    int blob_bytes;
    CoopLayout lay;
};

cudaError_t coop_step(const CoopLaunch& c, const StepParams& p);
cudaError_t coop_p2(const CoopLaunch& c, const P2Params& p);
cudaError_t coop_lin(const CoopLaunch& c, const LinParams& p, const AuxLayout& al);
cudaError_t coop_kernel_info(int which, KernelInfo* info);  // 0 step, 1 p2, 2 lin

}  // namespace trepb
