// Table-driven general kernels: any system the description format can express.
#include "trepb_kernels.cuh"

namespace trepb {
const KernelSet* general_kernels() {
    static const KernelSet ks = make_kernelset<RtSys>("general", 0ull, 0, 0, 0);
    return &ks;
}
}  // namespace trepb
