// Team-cooperative kernels, run-time-size flavour (any system the link tables can express) and
// the registry of compile-time-size flavours (generated gen/coop_<name>.cu files).
#include "trepb_coop_kernels.cuh"

namespace trepb {

CoopRegistry& coop_registry() {
    static CoopRegistry r = {{nullptr}, 0};
    return r;
}

const CoopKernelSet* coop_general_kernels() {
    static const CoopKernelSet ks = make_coop_kernelset<RtDims>("cooperative");
    return &ks;
}

const CoopKernelSet* coop_select(const CoopSys& s, bool allow_specialized) {
    if (allow_specialized) {
        CoopRegistry& r = coop_registry();
        for (int i = 0; i < r.n; ++i)
            if (r.sets[i]->matches(s)) return r.sets[i];
    }
    return coop_general_kernels();
}

}  // namespace trepb
