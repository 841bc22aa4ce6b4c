// Team-cooperative kernels, run-time-size flavour (any system the link tables can express) and
// the registry of compile-time-size flavours (generated gen/coop_<name>.cu files).
#include <stdlib.h>
#include "trepb_coop_kernels.cuh"

namespace trepb {

CoopRegistry& coop_registry() {
    static CoopRegistry r = {{nullptr}, 0};
    return r;
}

bool coop_register(const CoopKernelSet* ks, unsigned abi) {
    CoopRegistry& r = coop_registry();
    if (abi != kCoopAbi || r.n >= CoopRegistry::kMax) return false;
    r.sets[r.n++] = ks;
    return true;
}

const CoopKernelSet* coop_general_kernels() {
    static const CoopKernelSet ks = make_coop_kernelset<RtDims>("cooperative");
    return &ks;
}

const CoopKernelSet* coop_select(const CoopSys& s, bool allow_specialized, int team_warps) {
    if (allow_specialized) {
        // several flavours may be registered for one shape (one warp per instance, two warps, one warp with the
        // external-slab layout): the external-slab one wins, then the widest team, unless the caller
        // (TREPB_FLAG_COOP_ONE_WARP: the plain one-warp flavour) or TREPB_COOP_TEAM=<warps> (diagnostic; 3 = external
        // slab) asks for another
        CoopRegistry& r = coop_registry();
        const char* e = getenv("TREPB_COOP_TEAM");
        const int want = team_warps > 0 ? team_warps : (e ? atoi(e) : 0);
        const CoopKernelSet* best = nullptr;
        auto rank = [](const CoopKernelSet* k) { return k->ext ? 100 : k->team_warps; };
        for (int i = 0; i < r.n; ++i) {
            if (!r.sets[i]->matches(s)) continue;
            if (want > 0 && (want == 3 ? r.sets[i]->ext != 0 : (r.sets[i]->team_warps == want && !r.sets[i]->ext))) return r.sets[i];
            if (!best || rank(r.sets[i]) > rank(best)) best = r.sets[i];
        }
        if (best) return best;
    }
    return coop_general_kernels();
}

}  // namespace trepb
