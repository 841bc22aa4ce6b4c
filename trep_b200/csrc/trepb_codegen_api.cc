// Host-only part of the C ABI: code generation and description hashing (include/trepb.h).
// Built twice: into libtrepb.so, and on its own (g++, no CUDA) as the build-time generator that
// trep_b200/build.py loads to emit gen/spec_*.cu for the ahead-of-time specialised systems.
#include <string.h>
#include <string>

#include "../../include/trepb.h"
#include "trepb_codegen.h"
#include "trepb_coop_sys.h"
#include "trepb_err.h"
#include "trepb_pack.h"

using namespace trepb;

extern "C" {

static int codegen_impl(const trepb_sysdesc* desc, const char* struct_name, char* buf, int cap, bool literal) {
    PackedSys P;
    std::string err;
    if (!pack_system(desc, &P, &err)) { last_error() = err; return -1; }
    const std::string text = codegen_system(P, struct_name && *struct_name ? struct_name : "CtSys", literal);
    const int need = (int)text.size() + 1;
    if (buf && cap > 0) {
        const int n = need <= cap ? need - 1 : cap - 1;
        memcpy(buf, text.data(), n);
        buf[n] = 0;
    }
    return need;
}

int trepb_codegen(const trepb_sysdesc* desc, const char* struct_name, char* buf, int cap) {
    return codegen_impl(desc, struct_name, buf, cap, false);
}
int trepb_codegen_literal(const trepb_sysdesc* desc, const char* struct_name, char* buf, int cap) {
    return codegen_impl(desc, struct_name, buf, cap, true);
}

uint64_t trepb_desc_hash(const trepb_sysdesc* desc) {
    PackedSys P;
    std::string err;
    if (!pack_system(desc, &P, &err)) { last_error() = err; return 0; }
    return desc_hash(P);
}

uint64_t trepb_struct_hash(const trepb_sysdesc* desc) {
    PackedSys P;
    std::string err;
    if (!pack_system(desc, &P, &err)) { last_error() = err; return 0; }
    return struct_hash(P);
}

int trepb_coop_dims(const trepb_sysdesc* desc, int32_t* out) {
    PackedSys P;
    std::string err;
    if (!pack_system(desc, &P, &err)) { last_error() = err; return TREPB_ERR_INVALID; }
    CoopPack C = coop_pack(desc);
    if (!C.ok) { last_error() = "cooperative kernels do not apply: " + C.why; return TREPB_ERR_UNSUPPORTED; }
    const CoopSys& s = C.proto;
    // the eleventh entry: 1 when the system has LinearSprings / LinearDampers / spline springs / wrenches (CtDims<..., 1>)
    const int32_t v[11] = {s.nd, s.nk, s.nu, s.nc, s.nl, s.np, s.npairs, s.nlevels, s.ndc, s.nqc,
                           (s.ns + s.nfd + s.nns + s.nw) > 0 ? 1 : 0};
    for (int i = 0; i < 11; ++i) out[i] = v[i];
    return TREPB_OK;
}

int trepb_validate(const trepb_sysdesc* desc) {
    PackedSys P;
    std::string err;
    if (!pack_system(desc, &P, &err)) { last_error() = err; return TREPB_ERR_INVALID; }
    return TREPB_OK;
}

}  // extern "C"
