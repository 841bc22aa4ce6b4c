// Host-only: emit a compile-time ("constexpr") system type for one concrete trepb_sysdesc.
//
// The generated struct has the same accessor names as RtSys (trepb_sys.h) but every table is a
// constexpr switch, so that once the math of trepb_math.cuh is instantiated on it and the
// system-sized loops are fully unrolled (TREPB_UNROLL_SYS) every lookup folds to a literal:
// the frame-tree walk of the reference (trep/_trep/frame.c:1078-2081, recursive cache builders
// over Python-owned structs) becomes straight-line register code for that system.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "trepb_pack.h"

namespace trepb {

// FNV-1a over the whole description (topology AND parameters): identifies one concrete system.
inline uint64_t desc_hash(const PackedSys& P) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* c = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
    };
    const RtSys& s = P.proto;
    int dims[8] = {s.nf, s.nd, s.nk, s.nu, s.nc, s.npot, s.nforce, s.max_depth};
    mix(dims, sizeof(dims));
    mix(P.blob.data(), P.blob.size());
    return h;
}

// ---- structure vs parameters -------------------------------------------------------------------------
// A specialised kernel is compiled for the STRUCTURE of a system: sizes, topology, plugin kinds and their index
// arguments, and which numeric entries are exactly zero (or, for the rows of a constant SE(3) transform, exactly
// +-1) - those fold away at compile time.  Every other number (masses, inertias, lengths, gravity, spring and
// damping constants, tolerances, spline tables) is a run-time PARAMETER: slot k of the block the launcher passes
// as a kernel argument (constant bank).  param_map lists the slots in a fixed order that both the generator and
// trepb_system_create derive from the description alone.
enum ParTable { PT_VALUE = 0, PT_SE3, PT_MASS, PT_POT_D, PT_FORCE_D, PT_CON_D, PT_DPOOL, PT_GRAV, PT_COUNT };
struct ParamMap {
    std::vector<int> slot[PT_COUNT];      // per table entry: -1 literal zero, -2 literal one, -3 literal minus one, >= 0 slot
    std::vector<double> values;           // slot -> value
};
inline size_t dpool_size(const RtSys& s) { return ((const char*)s.dep_ - (const char*)s.dpool_) / sizeof(double); }
inline size_t ipool_size(const RtSys& s) { return ((const char*)s.frame_value - (const char*)s.ipool_) / sizeof(int32_t); }
inline ParamMap param_map(const PackedSys& P) {
    const RtSys s = P.view(P.blob.data());
    ParamMap M;
    auto add = [&](int t, const double* data, size_t n, bool unit_literals) {
        M.slot[t].assign(n, -1);
        for (size_t i = 0; i < n; ++i) {
            const double v = data[i];
            if (v == 0.0) continue;
            if (unit_literals && v == 1.0) { M.slot[t][i] = -2; continue; }
            if (unit_literals && v == -1.0) { M.slot[t][i] = -3; continue; }
            M.slot[t][i] = (int)M.values.size();
            M.values.push_back(v);
        }
    };
    add(PT_VALUE, s.frame_value, s.nf, false);
    add(PT_SE3, s.frame_se3, (size_t)s.nf * 12, true);
    add(PT_MASS, s.frame_mass, (size_t)s.nf * 4, false);
    add(PT_POT_D, s.pot_d_, (size_t)s.npot * 4, false);
    add(PT_FORCE_D, s.force_d_, (size_t)s.nforce * 4, false);
    add(PT_CON_D, s.con_d_, (size_t)s.nc * 4, false);
    add(PT_DPOOL, s.dpool_, dpool_size(s), false);
    add(PT_GRAV, s.grav, 3, false);
    return M;
}
// FNV-1a over the structure only: two systems with the same struct_hash run the same specialised kernel.
inline uint64_t struct_hash(const PackedSys& P) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* c = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
    };
    const RtSys s = P.view(P.blob.data());
    const int nq = s.nd + s.nk;
    int dims[11] = {s.nf, s.nd, s.nk, s.nu, s.nc, s.npot, s.nforce, s.max_depth, s.has_gravity, s.has_pairs, s.has_pairs_mid};
    mix(dims, sizeof(dims));
    mix(s.frame_parent, 4 * s.nf); mix(s.frame_kind, 4 * s.nf); mix(s.frame_config, 4 * s.nf);
    mix(s.cfg_frame_, 4 * nq);
    mix(s.pot_kind_, 4 * s.npot); mix(s.pot_i_, 16 * s.npot);
    mix(s.force_kind_, 4 * s.nforce); mix(s.force_i_, 16 * s.nforce);
    mix(s.con_kind_, 4 * s.nc); mix(s.con_i_, 16 * s.nc);
    mix(s.ipool_, 4 * ipool_size(s));
    const ParamMap M = param_map(P);
    for (int t = 0; t < PT_COUNT; ++t)
        for (int v : M.slot[t]) { const signed char c = v >= 0 ? 1 : (signed char)v; mix(&c, 1); }
    return h;
}

namespace cg {
inline std::string dlit(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";  // "3" -> "3.0"; inf/nan never occur
    return s;
}
template <class T, class F>
inline std::string table(const char* rtype, const char* name, const char* args, const char* key,
                         size_t n, F lit, const char* dflt, const T* data) {
    std::string o = std::string("    TREPB_HD static constexpr ") + rtype + " " + name + "(" + args + ") {\n";
    if (n == 0) { o += std::string("        return ") + dflt + ";\n    }\n"; return o; }
    o += std::string("        switch (") + key + ") {\n";
    for (size_t i = 0; i < n; ++i) {
        std::string v = lit(data[i]);
        if (v == dflt) continue;
        o += "        case " + std::to_string(i) + ": return " + v + ";\n";
    }
    o += std::string("        default: return ") + dflt + ";\n        }\n    }\n";
    return o;
}
inline std::string ptable(const char* name, const char* args, const char* key, const std::vector<int>& slot) {
    std::string o = std::string("    TREPB_HD double ") + name + "(" + args + ") const {\n";
    bool any = false;
    for (int v : slot) any = any || v != -1;
    if (!any) { o += "        return 0.0;\n    }\n"; return o; }
    o += std::string("        switch (") + key + ") {\n";
    for (size_t i = 0; i < slot.size(); ++i) {
        const int v = slot[i];
        if (v == -1) continue;
        o += "        case " + std::to_string(i) + ": return " +
             (v == -2 ? std::string("1.0") : v == -3 ? std::string("-1.0") : "par.v[" + std::to_string(v) + "]") + ";\n";
    }
    o += "        default: return 0.0;\n        }\n    }\n";
    return o;
}
inline std::string ilit(int32_t v) { return std::to_string(v); }
inline std::string blit(uint8_t v) { return v ? "true" : "false"; }
}  // namespace cg

// literal = true: every number of THIS description becomes a literal too (hash = desc_hash, no parameters): the
// compiler folds products of constants; worth a second instantiation for the few systems the benchmark configs
// name exactly (damped pendulum: 7.3e10 against 6.8e10 DEL steps/s with run-time parameters).
inline std::string codegen_system(const PackedSys& P, const std::string& struct_name, bool literal = false) {
    using namespace cg;
    const RtSys s = P.view(P.blob.data());
    const int nf = s.nf, nq = s.nd + s.nk;
    ParamMap M = param_map(P);
    const int npar = literal ? 0 : (int)M.values.size();

    std::string o;
    // unroll count: just above the longest system-sized loop (frames, 2*nq rows of A, ...), so
    // that "#pragma unroll (kUnroll)" means full unrolling without tripping size thresholds
    int unroll = nf;
    if (2 * nq > unroll) unroll = 2 * nq;
    if (s.nd + s.nc > unroll) unroll = s.nd + s.nc;
    if (s.npot > unroll) unroll = s.npot;
    if (s.nforce > unroll) unroll = s.nforce;
    if (unroll < 4) unroll = 4;
    char head[2048];
    snprintf(head, sizeof(head),
             "// generated by trepb_codegen (trep_b200/csrc/trepb_codegen.h) - do not edit\n"
             "struct %s {\n"
             "    static constexpr bool kStatic = true;\n"
             "    static constexpr int kUnroll = %d;\n"
             "    static constexpr unsigned long long kHash = 0x%016llxull;   // struct_hash: structure only\n"
             "    static constexpr int kNPAR = %d;\n"
             "    struct Params { double v[kNPAR > 0 ? kNPAR : 1]; };\n"
             "    Params par;   // run-time parameters (kernel argument), slots in trepb::param_map order\n"
             "    static constexpr int kNF = %d, kND = %d, kNK = %d, kNQ = %d, kNU = %d, kNC = %d,\n"
             "                         kNPOT = %d, kNFORCE = %d, kMAXDEPTH = %d;\n",
             struct_name.c_str(), unroll, (unsigned long long)(literal ? desc_hash(P) : struct_hash(P)), npar, nf, s.nd, s.nk, nq, s.nu, s.nc,
             s.npot, s.nforce, s.max_depth);
    o += head;
    const char* sizes[][2] = {{"NF", "kNF"}, {"ND", "kND"}, {"NK", "kNK"}, {"NQ", "kNQ"}, {"NU", "kNU"},
                              {"NC", "kNC"}, {"NPOT", "kNPOT"}, {"NFORCE", "kNFORCE"},
                              {"MAXDEPTH", "kMAXDEPTH"}};
    for (auto& z : sizes)
        o += std::string("    TREPB_HD static constexpr int ") + z[0] + "() { return " + z[1] + "; }\n";
    o += table<int32_t>("int", "parent", "int f", "f", nf, ilit, "0", s.frame_parent);
    o += table<int32_t>("int", "kind", "int f", "f", nf, ilit, "0", s.frame_kind);
    o += table<int32_t>("int", "config", "int f", "f", nf, ilit, "-1", s.frame_config);
    o += ptable("value", "int f", "f", M.slot[PT_VALUE]);
    o += ptable("se3", "int f, int k", "f * 12 + k", M.slot[PT_SE3]);
    o += ptable("mass", "int f, int k", "f * 4 + k", M.slot[PT_MASS]);
    o += table<int32_t>("int", "cfg_frame", "int c", "c", nq, ilit, "-1", s.cfg_frame_);
    o += table<uint8_t>("bool", "dep", "int f, int c", ("f * " + std::to_string(nq) + " + c").c_str(),
                        (size_t)nf * nq, blit, "false", s.dep_);
    o += table<uint8_t>("bool", "mass_below", "int f", "f", nf, blit, "false", s.mass_below_);
    o += table<uint8_t>("bool", "need_world", "int f", "f", nf, blit, "false", s.need_world_);
    o += table<uint8_t>("bool", "vzero", "int f", "f", nf, blit, "false", s.vzero_);
    {
        std::vector<uint8_t> hm(nf, 0);
        for (int f = 0; f < nf; ++f) hm[f] = s.has_mass(f) ? 1 : 0;
        o += table<uint8_t>("bool", "has_mass", "int f", "f", nf, blit, "false", hm.data());
    }
    o += table<int32_t>("int", "pot_kind", "int i", "i", s.npot, ilit, "0", s.pot_kind_);
    o += table<int32_t>("int", "pot_i", "int i, int k", "i * 4 + k", (size_t)s.npot * 4, ilit, "-1", s.pot_i_);
    o += ptable("pot_d", "int i, int k", "i * 4 + k", M.slot[PT_POT_D]);
    o += table<int32_t>("int", "force_kind", "int i", "i", s.nforce, ilit, "0", s.force_kind_);
    o += table<int32_t>("int", "force_i", "int i, int k", "i * 4 + k", (size_t)s.nforce * 4, ilit, "-1", s.force_i_);
    o += ptable("force_d", "int i, int k", "i * 4 + k", M.slot[PT_FORCE_D]);
    o += table<int32_t>("int", "con_kind", "int i", "i", s.nc, ilit, "0", s.con_kind_);
    o += table<int32_t>("int", "con_i", "int i, int k", "i * 4 + k", (size_t)s.nc * 4, ilit, "-1", s.con_i_);
    o += ptable("con_d", "int i, int k", "i * 4 + k", M.slot[PT_CON_D]);
    // pools: sizes are not in RtSys; recover them from the blob layout (ipool is followed by
    // frame_value, dpool by dep) - both 16-byte padded, so emit up to the padded length.
    {
        o += table<int32_t>("int", "ipool", "int k", "k", ipool_size(s), ilit, "0", s.ipool_);
        o += ptable("dpool", "int k", "k", M.slot[PT_DPOOL]);
    }
    o += ptable("gravity", "int k", "k", M.slot[PT_GRAV]);
    o += std::string("    TREPB_HD static constexpr bool gravity_on() { return ") + (s.has_gravity ? "true" : "false") + "; }\n";
    o += std::string("    TREPB_HD static constexpr bool pairs_on() { return ") + (s.has_pairs ? "true" : "false") + "; }\n";
    o += std::string("    TREPB_HD static constexpr bool pairs_mid() { return ") + (s.has_pairs_mid ? "true" : "false") + "; }\n";
    o += "};\n";
    if (literal) {
        // replace every parameter slot by its value
        for (int k = (int)M.values.size() - 1; k >= 0; --k) {
            const std::string key = "par.v[" + std::to_string(k) + "]", val = dlit(M.values[k]);
            size_t pos = 0;
            while ((pos = o.find(key, pos)) != std::string::npos) { o.replace(pos, key.size(), val); pos += val.size(); }
        }
    }
    return o;
}

}  // namespace trepb
