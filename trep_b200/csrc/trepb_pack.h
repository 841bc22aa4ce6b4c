// Host-only: validate a trepb_sysdesc, derive the structure tables (what the reference calls
// cache_index / config.masses, trep/frame.py:683-691, trep/system.py:769-771) and pack
// everything into one relocatable blob whose RtSys view can point at host or device memory.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/trepb.h"
#include "trepb_sys.h"

namespace trepb {

struct PackedSys {
    std::vector<char> blob;
    RtSys proto;                 // sizes + scalars; pointers filled by view()
    size_t off[32];
    int nf, nq;
    std::vector<int32_t> cfg_frame;
    std::vector<uint8_t> dep, mass_below, need_world, vzero;

    RtSys view(const char* base) const {
        RtSys s = proto;
        int k = 0;
        s.frame_parent = (const int32_t*)(base + off[k++]);
        s.frame_kind = (const int32_t*)(base + off[k++]);
        s.frame_config = (const int32_t*)(base + off[k++]);
        s.cfg_frame_ = (const int32_t*)(base + off[k++]);
        s.pot_kind_ = (const int32_t*)(base + off[k++]);
        s.pot_i_ = (const int32_t*)(base + off[k++]);
        s.force_kind_ = (const int32_t*)(base + off[k++]);
        s.force_i_ = (const int32_t*)(base + off[k++]);
        s.con_kind_ = (const int32_t*)(base + off[k++]);
        s.con_i_ = (const int32_t*)(base + off[k++]);
        s.ipool_ = (const int32_t*)(base + off[k++]);
        s.frame_value = (const double*)(base + off[k++]);
        s.frame_se3 = (const double*)(base + off[k++]);
        s.frame_mass = (const double*)(base + off[k++]);
        s.pot_d_ = (const double*)(base + off[k++]);
        s.force_d_ = (const double*)(base + off[k++]);
        s.con_d_ = (const double*)(base + off[k++]);
        s.dpool_ = (const double*)(base + off[k++]);
        s.dep_ = (const uint8_t*)(base + off[k++]);
        s.mass_below_ = (const uint8_t*)(base + off[k++]);
        s.need_world_ = (const uint8_t*)(base + off[k++]);
        s.vzero_ = (const uint8_t*)(base + off[k++]);
        return s;
    }
};

inline bool pack_system(const trepb_sysdesc* d, PackedSys* out, std::string* err) {
    auto fail = [&](const std::string& m) { *err = m; return false; };
    if (!d) return fail("null description");
    const int nf = d->n_frames, nd = d->nd, nk = d->nk, nq = nd + nk, nu = d->nu;
    const int np = d->n_potentials, nfo = d->n_forces, nc = d->n_constraints;
    if (nf < 1 || nd < 0 || nk < 0 || nu < 0 || np < 0 || nfo < 0 || nc < 0) return fail("negative size");
    if (nd < 1) return fail("system has no dynamic configuration");
    if (d->frame_kind[0] != TREPB_WORLD || d->frame_parent[0] != -1) return fail("frame 0 must be the world frame");
    std::vector<int32_t> cfg_frame(nq > 0 ? nq : 1, -1);
    for (int f = 1; f < nf; ++f) {
        const int p = d->frame_parent[f], k = d->frame_kind[f], c = d->frame_config[f];
        if (p < 0 || p >= f) return fail("frames must be in pre-order (parent index < child index)");
        if (k < TREPB_TX || k > TREPB_CONST_SE3) return fail("unknown frame transform kind");
        if (c < -1 || c >= nq) return fail("frame config index out of range");
        if (c >= 0) {
            if (k == TREPB_CONST_SE3) return fail("CONST_SE3 frame cannot be driven by a config");
            if (cfg_frame[c] != -1) return fail("a config drives more than one frame");
            cfg_frame[c] = f;
        }
    }
    int max_depth = 1;
    {
        std::vector<int> depth(nf, 0);
        for (int f = 1; f < nf; ++f) {
            depth[f] = depth[d->frame_parent[f]] + 1;
            if (depth[f] > max_depth) max_depth = depth[f];
        }
    }
    auto frame_ok = [&](int f) { return f >= 0 && f < nf; };
    // dep[f][c]
    std::vector<uint8_t> dep((size_t)nf * (nq > 0 ? nq : 1), 0), mass_below(nf, 0), need_world(nf, 0);
    for (int f = 1; f < nf; ++f) {
        const int p = d->frame_parent[f];
        for (int c = 0; c < nq; ++c) dep[(size_t)f * nq + c] = dep[(size_t)p * nq + c];
        if (d->frame_config[f] >= 0) dep[(size_t)f * nq + d->frame_config[f]] = 1;
    }
    for (int f = nf - 1; f >= 1; --f) {
        const double* m = d->frame_mass + 4 * f;
        if (m[0] != 0.0 || m[1] != 0.0 || m[2] != 0.0 || m[3] != 0.0) mass_below[f] = 1;
        if (mass_below[f]) mass_below[d->frame_parent[f]] = 1;
    }
    std::vector<uint8_t> vzero(nf, 1);
    for (int f = 1; f < nf; ++f) {
        const int p = d->frame_parent[f];
        vzero[f] = (p == 0) ? 1 : (vzero[p] && d->frame_config[p] < 0 ? 1 : 0);
    }
    auto mark_world = [&](int f) {
        while (f > 0 && !need_world[f]) { need_world[f] = 1; f = d->frame_parent[f]; }
    };
    double grav[3] = {0, 0, 0};
    int has_grav = 0, has_pairs = 0, has_pairs_mid = 0;
    for (int i = 0; i < np; ++i) {
        const int k = d->pot_kind[i];
        const int32_t* ii = d->pot_i + 4 * i;
        if (k == TREPB_POT_GRAVITY) {
            for (int j = 0; j < 3; ++j) grav[j] += d->pot_d[4 * i + j];
            has_grav = 1;
        } else if (k == TREPB_POT_LINEAR_SPRING) {
            if (!frame_ok(ii[0]) || !frame_ok(ii[1])) return fail("LinearSpring frame index out of range");
            mark_world(ii[0]); mark_world(ii[1]); has_pairs = 1; has_pairs_mid = 1;
        } else if (k == TREPB_POT_CONFIG_SPRING) {
            if (ii[0] < 0 || ii[0] >= nq) return fail("ConfigSpring config index out of range");
        } else if (k == TREPB_POT_NONLINEAR_CONFIG_SPRING) {
            if (ii[0] < 0 || ii[0] >= nq) return fail("NonlinearConfigSpring config index out of range");
            if (ii[2] < 2 || ii[1] < 0 || ii[1] + ii[2] + 6 * (ii[2] - 1) > d->n_dpool)
                return fail("NonlinearConfigSpring spline pool out of range");
            for (int j = 1; j < ii[2]; ++j)
                if (!(d->dpool[ii[1] + j] > d->dpool[ii[1] + j - 1])) return fail("spline x points must increase");
        } else return fail("unknown potential kind (Python-defined potentials have no device implementation)");
    }
    for (int i = 0; i < nfo; ++i) {
        const int k = d->force_kind[i];
        const int32_t* ii = d->force_i + 4 * i;
        if (k == TREPB_FORCE_DAMPING) {
            if (ii[0] < 0 || ii[1] != nd || ii[0] + nd > d->n_dpool) return fail("Damping coefficient pool out of range");
        } else if (k == TREPB_FORCE_CONFIG) {
            if (ii[0] < 0 || ii[0] >= nq || ii[1] < 0 || ii[1] >= nu) return fail("ConfigForce index out of range");
        } else if (k == TREPB_FORCE_LINEAR_DAMPER) {
            if (ii[1] != 2) return fail("LinearDamper: only two-frame paths are supported");
            if (ii[0] < 0 || ii[0] + 2 > d->n_ipool) return fail("LinearDamper path pool out of range");
            for (int j = 0; j < 2; ++j) {
                if (!frame_ok(d->ipool[ii[0] + j])) return fail("LinearDamper frame index out of range");
                mark_world(d->ipool[ii[0] + j]);
            }
            has_pairs = 1; has_pairs_mid = 1;
        } else if (k == TREPB_FORCE_BODY_WRENCH || k == TREPB_FORCE_HYBRID_WRENCH || k == TREPB_FORCE_SPATIAL_WRENCH) {
            if (!frame_ok(ii[0])) return fail("wrench frame index out of range");
            if (ii[1] < 0 || ii[1] + 6 > d->n_ipool) return fail("wrench input pool out of range");
            if (ii[2] < 0 || ii[2] + 6 > d->n_dpool) return fail("wrench constant pool out of range");
            for (int j = 0; j < 6; ++j)
                if (d->ipool[ii[1] + j] < -1 || d->ipool[ii[1] + j] >= nu) return fail("wrench input index out of range");
            mark_world(ii[0]); has_pairs = 1; has_pairs_mid = 1;
        } else return fail("unknown force kind (Python-defined forces have no device implementation)");
    }
    for (int i = 0; i < nc; ++i) {
        const int k = d->con_kind[i];
        const int32_t* ii = d->con_i + 4 * i;
        if (k != TREPB_CON_DISTANCE && k != TREPB_CON_POINT1D && k != TREPB_CON_PLANE)
            return fail("unknown constraint kind (Python-defined constraints have no device implementation)");
        if (!frame_ok(ii[0]) || !frame_ok(ii[1])) return fail("constraint frame index out of range");
        if (k == TREPB_CON_DISTANCE && (ii[2] < -1 || ii[2] >= nq)) return fail("Distance config out of range");
        if (k == TREPB_CON_POINT1D && (ii[2] < 0 || ii[2] > 2)) return fail("PointToPoint component out of range");
        mark_world(ii[0]); mark_world(ii[1]); has_pairs = 1;
    }

    // ---- pack
    PackedSys& P = *out;
    P.nf = nf; P.nq = nq;
    P.cfg_frame = cfg_frame; P.dep = dep; P.mass_below = mass_below; P.need_world = need_world; P.vzero = vzero;
    memset(&P.proto, 0, sizeof(P.proto));
    P.proto.nf = nf; P.proto.nd = nd; P.proto.nk = nk; P.proto.nu = nu; P.proto.nc = nc;
    P.proto.npot = np; P.proto.nforce = nfo; P.proto.max_depth = max_depth;
    for (int j = 0; j < 3; ++j) P.proto.grav[j] = grav[j];
    P.proto.has_gravity = has_grav; P.proto.has_pairs = has_pairs; P.proto.has_pairs_mid = has_pairs_mid;
    P.blob.clear();
    int k = 0;
    auto put = [&](const void* src, size_t bytes) {
        size_t o = (P.blob.size() + 15) & ~size_t(15);
        P.blob.resize(o + (bytes ? bytes : 16), 0);
        if (bytes && src) memcpy(P.blob.data() + o, src, bytes);
        P.off[k++] = o;
    };
    put(d->frame_parent, sizeof(int32_t) * nf);
    put(d->frame_kind, sizeof(int32_t) * nf);
    put(d->frame_config, sizeof(int32_t) * nf);
    put(cfg_frame.data(), sizeof(int32_t) * nq);
    put(d->pot_kind, sizeof(int32_t) * np);
    put(d->pot_i, sizeof(int32_t) * 4 * np);
    put(d->force_kind, sizeof(int32_t) * nfo);
    put(d->force_i, sizeof(int32_t) * 4 * nfo);
    put(d->con_kind, sizeof(int32_t) * nc);
    put(d->con_i, sizeof(int32_t) * 4 * nc);
    put(d->ipool, sizeof(int32_t) * d->n_ipool);
    put(d->frame_value, sizeof(double) * nf);
    put(d->frame_se3, sizeof(double) * 12 * nf);
    put(d->frame_mass, sizeof(double) * 4 * nf);
    put(d->pot_d, sizeof(double) * 4 * np);
    put(d->force_d, sizeof(double) * 4 * nfo);
    put(d->con_d, sizeof(double) * 4 * nc);
    put(d->dpool, sizeof(double) * d->n_dpool);
    put(dep.data(), dep.size());
    put(mass_below.data(), mass_below.size());
    put(need_world.data(), need_world.size());
    put(vzero.data(), vzero.size());
    return true;
}

}  // namespace trepb
