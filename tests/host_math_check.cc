// TEST INFRASTRUCTURE (not shipped, not a fallback): compiles the shared host/device math of
// trep_b200/csrc/trepb_math.cuh with g++ so that the algebra can be checked against the golden
// vectors on a machine without a GPU.  The product path (libtrepb.so) never links this file.
#include <stdlib.h>
#include <string>
#include <vector>
#include "../include/trepb.h"
#include "../trep_b200/csrc/trepb_kernels.cuh"
#include "../trep_b200/csrc/trepb_pack.h"
#include "../trep_b200/csrc/trepb_coop_math.cuh"
#include "../trep_b200/csrc/trepb_d2jac.cuh"

using namespace trepb;

namespace {
struct Host {
    PackedSys P;
    RtSys sys;
    WsStrided ws;
    std::vector<double> slab;
    bool init(const trepb_sysdesc* d) {
        std::string err;
        if (!pack_system(d, &P, &err)) return false;
        sys = P.view(P.blob.data());
        int n = ws.layout(sys.nf, sys.nd, sys.nk, sys.nu, sys.nc);
        slab.assign(n, 0.0);
        ws.base = slab.data();
        ws.stride = 1;
        return true;
    }
};
}  // namespace

extern "C" {

int th_step(const trepb_sysdesc* d, int nsteps, double t0, double dt, double tol, int maxit,
            const double* q1, const double* p1, const double* u1, const double* k2,
            const double* q2_guess, const double* lam_guess, double* q2, double* p2, double* lam,
            int* iters) {
    Host h;
    if (!h.init(d)) return -100;
    RtSys& s = h.sys;
    WsStrided& ws = h.ws;
    const int nd = s.nd, nk = s.nk, nq = nd + nk, nu = s.nu, nc = s.nc;
    for (int i = 0; i < nq; ++i) { ws.q1(i) = q1[i]; ws.q2(i) = q1[i]; }
    for (int i = 0; i < nd; ++i) { ws.p1(i) = p1[i]; if (q2_guess) ws.q2(i) = q2_guess[i]; }
    for (int c = 0; c < nc; ++c) ws.lam(c) = lam_guess ? lam_guess[c] : 0.0;
    int total = 0;
    double t1 = t0;
    for (int st = 0; st < nsteps; ++st) {
        if (st > 0) {
            for (int i = 0; i < nq; ++i) ws.q1(i) = ws.q2(i);
            for (int i = 0; i < nd; ++i) ws.p1(i) = ws.p2(i);
        }
        for (int i = 0; i < nu; ++i) ws.u1(i) = u1 ? u1[st * nu + i] : 0.0;
        for (int i = 0; i < nk; ++i) ws.q2(nd + i) = k2[st * nk + i];
        const double t2 = t1 + dt;
        int it = solve_del(s, ws, t1, t2, tol, maxit, sqrt_threshold(tol));
        if (it < 0) return it;
        total += it;
        t1 = t2;
    }
    for (int i = 0; i < nq; ++i) q2[i] = ws.q2(i);
    for (int i = 0; i < nd; ++i) p2[i] = ws.p2(i);
    for (int c = 0; c < nc; ++c) lam[c] = ws.lam(c);
    *iters = total;
    return 0;
}

// doubles of the thread-per-instance workspace (what trepb_system_create compares with its cooperative threshold)
int th_ws_doubles(const trepb_sysdesc* d) {
    Host h;
    if (!h.init(d)) return -100;
    return (int)h.slab.size();
}

int th_calc_p2(const trepb_sysdesc* d, double dt, const double* q0, const double* q1, double* p) {
    Host h;
    if (!h.init(d)) return -100;
    for (int i = 0; i < h.sys.nd + h.sys.nk; ++i) { h.ws.q1(i) = q0[i]; h.ws.q2(i) = q1[i]; }
    for (int i = 0; i < h.sys.nu; ++i) h.ws.u1(i) = 0.0;
    calc_p2(h.sys, h.ws, 0.0, dt);
    for (int i = 0; i < h.sys.nd; ++i) p[i] = h.ws.p2(i);
    return 0;
}

// raw[12] = q2_dq1 q2_dp1 q2_du1 q2_dk2 p2_dq1 ... l1_dk2 ; any may be null
int th_linearize(const trepb_sysdesc* d, double t1, double t2, double tol, int maxit,
                 const double* q1, const double* p1, const double* u1, const double* k2,
                 const double* q2_guess, const double* lam_guess, double* q2, double* p2,
                 double* lam, int* iters, double* A, double* B, double** raw) {
    Host h;
    if (!h.init(d)) return -100;
    RtSys& s = h.sys;
    WsStrided& ws = h.ws;
    const int nd = s.nd, nk = s.nk, nq = nd + nk, nu = s.nu, nc = s.nc;
    for (int i = 0; i < nq; ++i) { ws.q1(i) = q1[i]; ws.q2(i) = q1[i]; }
    for (int i = 0; i < nd; ++i) { ws.p1(i) = p1[i]; if (q2_guess) ws.q2(i) = q2_guess[i]; }
    for (int i = 0; i < nk; ++i) ws.q2(nd + i) = k2[i];
    for (int i = 0; i < nu; ++i) ws.u1(i) = u1[i];
    for (int c = 0; c < nc; ++c) ws.lam(c) = lam_guess ? lam_guess[c] : 0.0;
    int it = solve_del(s, ws, t1, t2, tol, maxit, sqrt_threshold(tol));
    if (it < 0) return it;
    *iters = it;
    for (int i = 0; i < nq; ++i) q2[i] = ws.q2(i);
    for (int i = 0; i < nd; ++i) p2[i] = ws.p2(i);
    for (int c = 0; c < nc; ++c) lam[c] = ws.lam(c);
    Deriv1Out o;
    o.q2_dq1 = raw[0]; o.q2_dp1 = raw[1]; o.q2_du1 = raw[2]; o.q2_dk2 = raw[3];
    o.p2_dq1 = raw[4]; o.p2_dp1 = raw[5]; o.p2_du1 = raw[6]; o.p2_dk2 = raw[7];
    o.l1_dq1 = raw[8]; o.l1_dp1 = raw[9]; o.l1_du1 = raw[10]; o.l1_dk2 = raw[11];
    o.A = A; o.B = B; o.es = 1;
    return deriv1(s, ws, t1, t2, o, true);  // same call sequence as lin_kernel
}

// Second derivatives through the same per-pair function the d2 kernel runs (trepb_d2.cuh), on the
// host: solve + deriv1 (+ aux export), then every pair s <= t.  d2[30] as in trepb_d2_args.
int th_deriv2(const trepb_sysdesc* d, double t1, double t2, double tol, int maxit,
              const double* q1, const double* p1, const double* u1, const double* k2,
              const double* q2_guess, const double* lam_guess, double** d2) {
    Host h;
    if (!h.init(d)) return -100;
    RtSys& s = h.sys;
    WsStrided& ws = h.ws;
    const int nd = s.nd, nk = s.nk, nq = nd + nk, nu = s.nu, nc = s.nc;
    for (int i = 0; i < nq; ++i) { ws.q1(i) = q1[i]; ws.q2(i) = q1[i]; }
    for (int i = 0; i < nd; ++i) { ws.p1(i) = p1[i]; if (q2_guess) ws.q2(i) = q2_guess[i]; }
    for (int i = 0; i < nk; ++i) ws.q2(nd + i) = k2[i];
    for (int i = 0; i < nu; ++i) ws.u1(i) = u1[i];
    for (int c = 0; c < nc; ++c) ws.lam(c) = lam_guess ? lam_guess[c] : 0.0;
    int it = solve_del(s, ws, t1, t2, tol, maxit, sqrt_threshold(tol));
    if (it < 0) return it;
    std::vector<double> q2(nq), lam(nc + 1), uu(nu + 1);
    for (int i = 0; i < nq; ++i) q2[i] = ws.q2(i);
    for (int c = 0; c < nc; ++c) lam[c] = ws.lam(c);
    for (int i = 0; i < nu; ++i) uu[i] = u1[i];
    const int cnt[4] = {nq, nd, nu, nk};
    std::vector<double> qd[4], ld[4], pd[4];
    for (int i = 0; i < 4; ++i) { qd[i].assign(cnt[i] * nd + 1, 0.0); pd[i].assign(cnt[i] * nd + 1, 0.0); ld[i].assign(cnt[i] * nc + 1, 0.0); }
    Deriv1Out o;
    o.q2_dq1 = qd[0].data(); o.q2_dp1 = qd[1].data(); o.q2_du1 = qd[2].data(); o.q2_dk2 = qd[3].data();
    o.p2_dq1 = pd[0].data(); o.p2_dp1 = pd[1].data(); o.p2_du1 = pd[2].data(); o.p2_dk2 = pd[3].data();
    o.l1_dq1 = ld[0].data(); o.l1_dp1 = ld[1].data(); o.l1_du1 = ld[2].data(); o.l1_dk2 = ld[3].data();
    o.A = nullptr; o.B = nullptr; o.es = 1;
    int rc = deriv1(s, ws, t1, t2, o, true);
    if (rc) return rc;
    AuxLayout al;
    al.set(nd, nc);
    std::vector<double> aux(al.size + 1);
    export_aux(s, ws, aux.data());
    // hyper-dual workspace
    // several directions per residual evaluation (the kernels use HDG, trepb_hd.h; the host check
    // exercises the N-direction arithmetic with N = 7)
    WsStridedT<HDn<7>> wh;
    const int n = wh.layout(s.nf, nd, nk, nu, nc, 1);
    std::vector<HDn<7>> slab(n + 1);
    wh.base = slab.data();
    wh.stride = 1;
    D2Params p;
    p.batch = 1; p.nx = nq + nd + nu + nk; p.npairs = p.nx * (p.nx + 1) / 2;
    p.t1s = t1; p.dts = t2 - t1; p.t1 = &t1; p.t2 = &t2;
    p.q1 = q1; p.u1 = uu.data(); p.q2 = q2.data(); p.lam = lam.data();
    for (int i = 0; i < 4; ++i) { p.q2_d[i] = qd[i].data(); p.l1_d[i] = ld[i].data(); }
    p.aux = aux.data(); p.auxl = al; p.status = nullptr;
    p.z = nullptr; p.zxx = p.zxu = p.zuu = nullptr;
    for (int w = 0; w < 3; ++w) for (int k = 0; k < 10; ++k) p.out[w][k] = d2[10 * w + k];
    for (int a = 0; a < p.nx; ++a)
        for (int b0 = a; b0 < p.nx; b0 += 7) deriv2_block(s, wh, p, 0, a, b0, p.nx - b0 < 7 ? p.nx - b0 : 7);
    return 0;
}

// Second derivatives through the O(nx) formulation (trepb_d2jac.cuh): per parameter s one dual
// evaluation of the Jacobian tables (pass A), then per pair the same contraction + solve function
// the d2solve kernel runs (pass B).  method-independent outputs: d2[30] as in trepb_d2_args.
namespace {
struct VecSink {
    std::vector<double>* v;
    void push(double x) { v->push_back(x); }
};
}  // namespace
int th_deriv2_jac(const trepb_sysdesc* d, double t1, double t2, double tol, int maxit,
                  const double* q1, const double* p1, const double* u1, const double* k2,
                  const double* q2_guess, const double* lam_guess, double** d2) {
    Host h;
    if (!h.init(d)) return -100;
    RtSys& s = h.sys;
    WsStrided& ws = h.ws;
    const int nd = s.nd, nk = s.nk, nq = nd + nk, nu = s.nu, nc = s.nc;
    for (int i = 0; i < nq; ++i) { ws.q1(i) = q1[i]; ws.q2(i) = q1[i]; }
    for (int i = 0; i < nd; ++i) { ws.p1(i) = p1[i]; if (q2_guess) ws.q2(i) = q2_guess[i]; }
    for (int i = 0; i < nk; ++i) ws.q2(nd + i) = k2[i];
    for (int i = 0; i < nu; ++i) ws.u1(i) = u1[i];
    for (int c = 0; c < nc; ++c) ws.lam(c) = lam_guess ? lam_guess[c] : 0.0;
    int it = solve_del(s, ws, t1, t2, tol, maxit, sqrt_threshold(tol));
    if (it < 0) return it;
    std::vector<double> q2(nq), lam(nc + 1), uu(nu + 1);
    for (int i = 0; i < nq; ++i) q2[i] = ws.q2(i);
    for (int c = 0; c < nc; ++c) lam[c] = ws.lam(c);
    for (int i = 0; i < nu; ++i) uu[i] = u1[i];
    const int cnt[4] = {nq, nd, nu, nk};
    std::vector<double> qd[4], ld[4], pd[4];
    for (int i = 0; i < 4; ++i) { qd[i].assign(cnt[i] * nd + 1, 0.0); pd[i].assign(cnt[i] * nd + 1, 0.0); ld[i].assign(cnt[i] * nc + 1, 0.0); }
    Deriv1Out o;
    o.q2_dq1 = qd[0].data(); o.q2_dp1 = qd[1].data(); o.q2_du1 = qd[2].data(); o.q2_dk2 = qd[3].data();
    o.p2_dq1 = pd[0].data(); o.p2_dp1 = pd[1].data(); o.p2_du1 = pd[2].data(); o.p2_dk2 = pd[3].data();
    o.l1_dq1 = ld[0].data(); o.l1_dp1 = ld[1].data(); o.l1_du1 = ld[2].data(); o.l1_dk2 = ld[3].data();
    o.A = nullptr; o.B = nullptr; o.es = 1;
    int rc = deriv1(s, ws, t1, t2, o, true);
    if (rc) return rc;
    AuxLayout al;
    al.set(nd, nc);
    std::vector<double> aux(al.size + 1);
    export_aux(s, ws, aux.data());
    WsStridedT<Dual> wd;
    const int n = wd.layout(s.nf, nd, nk, nu, nc, 2);
    std::vector<Dual> slab(n + 1);
    wd.base = slab.data();
    wd.stride = 1;
    D2Params p;
    p.batch = 1; p.nx = nq + nd + nu + nk; p.npairs = p.nx * (p.nx + 1) / 2;
    p.t1s = t1; p.dts = t2 - t1; p.t1 = &t1; p.t2 = &t2;
    p.q1 = q1; p.u1 = uu.data(); p.q2 = q2.data(); p.lam = lam.data();
    for (int i = 0; i < 4; ++i) { p.q2_d[i] = qd[i].data(); p.l1_d[i] = ld[i].data(); }
    p.aux = aux.data(); p.auxl = al; p.status = nullptr;
    p.z = nullptr; p.zxx = p.zxu = p.zuu = nullptr;
    for (int w = 0; w < 3; ++w) for (int k = 0; k < 10; ++k) p.out[w][k] = d2[10 * w + k];
    JacLayout jl;
    jl.set(nd, nk, nu, nc);
    std::vector<double> G, vec(3 * nd + 3 * nc + 1);
    std::vector<uint8_t> nzb(NzMaps::bytes(nd, nk) + 8);
    NzMaps nz;
    nz.place(nzb.data(), nd, nk);
    d2jac_build_nz(s, nz, 0, 1);
    // poison the workspace: entries outside the structure maps must never be read
    for (auto& e : slab) e = Dual(1e300, -1e300);
    for (int a = 0; a < p.nx; ++a) {
        d2jac_eval(s, wd, nz, p, 0, a);
        G.clear();
        VecSink sink{&G};
        d2jac_emit(s, wd, nz, t2 - t1, sink);
        if ((int)G.size() != jl.size) return -300;
        D2Pair P;
        std::vector<double> Jd(nd * nd + 1), Hd(nd * nd + 1);
        for (int i = 0; i < nd * nd; ++i) { Jd[i] = G[jl.o_q + 4 * i + 1]; Hd[i] = G[jl.o_q + 4 * i + 3]; }
        P.Jd = Jd.data(); P.Hd = Hd.data(); P.JL = G.data() + jl.o_jl; P.JC = G.data() + jl.o_jc;
        P.Grow = G.data(); P.aux = aux.data(); P.z = nullptr; P.jl = jl; P.al = al;
        double* y = vec.data();
        double *lt = y + nd, *c = lt + nc, *x = c + nd, *hh = x + nd, *hx = hh + nc;
        for (int b = a; b < p.nx; ++b) d2_pair(P, p, 0, nd, nk, nu, nc, a, b, y, lt, c, x, hh, hx, 1);
    }
    return 0;
}

// div_dt against the division it replaces, on n pseudo-random numerators of mixed magnitude
long th_div_dt_mismatches(double dt, long n) {
    unsigned long long s = 88172645463325252ull;
    const Dt d(dt);
    long bad = 0;
    for (long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double x = ((double)(s >> 11) / 9007199254740992.0 - 0.5) * ((i & 7) == 0 ? 1e-3 : ((i & 7) == 1 ? 100.0 : 1.0));
        if (div_dt(x, d) != x / dt) ++bad;
    }
    return bad;
}
double th_sqrt_threshold(double tol) { return sqrt_threshold(tol); }

// ---- team-cooperative path (trepb_coop_math.cuh) with a one-lane host team
// returns -200 when the cooperative path does not apply to this system, -201 when the
// compile-time-size flavour was asked for a system of another shape
}  // extern "C"
namespace {
template <class D>
int coop_linearize(const CoopSys& S, int nsteps, double t1, double dt, double tol, int maxit,
                   const double* q1, const double* p1, const double* u1, const double* k2,
                   const double* q2_guess, const double* lam_guess, double* q2, double* p2,
                   double* lam, int* iters, double* A, double* B, double** raw, double* aux) {
    CoopLayout L;
    L.set(S, D::kStatic, D::kExtS && !D::kExt, D::kExt || D::kExtS);   // ExtSolveDims: the solve-only layout (callers pass raw = null)
    std::vector<double> slab(L.total + 8, 0.0);
    // poisoned external slab: every entry read must have been written for this instance
    std::vector<double> xslab(L.xtotal + 8, 1e300);
    Coop<HostTeam, D> c(S, L, slab.data(), HostTeam(), xslab.data());
    double* w = slab.data();
    const int nd = S.nd, nk = S.nk, nq = S.nq, nu = S.nu, nc = S.nc;
    for (int i = 0; i < nq; ++i) { w[L.q1 + i] = q1[i]; w[L.q2 + i] = q1[i]; }
    for (int i = 0; i < nd; ++i) { w[L.p1 + i] = p1[i]; if (q2_guess) w[L.q2 + i] = q2_guess[i]; }
    for (int i = 0; i < nc; ++i) w[L.lam + i] = lam_guess ? lam_guess[i] : 0.0;
    int total = 0;
    double ta = t1;
    for (int st = 0; st < nsteps; ++st) {
        if (st > 0) {
            for (int i = 0; i < nq; ++i) w[L.q1 + i] = w[L.q2 + i];
            for (int i = 0; i < nd; ++i) w[L.p1 + i] = w[L.p2 + i];
        }
        for (int i = 0; i < nu; ++i) w[L.u1 + i] = u1 ? u1[st * nu + i] : 0.0;
        for (int i = 0; i < nk; ++i) w[L.q2 + nd + i] = k2[st * nk + i];
        const int it = c.solve(ta, ta + dt, tol, maxit);
        if (it < 0) return it;
        total += it;
        ta += dt;
    }
    *iters = total;
    for (int i = 0; i < nq; ++i) q2[i] = w[L.q2 + i];
    for (int i = 0; i < nd; ++i) p2[i] = w[L.p2 + i];
    for (int i = 0; i < nc; ++i) lam[i] = w[L.lam + i];
    if (!raw) return 0;
    Deriv1Out o;
    o.q2_dq1 = raw[0]; o.q2_dp1 = raw[1]; o.q2_du1 = raw[2]; o.q2_dk2 = raw[3];
    o.p2_dq1 = raw[4]; o.p2_dp1 = raw[5]; o.p2_du1 = raw[6]; o.p2_dk2 = raw[7];
    o.l1_dq1 = raw[8]; o.l1_dp1 = raw[9]; o.l1_du1 = raw[10]; o.l1_dk2 = raw[11];
    o.A = A; o.B = B; o.es = 1;
    AuxLayout al;
    al.set(nd, nc);
    const int auxo[7] = {al.o_m2, al.o_m2p, al.o_pj, al.o_pjp, al.o_dh1, al.o_dh2, al.o_t22};
    return c.deriv1(ta - dt, ta, o, aux, auxo);
}
// the shapes the build specialises ahead of time (trep_b200/build.py COOP_AOT_SYSTEMS)
using PuppetDims = CtDims<22, 18, 0, 6, 34, 12, 157, 10, 20, 38>;
}  // namespace
extern "C" {
int th_coop_linearize(const trepb_sysdesc* d, int static_dims, int nsteps, double t1, double dt, double tol, int maxit,
                      const double* q1, const double* p1, const double* u1, const double* k2,
                      const double* q2_guess, const double* lam_guess, double* q2, double* p2,
                      double* lam, int* iters, double* A, double* B, double** raw, double* aux) {
    std::string err;
    PackedSys chk;
    if (!pack_system(d, &chk, &err)) return -100;
    CoopPack P = coop_pack(d);
    if (!P.ok) return -200;
    CoopSys S = P.view(P.blob.data());
    if (static_dims) {
        if (!PuppetDims::matches(S)) return -201;
        if (static_dims == 3) {  // solve-only external-slab layout (step / project / p2 kernels): no derivatives
            if (raw) return -202;
            return coop_linearize<ExtSolveDims<PuppetDims>>(S, nsteps, t1, dt, tol, maxit, q1, p1, u1, k2, q2_guess, lam_guess,
                                                            q2, p2, lam, iters, A, B, raw, aux);
        }
        if (static_dims == 2)   // external-slab layout of the first-derivative workspace
            return coop_linearize<ExtDims<PuppetDims>>(S, nsteps, t1, dt, tol, maxit, q1, p1, u1, k2, q2_guess, lam_guess,
                                                       q2, p2, lam, iters, A, B, raw, aux);
        return coop_linearize<PuppetDims>(S, nsteps, t1, dt, tol, maxit, q1, p1, u1, k2, q2_guess, lam_guess, q2, p2,
                                          lam, iters, A, B, raw, aux);
    }
    return coop_linearize<RtDims>(S, nsteps, t1, dt, tol, maxit, q1, p1, u1, k2, q2_guess, lam_guess, q2, p2, lam,
                                  iters, A, B, raw, aux);
}

int th_coop_calc_p2(const trepb_sysdesc* d, double dt, const double* q0, const double* q1, double* p) {
    CoopPack P = coop_pack(d);
    if (!P.ok) return -200;
    CoopSys S = P.view(P.blob.data());
    CoopLayout L;
    L.set(S);
    std::vector<double> slab(L.total + 8, 0.0);
    Coop<HostTeam> c(S, L, slab.data(), HostTeam());
    for (int i = 0; i < S.nq; ++i) { slab[L.q1 + i] = q0[i]; slab[L.q2 + i] = q1[i]; }
    c.calc_p2(dt);
    for (int i = 0; i < S.nd; ++i) p[i] = slab[L.p2 + i];
    return 0;
}

int th_coop_info(const trepb_sysdesc* d, int* out) {
    CoopPack P = coop_pack(d);
    if (!P.ok) return -200;
    CoopSys S = P.view(P.blob.data());
    CoopLayout L;
    L.set(S);
    out[0] = S.nl; out[1] = S.nlevels; out[2] = S.npairs; out[3] = S.np; out[4] = L.total; out[5] = (int)P.blob.size();
    CoopLayout Ls;
    Ls.set(S, true);
    out[6] = Ls.total;
    return 0;
}
}
