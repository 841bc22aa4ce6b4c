// Operation-counted floating-point work of the thread-per-instance math (trepb_math.cuh) on the host: the same
// templates the kernels instantiate, compiled with `double` replaced by a counting wrapper (SURVEY.md 8d:
// "algorithmic flops ... by compiling the restatement with an operation-counting double wrapper").
// add / sub / mul / div / sqrt / compare-free: 1 each, fma 2, sincos counted separately.  Test infrastructure
// (tests/test_opcount.py, bench.py's roofline note); never linked into the product.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <type_traits>
#include <vector>
#include "../include/trepb.h"
#include "../trep_b200/csrc/trepb_sys.h"    // the system tables stay plain doubles
#include "../trep_b200/csrc/trepb_pack.h"

typedef double real_t;
struct OpCounts { unsigned long long add, mul, div, sqrt_, fma_, sincos_, cmp; };
static OpCounts g_ops;
struct CountD {
    real_t v;
    CountD() = default;
    CountD(real_t x) : v(x) {}
    CountD(int x) : v(x) {}
    explicit operator bool() const { return v != 0.0; }
    explicit operator int() const { return (int)v; }   // pivot indices kept in the workspace
};
inline CountD operator-(CountD a) { return CountD(-a.v); }
inline CountD operator+(CountD a, CountD b) { ++g_ops.add; return CountD(a.v + b.v); }
inline CountD operator-(CountD a, CountD b) { ++g_ops.add; return CountD(a.v - b.v); }
inline CountD operator*(CountD a, CountD b) { ++g_ops.mul; return CountD(a.v * b.v); }
inline CountD operator/(CountD a, CountD b) { ++g_ops.div; return CountD(a.v / b.v); }
#define MIX(op) \
    inline CountD operator op(CountD a, real_t b) { return a op CountD(b); } \
    inline CountD operator op(real_t a, CountD b) { return CountD(a) op b; } \
    inline CountD operator op(CountD a, int b) { return a op CountD((real_t)b); } \
    inline CountD operator op(int a, CountD b) { return CountD((real_t)a) op b; }
MIX(+) MIX(-) MIX(*) MIX(/)
#undef MIX
inline CountD& operator+=(CountD& a, CountD b) { a = a + b; return a; }
inline CountD& operator-=(CountD& a, CountD b) { a = a - b; return a; }
inline CountD& operator*=(CountD& a, CountD b) { a = a * b; return a; }
inline CountD& operator/=(CountD& a, CountD b) { a = a / b; return a; }
inline CountD& operator+=(CountD& a, real_t b) { a = a + CountD(b); return a; }
inline CountD& operator-=(CountD& a, real_t b) { a = a - CountD(b); return a; }
inline CountD& operator*=(CountD& a, real_t b) { a = a * CountD(b); return a; }
#define CMP(op) \
    inline bool operator op(CountD a, CountD b) { ++g_ops.cmp; return a.v op b.v; } \
    inline bool operator op(CountD a, real_t b) { ++g_ops.cmp; return a.v op b; } \
    inline bool operator op(real_t a, CountD b) { ++g_ops.cmp; return a op b.v; }
CMP(<) CMP(>) CMP(<=) CMP(>=) CMP(==) CMP(!=)
#undef CMP
inline CountD fabs(CountD a) { return CountD(::fabs(a.v)); }
inline CountD sqrt(CountD a) { ++g_ops.sqrt_; return CountD(::sqrt(a.v)); }
inline CountD fma(CountD a, CountD b, CountD c) { ++g_ops.fma_; return CountD(::fma(a.v, b.v, c.v)); }
inline bool isnan(CountD a) { return std::isnan(a.v); }
inline CountD sin(CountD a) { ++g_ops.sincos_; return CountD(::sin(a.v)); }
inline CountD cos(CountD a) { ++g_ops.sincos_; return CountD(::cos(a.v)); }
inline void sincos(CountD x, CountD* s, CountD* c) { g_ops.sincos_ += 2; real_t a, b; ::sincos(x.v, &a, &b); *s = CountD(a); *c = CountD(b); }
inline CountD nextafter(CountD a, real_t b) { return CountD(::nextafter(a.v, b)); }

#define double CountD
#include "../trep_b200/csrc/trepb_ws.h"
#include "../trep_b200/csrc/trepb_math.cuh"
#undef double

using namespace trepb;

namespace {
struct Host {
    PackedSys P;
    RtSys sys;
    WsStridedT<CountD> ws;
    std::vector<CountD> slab;
    bool init(const trepb_sysdesc* d) {
        std::string err;
        if (!pack_system(d, &P, &err)) return false;
        sys = P.view(P.blob.data());
        int n = ws.layout(sys.nf, sys.nd, sys.nk, sys.nu, sys.nc);
        slab.assign(n, CountD(0.0));
        ws.base = slab.data();
        ws.stride = 1;
        return true;
    }
};
void report(unsigned long long* out) {
    out[0] = g_ops.add; out[1] = g_ops.mul; out[2] = g_ops.div; out[3] = g_ops.sqrt_; out[4] = g_ops.fma_;
    out[5] = g_ops.sincos_; out[6] = g_ops.cmp;
}
}  // namespace

extern "C" {

// nsteps DEL steps from (q1, p1); counts[7] = add, mul, div, sqrt, fma, sin + cos evaluations, comparisons summed over the steps;
// returns the summed Newton iterations or a negative status
int oc_step(const trepb_sysdesc* d, int nsteps, real_t t0, real_t dt, real_t tol, int maxit, const real_t* q1,
            const real_t* p1, const real_t* u1, const real_t* k2, const real_t* lam_guess, real_t* q2, real_t* p2,
            unsigned long long* counts) {
    Host h;
    if (!h.init(d)) return -100;
    RtSys& s = h.sys;
    WsStridedT<CountD>& ws = h.ws;
    const int nd = s.nd, nk = s.nk, nq = nd + nk, nu = s.nu, nc = s.nc;
    for (int i = 0; i < nq; ++i) { ws.q1(i) = q1[i]; ws.q2(i) = q1[i]; }
    for (int i = 0; i < nd; ++i) ws.p1(i) = p1[i];
    for (int c = 0; c < nc; ++c) ws.lam(c) = lam_guess ? lam_guess[c] : 0.0;
    int total = 0;
    real_t t1 = t0;
    const CountD thr = sqrt_threshold(tol);
    memset(&g_ops, 0, sizeof g_ops);
    for (int st = 0; st < nsteps; ++st) {
        if (st > 0) {
            for (int i = 0; i < nq; ++i) ws.q1(i) = ws.q2(i);
            for (int i = 0; i < nd; ++i) ws.p1(i) = ws.p2(i);
        }
        for (int i = 0; i < nu; ++i) ws.u1(i) = u1 ? u1[st * nu + i] : 0.0;
        for (int i = 0; i < nk; ++i) ws.q2(nd + i) = k2[st * nk + i];
        const real_t t2 = t1 + dt;
        int it = solve_del(s, ws, t1, t2, tol, maxit, thr);
        if (it < 0) return it;
        total += it;
        t1 = t2;
    }
    report(counts);
    for (int i = 0; i < nq; ++i) q2[i] = ws.q2(i).v;
    for (int i = 0; i < nd; ++i) p2[i] = ws.p2(i).v;
    return total;
}

// one linearization (solve + deriv1 -> A, B): counts as above; returns the Newton iterations
int oc_linearize(const trepb_sysdesc* d, real_t t1, real_t t2, real_t tol, int maxit, const real_t* q1, const real_t* p1,
                 const real_t* u1, const real_t* k2, const real_t* q2_guess, const real_t* lam_guess, real_t* A, real_t* B,
                 unsigned long long* counts) {
    Host h;
    if (!h.init(d)) return -100;
    RtSys& s = h.sys;
    WsStridedT<CountD>& ws = h.ws;
    const int nd = s.nd, nk = s.nk, nq = nd + nk, nu = s.nu, nc = s.nc;
    for (int i = 0; i < nq; ++i) { ws.q1(i) = q1[i]; ws.q2(i) = q1[i]; }
    for (int i = 0; i < nd; ++i) { ws.p1(i) = p1[i]; if (q2_guess) ws.q2(i) = q2_guess[i]; }
    for (int i = 0; i < nk; ++i) ws.q2(nd + i) = k2[i];
    for (int i = 0; i < nu; ++i) ws.u1(i) = u1[i];
    for (int c = 0; c < nc; ++c) ws.lam(c) = lam_guess ? lam_guess[c] : 0.0;
    const CountD thr = sqrt_threshold(tol);
    memset(&g_ops, 0, sizeof g_ops);
    int it = solve_del(s, ws, t1, t2, tol, maxit, thr);
    if (it < 0) return it;
    const int nX = 2 * nq, nU = nu + nk;
    std::vector<CountD> Ac((size_t)nX * nX + 1), Bc((size_t)nX * (nU > 0 ? nU : 1) + 1);
    Deriv1Out o;
    memset(&o, 0, sizeof o);
    o.A = Ac.data(); o.B = nU > 0 ? Bc.data() : nullptr; o.es = 1;
    int rc = deriv1(s, ws, t1, t2, o, true);
    if (rc) return rc;
    report(counts);
    for (int i = 0; i < nX * nX; ++i) A[i] = Ac[i].v;
    for (int i = 0; i < nX * nU; ++i) B[i] = Bc[i].v;
    return it;
}

}  // extern "C"
