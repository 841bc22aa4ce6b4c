// Hyper-dual numbers: exact second-order forward differentiation of the DEL residual.
//
//   x = v + a e1 + sum_k ( b_k e2k + ab_k e1 e2k ) ,   e1^2 = e2k^2 = e2k e2l = 0      (k < N)
//
// Evaluating a function at  z + e1 xi_s + sum_k e2k xi_tk  yields in the e1 e2k components the second
// directional derivatives  D^2 f(z)[xi_s, xi_tk]  for N directions t_k at once (plus f_z . z_st if
// the input carries an ab part), exact to rounding.  N = 1 is the classical hyper-dual number; with
// N directions the value and e1 parts - and all the table look-ups, trigonometry and index work of the
// evaluation - are shared by N parameter pairs: (2 + 2N) components per number instead of 4N.
// The second-derivative kernel (trepb_d2.cuh) instantiates the first-order residual path of
// trepb_math.cuh on this type; that replaces the reference's hand-expanded third-order tables
// (trep/_trep/midpointvi.c:1122-1532 calc_deriv2_cache_*, system.c:204-268, 336-393, 514-557
// L_dqdqdq / L_ddqdqdq / L_ddqddqdq, constraint h_dqdqdq) by one generic rule.
#pragma once
#include <math.h>
#include "trepb_sys.h"

namespace trepb {

#if defined(__CUDACC__)
#define TREPB_HDU _Pragma("unroll")
#else
#define TREPB_HDU
#endif

template <int N>
struct HDn {
    double v, a, b[N], ab[N];
    HDn() = default;
    TREPB_HD HDn(double x) : v(x), a(0.0) { TREPB_HDU for (int k = 0; k < N; ++k) { b[k] = 0.0; ab[k] = 0.0; } }
};
using HD = HDn<1>;        // one direction pair per evaluation (register-resident specialised systems)
// table-driven systems.  Measured on the marionette (B200, thread per work item): N = 7 (16 doubles =
// one 128-byte line per workspace element) runs at 1.9e3 evaluations/s against 3.6e3 for N = 1 - the
// 6-vector / 3x3 temporaries of the tree passes no longer fit the register file (6.4 KB of local
// memory per thread, 20 KB of spill code), which costs more than the shared work saves; N = 2 gives
// 3.8e3 (+3 %, 1.4 KB of local memory).
constexpr int kD2Dirs = 1;
using HDG = HDn<kD2Dirs>;

template <int N> TREPB_HD HDn<N> operator-(const HDn<N>& x) {
    HDn<N> r; r.v = -x.v; r.a = -x.a;
    TREPB_HDU for (int k = 0; k < N; ++k) { r.b[k] = -x.b[k]; r.ab[k] = -x.ab[k]; }
    return r;
}
template <int N> TREPB_HD HDn<N> operator+(const HDn<N>& x, const HDn<N>& y) {
    HDn<N> r; r.v = x.v + y.v; r.a = x.a + y.a;
    TREPB_HDU for (int k = 0; k < N; ++k) { r.b[k] = x.b[k] + y.b[k]; r.ab[k] = x.ab[k] + y.ab[k]; }
    return r;
}
template <int N> TREPB_HD HDn<N> operator-(const HDn<N>& x, const HDn<N>& y) {
    HDn<N> r; r.v = x.v - y.v; r.a = x.a - y.a;
    TREPB_HDU for (int k = 0; k < N; ++k) { r.b[k] = x.b[k] - y.b[k]; r.ab[k] = x.ab[k] - y.ab[k]; }
    return r;
}
template <int N> TREPB_HD HDn<N> operator*(const HDn<N>& x, const HDn<N>& y) {
    HDn<N> r; r.v = x.v * y.v; r.a = x.a * y.v + x.v * y.a;
    TREPB_HDU
    for (int k = 0; k < N; ++k) {
        r.b[k] = x.b[k] * y.v + x.v * y.b[k];
        r.ab[k] = x.ab[k] * y.v + x.a * y.b[k] + x.b[k] * y.a + x.v * y.ab[k];
    }
    return r;
}
template <int N> TREPB_HD HDn<N> operator+(const HDn<N>& x, double y) { HDn<N> r = x; r.v += y; return r; }
template <int N> TREPB_HD HDn<N> operator+(double y, const HDn<N>& x) { HDn<N> r = x; r.v += y; return r; }
template <int N> TREPB_HD HDn<N> operator-(const HDn<N>& x, double y) { HDn<N> r = x; r.v -= y; return r; }
template <int N> TREPB_HD HDn<N> operator-(double y, const HDn<N>& x) { HDn<N> r = -x; r.v += y; return r; }
template <int N> TREPB_HD HDn<N> operator*(const HDn<N>& x, double y) {
    HDn<N> r; r.v = x.v * y; r.a = x.a * y;
    TREPB_HDU for (int k = 0; k < N; ++k) { r.b[k] = x.b[k] * y; r.ab[k] = x.ab[k] * y; }
    return r;
}
template <int N> TREPB_HD HDn<N> operator*(double y, const HDn<N>& x) { return x * y; }
// g(x) with g' = d1, g'' = d2 at x.v
template <int N> TREPB_HD HDn<N> hd_chain(const HDn<N>& x, double g, double d1, double d2) {
    HDn<N> r; r.v = g; r.a = d1 * x.a;
    TREPB_HDU for (int k = 0; k < N; ++k) { r.b[k] = d1 * x.b[k]; r.ab[k] = d1 * x.ab[k] + d2 * x.a * x.b[k]; }
    return r;
}
template <int N> TREPB_HD HDn<N> hd_inv(const HDn<N>& y) {
    const double r = 1.0 / y.v;
    return hd_chain(y, r, -r * r, 2.0 * r * r * r);
}
template <int N> TREPB_HD HDn<N> operator/(const HDn<N>& x, const HDn<N>& y) { return x * hd_inv(y); }
template <int N> TREPB_HD HDn<N> operator/(double x, const HDn<N>& y) { return x * hd_inv(y); }
template <int N> TREPB_HD HDn<N> operator/(const HDn<N>& x, double y) { return x * (1.0 / y); }
template <int N> TREPB_HD HDn<N>& operator+=(HDn<N>& x, const HDn<N>& y) { x = x + y; return x; }
template <int N> TREPB_HD HDn<N>& operator-=(HDn<N>& x, const HDn<N>& y) { x = x - y; return x; }
template <int N> TREPB_HD HDn<N>& operator*=(HDn<N>& x, const HDn<N>& y) { x = x * y; return x; }
template <int N> TREPB_HD HDn<N>& operator+=(HDn<N>& x, double y) { x.v += y; return x; }
template <int N> TREPB_HD HDn<N>& operator-=(HDn<N>& x, double y) { x.v -= y; return x; }
template <int N> TREPB_HD HDn<N>& operator*=(HDn<N>& x, double y) { x = x * y; return x; }

TREPB_HD void sincos_(double x, double* s, double* c);   // trepb_math.cuh
template <int N> TREPB_HD void sincos_(const HDn<N>& x, HDn<N>* s, HDn<N>* c) {
    double sn, cs;
    sincos_(x.v, &sn, &cs);
    *s = hd_chain(x, sn, cs, -sn);
    *c = hd_chain(x, cs, -sn, -cs);
}
template <int N> TREPB_HD HDn<N> sqrt_(const HDn<N>& x) {
    const double r = sqrt(x.v);
    return hd_chain(x, r, 0.5 / r, -0.25 / (r * x.v));
}
template <int N> TREPB_HD bool isnan_(const HDn<N>& x) { return isnan(x.v); }

}  // namespace trepb
