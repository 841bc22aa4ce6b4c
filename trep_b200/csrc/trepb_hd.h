// Hyper-dual numbers: exact second-order forward differentiation of the DEL residual.
//
//   x = v + a e1 + b e2 + ab e1 e2 ,   e1^2 = e2^2 = 0
//
// Evaluating a function at  z + e1 xi_s + e2 xi_t  yields in the e1e2 component the second
// directional derivative  D^2 f(z)[xi_s, xi_t]  (plus f_z . z_st if the input carries an ab part),
// exact to rounding.  The second-derivative kernel (trepb_d2.cuh) instantiates the first-order
// residual path of trepb_math.cuh on this type; that replaces the reference's hand-expanded
// third-order tables (trep/_trep/midpointvi.c:1122-1532 calc_deriv2_cache_*, system.c:204-268,
// 336-393, 514-557 L_dqdqdq / L_ddqdqdq / L_ddqddqdq, constraint h_dqdqdq) by one generic rule.
#pragma once
#include <math.h>
#include "trepb_sys.h"

namespace trepb {

struct HD {
    double v, a, b, ab;
    HD() = default;
    TREPB_HD HD(double x) : v(x), a(0.0), b(0.0), ab(0.0) {}
    TREPB_HD HD(double v_, double a_, double b_, double ab_) : v(v_), a(a_), b(b_), ab(ab_) {}
};

TREPB_HD HD operator-(const HD& x) { return HD(-x.v, -x.a, -x.b, -x.ab); }
TREPB_HD HD operator+(const HD& x, const HD& y) { return HD(x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab); }
TREPB_HD HD operator-(const HD& x, const HD& y) { return HD(x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab); }
TREPB_HD HD operator*(const HD& x, const HD& y) {
    return HD(x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b,
              x.ab * y.v + x.a * y.b + x.b * y.a + x.v * y.ab);
}
TREPB_HD HD operator+(const HD& x, double y) { return HD(x.v + y, x.a, x.b, x.ab); }
TREPB_HD HD operator+(double y, const HD& x) { return HD(x.v + y, x.a, x.b, x.ab); }
TREPB_HD HD operator-(const HD& x, double y) { return HD(x.v - y, x.a, x.b, x.ab); }
TREPB_HD HD operator-(double y, const HD& x) { return HD(y - x.v, -x.a, -x.b, -x.ab); }
TREPB_HD HD operator*(const HD& x, double y) { return HD(x.v * y, x.a * y, x.b * y, x.ab * y); }
TREPB_HD HD operator*(double y, const HD& x) { return HD(x.v * y, x.a * y, x.b * y, x.ab * y); }
// g(x) with g' = d1, g'' = d2 at x.v
TREPB_HD HD hd_chain(const HD& x, double g, double d1, double d2) {
    return HD(g, d1 * x.a, d1 * x.b, d1 * x.ab + d2 * x.a * x.b);
}
TREPB_HD HD hd_inv(const HD& y) {
    const double r = 1.0 / y.v;
    return hd_chain(y, r, -r * r, 2.0 * r * r * r);
}
TREPB_HD HD operator/(const HD& x, const HD& y) { return x * hd_inv(y); }
TREPB_HD HD operator/(double x, const HD& y) { return x * hd_inv(y); }
TREPB_HD HD operator/(const HD& x, double y) { return x * (1.0 / y); }
TREPB_HD HD& operator+=(HD& x, const HD& y) { x.v += y.v; x.a += y.a; x.b += y.b; x.ab += y.ab; return x; }
TREPB_HD HD& operator-=(HD& x, const HD& y) { x.v -= y.v; x.a -= y.a; x.b -= y.b; x.ab -= y.ab; return x; }
TREPB_HD HD& operator*=(HD& x, const HD& y) { x = x * y; return x; }
TREPB_HD HD& operator+=(HD& x, double y) { x.v += y; return x; }
TREPB_HD HD& operator-=(HD& x, double y) { x.v -= y; return x; }
TREPB_HD HD& operator*=(HD& x, double y) { x.v *= y; x.a *= y; x.b *= y; x.ab *= y; return x; }

TREPB_HD void sincos_(const HD& x, HD* s, HD* c) {
    double sn, cs;
#if defined(__CUDA_ARCH__)
    sincos(x.v, &sn, &cs);
#else
    sn = sin(x.v);
    cs = cos(x.v);
#endif
    *s = hd_chain(x, sn, cs, -sn);
    *c = hd_chain(x, cs, -sn, -cs);
}
TREPB_HD HD sqrt_(const HD& x) {
    const double r = sqrt(x.v);
    return hd_chain(x, r, 0.5 / r, -0.25 / (r * x.v));
}
TREPB_HD bool isnan_(const HD& x) { return isnan(x.v); }

}  // namespace trepb
