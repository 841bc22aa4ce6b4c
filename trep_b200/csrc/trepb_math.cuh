// Batched MidpointVI math: one *instance* per call, host+device, no Python, no allocation.
//
// What is computed is what the reference computes in trep/_trep/midpointvi.c:391-1120
// (set_state / set_midpoint, calc_f, calc_bar_Df_*, DEL_solved, MidpointVI_solve_DEL,
// calc_deriv1_cache, calc_M2, calc_proj_inv, calc_deriv1) on top of the Lagrangian sums of
// trep/_trep/system.c:129-557 and the frame caches of trep/_trep/frame.c:839-2081.
//
// HOW it is computed is different (this is not a port).  The reference caches 4x4 matrices
// g, dg/dq, d2g/dqdq, vb, dvb/dq, ... per frame and per ancestor tuple and sums
// <vb_x, vb_y> over masses for every index pair.  Here the same quantities come from a
// two-pass recursion in link coordinates with 6-vectors (v, w) in the reference's `unhat`
// order (math-code.c:275-284):
//
//   pass 1 (root->leaf)  V_f  = Ad(lg_f^-1) V_parent + s_f dq_f            body velocity
//                        W_f  = [Ad(lg_f^-1) V_parent, s_f]                = d vb / d q_f
//   pass 2 (leaf->root)  Ic_f = I_f + sum_children Ad^T Ic_child Ad        composite inertia
//                        mu_f = I_f V_f + sum_children Ad^T mu_child       composite momentum
//   per joint j          L_ddq(j)   = s_j . mu_j          L_dq(j) = W_j . mu_j - dV/dq_j
//                        H_j = Ic_j s_j ,  G_j = Ic_j W_j - ad*_{s_j} mu_j
//   per ancestor i of j  (H, G carried up the chain as force vectors)
//                        L_ddqddq(i,j) = s_i.H   L_ddqdq(i,j) = s_i.G   L_ddqdq(j,i) = W_i.H
//                        L_dqdq(i,j)   = W_i.G  (+ gravity: w_i . (P_j x g))
//
// using  d s_i/d q_j = [s_i, s_j] for i above j (Lie bracket of twists) — the identities
// behind Johnson & Murphey's eqs. (3)-(10) that frame.c implements entry by entry.
// Cost per evaluation is O(frames + sum_j depth(j)) instead of O(masses * depth^2) 4x4 products.
// Points (constraints, springs, dampers) use world-frame kinematics:
//   dp_F/dq_j = a_j (prismatic) | a_j x (p_F - o_j) (revolute),   d2p_F/dq_i dq_j = w_i x dp_F/dq_j.
//
// Everything is templated on  Sys (RtSys | generated constexpr system)  and
// Ws (WsStrided | WsStatic<Sys>) so the same source is the general kernel and the
// fully-unrolled specialised kernel.
#pragma once
#include <math.h>
#include <type_traits>
#include "trepb_sys.h"
#include "trepb_ws.h"

namespace trepb {

// TREPB_UNROLL      small fixed-count loops (3/6/9): always fully unrolled so the 3- and 6-vectors
//                   stay in registers.
// TREPB_UNROLL_SYS  loops over system-sized ranges (frames, configs, constraints ...): fully
//                   unrolled for a compile-time system (Sys::kUnroll large), left rolled for the
//                   table-driven system (Sys::kUnroll == 1).
#if defined(__CUDACC__)
#define TREPB_UNROLL _Pragma("unroll")
#define TREPB_UNROLL_SYS _Pragma("unroll (Sys::kUnroll)")
#else
#define TREPB_UNROLL
#define TREPB_UNROLL_SYS
#endif
// for (j = lo; j < n; ++j) with lo depending on an outer loop counter.  The unroller works inner loop
// first, so at that time such a bound is unknown and a rolled remainder loop stays behind - and with it
// dynamic indices that push the whole compile-time workspace into local memory.  The compile-time
// flavour therefore runs over the full range [0, n) and guards the body; the guard folds after unrolling.
#define TREPB_FOR_FROM(j, lo, n) \
    TREPB_UNROLL_SYS for (int j = Sys::kStatic ? 0 : (lo); j < (n); ++j) if (!Sys::kStatic || j >= (lo))


// Optional phase timing (development builds only: -DTREPB_PHASE_TIMING): clock64 deltas of one
// lane per warp accumulated into a device array, read back through trepb_phase_ticks().
#if defined(TREPB_PHASE_TIMING) && defined(__CUDA_ARCH__)
extern __device__ unsigned long long g_phase_ticks[32];
#define TREPB_TICK_INIT long long tick_ = clock64();
#define TREPB_TICK(i) do { const long long now_ = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase_ticks[i], (unsigned long long)(now_ - tick_)); tick_ = clock64(); } while (0)
#else
#define TREPB_TICK_INIT
#define TREPB_TICK(i)
#endif

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
struct Axis {
    int a, b, c;
    bool rot;
};
TREPB_HD Axis axis_of(int kind) {
    Axis x;
    int k = kind - K_TX;
    x.rot = k >= 3;
    x.a = x.rot ? k - 3 : k;
    x.b = (x.a + 1) % 3;
    x.c = (x.a + 2) % 3;
    return x;
}
TREPB_HD int symi(int r, int s) {  // index into (00,11,22,01,02,12)
    return r == s ? r : (r + s + 2);
}
template <class T>
TREPB_HD void cross3(const T* x, const T* y, T* o) {
    o[0] = x[1] * y[2] - x[2] * y[1];
    o[1] = x[2] * y[0] - x[0] * y[2];
    o[2] = x[0] * y[1] - x[1] * y[0];
}
template <class T>
TREPB_HD T dot3(const T* x, const T* y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }
template <class T>
TREPB_HD T dot6(const T* x, const T* y) {
    return x[0] * y[0] + x[1] * y[1] + x[2] * y[2] + x[3] * y[3] + x[4] * y[4] + x[5] * y[5];
}
// planar rotation of components (b,c):  R^T x  (parent->child)  and  R x  (child->parent)
template <class T>
TREPB_HD void rotT(T* x, int b, int c, T cs, T sn) {
    T xb = x[b], xc = x[c];
    x[b] = cs * xb + sn * xc;
    x[c] = -sn * xb + cs * xc;
}
template <class T>
TREPB_HD void rotF(T* x, int b, int c, T cs, T sn) {
    T xb = x[b], xc = x[c];
    x[b] = cs * xb - sn * xc;
    x[c] = sn * xb + cs * xc;
}
#if defined(__CUDACC__)
// sin and cos of one angle for the device.  Same structure as the library routine (Cody-Waite reduction by
// multiples of pi/2 in three pieces, minimax polynomials in r^2 on [-pi/4, pi/4], quadrant swap), with the
// coefficients in constant memory: they then reach the FP64 pipe as constant-bank operands of the
// DFMAs, where the inlined library routine materialised every coefficient with two uniform-register
// moves in the instruction stream (98 UMOV of the 429 warp instructions of a damped-pendulum DEL step,
// profiles/r01i_step_raw.txt).  Coefficients: the classical fdlibm kernels (k_sin.c / k_cos.c, |error| < 1 ulp
// on the reduced argument).  Arguments beyond 1e5 in magnitude take the library routine.
static __constant__ double kSinCos[16] = {
    6.36619772367581382433e-01,    //  0  2/pi
    1.57079632673412561417e+00,    //  1  pi/2, first 33 bits
    6.07710050630396597660e-11,    //  2  pi/2, next 33 bits
    2.02226624871116645580e-21,    //  3  pi/2, next 33 bits
    -1.66666666666666324348e-01,   //  4  S1
    8.33333333332248946124e-03,    //  5  S2
    -1.98412698298579493134e-04,   //  6  S3
    2.75573137070700676789e-06,    //  7  S4
    -2.50507602534068634195e-08,   //  8  S5
    1.58969099521155010221e-10,    //  9  S6
    4.16666666666666019037e-02,    // 10  C1
    -1.38888888888741095749e-03,   // 11  C2
    2.48015872894767294178e-05,    // 12  C3
    -2.75573143513906633035e-07,   // 13  C4
    2.08757232129817482790e-09,    // 14  C5
    -1.13596475577881948265e-11};  // 15  C6
__device__ __forceinline__ void sincos_dev(double x, double* s, double* c) {
    if (!(fabs(x) <= 1.0e5)) { sincos(x, s, c); return; }
    const double j = rint(x * kSinCos[0]);
    double r = fma(-j, kSinCos[1], x);
    r = fma(-j, kSinCos[2], r);
    r = fma(-j, kSinCos[3], r);
    const double z = r * r;
    double ps = fma(z, kSinCos[9], kSinCos[8]);
    double pc = fma(z, kSinCos[15], kSinCos[14]);
    ps = fma(z, ps, kSinCos[7]);  pc = fma(z, pc, kSinCos[13]);
    ps = fma(z, ps, kSinCos[6]);  pc = fma(z, pc, kSinCos[12]);
    ps = fma(z, ps, kSinCos[5]);  pc = fma(z, pc, kSinCos[11]);
    ps = fma(z, ps, kSinCos[4]);  pc = fma(z, pc, kSinCos[10]);
    const double sr = fma(z * r, ps, r);                       // r + r^3 S(r^2)
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));       // 1 - r^2/2 + r^4 C(r^2)
    const int q = (int)j;
    const double a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    *s = (q & 2) ? -a : a;
    *c = ((q + 1) & 2) ? -b : b;
}
#endif
TREPB_HD void sincos_(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    sincos_dev(x, s, c);
#else
    *s = sin(x);
    *c = cos(x);
#endif
}
TREPB_HD double div_r(double x, double d, double r);
TREPB_HD bool isnan_(double x) { return isnan(x); }
// value part of a (hyper-)dual number / the number itself
TREPB_HD double val_(double x) { return x; }
template <class T> TREPB_HD double val_(const T& x) { return x.v; }
TREPB_HD double sqrt_(double x) { return sqrt(x); }

// ---------------------------------------------------------------------------------------------
// pass 1: root -> leaf.  Inputs ws.qe (evaluation configuration) and ws.dq.
//   with_vel   : body velocities V, W and gravity direction gf (needed at the midpoint)
//   with_world : world pose Rw, pw of the frames that carry points (constraints/springs)
// ---------------------------------------------------------------------------------------------
template <class Sys, class Ws>
TREPB_HD void pass1(const Sys& sys, Ws& ws, bool with_vel, bool with_world) {
    using Real = typename Ws::Real;
    TREPB_UNROLL_SYS
    for (int f = 1; f < sys.NF(); ++f) {
        const int par = sys.parent(f), kind = sys.kind(f), cfg = sys.config(f);
        const bool vel = with_vel && sys.mass_below(f);
        const bool wrl = with_world && sys.need_world(f);
        if (!vel && !wrl) continue;
        const Real x = cfg >= 0 ? ws.qe(cfg) : sys.value(f);
        Real g[3], V[6], R[9], p[3];
        // vz: no variable joint above f, so the velocity carried in from the parent is exactly zero
        // and every operation on it can be dropped without changing a single bit of the result
        const bool vz = sys.vzero(f);
        if (vel) {
            if (par == 0) {
                TREPB_UNROLL for (int k = 0; k < 3; ++k) g[k] = sys.gravity(k);
            } else {
                TREPB_UNROLL for (int k = 0; k < 3; ++k) g[k] = ws.gf(par, k);
            }
            if (vz) {
                TREPB_UNROLL for (int k = 0; k < 6; ++k) V[k] = 0.0;
            } else {
                TREPB_UNROLL for (int k = 0; k < 6; ++k) V[k] = ws.V(par, k);
            }
        }
        if (wrl) {
            if (par == 0) {
                TREPB_UNROLL for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
                TREPB_UNROLL for (int k = 0; k < 3; ++k) p[k] = 0.0;
            } else {
                TREPB_UNROLL for (int k = 0; k < 9; ++k) R[k] = ws.Rw(par, k);
                TREPB_UNROLL for (int k = 0; k < 3; ++k) p[k] = ws.pw(par, k);
            }
        }
        if (kind == K_CONST_SE3) {
            Real lR[9], lp[3];
            TREPB_UNROLL for (int r = 0; r < 3; ++r) {
                TREPB_UNROLL for (int c = 0; c < 3; ++c) lR[r * 3 + c] = sys.se3(f, r * 4 + c);
                lp[r] = sys.se3(f, r * 4 + 3);
            }
            if (vel) {
                Real t[3], u[3];
                // v' = R^T (v + w x p), w' = R^T w, g' = R^T g
                if (!vz) {
                    cross3(V + 3, lp, t);
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) t[k] += V[k];
                    TREPB_UNROLL for (int c = 0; c < 3; ++c) u[c] = lR[c] * t[0] + lR[3 + c] * t[1] + lR[6 + c] * t[2];
                    TREPB_UNROLL for (int c = 0; c < 3; ++c) t[c] = lR[c] * V[3] + lR[3 + c] * V[4] + lR[6 + c] * V[5];
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { V[k] = u[k]; V[3 + k] = t[k]; }
                }
                TREPB_UNROLL for (int c = 0; c < 3; ++c) u[c] = lR[c] * g[0] + lR[3 + c] * g[1] + lR[6 + c] * g[2];
                TREPB_UNROLL for (int k = 0; k < 3; ++k) g[k] = u[k];
            }
            if (wrl) {
                Real Rn[9];
                TREPB_UNROLL for (int r = 0; r < 3; ++r) {
                    p[r] += R[r * 3] * lp[0] + R[r * 3 + 1] * lp[1] + R[r * 3 + 2] * lp[2];
                }
                TREPB_UNROLL for (int r = 0; r < 3; ++r)
                    TREPB_UNROLL for (int c = 0; c < 3; ++c)
                        Rn[r * 3 + c] = R[r * 3] * lR[c] + R[r * 3 + 1] * lR[3 + c] + R[r * 3 + 2] * lR[6 + c];
                TREPB_UNROLL for (int k = 0; k < 9; ++k) R[k] = Rn[k];
            }
        } else {
            const Axis ax = axis_of(kind);
            if (ax.rot) {
                Real sn, cs;
                sincos_(x, &sn, &cs);
                if (vel) {
                    // kept for pass 2 (force transforms); world-only passes at q1/q2 must not
                    // disturb the midpoint values (eval_mid_again)
                    ws.cs(f, 0) = cs;
                    ws.cs(f, 1) = sn;
                    rotT(g, ax.b, ax.c, cs, sn);
                    if (!vz) {
                        rotT(V, ax.b, ax.c, cs, sn);
                        rotT(V + 3, ax.b, ax.c, cs, sn);
                    }
                }
                if (wrl) {
                    TREPB_UNROLL for (int r = 0; r < 3; ++r) {
                        Real rb = R[r * 3 + ax.b], rc = R[r * 3 + ax.c];
                        R[r * 3 + ax.b] = cs * rb + sn * rc;
                        R[r * 3 + ax.c] = -sn * rb + cs * rc;
                    }
                }
            } else {
                if (vel && !vz) {
                    // v' = v + w x (x e_a)
                    V[ax.b] += x * V[3 + ax.c];
                    V[ax.c] -= x * V[3 + ax.b];
                }
                if (wrl) {
                    TREPB_UNROLL for (int r = 0; r < 3; ++r) p[r] += x * R[r * 3 + ax.a];
                }
            }
            if (vel && cfg >= 0) {
                // W = [S, s] with S the velocity carried in from the parent
                // (W is identically zero when vz: neither stored here nor read in pass 2)
                if (!vz) {
                    Real W[6];
                    TREPB_UNROLL for (int k = 0; k < 6; ++k) W[k] = 0.0;
                    if (ax.rot) {
                        W[ax.b] = V[ax.c];          // v_S x e_a
                        W[ax.c] = -V[ax.b];
                        W[3 + ax.b] = V[3 + ax.c];  // w_S x e_a
                        W[3 + ax.c] = -V[3 + ax.b];
                    } else {
                        W[ax.b] = V[3 + ax.c];      // w_S x e_a
                        W[ax.c] = -V[3 + ax.b];
                    }
                    TREPB_UNROLL for (int k = 0; k < 6; ++k) ws.W(f, k) = W[k];
                }
                if (ax.rot) V[3 + ax.a] += ws.dq(cfg);
                else V[ax.a] += ws.dq(cfg);
            }
        }
        if (vel) {
            TREPB_UNROLL for (int k = 0; k < 3; ++k) ws.gf(f, k) = g[k];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) ws.V(f, k) = V[k];
        }
        if (wrl) {
            TREPB_UNROLL for (int k = 0; k < 9; ++k) ws.Rw(f, k) = R[k];
            TREPB_UNROLL for (int k = 0; k < 3; ++k) ws.pw(f, k) = p[k];
        }
    }
}

// force-vector transform child -> parent through frame f:  f' = R f, n' = R n + p x (R f)
template <class Sys, class Ws>
TREPB_HD void force_up(const Sys& sys, Ws& ws, int f, typename Ws::Real* F) {
    using Real = typename Ws::Real;
    const int kind = sys.kind(f);
    if (kind == K_CONST_SE3) {
        Real a[3], b[3], lp[3], t[3];
        TREPB_UNROLL for (int r = 0; r < 3; ++r) {
            a[r] = sys.se3(f, r * 4) * F[0] + sys.se3(f, r * 4 + 1) * F[1] + sys.se3(f, r * 4 + 2) * F[2];
            b[r] = sys.se3(f, r * 4) * F[3] + sys.se3(f, r * 4 + 1) * F[4] + sys.se3(f, r * 4 + 2) * F[5];
            lp[r] = sys.se3(f, r * 4 + 3);
        }
        cross3(lp, a, t);
        TREPB_UNROLL for (int k = 0; k < 3; ++k) { F[k] = a[k]; F[3 + k] = b[k] + t[k]; }
    } else {
        const Axis ax = axis_of(kind);
        if (ax.rot) {
            const Real cs = ws.cs(f, 0), sn = ws.cs(f, 1);
            rotF(F, ax.b, ax.c, cs, sn);
            rotF(F + 3, ax.b, ax.c, cs, sn);
        } else {
            const int cfg = sys.config(f);
            const Real x = cfg >= 0 ? ws.qe(cfg) : sys.value(f);
            F[3 + ax.b] -= x * F[ax.c];
            F[3 + ax.c] += x * F[ax.b];
        }
    }
}
// free-vector transform child -> parent
template <class Sys, class Ws>
TREPB_HD void vec_up(const Sys& sys, Ws& ws, int f, typename Ws::Real* N) {
    using Real = typename Ws::Real;
    const int kind = sys.kind(f);
    if (kind == K_CONST_SE3) {
        Real a[3];
        TREPB_UNROLL for (int r = 0; r < 3; ++r)
            a[r] = sys.se3(f, r * 4) * N[0] + sys.se3(f, r * 4 + 1) * N[1] + sys.se3(f, r * 4 + 2) * N[2];
        TREPB_UNROLL for (int k = 0; k < 3; ++k) N[k] = a[k];
    } else {
        const Axis ax = axis_of(kind);
        if (ax.rot) rotF(N, ax.b, ax.c, ws.cs(f, 0), ws.cs(f, 1));
    }
}

// (f, n) = I (v, w) for I = (m, h, Isym6):  f = m v + w x h ,  n = Ibar w + h x v
template <class Real>
TREPB_HD void inertia_apply(Real m, const Real* h, const Real* I, const Real* X, Real* F) {
    Real t[3];
    cross3(X + 3, h, t);
    TREPB_UNROLL for (int k = 0; k < 3; ++k) F[k] = m * X[k] + t[k];
    cross3(h, X, t);
    F[3] = I[0] * X[3] + I[3] * X[4] + I[4] * X[5] + t[0];
    F[4] = I[3] * X[3] + I[1] * X[4] + I[5] * X[5] + t[1];
    F[5] = I[4] * X[3] + I[5] * X[4] + I[2] * X[5] + t[2];
}

// ---------------------------------------------------------------------------------------------
// pass 2: leaf -> root.  order 1: Lq, Lv.  order 2: also Lqq, Lvq, Lvv (all nq x nq).
// Potentials other than gravity are added by add_potentials().
// ---------------------------------------------------------------------------------------------
// zero_tables = false: the caller has already cleared the entries of Lqq / Lvq / Lvv that can be
// written (trepb_d2jac.cuh clears only the structurally non-zero ones).
template <class Sys, class Ws>
TREPB_HD void pass2(const Sys& sys, Ws& ws, int order, bool zero_tables = true) {
    using Real = typename Ws::Real;
    const int nq = sys.NQ();
    TREPB_UNROLL_SYS for (int i = 0; i < nq; ++i) {
        ws.Lq(i) = 0.0;
        ws.Lv(i) = 0.0;
    }
    if (order >= 2 && zero_tables) {
        TREPB_UNROLL_SYS for (int i = 0; i < nq; ++i)
            TREPB_UNROLL_SYS for (int j = 0; j < nq; ++j) {
                ws.Lqq(i, j) = 0.0;
                ws.Lvq(i, j) = 0.0;
                ws.Lvv(i, j) = 0.0;
            }
    }
    // own inertia / momentum
    TREPB_UNROLL_SYS
    for (int f = 1; f < sys.NF(); ++f) {
        if (!sys.mass_below(f)) continue;
        const double m = sys.mass(f, 0);
        ws.Im(f) = m;
        TREPB_UNROLL for (int k = 0; k < 3; ++k) ws.Ih(f, k) = 0.0;
        if (order >= 2) {
            TREPB_UNROLL for (int k = 0; k < 3; ++k) {
                ws.II(f, k) = sys.mass(f, 1 + k);
                ws.II(f, 3 + k) = 0.0;
            }
        }
        if (sys.has_mass(f)) {
            TREPB_UNROLL for (int k = 0; k < 3; ++k) {
                ws.mu(f, k) = m * ws.V(f, k);
                ws.mu(f, 3 + k) = sys.mass(f, 1 + k) * ws.V(f, 3 + k);
            }
        } else {
            TREPB_UNROLL for (int k = 0; k < 6; ++k) ws.mu(f, k) = 0.0;
        }
    }
    TREPB_UNROLL_SYS
    for (int f = sys.NF() - 1; f >= 1; --f) {
        if (!sys.mass_below(f)) continue;
        const int kind = sys.kind(f), cfg = sys.config(f), par = sys.parent(f);
        Real m = ws.Im(f), h[3], I[6], mu[6], g[3];
        TREPB_UNROLL for (int k = 0; k < 3; ++k) { h[k] = ws.Ih(f, k); g[k] = ws.gf(f, k); }
        TREPB_UNROLL for (int k = 0; k < 6; ++k) mu[k] = ws.mu(f, k);
        if (order >= 2) { TREPB_UNROLL for (int k = 0; k < 6; ++k) I[k] = ws.II(f, k); }

        if (cfg >= 0) {
            const Axis ax = axis_of(kind);
            const bool wz = sys.vzero(f);   // W_f == 0 identically (first variable joint of its chain)
            Real W[6];
            if (wz) { TREPB_UNROLL for (int k = 0; k < 6; ++k) W[k] = 0.0; }
            else { TREPB_UNROLL for (int k = 0; k < 6; ++k) W[k] = ws.W(f, k); }
            // first order: L_ddq = s.mu ; L_dq = W.mu + g.(m v_s + w_s x h)
            Real hxg[3];
            cross3(h, g, hxg);
            ws.Lv(cfg) = ax.rot ? mu[3 + ax.a] : mu[ax.a];
            const Real grav1 = sys.gravity_on() ? (ax.rot ? hxg[ax.a] : m * g[ax.a]) : Real(0.0);
            if (wz) ws.Lq(cfg) = grav1;
            else ws.Lq(cfg) = dot6(W, mu) + grav1;
            if (order >= 2) {
                Real H[6], G[6], N[3], P[3], t[3];
                // H = Ic s
                TREPB_UNROLL for (int k = 0; k < 6; ++k) H[k] = 0.0;
                if (ax.rot) {
                    H[ax.b] = -h[ax.c];  // e_a x h
                    H[ax.c] = h[ax.b];
                    H[3 + ax.a] = I[symi(ax.a, ax.a)];
                    H[3 + ax.b] = I[symi(ax.a, ax.b)];
                    H[3 + ax.c] = I[symi(ax.a, ax.c)];
                } else {
                    H[ax.a] = m;
                    H[3 + ax.b] = h[ax.c];  // h x e_a
                    H[3 + ax.c] = -h[ax.b];
                }
                // G = Ic W - ad*_s mu ;  ad*_s mu = (f x w_s, n x w_s + f x v_s)
                if (wz) { TREPB_UNROLL for (int k = 0; k < 6; ++k) G[k] = 0.0; }
                else inertia_apply(m, h, I, W, G);
                if (ax.rot) {
                    G[ax.b] -= mu[ax.c];
                    G[ax.c] += mu[ax.b];
                    G[3 + ax.b] -= mu[3 + ax.c];
                    G[3 + ax.c] += mu[3 + ax.b];
                } else {
                    G[3 + ax.b] -= mu[ax.c];
                    G[3 + ax.c] += mu[ax.b];
                }
                // gravity second derivative:  N = P x g ,  P = m v_s + w_s x h
                TREPB_UNROLL for (int k = 0; k < 3; ++k) P[k] = 0.0;
                if (ax.rot) {
                    P[ax.b] = -h[ax.c];
                    P[ax.c] = h[ax.b];
                } else {
                    P[ax.a] = m;
                }
                cross3(P, g, N);
                (void)t;
                // walk to the root
                int cur = f;
                TREPB_UNROLL_SYS
                for (int lvl = 0; lvl < sys.MAXDEPTH(); ++lvl) {
                    const int ci = sys.config(cur);
                    if (ci >= 0) {
                        const Axis ai = axis_of(sys.kind(cur));
                        const Real sH = ai.rot ? H[3 + ai.a] : H[ai.a];
                        const Real sG = ai.rot ? G[3 + ai.a] : G[ai.a];
                        const bool wzi = sys.vzero(cur);
                        Real Wi[6];
                        if (wzi) { TREPB_UNROLL for (int k = 0; k < 6; ++k) Wi[k] = 0.0; }
                        else { TREPB_UNROLL for (int k = 0; k < 6; ++k) Wi[k] = ws.W(cur, k); }
                        Real WG = wzi ? Real(0.0) : dot6(Wi, G);
                        if (sys.gravity_on() && ai.rot) WG += N[ai.a];
                        if (ci == cfg) {
                            ws.Lvv(cfg, cfg) = sH;
                            ws.Lvq(cfg, cfg) = sG;
                            ws.Lqq(cfg, cfg) = WG;
                        } else {
                            ws.Lvv(ci, cfg) = sH;
                            ws.Lvv(cfg, ci) = sH;
                            ws.Lvq(ci, cfg) = sG;
                            ws.Lvq(cfg, ci) = wzi ? Real(0.0) : dot6(Wi, H);
                            ws.Lqq(ci, cfg) = WG;
                            ws.Lqq(cfg, ci) = WG;
                        }
                    }
                    const int up = sys.parent(cur);
                    if (up == 0) break;
                    force_up(sys, ws, cur, H);
                    force_up(sys, ws, cur, G);
                    if (sys.gravity_on()) vec_up(sys, ws, cur, N);
                    cur = up;
                }
            }
        }
        if (par == 0) continue;
        // accumulate composite momentum / inertia into the parent
        force_up(sys, ws, f, mu);
        TREPB_UNROLL for (int k = 0; k < 6; ++k) ws.mu(par, k) += mu[k];
        if (kind == K_CONST_SE3) {
            Real R[9], lp[3], hr[3];
            TREPB_UNROLL for (int r = 0; r < 3; ++r) {
                TREPB_UNROLL for (int c = 0; c < 3; ++c) R[r * 3 + c] = sys.se3(f, r * 4 + c);
                lp[r] = sys.se3(f, r * 4 + 3);
            }
            TREPB_UNROLL for (int r = 0; r < 3; ++r) hr[r] = R[r * 3] * h[0] + R[r * 3 + 1] * h[1] + R[r * 3 + 2] * h[2];
            if (order >= 2) {
                // Ibar' = R Ibar R^T + 2(hr.p)1 - hr p^T - p hr^T + m(|p|^2 1 - p p^T)
                Real M[9], T[9];
                M[0] = I[0]; M[4] = I[1]; M[8] = I[2];
                M[1] = M[3] = I[3]; M[2] = M[6] = I[4]; M[5] = M[7] = I[5];
                TREPB_UNROLL for (int r = 0; r < 3; ++r)
                    TREPB_UNROLL for (int c = 0; c < 3; ++c)
                        T[r * 3 + c] = R[r * 3] * M[c] + R[r * 3 + 1] * M[3 + c] + R[r * 3 + 2] * M[6 + c];
                const Real hp = dot3(hr, lp), pp = dot3(lp, lp);
                TREPB_UNROLL for (int r = 0; r < 3; ++r)
                    TREPB_UNROLL for (int c = r; c < 3; ++c) {
                        Real v = T[r * 3] * R[c * 3] + T[r * 3 + 1] * R[c * 3 + 1] + T[r * 3 + 2] * R[c * 3 + 2];
                        v += -hr[r] * lp[c] - lp[r] * hr[c] - m * lp[r] * lp[c];
                        if (r == c) v += 2.0 * hp + m * pp;
                        ws.II(par, symi(r, c)) += v;
                    }
            }
            TREPB_UNROLL for (int k = 0; k < 3; ++k) ws.Ih(par, k) += hr[k] + m * lp[k];
            ws.Im(par) += m;
        } else {
            const Axis ax = axis_of(kind);
            if (ax.rot) {
                const Real cs = ws.cs(f, 0), sn = ws.cs(f, 1);
                if (order >= 2) {
                    const int iaa = symi(ax.a, ax.a), ibb = symi(ax.b, ax.b), icc = symi(ax.c, ax.c);
                    const int iab = symi(ax.a, ax.b), iac = symi(ax.a, ax.c), ibc = symi(ax.b, ax.c);
                    const Real c2 = cs * cs, s2 = sn * sn, sc = sn * cs;
                    ws.II(par, iaa) += I[iaa];
                    ws.II(par, iab) += cs * I[iab] - sn * I[iac];
                    ws.II(par, iac) += sn * I[iab] + cs * I[iac];
                    ws.II(par, ibb) += c2 * I[ibb] - 2.0 * sc * I[ibc] + s2 * I[icc];
                    ws.II(par, icc) += s2 * I[ibb] + 2.0 * sc * I[ibc] + c2 * I[icc];
                    ws.II(par, ibc) += sc * (I[ibb] - I[icc]) + (c2 - s2) * I[ibc];
                }
                rotF(h, ax.b, ax.c, cs, sn);
                TREPB_UNROLL for (int k = 0; k < 3; ++k) ws.Ih(par, k) += h[k];
                ws.Im(par) += m;
            } else {
                const Real x = cfg >= 0 ? ws.qe(cfg) : sys.value(f);
                if (order >= 2) {
                    const int ibb = symi(ax.b, ax.b), icc = symi(ax.c, ax.c);
                    const int iab = symi(ax.a, ax.b), iac = symi(ax.a, ax.c);
                    const Real d = 2.0 * h[ax.a] * x + m * x * x;
                    TREPB_UNROLL for (int k = 0; k < 6; ++k) ws.II(par, k) += I[k];
                    ws.II(par, ibb) += d;
                    ws.II(par, icc) += d;
                    ws.II(par, iab) -= x * h[ax.b];
                    ws.II(par, iac) -= x * h[ax.c];
                }
                h[ax.a] += m * x;
                TREPB_UNROLL for (int k = 0; k < 3; ++k) ws.Ih(par, k) += h[k];
                ws.Im(par) += m;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// world-frame point kinematics
// ---------------------------------------------------------------------------------------------
template <class Sys, class Ws>
TREPB_HD void joint_axis_w(const Sys& sys, Ws& ws, int j, typename Ws::Real* aw, bool* rot, int* fj) {
    using Real = typename Ws::Real;
    const int f = sys.cfg_frame(j);
    const Axis ax = axis_of(sys.kind(f));
    TREPB_UNROLL for (int r = 0; r < 3; ++r) aw[r] = ws.Rw(f, r * 3 + ax.a);
    *rot = ax.rot;
    *fj = f;
}
template <class Sys, class Ws>
TREPB_HD void frame_pos(const Sys& sys, Ws& ws, int F, typename Ws::Real* p) {
    using Real = typename Ws::Real;
    if (F == 0) { p[0] = p[1] = p[2] = 0.0; }
    else { TREPB_UNROLL for (int k = 0; k < 3; ++k) p[k] = ws.pw(F, k); }
}
// d p_F / d q_j   (zero when F does not depend on q_j; trep/_trep/frame.c:2247-2262)
template <class Sys, class Ws>
TREPB_HD void dpoint(const Sys& sys, Ws& ws, int F, int j, typename Ws::Real* out) {
    using Real = typename Ws::Real;
    out[0] = out[1] = out[2] = 0.0;
    if (F == 0 || sys.cfg_frame(j) < 0 || !sys.dep(F, j)) return;
    Real aw[3];
    bool rot;
    int fj;
    joint_axis_w(sys, ws, j, aw, &rot, &fj);
    if (rot) {
        Real r[3];
        TREPB_UNROLL for (int k = 0; k < 3; ++k) r[k] = ws.pw(F, k) - ws.pw(fj, k);
        cross3(aw, r, out);
    } else {
        TREPB_UNROLL for (int k = 0; k < 3; ++k) out[k] = aw[k];
    }
}
// d2 p_F / d q_i d q_j
template <class Sys, class Ws>
TREPB_HD void ddpoint(const Sys& sys, Ws& ws, int F, int i, int j, typename Ws::Real* out) {
    using Real = typename Ws::Real;
    out[0] = out[1] = out[2] = 0.0;
    if (F == 0 || sys.cfg_frame(i) < 0 || sys.cfg_frame(j) < 0) return;
    if (!sys.dep(F, i) || !sys.dep(F, j)) return;
    // the upper joint's axis crosses the lower joint's first derivative
    int up = i, lo = j;
    if (!sys.dep(sys.cfg_frame(j), i)) { up = j; lo = i; }
    Real aw[3], d[3];
    bool rot;
    int fu;
    joint_axis_w(sys, ws, up, aw, &rot, &fu);
    if (!rot) return;
    dpoint(sys, ws, F, lo, d);
    cross3(aw, d, out);
}

// world orientation of frame F (identity for the world frame)
template <class Sys, class Ws>
TREPB_HD void frame_rot(const Sys& sys, Ws& ws, int F, typename Ws::Real* R) {
    if (F == 0) { TREPB_UNROLL for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0; }
    else { TREPB_UNROLL for (int k = 0; k < 9; ++k) R[k] = ws.Rw(F, k); }
}
// d (R_F n) / d q_j = a_j x (R_F n) for a revolute ancestor-or-self joint of F, zero otherwise
template <class Sys, class Ws>
TREPB_HD void dnormal(const Sys& sys, Ws& ws, int F, int j, const typename Ws::Real* nw, typename Ws::Real* out) {
    using Real = typename Ws::Real;
    out[0] = out[1] = out[2] = 0.0;
    if (F == 0 || sys.cfg_frame(j) < 0 || !sys.dep(F, j)) return;
    Real aw[3];
    bool rot;
    int fj;
    joint_axis_w(sys, ws, j, aw, &rot, &fj);
    if (rot) cross3(aw, nw, out);
}
// d2 (R_F n) / d q_i d q_j = a_up x (a_lo x R_F n)
template <class Sys, class Ws>
TREPB_HD void ddnormal(const Sys& sys, Ws& ws, int F, int i, int j, const typename Ws::Real* nw, typename Ws::Real* out) {
    using Real = typename Ws::Real;
    out[0] = out[1] = out[2] = 0.0;
    if (F == 0 || sys.cfg_frame(i) < 0 || sys.cfg_frame(j) < 0) return;
    if (!sys.dep(F, i) || !sys.dep(F, j)) return;
    int up = i, lo = j;
    if (!sys.dep(sys.cfg_frame(j), i)) { up = j; lo = i; }
    Real au[3], al[3], t[3];
    bool ru, rl;
    int f;
    joint_axis_w(sys, ws, up, au, &ru, &f);
    joint_axis_w(sys, ws, lo, al, &rl, &f);
    if (!ru || !rl) return;
    cross3(al, nw, t);
    cross3(au, t, out);
}

// v = pA - pB and dv_j = d(pA-pB)/dq_j for all configs into ws.dv
// dep_only: dv_j is left untouched for the configs (other than `also`) neither frame depends on
// (it is zero there and the caller does not read it)
template <class Sys, class Ws>
TREPB_HD void pair_first(const Sys& sys, Ws& ws, int A, int B, typename Ws::Real* v, bool dep_only = false,
                         int also = -1) {
    using Real = typename Ws::Real;
    Real pa[3], pb[3];
    frame_pos(sys, ws, A, pa);
    frame_pos(sys, ws, B, pb);
    TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
    TREPB_UNROLL_SYS
    for (int j = 0; j < sys.NQ(); ++j) {
        if (dep_only && !(sys.dep(A, j) || sys.dep(B, j) || j == also)) continue;
        Real da[3], db[3];
        dpoint(sys, ws, A, j, da);
        dpoint(sys, ws, B, j, db);
        TREPB_UNROLL for (int k = 0; k < 3; ++k) ws.dv(j, k) = da[k] - db[k];
    }
}
template <class Sys, class Ws>
TREPB_HD void pair_second(const Sys& sys, Ws& ws, int A, int B, int i, int j, typename Ws::Real* ddv) {
    using Real = typename Ws::Real;
    Real a[3], b[3];
    ddpoint(sys, ws, A, i, j, a);
    ddpoint(sys, ws, B, i, j, b);
    TREPB_UNROLL for (int k = 0; k < 3; ++k) ddv[k] = a[k] - b[k];
}

// ---------------------------------------------------------------------------------------------
// constraints (trep/_trep/constraints/distance.c:16-100, point.c:16-46)
//   mode bit 0: h -> ws.hc      bit 1: Dh -> dest (Dh1 or Dh2)     bit 2: DDhl += lam_c * h_c,ij
// World pose must be current (pass1 with_world at the wanted evaluation point).
// ---------------------------------------------------------------------------------------------
template <class Sys, class Ws>
TREPB_HD void constraints_eval(const Sys& sys, Ws& ws, int mode, int which_dh, bool zero_ddh = true) {
    using Real = typename Ws::Real;
    const int nq = sys.NQ();
    if ((mode & 4) && zero_ddh) {
        TREPB_UNROLL_SYS for (int i = 0; i < nq; ++i)
            TREPB_UNROLL_SYS for (int j = 0; j < nq; ++j) ws.DDhl(i, j) = 0.0;
    }
    TREPB_UNROLL_SYS
    for (int c = 0; c < sys.NC(); ++c) {
        const int kind = sys.con_kind(c);
        const int A = sys.con_i(c, 0), B = sys.con_i(c, 1), third = sys.con_i(c, 2);
        Real v[3];
        // distance and plane constraints read dv only at the configs they depend on
        pair_first(sys, ws, A, B, v, !Sys::kStatic && (kind == C_DISTANCE || kind == C_PLANE),
                   kind == C_DISTANCE ? third : -1);
        if (kind == C_DISTANCE) {
            const Real d = third >= 0 ? ws.qe(third) : sys.con_d(c, 0);
            if (mode & 1) ws.hc(c) = dot3(v, v) - d * d;
            if (mode & 2) {
                TREPB_UNROLL_SYS
                for (int j = 0; j < nq; ++j) {
                    Real val = 0.0;
                    if (sys.dep(A, j) || sys.dep(B, j) || third == j) {
                        val = v[0] * ws.dv(j, 0) + v[1] * ws.dv(j, 1) + v[2] * ws.dv(j, 2);
                        if (third == j) val -= d;
                        val *= 2.0;
                    }
                    if (which_dh == 1) ws.Dh1(c, j) = val; else ws.Dh2(c, j) = val;
                }
            }
            if (mode & 4) {
                const Real lam = ws.lam(c);
                TREPB_UNROLL_SYS
                for (int i = 0; i < nq; ++i) {
                    if (!(sys.dep(A, i) || sys.dep(B, i) || third == i)) continue;
                    TREPB_FOR_FROM(j, i, nq) {
                        if (!(sys.dep(A, j) || sys.dep(B, j) || third == j)) continue;
                        Real ddv[3];
                        pair_second(sys, ws, A, B, i, j, ddv);
                        Real val = ws.dv(i, 0) * ws.dv(j, 0) + ws.dv(i, 1) * ws.dv(j, 1) + ws.dv(i, 2) * ws.dv(j, 2)
                                   + v[0] * ddv[0] + v[1] * ddv[1] + v[2] * ddv[2];
                        if (third == i && third == j) val -= 1.0;
                        val *= 2.0 * lam;
                        ws.DDhl(i, j) += val;
                        if (j != i) ws.DDhl(j, i) += val;
                    }
                }
            }
        } else if (kind == C_PLANE) {
            // h = (R_A n) . (p_A - p_B): point B on the plane through frame A's origin with normal
            // n fixed in frame A (trep/_trep/constraints/plane.c:14-84).  d = (n0, tol, n1, n2).
            Real RA[9], nw[3];
            frame_rot(sys, ws, A, RA);
            const double n0 = sys.con_d(c, 0), n1 = sys.con_d(c, 2), n2 = sys.con_d(c, 3);
            TREPB_UNROLL for (int r = 0; r < 3; ++r) nw[r] = RA[r * 3] * n0 + RA[r * 3 + 1] * n1 + RA[r * 3 + 2] * n2;
            if (mode & 1) ws.hc(c) = dot3(nw, v);
            if (mode & 2) {
                TREPB_UNROLL_SYS
                for (int j = 0; j < nq; ++j) {
                    Real val = 0.0;
                    if (sys.dep(A, j) || sys.dep(B, j)) {
                        Real dn[3];
                        dnormal(sys, ws, A, j, nw, dn);
                        val = dot3(dn, v) + (nw[0] * ws.dv(j, 0) + nw[1] * ws.dv(j, 1) + nw[2] * ws.dv(j, 2));
                    }
                    if (which_dh == 1) ws.Dh1(c, j) = val; else ws.Dh2(c, j) = val;
                }
            }
            if (mode & 4) {
                const Real lam = ws.lam(c);
                TREPB_UNROLL_SYS
                for (int i = 0; i < nq; ++i) {
                    if (!(sys.dep(A, i) || sys.dep(B, i))) continue;
                    Real dni[3];
                    dnormal(sys, ws, A, i, nw, dni);
                    TREPB_FOR_FROM(j, i, nq) {
                        if (!(sys.dep(A, j) || sys.dep(B, j))) continue;
                        Real dnj[3], ddn[3], ddv[3];
                        dnormal(sys, ws, A, j, nw, dnj);
                        ddnormal(sys, ws, A, i, j, nw, ddn);
                        pair_second(sys, ws, A, B, i, j, ddv);
                        Real val = dot3(ddn, v) + dot3(nw, ddv)
                                   + (dni[0] * ws.dv(j, 0) + dni[1] * ws.dv(j, 1) + dni[2] * ws.dv(j, 2))
                                   + (dnj[0] * ws.dv(i, 0) + dnj[1] * ws.dv(i, 1) + dnj[2] * ws.dv(i, 2));
                        val *= lam;
                        ws.DDhl(i, j) += val;
                        if (j != i) ws.DDhl(j, i) += val;
                    }
                }
            }
        } else {  // C_POINT1D
            const int comp = third;
            if (mode & 1) ws.hc(c) = v[comp];
            if (mode & 2) {
                TREPB_UNROLL_SYS
                for (int j = 0; j < nq; ++j) {
                    const Real val = ws.dv(j, comp);
                    if (which_dh == 1) ws.Dh1(c, j) = val; else ws.Dh2(c, j) = val;
                }
            }
            if (mode & 4) {
                const Real lam = ws.lam(c);
                TREPB_UNROLL_SYS
                for (int i = 0; i < nq; ++i)
                    TREPB_FOR_FROM(j, i, nq) {
                        Real ddv[3];
                        pair_second(sys, ws, A, B, i, j, ddv);
                        const Real val = lam * ddv[comp];
                        ws.DDhl(i, j) += val;
                        if (j != i) ws.DDhl(j, i) += val;
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// potentials other than gravity, at the current point (potentials/linearspring.c:30-74,
// configspring.c:22-38).  order as in pass2.
// ---------------------------------------------------------------------------------------------
template <class Sys, class Ws>
TREPB_HD void add_potentials(const Sys& sys, Ws& ws, int order) {
    using Real = typename Ws::Real;
    const int nq = sys.NQ();
    TREPB_UNROLL_SYS
    for (int p = 0; p < sys.NPOT(); ++p) {
        const int kind = sys.pot_kind(p);
        if (kind == P_CONFIG_SPRING) {
            const int c = sys.pot_i(p, 0);
            const double k = sys.pot_d(p, 0), q0 = sys.pot_d(p, 1);
            ws.Lq(c) -= k * (ws.qe(c) - q0);
            if (order >= 2) ws.Lqq(c, c) -= k;
        } else if (kind == P_NONLINEAR_CONFIG_SPRING) {
            // dV/dq = -y(m q + b) with y a piecewise quintic (potentials/nonlinear_config_spring.c:23-47,
            // spline.c:7-62).  i = config, dpool offset, number of x points; d = m, b;
            // dpool: x points [n], then coefficients [n-1][6] (highest power first).
            const int c = sys.pot_i(p, 0), off = sys.pot_i(p, 1), n = sys.pot_i(p, 2);
            const double m = sys.pot_d(p, 0), b = sys.pot_d(p, 1);
            const Real x = m * ws.qe(c) + b;
            const double xv = val_(x);
            int seg = 0;                                   // get_index (spline.c:7-22)
            if (xv >= sys.dpool(off + n - 1)) seg = n - 2;
            else if (!(xv < sys.dpool(off))) { while (xv >= sys.dpool(off + seg + 1)) ++seg; }
            const int co = off + n + 6 * seg;
            const double c0 = sys.dpool(co), c1 = sys.dpool(co + 1), c2 = sys.dpool(co + 2), c3 = sys.dpool(co + 3),
                         c4 = sys.dpool(co + 4), c5 = sys.dpool(co + 5);
            const Real dx = x - sys.dpool(off + seg);
            const Real dx2 = dx * dx, dx3 = dx2 * dx, dx4 = dx3 * dx, dx5 = dx4 * dx;
            ws.Lq(c) += c0 * dx5 + c1 * dx4 + c2 * dx3 + c3 * dx2 + c4 * dx + c5;
            if (order >= 2)
                ws.Lqq(c, c) += (5.0 * c0 * dx4 + 4.0 * c1 * dx3 + 3.0 * c2 * dx2 + 2.0 * c3 * dx + c4) * m;
        } else if (kind == P_LINEAR_SPRING) {
            const int A = sys.pot_i(p, 0), B = sys.pot_i(p, 1);
            const double k = sys.pot_d(p, 0), x0 = sys.pot_d(p, 1);
            Real v[3];
            pair_first(sys, ws, A, B, v);
            const Real x = sqrt_(dot3(v, v));
            TREPB_UNROLL_SYS
            for (int j = 0; j < nq; ++j) {
                Real dx = (1.0 / x) * (v[0] * ws.dv(j, 0) + v[1] * ws.dv(j, 1) + v[2] * ws.dv(j, 2));
                ws.dxs(j) = dx;
                Real val = k * (x - x0) * dx;
                if (isnan_(dx) && x0 == 0.0) val = 0.0;
                ws.Lq(j) -= val;
            }
            if (order >= 2) {
                // 1 / x^2 once per spring: the reference divides by x * x for every pair (linearspring.c:60-74);
                // one division and a multiplication per pair differ from that by a rounding of one Hessian term
                const Real rxx = 1.0 / (x * x);
                TREPB_UNROLL_SYS
                for (int i = 0; i < nq; ++i)
                    TREPB_FOR_FROM(j, i, nq) {
                        if (!(sys.dep(A, i) || sys.dep(B, i)) || !(sys.dep(A, j) || sys.dep(B, j))) continue;
                        Real ddv[3];
                        pair_second(sys, ws, A, B, i, j, ddv);
                        const Real vdi = v[0] * ws.dv(i, 0) + v[1] * ws.dv(i, 1) + v[2] * ws.dv(i, 2);
                        const Real didj = ws.dv(i, 0) * ws.dv(j, 0) + ws.dv(i, 1) * ws.dv(j, 1) + ws.dv(i, 2) * ws.dv(j, 2);
                        const Real dix = ws.dxs(i), djx = ws.dxs(j);
                        const Real ddx = -djx * rxx * vdi + 1.0 / x * didj + 1.0 / x * dot3(v, ddv);
                        const Real val = k * dix * djx + k * (x - x0) * ddx;
                        ws.Lqq(i, j) -= val;
                        if (j != i) ws.Lqq(j, i) -= val;
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forces at the current point (forces/damping.c, configforce.c, lineardamper.c + tapemeasure.c)
//   order 1: Fo        order 2: also Fq, Fv, Fu
// ---------------------------------------------------------------------------------------------
template <class Sys, class Ws>
TREPB_HD void forces_eval(const Sys& sys, Ws& ws, int order, bool zero_tables = true) {
    using Real = typename Ws::Real;
    const int nq = sys.NQ(), nd = sys.ND();
    TREPB_UNROLL_SYS for (int j = 0; j < nd; ++j) ws.Fo(j) = 0.0;
    if (order >= 2 && zero_tables) {
        TREPB_UNROLL_SYS for (int j = 0; j < nd; ++j) {
            TREPB_UNROLL_SYS for (int i = 0; i < nq; ++i) { ws.Fq(j, i) = 0.0; ws.Fv(j, i) = 0.0; }
            TREPB_UNROLL_SYS for (int u = 0; u < sys.NU(); ++u) ws.Fu(j, u) = 0.0;
        }
    }
    TREPB_UNROLL_SYS
    for (int fo = 0; fo < sys.NFORCE(); ++fo) {
        const int kind = sys.force_kind(fo);
        if (kind == F_DAMPING) {
            const int off = sys.force_i(fo, 0);
            TREPB_UNROLL_SYS
            for (int j = 0; j < nd; ++j) {
                const double c = sys.dpool(off + j);
                ws.Fo(j) += -c * ws.dq(j);
                if (order >= 2) ws.Fv(j, j) += -c;
            }
        } else if (kind == F_CONFIG) {
            const int c = sys.force_i(fo, 0), u = sys.force_i(fo, 1);
            if (c < nd) {
                ws.Fo(c) += ws.u1(u);
                if (order >= 2) ws.Fu(c, u) += 1.0;
            }
        } else if (kind == F_BODY_WRENCH || kind == F_HYBRID_WRENCH || kind == F_SPATIAL_WRENCH) {
            // wrench on frame F (forces/bodywrench.c, hybridwrench.c, spatialwrench.c):
            //   f_j = J_j . w ,  J_j = unhat(g^-1 g_dq_j)                  body   (R^T dp_j, R^T a_j)
            //                        = unhat(g_dq_j g^-1)                  spatial (dp_j - a_j x p, a_j)
            //                        = (dp_j, angular part of the spatial) hybrid (dp_j, a_j)
            // with dp_j = d p_F / d q_j and a_j the world axis of a revolute joint (0: prismatic).
            // i = frame, ipool offset of the six input indices (-1: constant), dpool offset of the constants
            const int F = sys.force_i(fo, 0), io = sys.force_i(fo, 1), dof = sys.force_i(fo, 2);
            Real wr[6], RF[9], pF[3];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) {
                const int in = sys.ipool(io + k);
                if (in >= 0) wr[k] = ws.u1(in); else wr[k] = sys.dpool(dof + k);
            }
            frame_rot(sys, ws, F, RF);
            frame_pos(sys, ws, F, pF);
            TREPB_UNROLL_SYS
            for (int j = 0; j < nd; ++j) {
                if (F == 0 || !sys.dep(F, j)) continue;
                Real dp[3], aj[3], J[6];
                bool rotj;
                int fj;
                dpoint(sys, ws, F, j, dp);
                joint_axis_w(sys, ws, j, aj, &rotj, &fj);
                if (!rotj) { aj[0] = aj[1] = aj[2] = 0.0; }
                if (kind == F_HYBRID_WRENCH) {
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { J[k] = dp[k]; J[3 + k] = aj[k]; }
                } else if (kind == F_SPATIAL_WRENCH) {
                    Real t[3];
                    cross3(aj, pF, t);
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { J[k] = dp[k] - t[k]; J[3 + k] = aj[k]; }
                } else {
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) {
                        J[k] = RF[k] * dp[0] + RF[3 + k] * dp[1] + RF[6 + k] * dp[2];
                        J[3 + k] = RF[k] * aj[0] + RF[3 + k] * aj[1] + RF[6 + k] * aj[2];
                    }
                }
                ws.Fo(j) += dot6(J, wr);
                if (order >= 2) {
                    TREPB_UNROLL for (int k = 0; k < 6; ++k) {
                        const int in = sys.ipool(io + k);
                        if (in >= 0) ws.Fu(j, in) += J[k];
                    }
                    TREPB_UNROLL_SYS
                    for (int i = 0; i < nq; ++i) {
                        if (!sys.dep(F, i)) continue;
                        Real ddp[3], ai[3], da[3], dJ[6];
                        bool roti;
                        int fi;
                        ddpoint(sys, ws, F, j, i, ddp);
                        joint_axis_w(sys, ws, i, ai, &roti, &fi);
                        if (!roti) { ai[0] = ai[1] = ai[2] = 0.0; }
                        // d a_j / d q_i = a_i x a_j when joint i is above joint j
                        if (i != j && sys.dep(fj, i)) cross3(ai, aj, da);
                        else { da[0] = da[1] = da[2] = 0.0; }
                        if (kind == F_HYBRID_WRENCH) {
                            TREPB_UNROLL for (int k = 0; k < 3; ++k) { dJ[k] = ddp[k]; dJ[3 + k] = da[k]; }
                        } else if (kind == F_SPATIAL_WRENCH) {
                            Real t[3], u[3], dpi[3];
                            dpoint(sys, ws, F, i, dpi);
                            cross3(da, pF, t);
                            cross3(aj, dpi, u);
                            TREPB_UNROLL for (int k = 0; k < 3; ++k) { dJ[k] = ddp[k] - t[k] - u[k]; dJ[3 + k] = da[k]; }
                        } else {
                            // d (R^T x) / d q_i = R^T (dx/dq_i - a_i x x)
                            Real t[3], u[3];
                            cross3(ai, dp, t);
                            cross3(ai, aj, u);
                            TREPB_UNROLL for (int k = 0; k < 3; ++k) { t[k] = ddp[k] - t[k]; u[k] = da[k] - u[k]; }
                            TREPB_UNROLL for (int k = 0; k < 3; ++k) {
                                dJ[k] = RF[k] * t[0] + RF[3 + k] * t[1] + RF[6 + k] * t[2];
                                dJ[3 + k] = RF[k] * u[0] + RF[3 + k] * u[1] + RF[6 + k] * u[2];
                            }
                        }
                        ws.Fq(j, i) += dot6(dJ, wr);
                    }
                }
            }
        } else if (kind == F_LINEAR_DAMPER) {
            // single-segment tape measure between two frames (forces/lineardamper.py:34)
            const int off = sys.force_i(fo, 0);
            const int A = sys.ipool(off), B = sys.ipool(off + 1);
            const double cdamp = sys.force_d(fo, 0);
            Real v[3];
            pair_first(sys, ws, A, B, v);
            const Real x = sqrt_(dot3(v, v));
            Real vel = 0.0;
            // dx_j only where exactly one end depends on q_j (tapemeasure.py:97-110)
            TREPB_UNROLL_SYS
            for (int j = 0; j < nq; ++j) {
                Real dx = 0.0;
                if (sys.dep(A, j) != sys.dep(B, j))
                    dx = 1.0 / x * (v[0] * ws.dv(j, 0) + v[1] * ws.dv(j, 1) + v[2] * ws.dv(j, 2));
                ws.dxs(j) = dx;
                vel += dx * ws.dq(j);
            }
            TREPB_UNROLL_SYS
            for (int j = 0; j < nd; ++j) {
                if (sys.dep(A, j) == sys.dep(B, j)) continue;
                ws.Fo(j) += -cdamp * vel * ws.dxs(j);
            }
            if (order >= 2) {
                // ddx(j,i) = TapeMeasure_length_dqdq(q_j, q_i) ; vel_dq(i) = sum_k ddx(k,i) dq_k
                TREPB_UNROLL_SYS
                for (int i = 0; i < nq; ++i) {
                    if (sys.dep(A, i) == sys.dep(B, i)) continue;
                    Real veldq = 0.0;
                    TREPB_UNROLL_SYS
                    for (int k = 0; k < nq; ++k) {
                        Real ddx = 0.0;
                        if (sys.dep(A, k) != sys.dep(B, k)) {
                            Real ddv[3];
                            pair_second(sys, ws, A, B, k, i, ddv);
                            const Real dkdi = ws.dv(k, 0) * ws.dv(i, 0) + ws.dv(k, 1) * ws.dv(i, 1) + ws.dv(k, 2) * ws.dv(i, 2);
                            Real t = ws.dxs(k) * ws.dxs(i) - dkdi - dot3(v, ddv);
                            ddx = -1.0 / x * t;
                        }
                        veldq += ddx * ws.dq(k);
                        if (k < nd) ws.Fq(k, i) += -cdamp * vel * ddx;  // second term of f_dq
                    }
                    TREPB_UNROLL_SYS
                    for (int j = 0; j < nd; ++j) {
                        if (sys.dep(A, j) == sys.dep(B, j)) continue;
                        ws.Fq(j, i) += -cdamp * veldq * ws.dxs(j);
                        ws.Fv(j, i) += -cdamp * ws.dxs(i) * ws.dxs(j);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Crout LU with implicit scaling, exactly the reference's pivoting rules
// (trep/_trep/math-code.c:337-432) so that pivot choices - and therefore rounding - agree.
// A(i,j) accessor object, piv/scales arrays in the workspace.  Returns false if singular.
// ---------------------------------------------------------------------------------------------
// For a compile-time system (Sys::kStatic) every array index must be a loop counter so that the
// matrix lives in registers after full unrolling: the row swap and the permuted gather become
// predicated selects.  Arithmetic and pivot rule are the same in both flavours.
// rd(j) receives 1 / U(j,j); the divisions by the pivots (and those of lu_solve) go through div_r.
template <class Sys, class MatAcc, class PivAcc, class ScaleAcc, class RdAcc>
TREPB_HD bool lu_decomp(MatAcc A, int n, PivAcc piv, ScaleAcc scales, RdAcc rd, double tol) {
    TREPB_UNROLL_SYS
    for (int i = 0; i < n; ++i) {
        double s = -1.0;
        TREPB_UNROLL_SYS
        for (int j = 0; j < n; ++j) {
            const double a = fabs(A(i, j));
            if (a > s) s = a;
        }
        scales(i) = 1.0 / s;
        piv(i) = (double)i;
    }
    TREPB_UNROLL_SYS
    for (int j = 0; j < n; ++j) {
        double pv = -1.0;
        int pi = 0;
        if constexpr (Sys::kStatic) {
            // every loop runs over the full compile-time range with the triangle as a guard: a bound that
            // depends on an outer counter is unknown when the (inner-first) unroller reaches the loop,
            // which then leaves a rolled remainder loop behind and the matrix in local memory
            TREPB_UNROLL_SYS
            for (int i = 0; i < n; ++i) {
                if (i >= j) continue;
                double a = A(i, j);
                TREPB_UNROLL_SYS
                for (int k = 0; k < n; ++k) if (k < i) a -= A(i, k) * A(k, j);
                A(i, j) = a;
            }
            TREPB_UNROLL_SYS
            for (int i = 0; i < n; ++i) {
                if (i < j) continue;
                double a = A(i, j);
                TREPB_UNROLL_SYS
                for (int k = 0; k < n; ++k) if (k < j) a -= A(i, k) * A(k, j);
                A(i, j) = a;
                const double t = fabs(a * scales(i));
                if (t > pv) { pv = t; pi = i; }
            }
        } else {
            // Same inner products in the same order, four rows at a time: each element of column
            // j is fetched once per four rows (the matrix lives in the strided global workspace,
            // the fetches are what bounds this loop).
            constexpr int kR = 4;
            for (int i0 = 0; i0 < j; i0 += kR) {
                double a[kR];
                TREPB_UNROLL for (int g = 0; g < kR; ++g) a[g] = (i0 + g < j) ? A(i0 + g, j) : 0.0;
                // rows above i0 are final; within the block row i0+g also needs rows i0..i0+g-1
                for (int k = 0; k < i0; ++k) {
                    const double b = A(k, j);
                    TREPB_UNROLL for (int g = 0; g < kR; ++g) if (i0 + g < j) a[g] -= A(i0 + g, k) * b;
                }
                TREPB_UNROLL
                for (int g = 0; g < kR; ++g) {
                    if (i0 + g < j) {
                        TREPB_UNROLL
                        for (int h = 0; h < g; ++h) a[g] -= A(i0 + g, i0 + h) * a[h];
                        A(i0 + g, j) = a[g];
                    }
                }
            }
            for (int i0 = j; i0 < n; i0 += kR) {
                double a[kR];
                TREPB_UNROLL for (int g = 0; g < kR; ++g) a[g] = (i0 + g < n) ? A(i0 + g, j) : 0.0;
                for (int k = 0; k < j; ++k) {
                    const double b = A(k, j);
                    TREPB_UNROLL for (int g = 0; g < kR; ++g) if (i0 + g < n) a[g] -= A(i0 + g, k) * b;
                }
                TREPB_UNROLL
                for (int g = 0; g < kR; ++g) {
                    if (i0 + g < n) {
                        A(i0 + g, j) = a[g];
                        const double t = fabs(a[g] * scales(i0 + g));
                        if (t > pv) { pv = t; pi = i0 + g; }
                    }
                }
            }
        }
        if (pv <= tol) return false;
        if (pi != j) {
            if constexpr (Sys::kStatic) {
                // selects with unconditional loads/stores: a branch on (pi == i) would let the
                // optimiser substitute the dynamic `pi` for the loop counter and index memory
                TREPB_UNROLL_SYS
                for (int i = 0; i < n; ++i) {
                    if (i <= j) continue;
                    const bool sw = (pi == i);
                    const double pj = piv(j), pi_ = piv(i);
                    piv(j) = sw ? pi_ : pj;
                    piv(i) = sw ? pj : pi_;
                    TREPB_UNROLL_SYS
                    for (int k = 0; k < n; ++k) {
                        const double aj = A(j, k), ai = A(i, k);
                        A(j, k) = sw ? ai : aj;
                        A(i, k) = sw ? aj : ai;
                    }
                    const double sj = scales(j), si = scales(i);
                    scales(i) = sw ? sj : si;
                }
            } else {
                const double ti = piv(j); piv(j) = piv(pi); piv(pi) = ti;
                for (int k = 0; k < n; ++k) { const double t = A(j, k); A(j, k) = A(pi, k); A(pi, k) = t; }
                scales(pi) = scales(j);
            }
        }
        const double d = A(j, j), r = 1.0 / d;
        rd(j) = r;
        if constexpr (Sys::kStatic) {
            TREPB_UNROLL_SYS
            for (int i = 0; i < n; ++i) if (i > j) A(i, j) = div_r(A(i, j), d, r);
        } else {
            for (int i = j + 1; i < n; ++i) A(i, j) = div_r(A(i, j), d, r);
        }
    }
    return true;
}
// the reciprocal-diagonal accessor of a factorization that did not keep them: plain divisions
struct NoRd {
    TREPB_HD double operator()(int) const { return 0.0; }
};
// solves in place: b <- A^-1 b  (math-code.c:434-461); x is scratch of length n
template <class Sys, class MatAcc, class PivAcc, class BAcc, class XAcc, class RdAcc = NoRd>
TREPB_HD void lu_solve(MatAcc A, int n, PivAcc piv, BAcc b, XAcc x, RdAcc rd = RdAcc()) {
    TREPB_UNROLL_SYS
    for (int i = 0; i < n; ++i) {
        double t;
        if constexpr (Sys::kStatic) {
            t = 0.0;
            const int src = (int)piv(i);
            TREPB_UNROLL_SYS
            for (int k = 0; k < n; ++k) {
                const double bk = b(k);
                t = (src == k) ? bk : t;
            }
        } else {
            t = b((int)piv(i));
        }
        if constexpr (Sys::kStatic) {
            TREPB_UNROLL_SYS
            for (int j = 0; j < n; ++j) if (j < i) t -= A(i, j) * x(j);
        } else {
            for (int j = 0; j < i; ++j) t -= A(i, j) * x(j);
        }
        x(i) = t;
    }
    TREPB_UNROLL_SYS
    for (int i = n - 1; i >= 0; --i) {
        double t = x(i);
        if constexpr (Sys::kStatic) {
            TREPB_UNROLL_SYS
            for (int j = 0; j < n; ++j) if (j > i) t -= A(i, j) * x(j);
        } else {
            for (int j = i + 1; j < n; ++j) t -= A(i, j) * x(j);
        }
        if constexpr (std::is_same<RdAcc, NoRd>::value) t = t / A(i, i);
        else t = div_r(t, A(i, i), rd(i));
        x(i) = t;
    }
    TREPB_UNROLL_SYS
    for (int i = 0; i < n; ++i) b(i) = x(i);
}

// accessor adaptors over workspace arrays
#define TREPB_ACC2(NAME, FIELD)                                         \
    template <class Ws> struct NAME {                                   \
        Ws* w;                                                          \
        TREPB_HD double& operator()(int i, int j) const { return w->FIELD(i, j); } \
    };
#define TREPB_ACC1(NAME, FIELD)                                         \
    template <class Ws> struct NAME {                                   \
        Ws* w;                                                          \
        TREPB_HD double& operator()(int i) const { return w->FIELD(i); } \
    };
TREPB_ACC2(AccDf, Df) TREPB_ACC2(AccM2, M2) TREPB_ACC2(AccPJ, PJ)
TREPB_ACC1(AccPiv, piv) TREPB_ACC1(AccLus, lus) TREPB_ACC1(AccLux, lux) TREPB_ACC1(AccFr, fr)
TREPB_ACC1(AccM2p, M2p) TREPB_ACC1(AccPJp, PJp) TREPB_ACC1(AccTnd, tnd) TREPB_ACC1(AccTnc, tnc)
TREPB_ACC1(AccCol, col)
TREPB_ACC1(AccDfr, Dfr) TREPB_ACC1(AccM2r, M2r) TREPB_ACC1(AccPJr, PJr)
template <class Ws> struct AccTdcCol {  // column k of Tdc (nd x nc)
    Ws* w; int k;
    TREPB_HD double& operator()(int i) const { return w->Tdc(i, k); }
};

// ---------------------------------------------------------------------------------------------
// evaluation-point setters (midpointvi.c:401-457)
// ---------------------------------------------------------------------------------------------
// The time step and its reciprocal.  x / dt is needed for every configuration at every evaluation point;
// the double-precision division is an emulated instruction sequence with special-case branches, so it is
// replaced by  q = x r,  q' = q + r (x - dt q)  with r = RN(1/dt) computed once and two fused
// multiply-adds: by Markstein's theorem q' is the correctly rounded quotient (no mismatch against x / dt in
// 10^9 random trials per dt, tests/test_host_math.py::test_division_by_the_time_step_is_exact), i.e. the
// same bits as the division for finite operands.
struct Dt {
    double dt, rdt;
    TREPB_HD Dt(double d) : dt(d), rdt(1.0 / d) {}
};
TREPB_HD double div_dt(double x, const Dt& d) {
    const double q = x * d.rdt;
    const double e = fma(-q, d.dt, x);
    return fma(e, d.rdt, q);
}
template <class T> TREPB_HD T div_dt(const T& x, const Dt& d) { return x * d.rdt; }   // (hyper-)dual numbers: x * (1 / dt), as their operator/ does
// x / d with r = RN(1/d) at hand (the pivots of an LU factorization divide many numbers): the same correction
TREPB_HD double div_r(double x, double d, double r) {
    const double q = x * r;
    const double e = fma(-q, d, x);
    return fma(e, r, q);
}

// sqrt(x) > tol  <=>  x > sqrt_threshold(tol)  for x >= 0 (sqrt is correctly rounded and monotonic): the
// largest T with sqrt(T) <= tol.  The convergence test of every Newton iteration (midpointvi.c:672-689)
// then needs no square root.  Host only (the launchers pass it in the kernel parameters).
inline double sqrt_threshold(double tol) {
    if (!(tol >= 0.0)) return tol < 0.0 ? -1.0 : tol;      // negative: never converged; NaN: comparison false
    double T = tol * tol;
    for (int i = 0; i < 8 && sqrt(T) > tol; ++i) T = nextafter(T, 0.0);
    for (int i = 0; i < 8; ++i) {
        const double up = nextafter(T, INFINITY);
        if (!(sqrt(up) <= tol)) break;
        T = up;
    }
    return T;
}

template <class Sys, class Ws>
TREPB_HD void set_point(const Sys& sys, Ws& ws, int which, const Dt& dt) {  // 0 = midpoint, 1 = q1, 2 = q2
    using Real = typename Ws::Real;
    TREPB_UNROLL_SYS
    for (int i = 0; i < sys.NQ(); ++i) {
        const Real a = ws.q1(i), b = ws.q2(i);
        ws.qe(i) = which == 0 ? 0.5 * (b + a) : (which == 1 ? a : b);
        ws.dq(i) = div_dt(b - a, dt);
    }
}

// Lagrangian + forces at the midpoint.  order 1: residual terms.  order 2: Jacobian terms too.
template <class Sys, class Ws>
TREPB_HD void eval_mid(const Sys& sys, Ws& ws, const Dt& dt, int order) {
    using Real = typename Ws::Real;
    set_point(sys, ws, 0, dt);
    pass1(sys, ws, true, sys.pairs_mid());   // world poses at the midpoint only for springs / dampers / wrenches
    pass2(sys, ws, order);
    add_potentials(sys, ws, order);
    forces_eval(sys, ws, order);
}

// Second-order tables at the SAME midpoint as the preceding eval_mid(order 1): pass 1 (the
// sin/cos, velocities and W vectors) is still valid in the workspace, only the evaluation
// configuration (and, when pair elements are evaluated at the midpoint, the world poses) may
// have been moved to q1/q2 by the constraint passes in between.  The reference re-walks its
// whole cache here (set_midpoint clears every cache flag, midpointvi.c:437-457).
template <class Sys, class Ws>
TREPB_HD void eval_mid_again(const Sys& sys, Ws& ws, const Dt& dt) {
    if (sys.NC() > 0) {
        set_point(sys, ws, 0, dt);
        if (sys.pairs_mid()) pass1(sys, ws, false, true);
    }
    pass2(sys, ws, 2);
    add_potentials(sys, ws, 2);
    forces_eval(sys, ws, 2);
}

// ---------------------------------------------------------------------------------------------
// MidpointVI_solve_DEL (midpointvi.c:691-747).  Inputs in ws: q1, p1, u1, q2 (dyn part = Newton
// start, kin part = k2), lam (start).  Outputs in ws: q2, lam, p2.  Returns the iteration count
// (>= 0) or ST_NOT_CONVERGED / ST_SINGULAR.
// ---------------------------------------------------------------------------------------------
// tolT: sqrt_threshold(tol), computed on the host by the launchers (the convergence test then needs no
// square root); NaN: test sqrt(|f|^2) > tol as written in the reference
template <class Sys, class Ws>
TREPB_HD int solve_del(const Sys& sys, Ws& ws, double t1, double t2, double tol, int max_it, double tolT = NAN) {
    const int nd = sys.ND(), nc = sys.NC(), nr = nd + nc;
    const double dt = t2 - t1;
    const Dt dtt(dt);
    const bool thr = tolT == tolT;
    int iterations = 0;
    TREPB_TICK_INIT
    if (nc > 0) {
        set_point(sys, ws, 1, dtt);
        pass1(sys, ws, false, true);
        constraints_eval(sys, ws, 2, 1);  // Dh1 = Dh(q1)
    }
    TREPB_TICK(0);
    for (;;) {
        // ---- calc_f (midpointvi.c:533-565)
        eval_mid(sys, ws, dtt, 1);
        TREPB_TICK(1);
        TREPB_UNROLL_SYS
        for (int j = 0; j < nd; ++j) {
            double f = ws.p1(j) + (0.5 * dt * ws.Lq(j) - ws.Lv(j)) + dt * ws.Fo(j);
            TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c) f -= ws.Dh1(c, j) * ws.lam(c);
            ws.fr(j) = f;
        }
        if (nc > 0) {
            set_point(sys, ws, 2, dtt);
            pass1(sys, ws, false, true);
            constraints_eval(sys, ws, 1, 2);
            TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c) ws.fr(nd + c) = ws.hc(c);
        }
        TREPB_TICK(2);
        // ---- DEL_solved (midpointvi.c:672-689)
        double nrm = 0.0;
        TREPB_UNROLL_SYS for (int j = 0; j < nd; ++j) nrm += ws.fr(j) * ws.fr(j);
        bool solved = thr ? !(nrm > tolT) : !(sqrt(nrm) > tol);      // the same test, see sqrt_threshold
        TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c)
            if (fabs(ws.fr(nd + c)) > sys.con_d(c, 1)) solved = false;
        if (solved) break;
        if (iterations > max_it) return ST_NOT_CONVERGED;

        // ---- Jacobian (midpointvi.c:577-670): Df_11 entry by entry (the reference fills it with a
        // symmetric/antisymmetric split, :603-627, i.e. the same terms in a different order)
        eval_mid_again(sys, ws, dtt);
        TREPB_TICK(3);
        TREPB_UNROLL_SYS
        for (int k = 0; k < nd; ++k) {
            TREPB_UNROLL_SYS
            for (int i = 0; i < nd; ++i)
                ws.Df(k, i) = (0.25 * dt * ws.Lqq(k, i) - dtt.rdt * ws.Lvv(k, i)) + 0.5 * ws.Lvq(i, k)
                            - 0.5 * ws.Lvq(k, i) + (0.5 * dt * ws.Fq(k, i) + ws.Fv(k, i));
        }
        if (nc > 0) {
            set_point(sys, ws, 2, dtt);
            pass1(sys, ws, false, true);
            constraints_eval(sys, ws, 2, 2);  // Dh2 = Dh(q2)
            TREPB_UNROLL_SYS for (int i = 0; i < nd; ++i)
                TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c) {
                    ws.Df(i, nd + c) = -ws.Dh1(c, i);
                    ws.Df(nd + c, i) = ws.Dh2(c, i);
                }
            TREPB_UNROLL_SYS for (int a = 0; a < nc; ++a)
                TREPB_UNROLL_SYS for (int b = 0; b < nc; ++b) ws.Df(nd + a, nd + b) = 0.0;
        }
        TREPB_TICK(4);
        if (nr == 1) {
            // 1x1: the reference's LU reduces to a scaled-pivot test and one division
            const double a = ws.Df(0, 0);
            if (!(fabs(a) > 0.0)) return ST_SINGULAR;
            ws.fr(0) = ws.fr(0) / a;
        } else {
            if (!lu_decomp<Sys>(AccDf<Ws>{&ws}, nr, AccPiv<Ws>{&ws}, AccLus<Ws>{&ws}, AccDfr<Ws>{&ws}, 1e-20)) return ST_SINGULAR;
            lu_solve<Sys>(AccDf<Ws>{&ws}, nr, AccPiv<Ws>{&ws}, AccFr<Ws>{&ws}, AccLux<Ws>{&ws}, AccDfr<Ws>{&ws});
        }
        TREPB_TICK(5);
        TREPB_UNROLL_SYS for (int k = 0; k < nd; ++k) ws.q2(k) -= ws.fr(k);
        TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c) ws.lam(c) -= ws.fr(nd + c);
        iterations++;
    }
    // p2 = D2L2 at the midpoint of the converged (q1, q2): the order-1 tables of the last
    // residual evaluation are exactly that point (midpointvi.c:742-743, 491-504).
    TREPB_UNROLL_SYS for (int j = 0; j < nd; ++j) ws.p2(j) = 0.5 * dt * ws.Lq(j) + ws.Lv(j);
    return iterations;
}

// MidpointVI_calc_f alone (midpointvi.c:533-575): residual of the DEL equation at the (q1, q2, p1, u1, lam) in
// the workspace -> fr[0:nd] = p1 + D1L2 + fm2 - Dh(q1)^T lam, fr[nd:nd+nc] = h(q2).
template <class Sys, class Ws>
TREPB_HD void calc_f(const Sys& sys, Ws& ws, double t1, double t2) {
    const int nd = sys.ND(), nc = sys.NC();
    const double dt = t2 - t1;
    const Dt dtt(dt);
    if (nc > 0) {
        set_point(sys, ws, 1, dtt);
        pass1(sys, ws, false, true);
        constraints_eval(sys, ws, 2, 1);  // Dh1 = Dh(q1)
    }
    eval_mid(sys, ws, dtt, 1);
    TREPB_UNROLL_SYS
    for (int j = 0; j < nd; ++j) {
        double f = ws.p1(j) + (0.5 * dt * ws.Lq(j) - ws.Lv(j)) + dt * ws.Fo(j);
        TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c) f -= ws.Dh1(c, j) * ws.lam(c);
        ws.fr(j) = f;
    }
    if (nc > 0) {
        set_point(sys, ws, 2, dtt);
        pass1(sys, ws, false, true);
        constraints_eval(sys, ws, 1, 2);
        TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c) ws.fr(nd + c) = ws.hc(c);
    }
}

// calc_p2 alone (midpointvi.c:2702-2708) for initialize_from_configs
template <class Sys, class Ws>
TREPB_HD void calc_p2(const Sys& sys, Ws& ws, double t1, double t2) {
    const double dt = t2 - t1;
    eval_mid(sys, ws, Dt(dt), 1);
    TREPB_UNROLL_SYS for (int j = 0; j < sys.ND(); ++j) ws.p2(j) = 0.5 * dt * ws.Lq(j) + ws.Lv(j);
}

// ---------------------------------------------------------------------------------------------
// First derivatives (midpointvi.c:749-1120).  Requires a solved step in ws (q1,q2,lam,u1).
// Output arrays use the reference's raw storage layout [wrt][out] (midpointvi.py:59-70);
// any pointer may be null.  A / B follow DSystem.fdx / fdu (dsystem.py:284-317), row-major.
// ---------------------------------------------------------------------------------------------
struct Deriv1Out {
    double *q2_dq1, *q2_dp1, *q2_du1, *q2_dk2;
    double *p2_dq1, *p2_dp1, *p2_du1, *p2_dk2;
    double *l1_dq1, *l1_dp1, *l1_du1, *l1_dk2;
    double *A, *B;
    long es;  // element stride of every output array (1 = contiguous per instance)
};

template <class Sys, class Ws>
TREPB_HD int deriv1(const Sys& sys, Ws& ws, double t1, double t2, const Deriv1Out& o, bool mid_valid = false) {
    const int nd = sys.ND(), nk = sys.NK(), nq = nd + nk, nc = sys.NC(), nu = sys.NU();
    const double dt = t2 - t1;
    const Dt dtt(dt);
    const int nX = 2 * nq, nU = nu + nk;
    TREPB_TICK_INIT
    // ---- constraint derivatives at q1 and q2
    if (nc > 0) {
        set_point(sys, ws, 1, dtt);
        pass1(sys, ws, false, true);
        constraints_eval(sys, ws, 2 | 4, 1);  // Dh1, DDhl = sum_c lam_c DDh_c(q1)
        set_point(sys, ws, 2, dtt);
        pass1(sys, ws, false, true);
        constraints_eval(sys, ws, 2, 2);      // Dh2
    }
    TREPB_TICK(8);
    // ---- calc_deriv1_cache at the midpoint (midpointvi.c:749-861).
    // mid_valid: the workspace still holds pass 1 of the converged midpoint (solve_del just ran)
    if (mid_valid) {
        if (nc == 0) set_point(sys, ws, 0, dtt);
        eval_mid_again(sys, ws, dtt);
    } else {
        eval_mid(sys, ws, dtt, 2);
    }
    TREPB_TICK(9);
    // Entry by entry; the reference fills the same four tables with a
    // symmetric i<j pass plus a separate L_ddqdq pass (midpointvi.c:771-858), which only changes
    // the order of the six-term sums.
    TREPB_UNROLL_SYS
    for (int a = 0; a < nq; ++a) {
        TREPB_UNROLL_SYS
        for (int b = 0; b < nd; ++b) {
            const double qq = 0.25 * dt * ws.Lqq(a, b), vv = dtt.rdt * ws.Lvv(a, b);
            const double vab = 0.5 * ws.Lvq(a, b), vba = 0.5 * ws.Lvq(b, a);
            const double fq = 0.5 * dt * ws.Fq(b, a), fv = ws.Fv(b, a);
            ws.T11(a, b) = (qq + vv) - vab - vba + (fq - fv);
            ws.T21(a, b) = (qq - vv) + vab - vba + (fq + fv);
            ws.T12(a, b) = (qq - vv) - vab + vba;
            ws.T22(a, b) = (qq + vv) + vab + vba;
        }
    }
    TREPB_UNROLL_SYS for (int u = 0; u < nu; ++u)
        TREPB_UNROLL_SYS for (int j = 0; j < nd; ++j) ws.T3(u, j) = dt * ws.Fu(j, u);

    TREPB_TICK(10);
    // ---- calc_M2 (midpointvi.c:891-908)
    TREPB_UNROLL_SYS for (int a = 0; a < nd; ++a)
        TREPB_UNROLL_SYS for (int b = 0; b < nd; ++b) ws.M2(a, b) = ws.T21(b, a);
    if (!lu_decomp<Sys>(AccM2<Ws>{&ws}, nd, AccM2p<Ws>{&ws}, AccLus<Ws>{&ws}, AccM2r<Ws>{&ws}, 1e-20)) return ST_SINGULAR;
    // ---- calc_proj_inv (midpointvi.c:910-927): proj = -Dh2_d M2^-1 Dh1^T
    if (nc > 0) {
        TREPB_UNROLL_SYS for (int i = 0; i < nd; ++i)
            TREPB_UNROLL_SYS for (int c = 0; c < nc; ++c) ws.Tdc(i, c) = ws.Dh1(c, i);
        TREPB_UNROLL_SYS
        for (int c = 0; c < nc; ++c)
            lu_solve<Sys>(AccM2<Ws>{&ws}, nd, AccM2p<Ws>{&ws}, AccTdcCol<Ws>{&ws, c}, AccLux<Ws>{&ws}, AccM2r<Ws>{&ws});
        TREPB_UNROLL_SYS
        for (int a = 0; a < nc; ++a)
            TREPB_UNROLL_SYS
            for (int b = 0; b < nc; ++b) {
                double s = 0.0;
                TREPB_UNROLL_SYS
                for (int k = 0; k < nd; ++k) s += ws.Dh2(a, k) * ws.Tdc(k, b);
                ws.PJ(a, b) = -s;
            }
        if (!lu_decomp<Sys>(AccPJ<Ws>{&ws}, nc, AccPJp<Ws>{&ws}, AccLus<Ws>{&ws}, AccPJr<Ws>{&ws}, 1e-20)) return ST_SINGULAR;
    }

    TREPB_TICK(11);
    // ---- calc_deriv1 (midpointvi.c:929-1098): one right-hand side per wrt-variable
    // kind 0: q1_i   1: p1_i   2: u1_i   3: k2_i
    const long es = o.es;
    if constexpr (Sys::kStatic) {
        TREPB_UNROLL_SYS
        for (int kindv = 0; kindv < 4; ++kindv) {
            const int count = kindv == 0 ? nq : (kindv == 1 ? nd : (kindv == 2 ? nu : nk));
            // full compile-time range + guard (see TREPB_FOR_FROM): count depends on the outer counter
            const int cmax = nq > nu ? nq : nu;
            TREPB_UNROLL_SYS
            for (int i = 0; i < cmax; ++i) {
                if (i >= count) continue;
                // explicit part c
                TREPB_UNROLL_SYS
                for (int j = 0; j < nd; ++j) {
                    double c;
                    if (kindv == 0) {
                        c = -ws.T11(i, j);
                        if (nc > 0) c += ws.DDhl(i, j);
                    } else if (kindv == 1) {
                        c = (j == i) ? -1.0 : 0.0;
                    } else if (kindv == 2) {
                        c = -ws.T3(i, j);
                    } else {
                        c = -ws.T21(nd + i, j);
                    }
                    ws.tnd(j) = c;
                    ws.col(j) = c;
                }
                if (nc > 0) {
                    lu_solve<Sys>(AccM2<Ws>{&ws}, nd, AccM2p<Ws>{&ws}, AccTnd<Ws>{&ws}, AccLux<Ws>{&ws}, AccM2r<Ws>{&ws});
                    TREPB_UNROLL_SYS
                    for (int c = 0; c < nc; ++c) {
                        double s = 0.0;
                        TREPB_UNROLL_SYS
                        for (int j = 0; j < nd; ++j) s += ws.Dh2(c, j) * ws.tnd(j);
                        if (kindv == 3) s += ws.Dh2(c, nd + i);
                        ws.tnc(c) = s;
                    }
                    lu_solve<Sys>(AccPJ<Ws>{&ws}, nc, AccPJp<Ws>{&ws}, AccTnc<Ws>{&ws}, AccLux<Ws>{&ws}, AccPJr<Ws>{&ws});
                    TREPB_UNROLL_SYS
                    for (int j = 0; j < nd; ++j) {
                        double s = ws.col(j);
                        TREPB_UNROLL_SYS
                        for (int c = 0; c < nc; ++c) s += ws.Dh1(c, j) * ws.tnc(c);
                        ws.col(j) = s;
                    }
                }
                lu_solve<Sys>(AccM2<Ws>{&ws}, nd, AccM2p<Ws>{&ws}, AccCol<Ws>{&ws}, AccLux<Ws>{&ws}, AccM2r<Ws>{&ws});
                // p2 derivative row
                double* q2o = kindv == 0 ? o.q2_dq1 : (kindv == 1 ? o.q2_dp1 : (kindv == 2 ? o.q2_du1 : o.q2_dk2));
                double* p2o = kindv == 0 ? o.p2_dq1 : (kindv == 1 ? o.p2_dp1 : (kindv == 2 ? o.p2_du1 : o.p2_dk2));
                double* l1o = kindv == 0 ? o.l1_dq1 : (kindv == 1 ? o.l1_dp1 : (kindv == 2 ? o.l1_du1 : o.l1_dk2));
                TREPB_UNROLL_SYS
                for (int j = 0; j < nd; ++j) {
                    double pv = kindv == 0 ? ws.T12(i, j) : (kindv == 3 ? ws.T22(nd + i, j) : 0.0);
                    TREPB_UNROLL_SYS
                    for (int k = 0; k < nd; ++k) pv += ws.T22(k, j) * ws.col(k);
                    const double qv = ws.col(j);
                    if (q2o) q2o[(long)(i * nd + j) * es] = qv;
                    if (p2o) p2o[(long)(i * nd + j) * es] = pv;
                    // A / B blocks
                    if (kindv == 0) {
                        if (o.A) { o.A[(long)(j * nX + i) * es] = qv; o.A[(long)((nq + j) * nX + i) * es] = pv; }
                    } else if (kindv == 1) {
                        if (o.A) { o.A[(long)(j * nX + nq + i) * es] = qv; o.A[(long)((nq + j) * nX + nq + i) * es] = pv; }
                    } else if (kindv == 2) {
                        if (o.B) { o.B[(long)(j * nU + i) * es] = qv; o.B[(long)((nq + j) * nU + i) * es] = pv; }
                    } else {
                        if (o.B) { o.B[(long)(j * nU + nu + i) * es] = qv; o.B[(long)((nq + j) * nU + nu + i) * es] = pv; }
                    }
                }
                if (l1o) {
                    TREPB_UNROLL_SYS
                    for (int c = 0; c < nc; ++c) l1o[(long)(i * nc + c) * es] = ws.tnc(c);
                }
            }
        }
    } else {
        // Table-driven flavour: the right-hand sides are processed kG at a time so that every
        // element of the M2 factors and of D2D2L2 is fetched once per group, not once per
        // column (these fetches are what the kernel is bound by for a system of this size), and
        // the second M2 back-substitution of the reference (q2_s = M2^-1 (c + Dh1^T l_s)) is
        // replaced by  M2^-1 c + (M2^-1 Dh1^T) l_s  with the nd x nc matrix calc_proj_inv already
        // formed.  Same mathematics; rounding differs at the 1e-16 level.
        constexpr int kG = 4;
        const int nrhs = nq + nd + nu + nk;
        for (int r0 = 0; r0 < nrhs; r0 += kG) {
            const int ng = nrhs - r0 < kG ? nrhs - r0 : kG;
            int kv[kG], iv[kG];
            TREPB_UNROLL
            for (int g = 0; g < kG; ++g) {
                const int r = r0 + (g < ng ? g : 0);
                kv[g] = r < nq ? 0 : (r < nq + nd ? 1 : (r < nq + nd + nu ? 2 : 3));
                iv[g] = r < nq ? r : (r < nq + nd ? r - nq : (r < nq + nd + nu ? r - nq - nd : r - nq - nd - nu));
            }
            for (int j = 0; j < nd; ++j) {
                TREPB_UNROLL
                for (int g = 0; g < kG; ++g) {
                    const int kindv = kv[g], i = iv[g];
                    double c;
                    if (kindv == 0) {
                        c = -ws.T11(i, j);
                        if (nc > 0) c += ws.DDhl(i, j);
                    } else if (kindv == 1) {
                        c = (j == i) ? -1.0 : 0.0;
                    } else if (kindv == 2) {
                        c = -ws.T3(i, j);
                    } else {
                        c = -ws.T21(nd + i, j);
                    }
                    ws.tnd(j, g) = c;
                }
            }
            // tnd <- M2^-1 c   (forward / back substitution, kG columns at once)
            for (int i = 0; i < nd; ++i) {
                double t[kG];
                const int src = (int)ws.M2p(i);
                TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] = ws.tnd(src, g);
                for (int j = 0; j < i; ++j) {
                    const double a = ws.M2(i, j);
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] -= a * ws.lux(j, g);
                }
                TREPB_UNROLL for (int g = 0; g < kG; ++g) ws.lux(i, g) = t[g];
            }
            for (int i = nd - 1; i >= 0; --i) {
                double t[kG];
                TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] = ws.lux(i, g);
                for (int j = i + 1; j < nd; ++j) {
                    const double a = ws.M2(i, j);
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] -= a * ws.lux(j, g);
                }
                const double dg = ws.M2(i, i), rg = ws.M2r(i);
                TREPB_UNROLL for (int g = 0; g < kG; ++g) { t[g] = div_r(t[g], dg, rg); ws.lux(i, g) = t[g]; }
            }
            for (int i = 0; i < nd; ++i) {
                TREPB_UNROLL for (int g = 0; g < kG; ++g) ws.tnd(i, g) = ws.lux(i, g);
            }
            if (nc > 0) {
                for (int c = 0; c < nc; ++c) {
                    double sacc[kG];
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) sacc[g] = 0.0;
                    for (int j = 0; j < nd; ++j) {
                        const double a = ws.Dh2(c, j);
                        TREPB_UNROLL for (int g = 0; g < kG; ++g) sacc[g] += a * ws.tnd(j, g);
                    }
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) {
                        if (kv[g] == 3) sacc[g] += ws.Dh2(c, nd + iv[g]);
                        ws.tnc(c, g) = sacc[g];
                    }
                }
                // tnc <- proj^-1 tnc
                for (int i = 0; i < nc; ++i) {
                    double t[kG];
                    const int src = (int)ws.PJp(i);
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] = ws.tnc(src, g);
                    for (int j = 0; j < i; ++j) {
                        const double a = ws.PJ(i, j);
                        TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] -= a * ws.lux(j, g);
                    }
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) ws.lux(i, g) = t[g];
                }
                for (int i = nc - 1; i >= 0; --i) {
                    double t[kG];
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] = ws.lux(i, g);
                    for (int j = i + 1; j < nc; ++j) {
                        const double a = ws.PJ(i, j);
                        TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] -= a * ws.lux(j, g);
                    }
                    const double dg = ws.PJ(i, i), rg = ws.PJr(i);
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) { t[g] = div_r(t[g], dg, rg); ws.lux(i, g) = t[g]; }
                }
                for (int i = 0; i < nc; ++i) {
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) ws.tnc(i, g) = ws.lux(i, g);
                }
                // col = M2^-1 c + (M2^-1 Dh1^T) lambda_s
                for (int j = 0; j < nd; ++j) {
                    double t[kG];
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] = ws.tnd(j, g);
                    for (int c = 0; c < nc; ++c) {
                        const double a = ws.Tdc(j, c);
                        TREPB_UNROLL for (int g = 0; g < kG; ++g) t[g] += a * ws.tnc(c, g);
                    }
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) ws.col(j, g) = t[g];
                }
            } else {
                for (int j = 0; j < nd; ++j) {
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) ws.col(j, g) = ws.tnd(j, g);
                }
            }
            // outputs: q2_s = col, p2_s = explicit + D2D2L2^T col
            for (int j = 0; j < nd; ++j) {
                double pv[kG];
                TREPB_UNROLL for (int g = 0; g < kG; ++g)
                    pv[g] = kv[g] == 0 ? ws.T12(iv[g], j) : (kv[g] == 3 ? ws.T22(nd + iv[g], j) : 0.0);
                for (int k = 0; k < nd; ++k) {
                    const double a = ws.T22(k, j);
                    TREPB_UNROLL for (int g = 0; g < kG; ++g) pv[g] += a * ws.col(k, g);
                }
                TREPB_UNROLL
                for (int g = 0; g < kG; ++g) {
                    if (g >= ng) continue;
                    const int kindv = kv[g], i = iv[g];
                    const double qv = ws.col(j, g);
                    double* q2o = kindv == 0 ? o.q2_dq1 : (kindv == 1 ? o.q2_dp1 : (kindv == 2 ? o.q2_du1 : o.q2_dk2));
                    double* p2o = kindv == 0 ? o.p2_dq1 : (kindv == 1 ? o.p2_dp1 : (kindv == 2 ? o.p2_du1 : o.p2_dk2));
                    if (q2o) q2o[(long)(i * nd + j) * es] = qv;
                    if (p2o) p2o[(long)(i * nd + j) * es] = pv[g];
                    if (kindv == 0) {
                        if (o.A) { o.A[(long)(j * nX + i) * es] = qv; o.A[(long)((nq + j) * nX + i) * es] = pv[g]; }
                    } else if (kindv == 1) {
                        if (o.A) { o.A[(long)(j * nX + nq + i) * es] = qv; o.A[(long)((nq + j) * nX + nq + i) * es] = pv[g]; }
                    } else if (kindv == 2) {
                        if (o.B) { o.B[(long)(j * nU + i) * es] = qv; o.B[(long)((nq + j) * nU + i) * es] = pv[g]; }
                    } else {
                        if (o.B) { o.B[(long)(j * nU + nu + i) * es] = qv; o.B[(long)((nq + j) * nU + nu + i) * es] = pv[g]; }
                    }
                }
            }
            TREPB_UNROLL
            for (int g = 0; g < kG; ++g) {
                if (g >= ng) continue;
                const int kindv = kv[g], i = iv[g];
                double* l1o = kindv == 0 ? o.l1_dq1 : (kindv == 1 ? o.l1_dp1 : (kindv == 2 ? o.l1_du1 : o.l1_dk2));
                if (l1o) {
                    for (int c = 0; c < nc; ++c) l1o[(long)(i * nc + c) * es] = ws.tnc(c, g);
                }
            }
        }
    }
    TREPB_TICK(12);
    // constant blocks of A and B
    if (o.A) {
        TREPB_UNROLL_SYS
        for (int r = 0; r < nX; ++r)
            TREPB_UNROLL_SYS
            for (int c = 0; c < nX; ++c) {
                const bool dyn_row = r < nd || (r >= nq && r < nq + nd);
                if (dyn_row && c < nq + nd) continue;  // written above
                double v = 0.0;
                if (r >= nq + nd && c >= nd && c < nq && (r - nq - nd) == (c - nd)) v = -dtt.rdt;
                o.A[(long)(r * nX + c) * es] = v;
            }
    }
    if (o.B) {
        TREPB_UNROLL_SYS
        for (int r = 0; r < nX; ++r)
            TREPB_UNROLL_SYS
            for (int c = 0; c < nU; ++c) {
                const bool dyn_row = r < nd || (r >= nq && r < nq + nd);
                if (dyn_row) continue;
                double v = 0.0;
                if (r >= nd && r < nq && c >= nu && (r - nd) == (c - nu)) v = 1.0;
                if (r >= nq + nd && c >= nu && (r - nq - nd) == (c - nu)) v = dtt.rdt;
                o.B[(long)(r * nU + c) * es] = v;
            }
    }
    TREPB_TICK(13);
    return ST_OK;
}

}  // namespace trepb
