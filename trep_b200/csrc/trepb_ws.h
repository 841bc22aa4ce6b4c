// Per-instance workspace for the MidpointVI math (trepb_math.cuh).
//
// The same array list is instantiated two ways:
//   WsStatic<Sys, Real>  - fixed-size member arrays (compile-time system): after unrolling, every
//                          index is a constant and the arrays live in registers.
//   WsStridedT<Real>     - one slab of global memory shared by the whole grid, element e of thread t
//                          at base[e * stride + t]  (structure-of-arrays across threads, so a warp
//                          executing the same line touches 32 consecutive elements = coalesced).
// Real is double for the step / first-derivative kernels, a hyper-dual number (trepb_hd.h) for the
// per-pair second-derivative kernel, which only evaluates the first-order residual path and therefore
// lays out only the arrays flagged 1, and a dual number (trepb_d2jac.cuh) for the directional
// derivative of the Jacobian tables (arrays flagged 1 or 2).
#pragma once
#include "trepb_sys.h"

namespace trepb {

// X(name, rows, cols, ad) ; sizes use NF NQ ND NU NC NR (= ND+NC) ; ad: 1 residual path, 2 second-order tables, 0 solver only
#define TREPB_WS_ARRAYS(X)                                                                     \
    /* integrator state */                                                                     \
    X(q1, NQ, 1, 1) X(q2, NQ, 1, 1) X(p1, ND, 1, 0) X(p2, ND, 1, 1) X(u1, NU, 1, 1) X(lam, NC, 1, 1) \
    X(qe, NQ, 1, 1) X(dq, NQ, 1, 1) X(vk, NQ, 1, 0)                                            \
    /* frame pass 1 */                                                                         \
    X(cs, NF, 2, 1) X(gf, NF, 3, 1) X(V, NF, 6, 1) X(W, NF, 6, 1) X(Rw, NF, 9, 1) X(pw, NF, 3, 1) \
    /* frame pass 2: composite inertia (m, h, I sym6) and momentum */                          \
    X(Im, NF, 1, 1) X(Ih, NF, 3, 1) X(II, NF, 6, 2) X(mu, NF, 6, 1)                            \
    /* Lagrangian tables */                                                                    \
    X(Lq, NQ, 1, 1) X(Lv, NQ, 1, 1) X(Lqq, NQ, NQ, 2) X(Lvq, NQ, NQ, 2) X(Lvv, NQ, NQ, 2)      \
    /* forces */                                                                               \
    X(Fo, ND, 1, 1) X(Fq, ND, NQ, 2) X(Fv, ND, NQ, 2) X(Fu, ND, NU, 2)                         \
    /* constraints */                                                                          \
    X(hc, NC, 1, 1) X(Dh1, NC, NQ, 1) X(Dh2, NC, NQ, 2) X(DDhl, NQ, NQ, 2)                     \
    /* point-pair scratch: d(pA-pB)/dq_j for every config */                                   \
    X(dv, NQ, 3, 1) X(dxs, NQ, 1, 1)                                                           \
    /* Newton */                                                                               \
    X(fr, NR, 1, 1) X(Df, NR, NR, 0) X(piv, NR, 1, 0) X(lus, NR, 1, 0) X(lux, NR, 4, 0) X(Dfr, NR, 1, 0) \
    /* first-derivative tables and solves (tnd/tnc/col/lux: 4 right-hand sides at a time) */                                                   \
    X(T11, NQ, ND, 0) X(T21, NQ, ND, 0) X(T12, NQ, ND, 0) X(T22, NQ, ND, 0) X(T3, NU, ND, 0)   \
    X(M2, ND, ND, 0) X(M2p, ND, 1, 0) X(PJ, NC, NC, 0) X(PJp, NC, 1, 0) X(tnd, ND, 4, 0) X(tnc, NC, 4, 0) \
    /* reciprocal diagonals of the three factorizations (divisions by them: div_r) */          \
    X(M2r, ND, 1, 0) X(PJr, NC, 1, 0)                                                          \
    X(Tdc, ND, NC, 0) X(col, ND, 4, 0)

template <class Sys, class RealT = double>
struct WsStatic {
    using Real = RealT;
    static constexpr int NF = Sys::kNF, NQ = Sys::kNQ, ND = Sys::kND, NU = Sys::kNU,
                         NC = Sys::kNC, NR = Sys::kND + Sys::kNC;
#define X(name, rows, cols, ad)                               \
    Real name##_[((rows) * (cols)) > 0 ? (rows) * (cols) : 1]; \
    TREPB_HD Real& name(int i, int j = 0) { return name##_[i * (cols) + j]; }
    TREPB_WS_ARRAYS(X)
#undef X
};

template <class RealT>
struct WsStridedT {
    using Real = RealT;
    Real* base;
    unsigned stride;   // threads of the grid; (elements per thread) x stride < 2^32 (the launchers cap the grid: ws_grid_cap)
    int NF, NQ, ND, NU, NC, NR;
#define X(name, rows, cols, ad) int o_##name; int ld_##name;
    TREPB_WS_ARRAYS(X)
#undef X
#define X(name, rows, cols, ad) \
    TREPB_HD Real& name(int i, int j = 0) { return base[(size_t)((unsigned)(o_##name + i * ld_##name + j) * stride)]; }
    TREPB_WS_ARRAYS(X)
#undef X
    // returns number of elements per thread.  level 0: every array; 1: only the arrays of the residual
    // path (flag 1); 2: those plus the second-order tables (flags 1 and 2, trepb_d2jac.cuh)
    TREPB_HD int layout(int nf, int nd, int nk, int nu, int nc, int level = 0) {
        NF = nf; ND = nd; NQ = nd + nk; NU = nu; NC = nc; NR = nd + nc;
        int off = 0;
#define X(name, rows, cols, ad) \
        o_##name = off; ld_##name = (cols); if (level == 0 || ((ad) != 0 && (ad) <= level)) off += (rows) * (cols);
        TREPB_WS_ARRAYS(X)
#undef X
        return off;
    }
};
using WsStrided = WsStridedT<double>;

// The strided accessors index with 32-bit arithmetic (one IMAD and one IMAD.WIDE per access instead of a 64-bit
// multiply: a quarter of the table-driven kernels' instructions is this address arithmetic): the largest grid,
// in CTAs of `block` threads, for which (elements per thread) x (threads) stays below 2^32
inline long long ws_grid_cap(long long elems_per_thread, int block) {
    const long long threads = 0xffffffffLL / (elems_per_thread > 0 ? elems_per_thread : 1);
    const long long g = threads / block;
    return g < 1 ? 1 : g;
}

}  // namespace trepb
