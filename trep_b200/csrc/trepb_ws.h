// Per-instance workspace for the MidpointVI math (trepb_math.cuh).
//
// The same array list is instantiated two ways:
//   WsStatic<Sys>  - fixed-size member arrays (compile-time system): after unrolling, every index
//                    is a constant and the arrays live in registers.
//   WsStrided      - one slab of global memory shared by the whole batch, element e of instance t
//                    at base[e * stride + t]  (structure-of-arrays across instances, so a warp
//                    executing the same line touches 32 consecutive doubles = coalesced).
#pragma once
#include "trepb_sys.h"

namespace trepb {

// X(name, rows, cols) ; sizes use NF NQ ND NU NC NR (= ND+NC)
#define TREPB_WS_ARRAYS(X)                                                                     \
    /* integrator state */                                                                     \
    X(q1, NQ, 1) X(q2, NQ, 1) X(p1, ND, 1) X(p2, ND, 1) X(u1, NU, 1) X(lam, NC, 1)             \
    X(qe, NQ, 1) X(dq, NQ, 1)                                                                  \
    /* frame pass 1 */                                                                         \
    X(cs, NF, 2) X(gf, NF, 3) X(V, NF, 6) X(W, NF, 6) X(Rw, NF, 9) X(pw, NF, 3)                \
    /* frame pass 2: composite inertia (m, h, I sym6) and momentum */                          \
    X(Im, NF, 1) X(Ih, NF, 3) X(II, NF, 6) X(mu, NF, 6)                                        \
    /* Lagrangian tables */                                                                    \
    X(Lq, NQ, 1) X(Lv, NQ, 1) X(Lqq, NQ, NQ) X(Lvq, NQ, NQ) X(Lvv, NQ, NQ)                     \
    /* forces */                                                                               \
    X(Fo, ND, 1) X(Fq, ND, NQ) X(Fv, ND, NQ) X(Fu, ND, NU)                                     \
    /* constraints */                                                                          \
    X(hc, NC, 1) X(Dh1, NC, NQ) X(Dh2, NC, NQ) X(DDhl, NQ, NQ)                                 \
    /* point-pair scratch: d(pA-pB)/dq_j for every config */                                   \
    X(dv, NQ, 3) X(dxs, NQ, 1)                                                                 \
    /* Newton */                                                                               \
    X(fr, NR, 1) X(Df, NR, NR) X(piv, NR, 1) X(lus, NR, 1) X(lux, NR, 1)                       \
    /* first-derivative tables and solves */                                                   \
    X(T11, NQ, ND) X(T21, NQ, ND) X(T12, NQ, ND) X(T22, NQ, ND) X(T3, NU, ND)                  \
    X(M2, ND, ND) X(M2p, ND, 1) X(PJ, NC, NC) X(PJp, NC, 1) X(tnd, ND, 1) X(tnc, NC, 1)        \
    X(Tdc, ND, NC) X(col, ND, 1)

template <class Sys>
struct WsStatic {
    static constexpr int NF = Sys::kNF, NQ = Sys::kNQ, ND = Sys::kND, NU = Sys::kNU,
                         NC = Sys::kNC, NR = Sys::kND + Sys::kNC;
#define X(name, rows, cols)                                     \
    double name##_[((rows) * (cols)) > 0 ? (rows) * (cols) : 1]; \
    TREPB_HD double& name(int i, int j = 0) { return name##_[i * (cols) + j]; }
    TREPB_WS_ARRAYS(X)
#undef X
};

struct WsStrided {
    double* base;
    long stride;
    int NF, NQ, ND, NU, NC, NR;
#define X(name, rows, cols) int o_##name; int ld_##name;
    TREPB_WS_ARRAYS(X)
#undef X
#define X(name, rows, cols) \
    TREPB_HD double& name(int i, int j = 0) { return base[(long)(o_##name + i * ld_##name + j) * stride]; }
    TREPB_WS_ARRAYS(X)
#undef X
    // returns number of doubles per instance
    TREPB_HD int layout(int nf, int nd, int nk, int nu, int nc) {
        NF = nf; ND = nd; NQ = nd + nk; NU = nu; NC = nc; NR = nd + nc;
        int off = 0;
#define X(name, rows, cols) o_##name = off; ld_##name = (cols); off += (rows) * (cols);
        TREPB_WS_ARRAYS(X)
#undef X
        return off;
    }
};

}  // namespace trepb
