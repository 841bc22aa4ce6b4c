// Team-cooperative MidpointVI math: ONE instance is worked on by a team of lanes (a warp on the
// device, a single "lane" on the host for CPU checks) with the whole per-instance workspace in
// shared memory.  Same quantities as trepb_math.cuh / the reference
// (trep/_trep/midpointvi.c:391-1120 on top of system.c:129-742, frame.c:839-2081 and
// math-code.c:337-461), evaluated on the link-level tables of trepb_coop_sys.h in WORLD (spatial)
// coordinates so that every phase is either a short level-by-level sweep of the link tree or a
// flat loop the lanes of the team share:
//
//   pose sweep (root->leaf, one step per tree level)
//       T_l = T_parent Xc_l J_l(q_l)            pose of link l
//       s_l = (o_l x a_l, a_l) | (a_l, 0)       joint twist about the world origin (revolute | prismatic)
//       V_l = V_parent + s_l dq_l               spatial velocity
//   per link      I_l^w (m, h = m c, Ibar about the world origin),  mu_l = I_l^w V_l
//   up sweep      composite Ic_l, mu_l = sums over the subtree (plain additions in world coordinates)
//   per link      W_l = [V_parent, s_l] (= d vb / d q_l),  L_ddq = s.mu,  L_dq = W.mu + g.P,
//                 H = Ic s,  G = Ic W - ad*_s mu,  P = m v_s + w_s x h
//   per chain pair (i above-or-equal j)
//                 L_ddqddq(i,j) = s_i.H_j   L_ddqdq(i,j) = s_i.G_j   L_ddqdq(j,i) = W_i.H_j
//                 L_dqdq(i,j) = W_i.G_j + a_i.(P_j x g)
//   constraints   world points p = o_l + R_l r ; dp/dq_j = a_j x (p - o_j) | a_j ;
//                 d2p/dq_i dq_j = a_up x dp/dq_lo
//   linear algebra  right-looking LU with the reference's implicit-scaling pivot rule
//                 (math-code.c:337-432), lanes over columns; right-hand sides ride along as extra
//                 columns or are solved one column per lane (in registers when the sizes are
//                 compile-time constants).
//
// Two flavours from one source: D = RtDims (sizes read from the tables at run time: any system) and
// D = CtDims<...> (sizes are template constants: loops unroll, workspace offsets fold to immediates,
// the right-hand-side columns live in registers).  The link tables themselves are always run-time
// data, so one CtDims instantiation serves every system of that shape.
//
// Nothing here is a transcription of frame.c: there are no per-frame derivative caches at all.
#pragma once
#include <math.h>
#include "trepb_coop_sys.h"
#include "trepb_math.cuh"   // cross3 / dot3 / inertia_apply / Deriv1Out / Status

namespace trepb {

// ---------------------------------------------------------------------------------------------
// teams
// ---------------------------------------------------------------------------------------------
struct HostTeam {
    static constexpr int kSize = 1, kWarps = 1;
    static constexpr bool kWarp = false;
    using Sub = HostTeam;
    TREPB_HD int lane() const { return 0; }
    TREPB_HD int warp() const { return 0; }
    TREPB_HD Sub sub() const { return Sub(); }
    TREPB_HD void sync() const {}
    TREPB_HD void argmax(double&, int&) const {}
};
#if defined(__CUDACC__)
struct WarpTeam {
    static constexpr int kSize = 32, kWarps = 1;
    static constexpr bool kWarp = true;
    using Sub = WarpTeam;
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int warp() const { return 0; }
    __device__ __forceinline__ Sub sub() const { return Sub(); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    // largest value wins, ties go to the lowest index; every lane gets the result
    __device__ __forceinline__ void argmax(double& v, int& i) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
    }
};
// Two warps on one instance: the flat phases (pairs, matrix entries, right-hand-side columns, constraint
// lists) spread over 64 lanes and the midpoint / q2 pose sweeps run on one warp each, which doubles the
// warps an SM has to hide latencies with while the shared-memory footprint per instance stays the same.
// Synchronisation is a named barrier per team (bar.sync id, 64); the factorizations, which use warp
// collectives, run on the team's first warp (Sub).
struct PairTeam {
    static constexpr int kSize = 64, kWarps = 2;
    static constexpr bool kWarp = false;
    using Sub = WarpTeam;
    int bar;   // named barrier of this team (1 + team index in the CTA; 0 is __syncthreads)
    __device__ __forceinline__ int lane() const { return threadIdx.x & 63; }
    __device__ __forceinline__ int warp() const { return (threadIdx.x >> 5) & 1; }
    __device__ __forceinline__ Sub sub() const { return Sub(); }
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory"); }
    __device__ __forceinline__ void argmax(double&, int&) const {}
};
#endif

// ---------------------------------------------------------------------------------------------
// sizes: run-time or compile-time
// ---------------------------------------------------------------------------------------------
struct RtDims {
    static constexpr bool kStatic = false, kExtras = true, kExt = false, kExtS = false;
    using Solve = RtDims;
    static constexpr int ND = 1, NK = 0, NU = 0, NC = 0, NL = 1, NP = 0, NPAIRS = 1, NLEVELS = 1, NDC = 0, NQC = 0;
};

// workspace layout (offsets in doubles from the instance's base)
struct CoopLayout {
    int nls, ldf, ldm, ldp, ldy, ldd;
    int q1, q2, qe, dq, p1, p2, u1, lam, vk;
    int cs, R, p, V, comp, pts;      // link region
    int M2, T22;                     // first-derivative factors; alias the link region
    int Lq, Lv, VV, QQ, UP, DN;
    int XS;                          // sum of the LinearSpring Hessians d2V / dq dq at the midpoint [nqs][nqs]
    int FD, FX, FQ, FV;              // LinearDampers: force [nd], dx scratch [nq], f_dq / f_ddq blocks [nqf][nqf]
    int KS;                          // config stiffness at the evaluation point [nq]: ConfigSpring k + spline springs' -d2V/dq2
    int FUW;                         // wrenches: d f / d u at the evaluation point [nd][nu] (on top of the constant ConfigForce part)
    int SC;                          // scratch of the spring / damper Hessians: dA, dB, dx of one potential or force [7 nq]
    int Dh1, Dh2, hc;
    int N;                           // Newton augmented matrix [nr][ldf]; aliases the link region
    int Y;                           // DDh.lambda block / right-hand sides [nd][ldy] (first-derivative kernels only)
    int Z, PJ;
    int fr, scl, rdM, rdP;
    int ints;                        // pivM[nr] swpM[nr] pivP[nc] swpP[nc] flag  (int32)
    int total;
    int ext, xtotal;                 // blocks Y, UP, DN in the team's external slab of xtotal doubles (make(), ext)

    TREPB_HD static constexpr int odd(int n) { return n | 1; }
    // stat: the compile-time-size flavour (right-hand-side columns in registers: no Z block, the
    // DDh.lambda block compacted to the ndc x nqc configs some constraint depends on).
    // Aliases: cs (sin/cos of the sweep) sits in the tail of comp, which the sweeps never touch;
    // fr / scl reuse Lq / Lv, which are dead once the residual and p2 have been formed; M2 / T22 and
    // the Newton matrix N reuse the link region (poses, velocities, composite inertias are dead from
    // the moment the pair tables of an iteration exist until the next iteration recomputes them; the
    // converged iteration leaves before it would assemble N, so deriv1 finds the midpoint data intact).
    // solve_only: the layout of the kernels that never call deriv1 (step, project, p2 / f): no DDh.lambda
    // block, no projection factors, two pair arrays instead of four (dyn_second) - 18.0 instead of 27.0 KB
    // per marionette instance, 12 instead of 8 instances per SM.
    // ext: the layout of the flavours that keep the blocks read with one gather per entry or written once per
    // evaluation - the pair arrays VV, QQ (and UP, DN of deriv1), the constraint Jacobians Dh1, Dh2 and deriv1's
    // DDh.lambda block Y - in a per-team slab of global memory (L2-resident: 14.1 KB per marionette instance, 5.4 KB
    // for the solve-only kernels) instead of shared memory, and the projection factors over the midpoint / velocity
    // vectors that are dead by then: 12.3 instead of 26.4 KB of shared memory per marionette instance (12.2 instead
    // of 18.0 solve-only), 16 instances per SM instead of 8 (12).  Those offsets then index the external slab.
    TREPB_HD static constexpr CoopLayout make(int nd, int nk, int nu, int nc, int nl, int np, int npairs,
                                              bool stat = false, int ndc = 0, int nqc = 0, bool solve_only = false,
                                              int nqs = 0, int nqf = 0, int nns = 0, int nw = 0, bool ext = false) {
        CoopLayout L{};
        const bool sx = ext && !solve_only;
        int xo = 0;
        L.ext = ext ? 1 : 0;
        const int nq = nd + nk, nr = nd + nc;
        L.nls = nl;
        L.ldf = odd(nr + 1);
        L.ldm = odd(nd + nc);
        L.ldp = odd(nc > 0 ? nc : 1);
        L.ldy = nq + nu;
        L.ldd = stat ? nqc : L.ldy;
        int o = 0;
        L.q1 = o; o += nq; L.q2 = o; o += nq; L.qe = o; o += nq; L.dq = o; o += nq;
        L.p1 = o; o += nd; L.p2 = o; o += nd; L.u1 = o; o += nu; L.lam = o; o += nc; L.vk = o; o += nk;
        const int link0 = o;
        L.R = o; o += 9 * nl; L.p = o; o += 3 * nl; L.V = o; o += 6 * nl;
        L.comp = o; o += 16 * nl; L.cs = L.comp + 14 * nl; L.pts = o; o += 3 * np;
        const int nN = nr * L.ldf;
        const int need = nd * L.ldm + nd * nd > nN ? nd * L.ldm + nd * nd : nN;
        if (o - link0 < need) o = link0 + need;
        L.M2 = link0; L.T22 = link0 + nd * L.ldm;
        L.N = link0;
        const int nlq = nq > nr ? nq : nr;
        L.Lq = o; o += nlq; L.Lv = o; o += nlq;
        L.fr = L.Lq; L.scl = L.Lv;
        if (ext) { L.VV = xo; xo += npairs; L.QQ = xo; xo += npairs; }
        else { L.VV = o; o += npairs; L.QQ = o; o += npairs; }
        if (sx) { L.UP = xo; xo += npairs; L.DN = xo; xo += npairs; }
        else { L.UP = o; o += solve_only ? 0 : npairs; L.DN = o; o += solve_only ? 0 : npairs; }
        if (ext) { L.Dh1 = xo; xo += nc * nd; L.Dh2 = xo; xo += nc * nq; }
        else { L.Dh1 = o; o += nc * nd; L.Dh2 = o; o += nc * nq; }
        L.hc = o; o += nc;
        const int nY = stat ? ndc * nqc : nd * L.ldy;
        if (sx) { L.Y = xo; xo += nY; }
        else { L.Y = o; o += solve_only ? 0 : nY; }
        L.Z = o; o += (stat || solve_only) ? 0 : nc * L.ldy;
        if (sx && nc * L.ldp <= 2 * nq) L.PJ = L.qe;   // qe, dq: last read by the pose sweeps
        else { L.PJ = o; o += solve_only ? 0 : nc * L.ldp; }
        L.rdM = o; o += nr; L.rdP = o; o += solve_only ? 0 : nc;
        L.ints = o; o += (2 * nr + 2 * nc + 2) / 2 + 1;   // + the first-warp flag
        L.total = (o + 1) & ~1;
        L.xtotal = (xo + 1) & ~1;
        L.append_extras(nd, nq, nu, nqs, nqf, nns, nw);
        return L;
    }
    // The blocks of the plugin kinds that only some systems have (LinearSpring, LinearDamper, NonlinearConfigSpring,
    // wrenches) sit behind everything else, so that the offsets of a compile-time-size flavour stay constants and
    // only this tail is laid out from run-time counts (CtDims<..., 1>).
    TREPB_HD constexpr void append_extras(int nd, int nq, int nu, int nqs, int nqf, int nns, int nw) {
        int o = total;
        XS = o; o += nqs * nqs;
        KS = o; o += nns > 0 ? nq : 0;
        FUW = o; o += nw > 0 ? nd * nu : 0;
        FD = o; o += nqf > 0 ? nd : 0; FX = o; o += nqf > 0 ? nq : 0; FQ = o; o += nqf * nqf; FV = o; o += nqf * nqf;
        SC = o; o += (nqs > 0 || nqf > 0) ? 7 * nq : 0;
        total = (o + 1) & ~1;
    }
    TREPB_HD void set(const CoopSys& s, bool stat = false, bool solve_only = false, bool ext_ = false) {
        *this = make(s.nd, s.nk, s.nu, s.nc, s.nl, s.np, s.npairs, stat, s.ndc, s.nqc, solve_only, s.nqs, s.nqf, s.nns, s.nw, ext_);
    }
};

// EX_ = 1: the shape of a system that also has LinearSprings / LinearDampers / spline springs / wrenches: their
// counts stay run-time data (tail of the layout), everything else is compile-time as usual
template <int ND_, int NK_, int NU_, int NC_, int NL_, int NP_, int NPAIRS_, int NLEVELS_, int NDC_, int NQC_, int EX_ = 0>
struct CtDims {
    static constexpr bool kStatic = true;
    static constexpr bool kExtras = EX_ != 0;
    static constexpr bool kExt = false, kExtS = false;
    using Solve = CtDims;   // the flavour of the kernels that never call deriv1
    static constexpr int ND = ND_, NK = NK_, NU = NU_, NC = NC_, NL = NL_, NP = NP_, NPAIRS = NPAIRS_, NLEVELS = NLEVELS_,
                         NDC = NDC_, NQC = NQC_;
    TREPB_HD static constexpr CoopLayout layout(bool solve_only = false) {
        return CoopLayout::make(ND, NK, NU, NC, NL, NP, NPAIRS, true, NDC, NQC, solve_only);
    }
    TREPB_HD static bool matches(const CoopSys& s) {
        return s.nd == ND && s.nk == NK && s.nu == NU && s.nc == NC && s.nl == NL && s.np == NP && s.npairs == NPAIRS &&
               s.nlevels == NLEVELS && s.ndc == NDC && s.nqc == NQC && (kExtras || (s.ns == 0 && s.nfd == 0 && s.nns == 0 && s.nw == 0));
    }
};

// The same shape with the external-slab layouts (CoopLayout::make, ext): ExtSolveDims for the kernels that never
// call deriv1 (pair arrays VV, QQ and constraint Jacobians in the slab: kExtS), ExtDims for the first-derivative
// kernel (those plus the DDh.lambda block and the pair arrays UP, DN only deriv1 uses: kExt)
template <class Dims>
struct ExtSolveDims : Dims {
    static constexpr bool kExt = false, kExtS = true;
    using Solve = ExtSolveDims;
    TREPB_HD static constexpr CoopLayout layout(bool = true) {
        return CoopLayout::make(Dims::ND, Dims::NK, Dims::NU, Dims::NC, Dims::NL, Dims::NP, Dims::NPAIRS, true, Dims::NDC,
                                Dims::NQC, true, 0, 0, 0, 0, true);
    }
};
template <class Dims>
struct ExtDims : Dims {
    static constexpr bool kExt = true, kExtS = true;
    using Solve = ExtSolveDims<Dims>;
    TREPB_HD static constexpr CoopLayout layout(bool solve_only = false) {
        return CoopLayout::make(Dims::ND, Dims::NK, Dims::NU, Dims::NC, Dims::NL, Dims::NP, Dims::NPAIRS, true, Dims::NDC,
                                Dims::NQC, solve_only, 0, 0, 0, 0, true);
    }
};

TREPB_HD double sel3(const double* v, int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }

// ---------------------------------------------------------------------------------------------
// cooperative dense helpers
// ---------------------------------------------------------------------------------------------
// In-place LU of the leading n x n block of A (row-major, leading dimension ld) with the
// reference's pivot rule (math-code.c:337-432: implicit row scaling, first strict maximum, whole
// row swap, scales[pivot] = scales[j]); the same row operations are applied to nx extra columns
// (right-hand sides), which therefore leave forward-eliminated.
// Storage on return: U on and above the diagonal, rd[k] = 1/U(k,k), and BELOW the diagonal the
// unscaled multipliers m(i,k) = L(i,k) U(k,k)  (the update uses m(i,k) * (rd[k] a(k,j)), one
// multiplication per column instead of a division per element and no separate scaling phase).
// piv[] is the composed permutation of LU_decomp (x[i] = b[piv[i]]), swp[k] the row exchanged with k
// at step k.  Returns false when the scaled pivot is <= tol (the reference's "singular" ValueError).
template <class Team>
TREPB_HD bool team_lu(const Team& t, double* A, int ld, int n, int nx, int* piv, int* swp, double* scl,
                      double* rd, double tol) {
    const int lane = t.lane();
    for (int i = lane; i < n; i += Team::kSize) {
        double s = -1.0;
        for (int j = 0; j < n; ++j) {
            const double a = fabs(A[i * ld + j]);
            if (a > s) s = a;
        }
        scl[i] = 1.0 / s;
        piv[i] = i;
    }
    t.sync();
    for (int k = 0; k < n; ++k) {
        double best = -1.0;
        int bi = k;
#if defined(__CUDA_ARCH__)
        if (Team::kWarp && n - k <= 32) {
            // one candidate per lane; two 32-bit max-reductions over the bit pattern of |a| * scale
            // (non-negative doubles order like unsigned integers), lowest lane among the maxima wins
            const int i = k + lane;
            const bool act = i < n;
            double v = act ? fabs(A[i * ld + k] * scl[i]) : 0.0;
            if (!(v == v)) v = 0.0;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
            const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
            const unsigned cand = __ballot_sync(0xffffffffu, act && hi == mh && lo == ml);
            bi = k + __ffs(cand) - 1;
            best = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
        } else
#endif
        {
            for (int i = k + lane; i < n; i += Team::kSize) {
                const double v = fabs(A[i * ld + k] * scl[i]);
                if (v > best) { best = v; bi = i; }
            }
            t.argmax(best, bi);
        }
        if (!(best > tol)) return false;
        t.sync();   // every lane has read scl / column k before rows move
        if (bi != k) {
            for (int j = lane; j < n + nx; j += Team::kSize) {
                const double a = A[k * ld + j];
                A[k * ld + j] = A[bi * ld + j];
                A[bi * ld + j] = a;
            }
            if (lane == 0) {
                const int pk = piv[k];
                piv[k] = piv[bi];
                piv[bi] = pk;
                scl[bi] = scl[k];
            }
        }
        if (lane == 0) swp[k] = bi;
        t.sync();
        const double rdk = 1.0 / A[k * ld + k];
        if (lane == 0) rd[k] = rdk;
        for (int j = k + 1 + lane; j < n + nx; j += Team::kSize) {
            const double akj = A[k * ld + j] * rdk;
            const double* mk = A + (k + 1) * ld + k;
            double* aj = A + (k + 1) * ld + j;
            const int cnt = n - k - 1, head = cnt & 3;
            if (head) {
                // 1-3 leading rows, predicated; the rest goes four rows at a time without guards
                const bool g1 = head > 1, g2 = head > 2;
                const double m0 = mk[0], m1 = g1 ? mk[ld] : 0.0, m2 = g2 ? mk[2 * ld] : 0.0;
                const double a0 = aj[0], a1 = g1 ? aj[ld] : 0.0, a2 = g2 ? aj[2 * ld] : 0.0;
                aj[0] = a0 - m0 * akj;
                if (g1) aj[ld] = a1 - m1 * akj;
                if (g2) aj[2 * ld] = a2 - m2 * akj;
                mk += head * ld; aj += head * ld;
            }
            for (int i = head; i < cnt; i += 4) {
                const double m0 = mk[0], m1 = mk[ld], m2 = mk[2 * ld], m3 = mk[3 * ld];
                const double a0 = aj[0], a1 = aj[ld], a2 = aj[2 * ld], a3 = aj[3 * ld];
                aj[0] = a0 - m0 * akj; aj[ld] = a1 - m1 * akj; aj[2 * ld] = a2 - m2 * akj; aj[3 * ld] = a3 - m3 * akj;
                mk += 4 * ld; aj += 4 * ld;
            }
        }
        t.sync();
    }
    return true;
}

// Back substitution of ONE forward-eliminated column (column index c of A) by the whole team.
template <class Team>
TREPB_HD void team_backsolve_vec(const Team& t, double* A, int ld, int n, int c, const double* rd, double* x) {
    const int lane = t.lane();
    for (int i = n - 1; i >= 0; --i) {
        const double xi = A[i * ld + c] * rd[i];
        for (int r = lane; r < i; r += Team::kSize) A[r * ld + c] -= A[r * ld + i] * xi;
        if (lane == 0) x[i] = xi;
        t.sync();
    }
}

// One column per lane: y (stride ldy) <- U^-1 y for an already forward-eliminated column
TREPB_HD void col_backsolve(const double* A, int ld, int n, const double* rd, double* y, int ldy) {
    for (int i = n - 1; i >= 0; --i) {
        double v = y[i * ldy];
        for (int j = i + 1; j < n; ++j) v -= A[i * ld + j] * y[j * ldy];
        y[i * ldy] = v * rd[i];
    }
}
// One column per lane: y <- (LU)^-1 P y  with team_lu's storage (unscaled multipliers)
TREPB_HD void col_solve(const double* A, int ld, int n, const int* swp, const double* rd, double* y, int ldy) {
    for (int k = 0; k < n; ++k) {
        const int p = swp[k];
        if (p != k) {
            const double a = y[k * ldy];
            y[k * ldy] = y[p * ldy];
            y[p * ldy] = a;
        }
    }
    // forward: yt_i = (b_i - sum_{j<i} m_ij yt_j) rd_i ; backward: x_i = yt_i - rd_i sum_{j>i} U_ij x_j
    for (int i = 0; i < n; ++i) {
        double v = y[i * ldy];
        for (int j = 0; j < i; ++j) v -= A[i * ld + j] * y[j * ldy];
        y[i * ldy] = v * rd[i];
    }
    for (int i = n - 2; i >= 0; --i) {
        double v = 0.0;
        for (int j = i + 1; j < n; ++j) v += A[i * ld + j] * y[j * ldy];
        y[i * ldy] -= rd[i] * v;
    }
}
// the same on a register-resident column (N compile-time): y[] indices are constants after
// unrolling.  Rows are processed four at a time so that four independent accumulation chains are
// in flight (a single chain of dependent FMAs leaves the FP64 pipe idle most of the time).
template <int N, int I0>
TREPB_HD void reg_fwd(const double* A, int ld, const double* rd, double* y) {
    if constexpr (I0 < N) {
        constexpr int R = N - I0 < 4 ? N - I0 : 4;
        double v[R];
        TREPB_UNROLL for (int r = 0; r < R; ++r) v[r] = y[I0 + r];
        TREPB_UNROLL
        for (int j = 0; j < I0; ++j) {
            TREPB_UNROLL for (int r = 0; r < R; ++r) v[r] -= A[(I0 + r) * ld + j] * y[j];
        }
        TREPB_UNROLL
        for (int r = 0; r < R; ++r) {
            TREPB_UNROLL for (int q = 0; q < r; ++q) v[r] -= A[(I0 + r) * ld + I0 + q] * y[I0 + q];
            y[I0 + r] = v[r] * rd[I0 + r];
        }
        reg_fwd<N, I0 + R>(A, ld, rd, y);
    }
}
template <int N, int I1>
TREPB_HD void reg_bwd(const double* A, int ld, const double* rd, double* y) {
    if constexpr (I1 > 0) {
        constexpr int R = I1 < 4 ? I1 : 4;      // rows I1-R .. I1-1
        double v[R];
        TREPB_UNROLL for (int r = 0; r < R; ++r) v[r] = 0.0;
        TREPB_UNROLL
        for (int j = I1; j < N; ++j) {
            TREPB_UNROLL for (int r = 0; r < R; ++r) v[r] += A[(I1 - 1 - r) * ld + j] * y[j];
        }
        TREPB_UNROLL
        for (int r = 0; r < R; ++r) {
            TREPB_UNROLL for (int q = 0; q < r; ++q) v[r] += A[(I1 - 1 - r) * ld + I1 - 1 - q] * y[I1 - 1 - q];
            y[I1 - 1 - r] -= rd[I1 - 1 - r] * v[r];
        }
        reg_bwd<N, I1 - R>(A, ld, rd, y);
    }
}
// forward: yt_i = (b_i - sum_{j<i} m_ij yt_j) rd_i ; backward: x_i = yt_i - rd_i sum_{j>i} U_ij x_j
template <int N>
TREPB_HD void reg_solve(const double* A, int ld, const double* rd, double* y) {
    reg_fwd<N, 0>(A, ld, rd, y);
    reg_bwd<N, N>(A, ld, rd, y);
}

#if defined(__CUDACC__)
// The same factorization as team_lu for compile-time sizes N <= 32, with ONE ROW PER LANE held in
// registers (N + NX doubles): no shared-memory traffic and no __syncwarp inside the factorization.
// Rows are never moved: every lane keeps its row, its implicit scale and the POSITION its row would
// occupy after the reference's row exchanges (math-code.c:337-432); the pivot of step k is the largest
// |a_ik| scale_i over the rows at positions >= k, ties to the lowest position (= the first strict
// maximum of the reference's scan), the row that sat at position k takes the pivot row's old position.
// Step k broadcasts the pivot row by shuffles; arithmetic per element is team_lu's:
// a_ij -= m_ik (rd_k a_kj) with the unscaled multiplier m_ik left in place.
template <int N, int NX>
struct RowLU {
    static constexpr unsigned kFull = 0xffffffffu;
    double a[N + NX];
    double scl, rdp;
    int pos;
    __device__ __forceinline__ void load(const double* A, int ld, int lane) {
        pos = lane < N ? lane : 255;
        const double* row = A + (lane < N ? lane : 0) * ld;
        double s = -1.0;
#pragma unroll
        for (int j = 0; j < N + NX; ++j) {
            a[j] = row[j];
            if (j < N) { const double v = fabs(a[j]); if (v > s) s = v; }
        }
        scl = 1.0 / s;
        rdp = 0.0;
    }
    __device__ __forceinline__ bool factor(double tol) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const bool act = pos >= k && pos < N;
            double v = act ? fabs(a[k] * scl) : 0.0;
            if (!(v == v)) v = 0.0;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
            const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
            const unsigned mh = __reduce_max_sync(kFull, hi);
            const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
            const bool ismax = act && hi == mh && lo == ml;
            const unsigned bpos = __reduce_min_sync(kFull, ismax ? (unsigned)pos : 255u);
            const double best = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
            if (!(best > tol)) return false;
            const bool isp = ismax && (unsigned)pos == bpos;
            const int pl = __ffs(__ballot_sync(kFull, isp)) - 1;
            if (pos == k) pos = (int)bpos;
            if (isp) pos = k;
            const double rdk = 1.0 / __shfl_sync(kFull, a[k], pl);
            if (isp) rdp = rdk;
            // rows already used as pivots (and idle lanes) take a zero multiplier: an unconditional
            // fused multiply-add instead of a select per updated element
            const double m = (pos > k && pos < N) ? a[k] : 0.0;
#pragma unroll
            for (int j = 0; j < N + NX; ++j) {
                if (j <= k) continue;
                const double akj = __shfl_sync(kFull, a[j], pl) * rdk;
                a[j] -= m * akj;
            }
        }
        return true;
    }
    // back substitution of the forward-eliminated column N + c: x[i] = unknown at position i
    __device__ __forceinline__ void backsolve(int c, double* x, int lane) {
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            const int li = __ffs(__ballot_sync(kFull, pos == i)) - 1;
            const double xi = __shfl_sync(kFull, a[N + c] * rdp, li);
            a[N + c] -= (pos < i ? a[i] : 0.0) * xi;
            if (lane == li) x[i] = xi;
        }
    }
    // factors back to shared memory in team_lu's storage (row of position p at row p), rd, composed permutation
    __device__ __forceinline__ void store(double* A, int ld, double* rd, int* piv, int lane) const {
        if (pos < N) {
            double* row = A + pos * ld;
#pragma unroll
            for (int j = 0; j < N + NX; ++j) row[j] = a[j];
            rd[pos] = rdp;
            piv[pos] = lane;
        }
    }
};
#endif

// ---------------------------------------------------------------------------------------------
// the per-instance context
// ---------------------------------------------------------------------------------------------
template <class Team, class D = RtDims>
struct Coop {
    const CoopSys& S;
    const CoopLayout L;
    double* w;      // workspace base of this instance
    double* x;      // the team's external slab (D::kExt flavours: blocks Y, UP, DN), else unused
    Team t;

    TREPB_HD Coop(const CoopSys& s, const CoopLayout& l, double* base, Team team, double* ext = nullptr)
        : S(s), L(l), w(base), x(ext), t(team) {
        use_pose(false);
    }
    // base of a block that the ext layout keeps in the external slab
    TREPB_HD double* xw(int off) const { if constexpr (D::kExt) return x + off; else return w + off; }
    // ... and of one the solve-only ext layout keeps there (VV, QQ, Dh1, Dh2)
    TREPB_HD double* xv(int off) const { if constexpr (D::kExtS) return x + off; else return w + off; }

#define TREPB_DIM(fn, CT, rt) \
    TREPB_HD int fn() const { if constexpr (D::kStatic) return D::CT; else return S.rt; }
    TREPB_DIM(ND, ND, nd) TREPB_DIM(NK, NK, nk) TREPB_DIM(NU, NU, nu) TREPB_DIM(NC, NC, nc)
    TREPB_DIM(NL, NL, nl) TREPB_DIM(NP, NP, np) TREPB_DIM(NPAIRS, NPAIRS, npairs) TREPB_DIM(NLEVELS, NLEVELS, nlevels)
#undef TREPB_DIM
    TREPB_HD int NQ() const { return ND() + NK(); }
    // the plugin kinds only some systems have: flavours instantiated without them (CtDims<..., 0>) drop the code
    TREPB_HD int NS() const { if constexpr (!D::kExtras) return 0; else return S.ns; }
    TREPB_HD int NFD() const { if constexpr (!D::kExtras) return 0; else return S.nfd; }
    TREPB_HD int NNS() const { if constexpr (!D::kExtras) return 0; else return S.nns; }
    TREPB_HD int NW() const { if constexpr (!D::kExtras) return 0; else return S.nw; }
    TREPB_HD int NPF() const { return NFD() + NW(); }     // forces acting through world points: dampers, wrenches
    TREPB_HD double fu_eff(int j, int u) const {
        const double c = S.Fu()[j * NU() + u];
        return NW() > 0 ? c + w[L.FUW + j * NU() + u] : c;
    }
    // stiffness of config i seen by the second-order terms: the ConfigSpring constant, plus -d2V/dq2 of the spline
    // springs at the last evaluated midpoint (dyn_first)
    TREPB_HD double ks_eff(int i) const { return NNS() > 0 ? w[L.KS + i] : S.ks()[i]; }

    TREPB_HD int* ipivM() const { return (int*)(w + L.ints); }
    TREPB_HD int* iswpM() const { return ipivM() + (ND() + NC()); }
    TREPB_HD int* ipivP() const { return iswpM() + (ND() + NC()); }
    TREPB_HD int* iswpP() const { return ipivP() + NC(); }
    TREPB_HD int* iflag() const { return iswpP() + NC(); }   // result of a first-warp section (teams of several warps)
    // the outcome of a section only the team's first warp ran, made known to every lane of the team
    TREPB_HD bool first_warp_result(bool ok) {
        if constexpr (Team::kWarps > 1) {
            if (t.lane() == 0) iflag()[0] = ok ? 1 : 0;
            t.sync();
            ok = iflag()[0] != 0;
        }
        return ok;
    }

    // ---- evaluation point: which = 0 midpoint, 1 q1, 2 q2 (midpointvi.c:401-457)
    TREPB_HD void set_point(int which, double dt) {
        for (int i = t.lane(); i < NQ(); i += Team::kSize) {
            const double a = w[L.q1 + i], b = w[L.q2 + i];
            w[L.qe + i] = which == 0 ? 0.5 * (b + a) : (which == 1 ? a : b);
            w[L.dq + i] = (b - a) / dt;
        }
        t.sync();
    }

    template <int A_>
    TREPB_HD static void rot_cols(double* Rb, double c_, double s_) {
        constexpr int b = (A_ + 1) % 3, c = (A_ + 2) % 3;
        TREPB_UNROLL
        for (int r = 0; r < 3; ++r) {
            const double rb = Rb[r * 3 + b], rc = Rb[r * 3 + c];
            Rb[r * 3 + b] = c_ * rb + s_ * rc;
            Rb[r * 3 + c] = -s_ * rb + c_ * rc;
        }
    }
    // joint twist (v_O, w) in world coordinates about the world origin, axis aw through pl
    TREPB_HD static void twist(const double* aw, const double* pl, bool rot, double* s6) {
        if (rot) {
            cross3(pl, aw, s6);
            s6[3] = aw[0]; s6[4] = aw[1]; s6[5] = aw[2];
        } else {
            s6[0] = aw[0]; s6[1] = aw[1]; s6[2] = aw[2];
            s6[3] = s6[4] = s6[5] = 0.0;
        }
    }
    // Lie bracket [X, Y] of two twists (v, w):  (w_X x v_Y + v_X x w_Y , w_X x w_Y)
    TREPB_HD static void bracket(const double* X, const double* Y, double* o) {
        double a[3], b[3];
        cross3(X + 3, Y, a);
        cross3(X, Y + 3, b);
        o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
        cross3(X + 3, Y + 3, o + 3);
    }

    // ---- pose sweep, one step per tree level.
    //   mode 0: links that carry mass, at qe, with spatial velocities V          -> R, p, V
    //   mode 1: links that carry constraint points, at qe                        -> R, p
    //   mode 2: both at once on the two halves of the team: half 0 as mode 0 (the midpoint, qe),
    //           half 1 as mode 1 but at q2 into a second pose set (R2, p2) that lives in the comp
    //           region until dyn_first() overwrites it.  (The link tree is narrow - at most a few
    //           links per level - so a full team per sweep would leave most lanes idle.)
    TREPB_HD void pose_sweep(int mode) {
        const int nls = L.nls, lane = t.lane();
        constexpr int TS = Team::kSize;
        const int npass = (mode == 2 && TS == 1) ? 2 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            int half, sub, stride;
            if (mode != 2) { half = mode; sub = lane; stride = TS; }
            else if (TS == 1) { half = pass; sub = 0; stride = 1; }
            else { half = lane / (TS / 2); sub = lane % (TS / 2); stride = TS / 2; }
            const bool second = mode == 2 && half == 1;
            const int flag = half == 0 ? (16 | 64) : 32, qoff = second ? L.q2 : L.qe;
            double* cs = w + (second ? L.comp : L.cs);
            for (int l = sub; l < NL(); l += stride) {
                const int kind = S.l_kind()[l];
                if ((kind & flag) == 0 || (kind & 4) == 0) continue;
                double sn, c;
                sincos_(w[qoff + S.l_cfg()[l]], &sn, &c);
                cs[l] = c;
                cs[nls + l] = sn;
            }
        }
        t.sync();
        // one step per super level: a lane takes the head of a chain (a run of single-child links)
        // and follows it with the pose and velocity in registers
        for (int sl = 0; sl < S.nsl; ++sl) {
            const int h0 = S.sl_off()[sl], h1 = S.sl_off()[sl + 1];
            for (int pass = 0; pass < npass; ++pass) {
                int half, sub, stride;
                if (mode != 2) { half = mode; sub = lane; stride = TS; }
                else if (TS == 1) { half = pass; sub = 0; stride = 1; }
                else { half = lane / (TS / 2); sub = lane % (TS / 2); stride = TS / 2; }
                const bool second = mode == 2 && half == 1, vel = half == 0;
                const int flag = vel ? (16 | 64) : 32, qoff = second ? L.q2 : L.qe;
                double* cs = w + (second ? L.comp : L.cs);
                double* R = w + (second ? L.comp + 2 * nls : L.R);
                double* p = w + (second ? L.comp + 11 * nls : L.p);
                double* V = w + L.V;
                for (int h = h0 + sub; h < h1; h += stride) {
                    int l = S.sl_head()[h];
                    int kind = S.l_kind()[l];
                    if ((kind & flag) == 0) continue;
                    double Rb[9], pb[3], Vl[6];
                    {
                        const int par = S.l_par()[l];
                        if (par < 0) {
                            TREPB_UNROLL for (int k = 0; k < 9; ++k) Rb[k] = (k % 4 == 0) ? 1.0 : 0.0;
                            TREPB_UNROLL for (int k = 0; k < 3; ++k) pb[k] = 0.0;
                            TREPB_UNROLL for (int k = 0; k < 6; ++k) Vl[k] = 0.0;
                        } else {
                            TREPB_UNROLL for (int k = 0; k < 9; ++k) Rb[k] = R[k * nls + par];
                            TREPB_UNROLL for (int k = 0; k < 3; ++k) pb[k] = p[k * nls + par];
                            if (vel) { TREPB_UNROLL for (int k = 0; k < 6; ++k) Vl[k] = V[k * nls + par]; }
                        }
                    }
                    for (;;) {
                        const int a = kind & 3, cfg = S.l_cfg()[l];
                        const bool rot = (kind & 4) != 0;
                        if (kind & 8) {
                            const double* Rc = S.l_Rc() + 9 * l;
                            const double* pc = S.l_pc() + 3 * l;
                            double Rn[9];
                            TREPB_UNROLL
                            for (int r = 0; r < 3; ++r) {
                                TREPB_UNROLL
                                for (int q = 0; q < 3; ++q)
                                    Rn[r * 3 + q] = Rb[r * 3] * Rc[q] + Rb[r * 3 + 1] * Rc[3 + q] + Rb[r * 3 + 2] * Rc[6 + q];
                                pb[r] += Rb[r * 3] * pc[0] + Rb[r * 3 + 1] * pc[1] + Rb[r * 3 + 2] * pc[2];
                            }
                            TREPB_UNROLL for (int k = 0; k < 9; ++k) Rb[k] = Rn[k];
                        }
                        if (rot) {
                            const double c_ = cs[l], s_ = cs[nls + l];
                            if (a == 0) rot_cols<0>(Rb, c_, s_);
                            else if (a == 1) rot_cols<1>(Rb, c_, s_);
                            else rot_cols<2>(Rb, c_, s_);
                        }
                        double aw[3];
                        TREPB_UNROLL for (int r = 0; r < 3; ++r) aw[r] = sel3(Rb + 3 * r, a);
                        if (!rot) {
                            const double x = w[qoff + cfg];
                            TREPB_UNROLL for (int r = 0; r < 3; ++r) pb[r] += x * aw[r];
                        }
                        TREPB_UNROLL for (int k = 0; k < 9; ++k) R[k * nls + l] = Rb[k];
                        TREPB_UNROLL for (int k = 0; k < 3; ++k) p[k * nls + l] = pb[k];
                        if (vel) {
                            double s6[6];
                            twist(aw, pb, rot, s6);
                            const double d = w[L.dq + cfg];
                            TREPB_UNROLL for (int k = 0; k < 6; ++k) { Vl[k] += s6[k] * d; V[k * nls + l] = Vl[k]; }
                        }
                        const int nx = S.l_next()[l];
                        if (nx < 0) break;
                        const int knx = S.l_kind()[nx];
                        if ((knx & flag) == 0) break;
                        l = nx;
                        kind = knx;
                    }
                }
            }
            t.sync();
        }
    }
    // which pose set / configuration the constraint routines read: the primary one (R, p at qe) or
    // the second one of a dual sweep (R2, p2 at q2)
    int oR, oP, oQ;
    TREPB_HD void use_pose(bool second) {
        oR = second ? L.comp + 2 * L.nls : L.R;
        oP = second ? L.comp + 11 * L.nls : L.p;
        oQ = second ? L.q2 : L.qe;
    }

    TREPB_HD void load_link_twists(int l, double* s6, double* W6, double* aw) const {
        const int nls = L.nls, kind = S.l_kind()[l], a = kind & 3, par = S.l_par()[l];
        double pl[3];
        TREPB_UNROLL for (int r = 0; r < 3; ++r) { aw[r] = w[L.R + (r * 3 + a) * nls + l]; pl[r] = w[L.p + r * nls + l]; }
        twist(aw, pl, (kind & 4) != 0, s6);
        if (par < 0) { TREPB_UNROLL for (int k = 0; k < 6; ++k) W6[k] = 0.0; }
        else {
            double Vp[6];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) Vp[k] = w[L.V + k * nls + par];
            bracket(Vp, s6, W6);
        }
    }

    // ---- first-order dynamics at the midpoint: composite inertia/momentum, Lq, Lv
    // (system.c:129-168 L_dq, :270-300 L_ddq, gravity.c:30-50)
    TREPB_HD void dyn_first() {
        const int nls = L.nls, lane = t.lane();
        double* comp = w + L.comp;
        for (int i = lane; i < NQ(); i += Team::kSize) { w[L.Lq + i] = 0.0; w[L.Lv + i] = 0.0; }
        for (int l = lane; l < NL(); l += Team::kSize) {
            if (!S.dyn(l)) continue;
            double Rl[9], pl[3], Vl[6];
            TREPB_UNROLL for (int k = 0; k < 9; ++k) Rl[k] = w[L.R + k * nls + l];
            TREPB_UNROLL for (int k = 0; k < 3; ++k) pl[k] = w[L.p + k * nls + l];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) Vl[k] = w[L.V + k * nls + l];
            const double* in = S.l_in() + 10 * l;
            const double m = in[0];
            double hr[3], hw[3], I[6];
            TREPB_UNROLL
            for (int r = 0; r < 3; ++r) {
                hr[r] = Rl[r * 3] * in[1] + Rl[r * 3 + 1] * in[2] + Rl[r * 3 + 2] * in[3];
                hw[r] = hr[r] + m * pl[r];
            }
            // Ibar^w = R Ibar R^T + 2 (hr.p) 1 - hr p^T - p hr^T + m (|p|^2 1 - p p^T)
            {
                double M[9], T[9];
                M[0] = in[4]; M[4] = in[5]; M[8] = in[6];
                M[1] = M[3] = in[7]; M[2] = M[6] = in[8]; M[5] = M[7] = in[9];
                TREPB_UNROLL
                for (int r = 0; r < 3; ++r)
                    TREPB_UNROLL
                    for (int c = 0; c < 3; ++c)
                        T[r * 3 + c] = Rl[r * 3] * M[c] + Rl[r * 3 + 1] * M[3 + c] + Rl[r * 3 + 2] * M[6 + c];
                const double hp = dot3(hr, pl), pp = dot3(pl, pl);
                TREPB_UNROLL
                for (int r = 0; r < 3; ++r)
                    TREPB_UNROLL
                    for (int c = r; c < 3; ++c) {
                        double v = T[r * 3] * Rl[c * 3] + T[r * 3 + 1] * Rl[c * 3 + 1] + T[r * 3 + 2] * Rl[c * 3 + 2];
                        v += -hr[r] * pl[c] - pl[r] * hr[c] - m * pl[r] * pl[c];
                        if (r == c) v += 2.0 * hp + m * pp;
                        I[symi(r, c)] = v;
                    }
            }
            double mu[6];
            inertia_apply(m, hw, I, Vl, mu);
            comp[0 * nls + l] = m;
            TREPB_UNROLL for (int k = 0; k < 3; ++k) comp[(1 + k) * nls + l] = hw[k];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) comp[(4 + k) * nls + l] = I[k];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) comp[(10 + k) * nls + l] = mu[k];
        }
        t.sync();
        // up sweep: a parent gathers its children (plain sums in world coordinates)
        for (int lev = NLEVELS() - 1; lev >= 1; --lev) {
            const int lo = S.lvl_off()[lev - 1], cnt = S.lvl_off()[lev] - lo;
            for (int e = lane; e < cnt * 16; e += Team::kSize) {
                const int l = lo + (e >> 4), k = e & 15;
                if (!S.dyn(l)) continue;
                const int c0 = S.l_child0()[l], nch = S.l_nchild()[l];
                double acc = comp[k * nls + l];
                for (int ch = c0; ch < c0 + nch; ++ch)
                    if (S.dyn(ch)) acc += comp[k * nls + ch];
                comp[k * nls + l] = acc;
            }
            t.sync();
        }
        for (int l = lane; l < NL(); l += Team::kSize) {
            if (!S.dyn(l)) continue;
            double s6[6], W6[6], aw[3], mu[6], h[3];
            load_link_twists(l, s6, W6, aw);
            const double m = comp[l];
            TREPB_UNROLL for (int k = 0; k < 3; ++k) h[k] = comp[(1 + k) * nls + l];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) mu[k] = comp[(10 + k) * nls + l];
            const int cfg = S.l_cfg()[l];
            w[L.Lv + cfg] = dot6(s6, mu);
            double lq = dot6(W6, mu);
            if (S.has_gravity) {
                double P[3], t3[3];
                cross3(s6 + 3, h, t3);
                TREPB_UNROLL for (int k = 0; k < 3; ++k) P[k] = m * s6[k] + t3[k];
                lq += dot3(S.grav, P);
            }
            w[L.Lq + cfg] = lq;
        }
        t.sync();
        // ConfigSpring (configspring.c:22-30)
        for (int i = lane; i < NQ(); i += Team::kSize)
            if (S.ks()[i] != 0.0) w[L.Lq + i] -= S.ks()[i] * w[L.qe + i] - S.kq0()[i];
        if (NNS() > 0) {
            // NonlinearConfigSpring: dV/dq = -y(m q + b), y a piecewise quintic (nonlinear_config_spring.c:23-47,
            // spline.c:7-62: same segment search, same extrapolation segments)
            for (int i = lane; i < NQ(); i += Team::kSize) {
                double ksum = S.ks()[i], lq = 0.0;
                for (int sp = 0; sp < NNS(); ++sp) {
                    const int32_t* si = S.nsp_i() + 3 * sp;
                    if (si[0] != i) continue;
                    const double* tab = S.nsp_tab() + si[1];
                    const int n = si[2];
                    const double m = S.nsp_d()[2 * sp], b = S.nsp_d()[2 * sp + 1];
                    const double x = m * w[L.qe + i] + b;
                    int seg = 0;
                    if (x >= tab[n - 1]) seg = n - 2;
                    else if (!(x < tab[0])) { while (x >= tab[seg + 1]) ++seg; }
                    const double* cf = tab + n + 6 * seg;
                    const double dx = x - tab[seg];
                    const double dx2 = dx * dx, dx3 = dx2 * dx, dx4 = dx3 * dx, dx5 = dx4 * dx;
                    lq += cf[0] * dx5 + cf[1] * dx4 + cf[2] * dx3 + cf[3] * dx2 + cf[4] * dx + cf[5];
                    ksum -= (5.0 * cf[0] * dx4 + 4.0 * cf[1] * dx3 + 3.0 * cf[2] * dx2 + 2.0 * cf[3] * dx + cf[4]) * m;
                }
                w[L.Lq + i] += lq;
                w[L.KS + i] = ksum;
            }
        }
        t.sync();
        springs_first();
    }

    // ---- LinearSpring potentials (potentials/linearspring.c:30-74) at the midpoint pose: V = k/2 (x - x0)^2 with
    // x = |pA - pB|.  The end points are world points of the primary pose set; d(pA - pB)/dq_j and the second
    // derivatives come from dpoint() and the axis-cross-derivative rule the constraint Hessians use.
    TREPB_HD void spring_v(int sp, double* v, double& x) const {
        double pa[3], pb[3];
        point(S.sp_a()[sp], pa); point(S.sp_b()[sp], pb);
        TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
        x = sqrt(dot3(v, v));
    }
    // L_dq -= dV/dq (called by dyn_first; the primary pose set is selected)
    TREPB_HD void springs_first() {
        if (NS() == 0 && NPF() == 0) return;
        points(true);
        if (NPF() > 0) {
            for (int j = t.lane(); j < ND(); j += Team::kSize) w[L.FD + j] = 0.0;
            t.sync();
            dampers_first();
            wrenches_first();
        }
        if (NS() == 0) return;
        for (int j = t.lane(); j < NQ(); j += Team::kSize) {
            if (S.xs_idx()[j] < 0) continue;
            const int lj = S.cfg_link()[j];
            double acc = 0.0;
            for (int sp = 0; sp < NS(); ++sp) {
                double v[3], x, da[3], db[3];
                spring_v(sp, v, x);
                dpoint(S.sp_a()[sp], lj, da); dpoint(S.sp_b()[sp], lj, db);
                const double dx = (1.0 / x) * (v[0] * (da[0] - db[0]) + v[1] * (da[1] - db[1]) + v[2] * (da[2] - db[2]));
                double val = S.sp_k()[sp] * (x - S.sp_x0()[sp]) * dx;
                if (dx != dx && S.sp_x0()[sp] == 0.0) val = 0.0;      // linearspring.c:41
                acc += val;
            }
            w[L.Lq + j] -= acc;
        }
        t.sync();
    }
    // ---- LinearDamper forces (forces/lineardamper.c:14-107 over a single-segment TapeMeasure, tapemeasure.c:6-226)
    // at the midpoint: x = |pA - pB|, dx_j = (v . d(pA - pB)/dq_j) / x for the configs that move exactly one end,
    // vel = sum_k dx_k dq_k, f_j = -c vel dx_j.   FD [nd] <- sum over the dampers (spring / damper points are valid)
    TREPB_HD void dampers_first() {
        if (NFD() == 0) return;
        const int lane = t.lane();
        for (int f = 0; f < NFD(); ++f) {
            const int off = S.dp_off()[f], m = S.dp_off()[f + 1] - off;
            const int* list = S.dp_cfg() + off;
            const int A = S.da_a()[f], B = S.da_b()[f];
            double pa[3], pb[3], v[3];
            point(A, pa); point(B, pb);
            TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
            const double x = sqrt(dot3(v, v));
            t.sync();
            for (int a = lane; a < m; a += Team::kSize) {
                const int lj = S.cfg_link()[list[a]];
                double da[3], db[3];
                dpoint(A, lj, da); dpoint(B, lj, db);
                w[L.FX + a] = 1.0 / x * (v[0] * (da[0] - db[0]) + v[1] * (da[1] - db[1]) + v[2] * (da[2] - db[2]));
            }
            t.sync();
            double vel = 0.0;
            for (int a = 0; a < m; ++a) vel += w[L.FX + a] * w[L.dq + list[a]];
            for (int a = lane; a < m; a += Team::kSize)
                if (list[a] < ND()) w[L.FD + list[a]] += -S.da_c()[f] * vel * w[L.FX + a];
        }
        t.sync();
    }
    // FQ (j, i) = sum over dampers of d f_j / d q_i, FV (j, i) = d f_j / d dq_i on the compact block of configs that
    // move exactly one end of some damper (called by dyn_second: the comp region is free as scratch)
    TREPB_HD void dampers_second() {
        if (NPF() == 0) return;
        const int lane = t.lane(), nqf = S.nqf, nls = L.nls, nd = ND();
        for (int e = lane; e < nqf * nqf; e += Team::kSize) { w[L.FQ + e] = 0.0; w[L.FV + e] = 0.0; }
        double* DA = w + L.SC;
        for (int f = 0; f < NFD(); ++f) {
            const int off = S.dp_off()[f], m = S.dp_off()[f + 1] - off;
            const int* list = S.dp_cfg() + off;
            const int A = S.da_a()[f], B = S.da_b()[f];
            const double c = S.da_c()[f];
            double* DB = DA + 3 * m;
            double* DX = DB + 3 * m;
            double pa[3], pb[3], v[3];
            point(A, pa); point(B, pb);
            TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
            const double x = sqrt(dot3(v, v));
            t.sync();
            for (int a = lane; a < m; a += Team::kSize) {
                const int lj = S.cfg_link()[list[a]];
                double da[3], db[3];
                dpoint(A, lj, da); dpoint(B, lj, db);
                TREPB_UNROLL for (int k = 0; k < 3; ++k) { DA[k * m + a] = da[k]; DB[k * m + a] = db[k]; }
                DX[a] = 1.0 / x * (v[0] * (da[0] - db[0]) + v[1] * (da[1] - db[1]) + v[2] * (da[2] - db[2]));
            }
            t.sync();
            double vel = 0.0;
            for (int a = 0; a < m; ++a) vel += DX[a] * w[L.dq + list[a]];
            const int lA = S.pt_link()[A], lB = S.pt_link()[B];
            const unsigned long long ancA = lA >= 0 ? S.l_anc()[lA] : 0ull, ancB = lB >= 0 ? S.l_anc()[lB] : 0ull;
            // one column i per lane: ddx(k, i) = TapeMeasure_length_dqdq, vel_dq(i) = sum_k ddx(k, i) dq_k
            for (int ai = lane; ai < m; ai += Team::kSize) {
                const int i = list[ai], li = S.cfg_link()[i], xi = S.xf_idx()[i];
                double veldq = 0.0;
                for (int bk = 0; bk < m; ++bk) {
                    const int kc = list[bk], lk = S.cfg_link()[kc];
                    double ddv[3] = {0.0, 0.0, 0.0};
                    if (li >= 0 && lk >= 0) {
                        int up = lk, lo_idx = ai;
                        if (!((S.l_anc()[li] >> lk) & 1ull)) { up = li; lo_idx = bk; }
                        const int kup = S.l_kind()[up];
                        if (kup & 4) {
                            const int ax = kup & 3;
                            const double aw[3] = {w[oR + ax * nls + up], w[oR + (3 + ax) * nls + up], w[oR + (6 + ax) * nls + up]};
                            const bool onA = ((ancA >> li) & 1ull) && ((ancA >> lk) & 1ull);
                            const bool onB = ((ancB >> li) & 1ull) && ((ancB >> lk) & 1ull);
                            double d3[3] = {0.0, 0.0, 0.0};
                            if (onA) { TREPB_UNROLL for (int q = 0; q < 3; ++q) d3[q] += DA[q * m + lo_idx]; }
                            if (onB) { TREPB_UNROLL for (int q = 0; q < 3; ++q) d3[q] -= DB[q * m + lo_idx]; }
                            cross3(aw, d3, ddv);
                        }
                    }
                    double dk[3], di[3];
                    TREPB_UNROLL for (int q = 0; q < 3; ++q) { dk[q] = DA[q * m + bk] - DB[q * m + bk]; di[q] = DA[q * m + ai] - DB[q * m + ai]; }
                    const double tt = DX[bk] * DX[ai] - dot3(dk, di) - dot3(v, ddv);
                    const double ddx = -1.0 / x * tt;
                    veldq += ddx * w[L.dq + kc];
                    if (kc < nd) w[L.FQ + S.xf_idx()[kc] * nqf + xi] += -c * vel * ddx;
                }
                for (int bj = 0; bj < m; ++bj) {
                    const int j = list[bj];
                    if (j >= nd) continue;
                    w[L.FQ + S.xf_idx()[j] * nqf + xi] += -c * veldq * DX[bj];
                    w[L.FV + S.xf_idx()[j] * nqf + xi] += -c * DX[ai] * DX[bj];
                }
            }
        }
        t.sync();
        wrenches_second();
    }
    // ---- Body / Hybrid / Spatial wrenches (forces/bodywrench.c, hybridwrench.c, spatialwrench.c) on a frame F:
    // f_j = J_j . w with J_j = (dp_j, a_j) hybrid, (dp_j - a_j x p_F, a_j) spatial, (R_F^T dp_j, R_F^T a_j) body;
    // dp_j = d p_F / d q_j, a_j the world axis of a revolute joint (0: prismatic); d a_j / d q_i = a_i x a_j for a
    // joint i above j; each of the six components of w a constant or an input.
    TREPB_HD void wrench_setup(int wi, double* wr, double* RF, double* pF) const {
        const int32_t* ii = S.wr_i() + 8 * wi;
        const double* dd = S.wr_d() + 15 * wi;
        TREPB_UNROLL for (int k = 0; k < 6; ++k) wr[k] = ii[2 + k] >= 0 ? w[L.u1 + ii[2 + k]] : dd[k];
        const int pt = ii[1], l = S.pt_link()[pt], nls = L.nls;
        point(pt, pF);
        TREPB_UNROLL
        for (int r = 0; r < 3; ++r)
            TREPB_UNROLL
            for (int c = 0; c < 3; ++c)
                RF[r * 3 + c] = l < 0 ? dd[6 + r * 3 + c]
                                      : (w[oR + (r * 3) * nls + l] * dd[6 + c] + w[oR + (r * 3 + 1) * nls + l] * dd[9 + c] +
                                         w[oR + (r * 3 + 2) * nls + l] * dd[12 + c]);
    }
    TREPB_HD void link_axis(int lj, double* aj) const {   // world axis of a revolute link, zero otherwise
        const int kind = S.l_kind()[lj], a = kind & 3, nls = L.nls;
        if (kind & 4) { aj[0] = w[oR + a * nls + lj]; aj[1] = w[oR + (3 + a) * nls + lj]; aj[2] = w[oR + (6 + a) * nls + lj]; }
        else { aj[0] = aj[1] = aj[2] = 0.0; }
    }
    TREPB_HD void wrench_J(int kind, const double* RF, const double* pF, const double* dp, const double* aj, double* J) const {
        if (kind == F_HYBRID_WRENCH) {
            TREPB_UNROLL for (int k = 0; k < 3; ++k) { J[k] = dp[k]; J[3 + k] = aj[k]; }
        } else if (kind == F_SPATIAL_WRENCH) {
            double t3[3];
            cross3(aj, pF, t3);
            TREPB_UNROLL for (int k = 0; k < 3; ++k) { J[k] = dp[k] - t3[k]; J[3 + k] = aj[k]; }
        } else {
            TREPB_UNROLL
            for (int k = 0; k < 3; ++k) {
                J[k] = RF[k] * dp[0] + RF[3 + k] * dp[1] + RF[6 + k] * dp[2];
                J[3 + k] = RF[k] * aj[0] + RF[3 + k] * aj[1] + RF[6 + k] * aj[2];
            }
        }
    }
    // FD [j] += J_j . w and FUW [j][u] = J_j [k] for the input components (one row j per lane)
    TREPB_HD void wrenches_first() {
        if (NW() == 0) return;
        const int nu = NU();
        for (int j = t.lane(); j < ND(); j += Team::kSize) {
            for (int u = 0; u < nu; ++u) w[L.FUW + j * nu + u] = 0.0;
            const int lj = S.cfg_link()[j];
            if (S.xf_idx()[j] < 0 || lj < 0) continue;
            double acc = 0.0;
            for (int wi = 0; wi < NW(); ++wi) {
                if (!((S.wr_dep()[wi] >> j) & 1ull)) continue;
                const int32_t* ii = S.wr_i() + 8 * wi;
                double wr[6], RF[9], pF[3], dp[3], aj[3], J[6];
                wrench_setup(wi, wr, RF, pF);
                dpoint(ii[1], lj, dp);
                link_axis(lj, aj);
                wrench_J(ii[0], RF, pF, dp, aj, J);
                acc += dot6(J, wr);
                TREPB_UNROLL for (int k = 0; k < 6; ++k) if (ii[2 + k] >= 0) w[L.FUW + j * nu + ii[2 + k]] += J[k];
            }
            w[L.FD + j] += acc;
        }
        t.sync();
    }
    // FQ (j, i) += d J_j / d q_i . w (one entry of the compact block per lane)
    TREPB_HD void wrenches_second() {
        if (NW() == 0) return;
        const int nqf = S.nqf, nd = ND();
        for (int e = t.lane(); e < nqf * nqf; e += Team::kSize) {
            const int j = S.xf_cfg()[e / nqf], i = S.xf_cfg()[e % nqf];
            if (j >= nd) continue;
            const int lj = S.cfg_link()[j], li = S.cfg_link()[i];
            if (lj < 0 || li < 0) continue;
            double acc = 0.0;
            for (int wi = 0; wi < NW(); ++wi) {
                const unsigned long long dep = S.wr_dep()[wi];
                if (!((dep >> j) & 1ull) || !((dep >> i) & 1ull)) continue;
                const int32_t* ii = S.wr_i() + 8 * wi;
                const int kind = ii[0], pt = ii[1];
                double wr[6], RF[9], pF[3], dpj[3], dpi[3], aj[3], ai[3], ddp[3], da[3], dJ[6];
                wrench_setup(wi, wr, RF, pF);
                dpoint(pt, lj, dpj); dpoint(pt, li, dpi);
                link_axis(lj, aj); link_axis(li, ai);
                // d2 p_F / d q_j d q_i: the upper joint's axis crosses the lower joint's first derivative
                const bool j_above = ((S.l_anc()[li] >> lj) & 1ull) != 0;
                if (j_above) cross3(aj, dpi, ddp); else cross3(ai, dpj, ddp);
                // d a_j / d q_i = a_i x a_j when joint i is strictly above joint j
                if (i != j && ((S.l_anc()[lj] >> li) & 1ull)) cross3(ai, aj, da);
                else { da[0] = da[1] = da[2] = 0.0; }
                if (kind == F_HYBRID_WRENCH) {
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { dJ[k] = ddp[k]; dJ[3 + k] = da[k]; }
                } else if (kind == F_SPATIAL_WRENCH) {
                    double t3[3], u3[3];
                    cross3(da, pF, t3);
                    cross3(aj, dpi, u3);
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { dJ[k] = ddp[k] - t3[k] - u3[k]; dJ[3 + k] = da[k]; }
                } else {
                    // d (R^T x) / d q_i = R^T (dx/dq_i - a_i x x)
                    double t3[3], u3[3];
                    cross3(ai, dpj, t3);
                    cross3(ai, aj, u3);
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { t3[k] = ddp[k] - t3[k]; u3[k] = da[k] - u3[k]; }
                    TREPB_UNROLL
                    for (int k = 0; k < 3; ++k) {
                        dJ[k] = RF[k] * t3[0] + RF[3 + k] * t3[1] + RF[6 + k] * t3[2];
                        dJ[3 + k] = RF[k] * u3[0] + RF[3 + k] * u3[1] + RF[6 + k] * u3[2];
                    }
                }
                acc += dot6(dJ, wr);
            }
            w[L.FQ + e] += acc;
        }
        t.sync();
    }
    // generalized force on dynamic config j besides the per-config constants: Damping, ConfigForce, LinearDampers
    TREPB_HD double ext_force(int j) const {
        double fo = -S.damp()[j] * w[L.dq + j];
        for (int u = 0; u < NU(); ++u) fo += S.Fu()[j * NU() + u] * w[L.u1 + u];
        if (NPF() > 0) fo += w[L.FD + j];
        return fo;
    }

    // XS = sum over springs of d2V / dq_i dq_j on the configs any spring depends on (called by dyn_second once
    // the pair tables are done: the comp region is free to hold dA_i, dB_i, dx_i of one spring at a time)
    TREPB_HD void springs_second() {
        if (NS() == 0) return;
        const int lane = t.lane(), nqs = S.nqs, nls = L.nls;
        for (int e = lane; e < nqs * nqs; e += Team::kSize) w[L.XS + e] = 0.0;
        double* DA = w + L.SC;
        for (int sp = 0; sp < NS(); ++sp) {
            const int off = S.sp_off()[sp], m = S.sp_off()[sp + 1] - off;
            const int* list = S.sp_cfg() + off;
            const int A = S.sp_a()[sp], B = S.sp_b()[sp];
            double* DB = DA + 3 * m;
            double* DX = DB + 3 * m;
            double v[3], x;
            spring_v(sp, v, x);
            const double k = S.sp_k()[sp], x0 = S.sp_x0()[sp];
            t.sync();
            for (int a = lane; a < m; a += Team::kSize) {
                const int lj = S.cfg_link()[list[a]];
                double da[3], db[3];
                dpoint(A, lj, da); dpoint(B, lj, db);
                TREPB_UNROLL for (int c = 0; c < 3; ++c) { DA[c * m + a] = da[c]; DB[c * m + a] = db[c]; }
                DX[a] = (1.0 / x) * (v[0] * (da[0] - db[0]) + v[1] * (da[1] - db[1]) + v[2] * (da[2] - db[2]));
            }
            t.sync();
            const int lA = S.pt_link()[A], lB = S.pt_link()[B];
            const unsigned long long ancA = lA >= 0 ? S.l_anc()[lA] : 0ull, ancB = lB >= 0 ? S.l_anc()[lB] : 0ull;
            for (int e = lane; e < m * m; e += Team::kSize) {
                const int bj = e / m, ai = e - bj * m;
                const int i = list[ai], j = list[bj];
                const int li = S.cfg_link()[i], lj = S.cfg_link()[j];
                double ddv[3] = {0.0, 0.0, 0.0};
                if (li >= 0 && lj >= 0) {
                    // the upper joint's axis crosses the lower joint's first derivative
                    int up = li, lo_idx = bj;
                    if (!((S.l_anc()[lj] >> li) & 1ull)) { up = lj; lo_idx = ai; }
                    const int kup = S.l_kind()[up];
                    if (kup & 4) {
                        const int a = kup & 3;
                        const double aw[3] = {w[oR + a * nls + up], w[oR + (3 + a) * nls + up], w[oR + (6 + a) * nls + up]};
                        const bool onA = ((ancA >> li) & 1ull) && ((ancA >> lj) & 1ull);
                        const bool onB = ((ancB >> li) & 1ull) && ((ancB >> lj) & 1ull);
                        double d3[3] = {0.0, 0.0, 0.0};
                        if (onA) { TREPB_UNROLL for (int c = 0; c < 3; ++c) d3[c] += DA[c * m + lo_idx]; }
                        if (onB) { TREPB_UNROLL for (int c = 0; c < 3; ++c) d3[c] -= DB[c * m + lo_idx]; }
                        cross3(aw, d3, ddv);
                    }
                }
                double di[3], dj[3];
                TREPB_UNROLL for (int c = 0; c < 3; ++c) { di[c] = DA[c * m + ai] - DB[c * m + ai]; dj[c] = DA[c * m + bj] - DB[c * m + bj]; }
                const double vdi = dot3(v, di), didj = dot3(di, dj), dix = DX[ai], djx = DX[bj];
                const double ddx = -djx / (x * x) * vdi + 1.0 / x * didj + 1.0 / x * dot3(v, ddv);
                w[L.XS + S.xs_idx()[i] * nqs + S.xs_idx()[j]] += k * dix * djx + k * (x - x0) * ddx;
            }
        }
        t.sync();
    }

    // ---- second-order tables on the chain pairs (system.c:170-268, 302-393, 479-512)
    // newton_dt != 0: called from the Newton iteration, which needs two combinations per pair only - they
    // are stored instead of the four tables (VV <- dt/4 L_dqdq - 1/dt L_ddqddq, QQ <- the antisymmetric part),
    // so the solve-only layout carries two pair arrays.
    TREPB_HD void dyn_second(double newton_dt = 0.0) {
        const int nls = L.nls, lane = t.lane();
        double* comp = w + L.comp;
        for (int l = lane; l < NL(); l += Team::kSize) {
            if (!S.dyn(l)) continue;
            double s6[6], W6[6], aw[3], mu[6], h[3], I[6], H[6], G[6], P[3], t3[3];
            load_link_twists(l, s6, W6, aw);
            const double m = comp[l];
            TREPB_UNROLL for (int k = 0; k < 3; ++k) h[k] = comp[(1 + k) * nls + l];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) I[k] = comp[(4 + k) * nls + l];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) mu[k] = comp[(10 + k) * nls + l];
            inertia_apply(m, h, I, s6, H);
            inertia_apply(m, h, I, W6, G);
            // G -= ad*_s mu = (f x w_s , n x w_s + f x v_s)
            cross3(mu, s6 + 3, t3);
            TREPB_UNROLL for (int k = 0; k < 3; ++k) G[k] -= t3[k];
            cross3(mu + 3, s6 + 3, t3);
            TREPB_UNROLL for (int k = 0; k < 3; ++k) G[3 + k] -= t3[k];
            cross3(mu, s6, t3);
            TREPB_UNROLL for (int k = 0; k < 3; ++k) G[3 + k] -= t3[k];
            cross3(s6 + 3, h, t3);
            TREPB_UNROLL for (int k = 0; k < 3; ++k) P[k] = m * s6[k] + t3[k];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) comp[k * nls + l] = H[k];
            TREPB_UNROLL for (int k = 0; k < 6; ++k) comp[(6 + k) * nls + l] = G[k];
            TREPB_UNROLL for (int k = 0; k < 3; ++k) comp[(12 + k) * nls + l] = P[k];
        }
        t.sync();
        for (int e = lane; e < NPAIRS(); e += Team::kSize) {
            const int ij = S.pair_ij()[e], i = ij & 255, j = ij >> 8;
            double s6[6], W6[6], aw[3], H[6], G[6];
            load_link_twists(i, s6, W6, aw);
            TREPB_UNROLL for (int k = 0; k < 6; ++k) { H[k] = comp[k * nls + j]; G[k] = comp[(6 + k) * nls + j]; }
            const double sG = dot6(s6, G);
            double WG = dot6(W6, G);
            if (S.has_gravity && S.rot(i)) {
                double P[3], N[3];
                TREPB_UNROLL for (int k = 0; k < 3; ++k) P[k] = comp[(12 + k) * nls + j];
                cross3(P, S.grav, N);
                WG += dot3(aw, N);
            }
            const double vv = dot6(s6, H), dn = i == j ? sG : dot6(W6, H);
            if (newton_dt != 0.0) {
                xv(L.VV)[e] = 0.25 * newton_dt * WG - 1.0 / newton_dt * vv;
                xv(L.QQ)[e] = 0.5 * dn - 0.5 * sG;   // + 1/2 L_ddqdq(i,k) - 1/2 L_ddqdq(k,i)
            } else {
                xv(L.VV)[e] = vv;
                xw(L.UP)[e] = sG;
                xw(L.DN)[e] = dn;
                xv(L.QQ)[e] = WG;
            }
        }
        t.sync();
        springs_second();
        dampers_second();
    }

    // ---- first-derivative blocks from the chain-pair tables.  Every block of calc_deriv1
    // (midpointvi.c:749-861: D1D1L2, D2D1L2, D1D2L2, D2D2L2 with the force terms) is one of four
    // combinations of  Q = dt/4 L_dqdq,  V = 1/dt L_ddqddq,  U = 1/2 L_ddqdq(i,j),  D = 1/2 L_ddqdq(j,i)
    // (i above-or-equal j).  pair_combos() replaces the raw tables by the combinations once per pair,
    //     VV <- (Q+V)+U+D   QQ <- (Q+V)-U-D   UP <- (Q-V)+U-D   DN <- (Q-V)-U+D
    // (ConfigSpring's -k on the diagonal folded into Q), so that each of the ~4000 block entries deriv1
    // reads is one lookup instead of four loads and the arithmetic.
    TREPB_HD void pair_combos(double dt) {
        for (int e = t.lane(); e < NPAIRS(); e += Team::kSize) {
            const int ij = S.pair_ij()[e], i = ij & 255, j = ij >> 8;
            double qq = xv(L.QQ)[e];
            if (i == j) qq -= ks_eff(S.l_cfg()[i]);
            const double Q = 0.25 * dt * qq, V = 1.0 / dt * xv(L.VV)[e], U = 0.5 * xw(L.UP)[e], Dn = 0.5 * xw(L.DN)[e];
            xv(L.VV)[e] = (Q + V) + U + Dn;
            xv(L.QQ)[e] = (Q + V) - U - Dn;
            xw(L.UP)[e] = (Q - V) + U - Dn;
            xw(L.DN)[e] = (Q - V) - U + Dn;
        }
        t.sync();
    }
    // combination for the config pair (a, b), b dynamic, after pair_combos():
    //   which 0: (Q+V)+vab+vba   1: (Q+V)-vab-vba   2: (Q-V)+vab-vba   3: (Q-V)-vab+vba
    // with vab = 1/2 L_ddqdq(a,b), vba = 1/2 L_ddqdq(b,a)  (first index of L_ddqdq is the velocity slot)
    TREPB_HD double comb(int which, int a, int b, double dt) const {
        const int m = S.pm()[a * ND() + b];
        double val;
        if (m == 0) val = a == b ? 0.25 * dt * -ks_eff(a) : 0.0;
        else {
            const int e = (m > 0 ? m : -m) - 1;
            if (which < 2) val = xv(which == 0 ? L.VV : L.QQ)[e];
            else val = xw(((which == 2) == (m > 0)) ? L.UP : L.DN)[e];
        }
        if (NS() > 0) {   // Q = dt/4 L_dqdq enters all four combinations with a plus sign; L_dqdq -= d2V/dqdq
            const int xa = S.xs_idx()[a], xb = S.xs_idx()[b];
            if (xa >= 0 && xb >= 0) val -= 0.25 * dt * w[L.XS + xa * S.nqs + xb];
        }
        if (NPF() > 0 && (which == 1 || which == 2)) {   // D1fm2 = dt/2 f_dq - f_ddq, D2fm2 = dt/2 f_dq + f_ddq (row b, column a)
            const int xa = S.xf_idx()[a], xb = S.xf_idx()[b];
            if (xa >= 0 && xb >= 0) {
                const double fq = 0.5 * dt * w[L.FQ + xb * S.nqf + xa], fv = w[L.FV + xb * S.nqf + xa];
                val += which == 1 ? fq - fv : fq + fv;
            }
        }
        return val;
    }

    // ---- world points of the constraints at the current pose
    TREPB_HD void points(bool springs = false) {
        const int nls = L.nls, np = NP();
        // constraint points [0, npc) at the selected pose; spring ends [npc, np) at the primary (midpoint) pose
        const int q0 = springs ? S.npc : 0, q1 = springs ? np : S.npc;
        for (int q = q0 + t.lane(); q < q1; q += Team::kSize) {
            const int l = S.pt_link()[q];
            const double* r = S.pt_r() + 3 * q;
            TREPB_UNROLL
            for (int k = 0; k < 3; ++k) {
                double v = r[k];
                if (l >= 0)
                    v = w[oP + k * nls + l] + (w[oR + (k * 3) * nls + l] * r[0] + w[oR + (k * 3 + 1) * nls + l] * r[1] +
                                             w[oR + (k * 3 + 2) * nls + l] * r[2]);
                w[L.pts + k * np + q] = v;
            }
        }
        t.sync();
    }
    TREPB_HD void point(int q, double* o) const { TREPB_UNROLL for (int k = 0; k < 3; ++k) o[k] = w[L.pts + k * NP() + q]; }
    TREPB_HD bool pdep(int q, int lj) const {
        const int l = S.pt_link()[q];
        return l >= 0 && lj >= 0 && ((S.l_anc()[l] >> lj) & 1ull) != 0;
    }
    // d p_q / d q_j for the link lj driven by config j  (zero unless the point hangs below lj)
    TREPB_HD void dpoint(int q, int lj, double* o) const {
        o[0] = o[1] = o[2] = 0.0;
        if (!pdep(q, lj)) return;
        const int nls = L.nls, kind = S.l_kind()[lj], a = kind & 3;
        const double aw[3] = {w[oR + a * nls + lj], w[oR + (3 + a) * nls + lj], w[oR + (6 + a) * nls + lj]};
        if (kind & 4) {
            double r[3];
            TREPB_UNROLL for (int k = 0; k < 3; ++k) r[k] = w[L.pts + k * NP() + q] - w[oP + k * nls + lj];
            cross3(aw, r, o);
        } else {
            o[0] = aw[0]; o[1] = aw[1]; o[2] = aw[2];
        }
    }

    // PointOnPlane (constraints/plane.c:14-84): world normal n_w = R_l n of constraint c (l = link of the
    // plane frame, n in link coordinates) and its derivatives d n_w / d q_j = a_j x n_w for a revolute
    // link lj above-or-equal l,  d2 n_w / d q_i d q_j = a_up x (a_lo x n_w)
    TREPB_HD void plane_normal(int c, double* nw) const {
        const int nls = L.nls, l = S.pt_link()[S.con_a()[c]];
        const double* n = S.con_n() + 3 * c;
        TREPB_UNROLL
        for (int k = 0; k < 3; ++k)
            nw[k] = l < 0 ? n[k] : (w[oR + (k * 3) * nls + l] * n[0] + w[oR + (k * 3 + 1) * nls + l] * n[1] +
                                    w[oR + (k * 3 + 2) * nls + l] * n[2]);
    }
    // world axis of a revolute link lj that carries the plane link l of constraint c, else zero
    TREPB_HD bool plane_axis(int c, int lj, double* aw) const {
        aw[0] = aw[1] = aw[2] = 0.0;
        const int l = S.pt_link()[S.con_a()[c]];
        if (l < 0 || lj < 0 || !((S.l_anc()[l] >> lj) & 1ull)) return false;
        const int kind = S.l_kind()[lj], a = kind & 3, nls = L.nls;
        if (!(kind & 4)) return false;
        aw[0] = w[oR + a * nls + lj]; aw[1] = w[oR + (3 + a) * nls + lj]; aw[2] = w[oR + (6 + a) * nls + lj];
        return true;
    }

    // ---- constraints at the current pose (distance.c:16-100, point.c:16-46, plane.c:14-84)
    //   want_h: hc ;  dh: 0 none, 1 -> Dh1 [nc][nd], 2 -> Dh2 [nc][nq]
    TREPB_HD void constraints(bool want_h, int dh) {
        const int lane = t.lane(), nq = NQ(), nd = ND(), nc = NC();
        if (want_h) {
            for (int c = lane; c < nc; c += Team::kSize) {
                double pa[3], pb[3], v[3];
                point(S.con_a()[c], pa); point(S.con_b()[c], pb);
                TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
                if (S.con_kind()[c] == C_DISTANCE) {
                    const int third = S.con_third()[c];
                    const double d = third >= 0 ? w[oQ + third] : S.con_dist()[c];
                    w[L.hc + c] = dot3(v, v) - d * d;
                } else if (S.con_kind()[c] == C_PLANE) {
                    double nw[3];
                    plane_normal(c, nw);
                    w[L.hc + c] = dot3(nw, v);
                } else {
                    w[L.hc + c] = sel3(v, S.con_third()[c]);
                }
            }
        }
        if (dh) {
            const int ncol = dh == 1 ? nd : nq;
            double* Dm = xv(dh == 1 ? L.Dh1 : L.Dh2);
            for (int e = lane; e < nc * ncol; e += Team::kSize) {
                const int c = e / ncol, j = e - c * ncol;
                double val = 0.0;
                if ((S.con_dep()[c] >> j) & 1ull) {
                    const int A = S.con_a()[c], B = S.con_b()[c], lj = S.cfg_link()[j];
                    double da[3], db[3], dv[3];
                    dpoint(A, lj, da); dpoint(B, lj, db);
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) dv[k] = da[k] - db[k];
                    if (S.con_kind()[c] == C_DISTANCE) {
                        double pa[3], pb[3], v[3];
                        point(A, pa); point(B, pb);
                        TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
                        const int third = S.con_third()[c];
                        val = dot3(v, dv);
                        if (third == j) val -= w[oQ + third];
                        val *= 2.0;
                    } else if (S.con_kind()[c] == C_PLANE) {
                        double pa[3], pb[3], v[3], nw[3], aw[3], dn[3];
                        point(A, pa); point(B, pb);
                        TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
                        plane_normal(c, nw);
                        plane_axis(c, lj, aw);
                        cross3(aw, nw, dn);
                        val = dot3(dn, v) + dot3(nw, dv);
                    } else {
                        val = sel3(dv, S.con_third()[c]);
                    }
                }
                Dm[e] = val;
            }
        }
        t.sync();
    }

    // Y[j][i] = sum_c lam_c d2h_c/dq_i dq_j   for i < nq, j < nd   (midpointvi.c:864-889).
    // Per constraint: first dA_i, dB_i (the two end points' first derivatives) for the configs the
    // constraint depends on, then the dependent (i, j) pairs only.  Scratch: the comp region.
    TREPB_HD void ddh_lambda(double* Y, int ldy) {
        const int lane = t.lane(), nq = NQ(), nd = ND(), nls = L.nls;
        if constexpr (D::kStatic) {
            for (int e = lane; e < D::NDC * D::NQC; e += Team::kSize) Y[e] = 0.0;
        } else {
            for (int e = lane; e < nd * nq; e += Team::kSize) {
                const int j = e / nq, i = e - j * nq;
                Y[j * ldy + i] = 0.0;
            }
        }
        double* DA = w + L.comp;
        for (int c = 0; c < NC(); ++c) {
            const int off = S.cd_off()[c], m = S.cd_off()[c + 1] - off, md = S.cd_nd()[c];
            const int* list = S.cd_cfg() + off;
            const int A = S.con_a()[c], B = S.con_b()[c];
            double* DB = DA + 3 * m;
            t.sync();
            for (int a = lane; a < 2 * m; a += Team::kSize) {
                const int which = a >= m, ai = which ? a - m : a;
                double d3[3];
                dpoint(which ? B : A, S.cfg_link()[list[ai]], d3);
                double* dst = which ? DB : DA;
                TREPB_UNROLL for (int k = 0; k < 3; ++k) dst[k * m + ai] = d3[k];
            }
            t.sync();
            const bool dist = S.con_kind()[c] == C_DISTANCE, plane = S.con_kind()[c] == C_PLANE;
            const int third = S.con_third()[c];
            const double lam = w[L.lam + c];
            double nw[3] = {0.0, 0.0, 0.0};
            if (plane) plane_normal(c, nw);
            double v[3];
            {
                double pa[3], pb[3];
                point(A, pa); point(B, pb);
                TREPB_UNROLL for (int k = 0; k < 3; ++k) v[k] = pa[k] - pb[k];
            }
            const int lA = S.pt_link()[A], lB = S.pt_link()[B];
            const unsigned long long ancA = lA >= 0 ? S.l_anc()[lA] : 0ull, ancB = lB >= 0 ? S.l_anc()[lB] : 0ull;
            for (int e = lane; e < m * md; e += Team::kSize) {
                const int bj = e / m, ai = e - bj * m;     // j = list[bj] dynamic, i = list[ai]
                const int i = list[ai], j = list[bj];
                const int li = S.cfg_link()[i], lj = S.cfg_link()[j];
                double ddv[3] = {0.0, 0.0, 0.0};
                if (li >= 0 && lj >= 0) {
                    // the upper joint's axis crosses the lower joint's first derivative
                    int up = li, lo_idx = bj;
                    if (!((S.l_anc()[lj] >> li) & 1ull)) { up = lj; lo_idx = ai; }
                    const int kup = S.l_kind()[up];
                    if (kup & 4) {
                        const int a = kup & 3;
                        const double aw[3] = {w[oR + a * nls + up], w[oR + (3 + a) * nls + up], w[oR + (6 + a) * nls + up]};
                        const bool onA = ((ancA >> li) & 1ull) && ((ancA >> lj) & 1ull);
                        const bool onB = ((ancB >> li) & 1ull) && ((ancB >> lj) & 1ull);
                        double d3[3] = {0.0, 0.0, 0.0};
                        if (onA) { TREPB_UNROLL for (int k = 0; k < 3; ++k) d3[k] += DA[k * m + lo_idx]; }
                        if (onB) { TREPB_UNROLL for (int k = 0; k < 3; ++k) d3[k] -= DB[k * m + lo_idx]; }
                        cross3(aw, d3, ddv);
                    }
                }
                double val;
                if (dist) {
                    double di[3], dj[3];
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { di[k] = DA[k * m + ai] - DB[k * m + ai]; dj[k] = DA[k * m + bj] - DB[k * m + bj]; }
                    val = dot3(di, dj) + dot3(v, ddv);
                    if (third == i && third == j) val -= 1.0;
                    val *= 2.0 * lam;
                } else if (plane) {
                    double di[3], dj[3], axi[3], axj[3], dni[3], dnj[3], ddn[3] = {0.0, 0.0, 0.0};
                    TREPB_UNROLL for (int k = 0; k < 3; ++k) { di[k] = DA[k * m + ai] - DB[k * m + ai]; dj[k] = DA[k * m + bj] - DB[k * m + bj]; }
                    const bool ri = plane_axis(c, li, axi), rj = plane_axis(c, lj, axj);
                    cross3(axi, nw, dni);
                    cross3(axj, nw, dnj);
                    if (ri && rj) {
                        // the upper joint's axis crosses the lower joint's derivative of the normal
                        const bool i_up = ((S.l_anc()[lj] >> li) & 1ull) != 0;
                        if (i_up) cross3(axi, dnj, ddn); else cross3(axj, dni, ddn);
                    }
                    val = lam * (dot3(ddn, v) + dot3(dni, dj) + dot3(dnj, di) + dot3(nw, ddv));
                } else {
                    val = lam * sel3(ddv, third);
                }
                if constexpr (D::kStatic) Y[S.dd_row()[j] * ldy + S.dd_col()[i]] += val;
                else Y[j * ldy + i] += val;
            }
        }
        t.sync();
    }

    // ---- MidpointVI_solve_DEL (midpointvi.c:691-747).  In: q1, p1, u1, q2 (start / k2), lam.
    // Out: q2, lam, p2.  Returns the iteration count or a negative Status.  On return the
    // workspace holds the first-order midpoint data of the converged step and (nc > 0) Dh1, Dh2.
    TREPB_HD int solve(double t1, double t2, double tol, int max_it) {
        const int nd = ND(), nc = NC(), nr = nd + nc, nq = NQ(), nu = NU(), lane = t.lane();
        const double dt = t2 - t1;
        int iterations = 0;
        TREPB_TICK_INIT
        if (nc > 0) {
            set_point(1, dt);
            pose_sweep(1);
            points();
            constraints(false, 1);
        }
        TREPB_TICK(16);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (;;) {
            set_point(0, dt);
            if (nc > 0) {
                // midpoint pose + velocities and the q2 pose of the constraint links in one sweep
                pose_sweep(2);
                TREPB_TICK(18);
                use_pose(true);
                points();
                constraints(true, 2);
                use_pose(false);
                TREPB_TICK(17);
            } else {
                pose_sweep(0);
                TREPB_TICK(18);
            }
            dyn_first();
            TREPB_TICK(19);
            // residual (midpointvi.c:533-565); forces: Damping (damping.c:13-22), ConfigForce (configforce.c:13-22)
            for (int j = lane; j < nd; j += Team::kSize) {
                const double fo = ext_force(j);
                double f = w[L.p1 + j] + (0.5 * dt * w[L.Lq + j] - w[L.Lv + j]) + dt * fo;
                for (int c = 0; c < nc; ++c) f -= xv(L.Dh1)[c * nd + j] * w[L.lam + c];
                // p2 = D2L2 of this midpoint (midpointvi.c:742-743, 491-504): the value of the last
                // (converged) evaluation is the step's result; fr aliases Lq, so form it here
                w[L.p2 + j] = 0.5 * dt * w[L.Lq + j] + w[L.Lv + j];
                w[L.fr + j] = f;
            }
            for (int c = lane; c < nc; c += Team::kSize) w[L.fr + nd + c] = w[L.hc + c];
            t.sync();
            // DEL_solved (midpointvi.c:672-689): every lane evaluates the same test
            double nrm = 0.0;
            for (int j = 0; j < nd; ++j) nrm += w[L.fr + j] * w[L.fr + j];
            bool solved = !(sqrt(nrm) > tol);
            for (int c = 0; c < nc; ++c)
                if (fabs(w[L.fr + nd + c]) > S.con_tol()[c]) solved = false;
            TREPB_TICK(20);
            if (solved) break;
            if (iterations > max_it) return ST_NOT_CONVERGED;
            // Jacobian (midpointvi.c:577-670) with the residual as an extra column
            dyn_second(dt);
            TREPB_TICK(21);
            double* A = w + L.N;
            const int ld = L.ldf;
            // constant / constraint parts entry by entry, then the Lagrangian terms scattered from the
            // chain pairs (each (k, i) of the dynamic block receives at most one pair)
            for (int e = lane; e < nr * (nr + 1); e += Team::kSize) {
                const int k = e / (nr + 1), i = e - k * (nr + 1);
                double v = 0.0;
                if (i == nr) v = w[L.fr + k];
                else if (k < nd && i < nd) { if (k == i) v = -S.damp()[k] - 0.25 * dt * ks_eff(k); }
                else if (k < nd) v = -xv(L.Dh1)[(i - nd) * nd + k];
                else if (i < nd) v = xv(L.Dh2)[(k - nd) * nq + i];
                A[k * ld + i] = v;
            }
            t.sync();
            for (int e = lane; e < NPAIRS(); e += Team::kSize) {
                const int ij = S.pair_ij()[e];
                const int ci = S.l_cfg()[ij & 255], cj = S.l_cfg()[ij >> 8];   // ci above-or-equal cj
                if (ci >= nd || cj >= nd) continue;
                const double base = xv(L.VV)[e];   // dt/4 L_dqdq - 1/dt L_ddqddq (dyn_second, Newton form)
                if (ci == cj) A[ci * ld + ci] += base;
                else {
                    const double asym = xv(L.QQ)[e];
                    A[ci * ld + cj] += base + asym;
                    A[cj * ld + ci] += base - asym;
                }
            }
            t.sync();
            if (NPF() > 0) {
                // dt (1/2 f_dq + 1/dt f_ddq) of the LinearDampers and wrenches (midpointvi.c:577-670)
                const int nqf = S.nqf;
                for (int e = lane; e < nqf * nqf; e += Team::kSize) {
                    const int ca = S.xf_cfg()[e / nqf], cb = S.xf_cfg()[e % nqf];
                    if (ca < nd && cb < nd) A[ca * ld + cb] += 0.5 * dt * w[L.FQ + e] + w[L.FV + e];
                }
                t.sync();
            }
            if (NS() > 0) {
                // L_dqdq -= d2V/dqdq of the LinearSprings: cross-chain entries the pair tables do not carry
                const int nqs = S.nqs;
                for (int e = lane; e < nqs * nqs; e += Team::kSize) {
                    const int ca = S.xs_cfg()[e / nqs], cb = S.xs_cfg()[e % nqs];
                    if (ca < nd && cb < nd) A[ca * ld + cb] -= 0.25 * dt * w[L.XS + e];
                }
                t.sync();
            }
            TREPB_TICK(22);
            {
                // factorization + back substitution on the team's first warp (warp collectives)
                bool ok = true;
                if (t.warp() == 0) {
                    bool done = false;
#if defined(__CUDACC__)
                    if constexpr (D::kStatic && Team::Sub::kWarp && D::ND + D::NC <= 32) {
                        RowLU<D::ND + D::NC, 1> lu;
                        lu.load(A, ld, lane & 31);
                        ok = lu.factor(1e-20);
                        TREPB_TICK(23);
                        if (ok) lu.backsolve(0, w + L.fr, lane & 31);
                        done = true;
                    }
#endif
                    if (!done) {
                        ok = team_lu(t.sub(), A, ld, nr, 1, ipivM(), iswpM(), w + L.scl, w + L.rdM, 1e-20);
                        TREPB_TICK(23);
                        if (ok) team_backsolve_vec(t.sub(), A, ld, nr, nr, w + L.rdM, w + L.fr);
                    }
                }
                ok = first_warp_result(ok);
                if (!ok) return ST_SINGULAR;
                t.sync();
            }
            for (int k = lane; k < nd; k += Team::kSize) w[L.q2 + k] -= w[L.fr + k];
            for (int c = lane; c < nc; c += Team::kSize) w[L.lam + c] -= w[L.fr + nd + c];
            t.sync();
            TREPB_TICK(24);
            iterations++;
        }
        return iterations;
    }

    // calc_p2 alone (midpointvi.c:2702-2708)
    TREPB_HD void calc_p2(double dt) {
        set_point(0, dt);
        pose_sweep(0);
        dyn_first();
        for (int j = t.lane(); j < ND(); j += Team::kSize) w[L.p2 + j] = 0.5 * dt * w[L.Lq + j] + w[L.Lv + j];
        t.sync();
    }

    // MidpointVI_calc_f alone (midpointvi.c:533-575): fr[0:nd] = p1 + D1L2 + fm2 - Dh(q1)^T lam, fr[nd:] = h(q2)
    TREPB_HD void calc_f(double dt) {
        const int nd = ND(), nc = NC(), nu = NU(), lane = t.lane();
        if (nc > 0) {
            set_point(1, dt);
            pose_sweep(1);
            points();
            constraints(false, 1);
        }
        set_point(0, dt);
        if (nc > 0) {
            pose_sweep(2);
            use_pose(true);
            points();
            constraints(true, 0);
            use_pose(false);
        } else {
            pose_sweep(0);
        }
        dyn_first();
        double f = 0.0;
        // fr aliases Lq: every lane forms its entries before any is stored
        for (int j = lane; j < nd; j += Team::kSize) {
            const double fo = ext_force(j);
            f = w[L.p1 + j] + (0.5 * dt * w[L.Lq + j] - w[L.Lv + j]) + dt * fo;
            for (int c = 0; c < nc; ++c) f -= xv(L.Dh1)[c * nd + j] * w[L.lam + c];
            w[L.fr + j] = f;
        }
        for (int c = lane; c < nc; c += Team::kSize) w[L.fr + nd + c] = w[L.hc + c];
        t.sync();
    }
    // discrete_fm2 (midpointvi.c:474-478, 2710-2727): dt F(q_mid, dq, u1) -> fr[0:nd]
    TREPB_HD void calc_fm2(double dt) {
        const int nd = ND(), nu = NU();
        set_point(0, dt);
        if (NPF() > 0) {       // dampers and wrenches need the midpoint pose of their end points
            pose_sweep(0);
            points(true);
            for (int j = t.lane(); j < nd; j += Team::kSize) w[L.FD + j] = 0.0;
            t.sync();
            dampers_first();
            wrenches_first();
        }
        for (int j = t.lane(); j < nd; j += Team::kSize) w[L.fr + j] = dt * ext_force(j);
        t.sync();
    }

    // explicit right-hand-side entry of column `col` (nq: q1 | nd: p1 | nu: u1 | nk: k2) at
    // dynamic row j  (midpointvi.c:929-1098: -D1D1L2_D1fm2 + DDh1.lambda | -e | -D3fm2 | -D2D1L2_D2fm2)
    TREPB_HD double rhs_c(int col, int j, double dt, const double* Y, int ldy) const {
        const int nq = NQ(), nd = ND(), nu = NU();
        if (col < nq) {
            const double fv = col == j ? -S.damp()[j] : 0.0;
            double c = -(comb(1, col, j, dt) - fv);   // T11(col, j)
            if (NC() > 0) {
                if constexpr (D::kStatic) {
                    // unconditional load (the column phase issues its 22 gathers back to back: with D::kExt they
                    // come from the L2-resident slab)
                    const int r = S.dd_row()[j], cc = S.dd_col()[col];
                    const bool in = r >= 0 && cc >= 0;
                    const double yv = Y[in ? r * ldy + cc : 0];
                    c += in ? yv : 0.0;
                } else {
                    c += Y[j * ldy + col];
                }
            }
            return c;
        }
        col -= nq;
        if (col < nd) return col == j ? -1.0 : 0.0;
        if (col < nd + nu) return -dt * fu_eff(j, col - nd);
        return -comb(2, nd + (col - nd - nu), j, dt);     // T21(a, j), a kinematic
    }
    // explicit part of d p2 / d (column) at dynamic row j: D1D2L2 for q1 columns, D2D2L2 for k2 columns
    TREPB_HD double rhs_e(int kindv, int i, int j, double dt) const {
        if (kindv != 0 && kindv != 3) return 0.0;
        const int a = kindv == 0 ? i : ND() + i;
        return comb(kindv == 0 ? 3 : 0, a, j, dt);
    }
    TREPB_HD void col_kind(int col, int& kindv, int& i) const {
        const int nq = NQ(), nd = ND(), nu = NU();
        if (col < nq) { kindv = 0; i = col; }
        else if (col < nq + nd) { kindv = 1; i = col - nq; }
        else if (col < nq + nd + nu) { kindv = 2; i = col - nq - nd; }
        else { kindv = 3; i = col - nq - nd - nu; }
    }
    // n doubles of zeros to global memory, 16 bytes per store when the block is 16-byte aligned
    TREPB_HD void fill_zero(double* dst, int n) const {
        const int lane = t.lane();
#if defined(__CUDA_ARCH__)
        if ((((unsigned long long)dst) & 15ull) == 0ull) {
            double2* d2 = (double2*)dst;
            const int n2 = n >> 1;
            for (int e = lane; e < n2; e += Team::kSize) d2[e] = make_double2(0.0, 0.0);
            if ((n & 1) && lane == 0) dst[n - 1] = 0.0;
            return;
        }
#endif
        for (int e = lane; e < n; e += Team::kSize) dst[e] = 0.0;
    }

    // where the results of one right-hand-side column go: raw arrays [wrt][out] and the A / B blocks
    // of DSystem.fdx / fdu (dsystem.py:284-317); resolved once per column
    struct ColOut {
        double *q2o, *p2o, *mq, *mp;   // raw q2_d*, p2_d* rows ; matrix (A or B) entries of the q2 / p2 rows
        int ms;                        // row stride of the matrix
    };
    TREPB_HD ColOut col_out(const Deriv1Out& o, int kindv, int i) const {
        const int nq = NQ(), nd = ND(), nu = NU(), nX = 2 * nq, nU = nu + NK();
        double* q2o = kindv == 0 ? o.q2_dq1 : (kindv == 1 ? o.q2_dp1 : (kindv == 2 ? o.q2_du1 : o.q2_dk2));
        double* p2o = kindv == 0 ? o.p2_dq1 : (kindv == 1 ? o.p2_dp1 : (kindv == 2 ? o.p2_du1 : o.p2_dk2));
        ColOut c;
        c.q2o = q2o ? q2o + i * nd : nullptr;
        c.p2o = p2o ? p2o + i * nd : nullptr;
        double* M = kindv < 2 ? o.A : o.B;
        c.ms = kindv < 2 ? nX : nU;
        const int colm = kindv == 0 ? i : (kindv == 1 ? nq + i : (kindv == 2 ? i : nu + i));
        c.mq = M ? M + colm : nullptr;
        c.mp = M ? M + nq * c.ms + colm : nullptr;
        return c;
    }
    TREPB_HD static void store_col(const ColOut& c, int j, double qv, double pv) {
        if (c.q2o) c.q2o[j] = qv;
        if (c.p2o) c.p2o[j] = pv;
        if (c.mq) { c.mq[j * c.ms] = qv; c.mp[j * c.ms] = pv; }
    }

    // ---- MidpointVI_calc_deriv1 (midpointvi.c:749-1120) right after solve() on the same workspace.
    // Output layout as trepb_math.cuh::deriv1.  aux: optional export for the second-derivative
    // kernel (AuxLayout of trepb_kernels.cuh: M2 LU, M2 piv, PJ LU, PJ piv, Dh1, Dh2, T22).
    TREPB_HD int deriv1(double t1, double t2, const Deriv1Out& o, double* aux, const int* auxo) {
        const int nd = ND(), nk = NK(), nq = NQ(), nc = NC(), nu = NU(), lane = t.lane();
        const int nX = 2 * nq, nU = nu + nk;
        const double dt = t2 - t1;
        double* Y = xw(L.Y);
        const int ldy = L.ldd;   // leading dimension of the DDh.lambda block (== L.ldy for run-time sizes)
        TREPB_TICK_INIT
        dyn_second();   // tables at the converged midpoint
        pair_combos(dt);
        TREPB_TICK(25);
        if (nc > 0) {
            set_point(1, dt);
            pose_sweep(1);
            points();
            ddh_lambda(Y, ldy);
        }
        TREPB_TICK(26);
        // M2 = (D2D1L2 + D2fm2)^T on the dynamic block, with Dh1^T as extra columns; D2D2L2 dynamic block
        double* M2 = w + L.M2;
        double* T22 = w + L.T22;
        const int ldm = L.ldm;
        for (int e = lane; e < nd * nd; e += Team::kSize) {
            const int a = e / nd, b = e - a * nd;
            const double fv = a == b ? -S.damp()[a] : 0.0;
            M2[a * ldm + b] = comb(2, b, a, dt) + fv;   // T21(b, a)
            // T22(a, b): symmetric in the table terms -> same lookup transposed
            T22[b * nd + a] = comb(0, b, a, dt);
        }
        for (int e = lane; e < nd * nc; e += Team::kSize) {
            const int a = e / nc, c = e - a * nc;
            M2[a * ldm + nd + c] = xv(L.Dh1)[c * nd + a];
        }
        t.sync();
        TREPB_TICK(27);
        {
            bool ok = true;
            if (t.warp() == 0) {
                bool done = false;
#if defined(__CUDACC__)
                if constexpr (D::kStatic && Team::Sub::kWarp && D::ND <= 32) {
                    RowLU<D::ND, D::NC> lu;
                    lu.load(M2, ldm, lane & 31);
                    ok = lu.factor(1e-20);
                    t.sub().sync();   // every lane has read its row before rows are written back in pivot order
                    if (ok) lu.store(M2, ldm, w + L.rdM, ipivM(), lane & 31);
                    done = true;
                }
#endif
                if (!done) ok = team_lu(t.sub(), M2, ldm, nd, nc, ipivM(), iswpM(), w + L.scl, w + L.rdM, 1e-20);
            }
            ok = first_warp_result(ok);
            if (!ok) return ST_SINGULAR;
            t.sync();
        }
        double* PJ = w + L.PJ;
        const int ldp = L.ldp;
        if (nc > 0) {
            for (int c = lane; c < nc; c += Team::kSize) col_backsolve(M2, ldm, nd, w + L.rdM, M2 + nd + c, ldm);
            t.sync();
            // proj = -Dh2_d M2^-1 Dh1^T  (midpointvi.c:910-927)
            for (int e = lane; e < nc * nc; e += Team::kSize) {
                const int a = e / nc, b = e - a * nc;
                double s = 0.0;
                for (int k = 0; k < nd; ++k) s += xv(L.Dh2)[a * nq + k] * M2[k * ldm + nd + b];
                PJ[a * ldp + b] = -s;
            }
            t.sync();
            {
                bool ok = true;
                if (t.warp() == 0) ok = team_lu(t.sub(), PJ, ldp, nc, 0, ipivP(), iswpP(), w + L.scl, w + L.rdP, 1e-20);
                ok = first_warp_result(ok);
                if (!ok) return ST_SINGULAR;
                t.sync();
            }
        }
        if (aux) {
            // auxo: o_m2, o_m2p, o_pj, o_pjp, o_dh1, o_dh2, o_t22 ; factors in LU_decomp's convention
            // (unit lower triangle holds the scaled multipliers)
            for (int e = lane; e < nd * nd; e += Team::kSize) {
                const int i = e / nd, j = e - i * nd;
                double v = M2[i * ldm + j];
                if (i > j) v *= w[L.rdM + j];
                aux[auxo[0] + e] = v;
                aux[auxo[6] + e] = T22[e];
            }
            for (int i = lane; i < nd; i += Team::kSize) aux[auxo[1] + i] = (double)ipivM()[i];
            for (int e = lane; e < nc * nc; e += Team::kSize) {
                const int i = e / nc, j = e - i * nc;
                double v = PJ[i * ldp + j];
                if (i > j) v *= w[L.rdP + j];
                aux[auxo[2] + e] = v;
            }
            for (int i = lane; i < nc; i += Team::kSize) aux[auxo[3] + i] = (double)ipivP()[i];
            for (int e = lane; e < nc * nd; e += Team::kSize) {
                aux[auxo[4] + e] = xv(L.Dh1)[e];
                aux[auxo[5] + e] = xv(L.Dh2)[(e / nd) * nq + e % nd];
            }
        }
        TREPB_TICK(28);
        // ---- constant blocks of A and B (dsystem.py:284-317): zeros with wide stores over what the column loop
        // below does not write - the kinematic rows of both halves and, in A's dynamic rows, the nk trailing
        // columns - then the few non-zero constants
        if (o.A) {
            fill_zero(o.A + (long)nd * nX, nk * nX);
            fill_zero(o.A + (long)(nq + nd) * nX, nk * nX);
            for (int e = lane; e < 2 * nd * nk; e += Team::kSize) {
                const int r = e / nk, c = e - r * nk;
                o.A[(long)(r < nd ? r : nq + r - nd) * nX + nq + nd + c] = 0.0;
            }
        }
        if (o.B) {
            fill_zero(o.B + (long)nd * nU, nk * nU);
            fill_zero(o.B + (long)(nq + nd) * nU, nk * nU);
        }
        t.sync();
        for (int i = lane; i < nk; i += Team::kSize) {
            if (o.A) o.A[(nq + nd + i) * nX + nd + i] = -1.0 / dt;
            if (o.B) { o.B[(nd + i) * nU + nu + i] = 1.0; o.B[(nq + nd + i) * nU + nu + i] = 1.0 / dt; }
        }
        TREPB_TICK(30);
        // ---- right-hand sides, one column per lane: q1 (nq) | p1 (nd) | u1 (nu) | k2 (nk)
        const int ncols = nq + nd + nu + nk;
        if constexpr (D::kStatic) {
            // register-resident columns: the permuted explicit part is built directly
            // (y[i] = c[piv[i]]), everything else is constant-index arithmetic
            constexpr int N = D::ND, C = D::NC > 0 ? D::NC : 1;
            const int* pivM = ipivM();
            const int* pivP = ipivP();
            for (int col = lane; col < ncols; col += Team::kSize) {
                int kindv, ci;
                col_kind(col, kindv, ci);
                double y[N];
                TREPB_UNROLL for (int i = 0; i < N; ++i) y[i] = rhs_c(col, pivM[i], dt, Y, ldy);
                reg_solve<N>(M2, ldm, w + L.rdM, y);
                double z[C];
                if (D::NC > 0) {
                    double zz[C];
                    TREPB_UNROLL
                    for (int c = 0; c < C; ++c) zz[c] = kindv == 3 ? xv(L.Dh2)[c * nq + nd + ci] : 0.0;
                    TREPB_UNROLL
                    for (int j = 0; j < N; ++j) {
                        TREPB_UNROLL for (int c = 0; c < C; ++c) zz[c] += xv(L.Dh2)[c * nq + j] * y[j];
                    }
                    TREPB_UNROLL
                    for (int c = 0; c < C; ++c) {
                        const int src = pivP[c];
                        double v = 0.0;
                        TREPB_UNROLL for (int k = 0; k < C; ++k) v = src == k ? zz[k] : v;
                        z[c] = v;
                    }
                    reg_solve<C>(PJ, ldp, w + L.rdP, z);
                    TREPB_UNROLL
                    for (int c = 0; c < C; ++c) {
                        TREPB_UNROLL for (int j = 0; j < N; ++j) y[j] += M2[j * ldm + nd + c] * z[c];
                    }
                }
                double pv[N];
                TREPB_UNROLL for (int j = 0; j < N; ++j) pv[j] = rhs_e(kindv, ci, j, dt);
                TREPB_UNROLL
                for (int k = 0; k < N; ++k) {
                    TREPB_UNROLL for (int j = 0; j < N; ++j) pv[j] += T22[k * nd + j] * y[k];
                }
                const ColOut co = col_out(o, kindv, ci);
                TREPB_UNROLL for (int j = 0; j < N; ++j) store_col(co, j, y[j], pv[j]);
                double* l1o = kindv == 0 ? o.l1_dq1 : (kindv == 1 ? o.l1_dp1 : (kindv == 2 ? o.l1_du1 : o.l1_dk2));
                if (D::NC > 0 && l1o) { TREPB_UNROLL for (int c = 0; c < C; ++c) l1o[ci * nc + c] = z[c]; }
            }
        } else {
            // run-time sizes: columns staged in shared memory, two groups sharing the Y block
            double* Z = w + L.Z;
            for (int grp = 0; grp < 2; ++grp) {
                const int c0 = grp == 0 ? 0 : nq, gcols = grp == 0 ? nq : ncols - nq;
                for (int e = lane; e < nd * gcols; e += Team::kSize) {
                    const int j = e / gcols, col = e - j * gcols;
                    Y[j * ldy + col] = rhs_c(c0 + col, j, dt, Y, ldy);
                }
                t.sync();
                for (int col = lane; col < gcols; col += Team::kSize) {
                    int kindv, ci;
                    col_kind(c0 + col, kindv, ci);
                    double* y = Y + col;
                    col_solve(M2, ldm, nd, iswpM(), w + L.rdM, y, ldy);
                    double* z = Z + col;
                    if (nc > 0) {
                        for (int c = 0; c < nc; ++c) {
                            double s = 0.0;
                            for (int j = 0; j < nd; ++j) s += xv(L.Dh2)[c * nq + j] * y[j * ldy];
                            if (kindv == 3) s += xv(L.Dh2)[c * nq + nd + ci];
                            z[c * ldy] = s;
                        }
                        col_solve(PJ, ldp, nc, iswpP(), w + L.rdP, z, ldy);
                        for (int j = 0; j < nd; ++j) {
                            double s = y[j * ldy];
                            for (int c = 0; c < nc; ++c) s += M2[j * ldm + nd + c] * z[c * ldy];
                            y[j * ldy] = s;
                        }
                    }
                    const ColOut co = col_out(o, kindv, ci);
                    for (int j = 0; j < nd; ++j) {
                        double pv = rhs_e(kindv, ci, j, dt);
                        for (int k = 0; k < nd; ++k) pv += T22[k * nd + j] * y[k * ldy];
                        store_col(co, j, y[j * ldy], pv);
                    }
                    double* l1o = kindv == 0 ? o.l1_dq1 : (kindv == 1 ? o.l1_dp1 : (kindv == 2 ? o.l1_du1 : o.l1_dk2));
                    if (l1o) for (int c = 0; c < nc; ++c) l1o[ci * nc + c] = z[c * ldy];
                }
                t.sync();
            }
        }
        TREPB_TICK(29);
        t.sync();
        TREPB_TICK(30);
        return ST_OK;
    }
};

}  // namespace trepb
