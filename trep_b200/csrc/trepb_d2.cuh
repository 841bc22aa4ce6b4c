// Second derivatives of the discrete flow (MidpointVI_calc_deriv2, trep/_trep/midpointvi.c:2516-2545):
//   q2, p2, lambda1  _d{q1q1, q1p1, q1u1, q1k2, p1p1, p1u1, p1k2, u1u1, u1k2, k2k2}
// in the reference's storage layout [wrt A][wrt B][output] (trep/_trep/trep.h:439-473).
//
// Formulation (not the reference's): with x = (q1, p1, u1, k2) the parameters and y = (q2_dyn,
// lambda) the unknowns of
//     F1 = p1 + D1L2(q1,q2) + fm2(q1,q2,u1) - Dh(q1)^T lambda = 0 ,   F2 = h(q2) = 0 ,
// the second-order implicit-function identity for a pair of parameters (s,t) is
//     F_y y_st + D^2F[xi_s, xi_t] = 0 ,   xi_s = (x_s, y_s)  the first-order tangent of deriv1,
// and  p2_st = D^2 H[xi_s, xi_t] + H_q2 q2_st  for  p2 = H = D2L2(q1,q2).
// D^2F[xi_s, xi_t] and D^2H[xi_s, xi_t] are obtained exactly by evaluating the first-order residual
// path of trepb_math.cuh on hyper-dual numbers (trepb_hd.h) along (xi_s, xi_t); the linear solve
// reuses deriv1's factorizations M2 = LU, proj = LU exactly as calc_deriv1 does for its right-hand
// sides (midpointvi.c:929-1098).  One thread per (instance, pair s<=t): the pairs of an instance are
// independent, which is the parallelism the reference spends its pthread row pool on
// (midpointvi.c:2536-2541).
#pragma once
// (included at the end of trepb_kernels.cuh)

namespace trepb {

struct AuxMat {
    const double* p; int ld;
    TREPB_HD double operator()(int i, int j) const { return p[i * ld + j]; }
};
struct AuxVec {
    const double* p;
    TREPB_HD double operator()(int i) const { return p[i]; }
};
// scratch vectors live in the components of hyper-dual workspace slots that are free after the
// residual evaluation
template <class Ws> struct FrV  { Ws* w; TREPB_HD double& operator()(int i) const { return w->fr(i).v; } };
template <class Ws> struct FrA  { Ws* w; TREPB_HD double& operator()(int i) const { return w->fr(i).a; } };
template <class Ws> struct FrB  { Ws* w; TREPB_HD double& operator()(int i) const { return w->fr(i).b; } };
template <class Ws> struct FrAB { Ws* w; TREPB_HD double& operator()(int i) const { return w->fr(i).ab; } };
template <class Ws> struct HcV  { Ws* w; TREPB_HD double& operator()(int i) const { return w->hc(i).v; } };
template <class Ws> struct HcA  { Ws* w; TREPB_HD double& operator()(int i) const { return w->hc(i).a; } };

// type (0 q1, 1 p1, 2 u1, 3 k2) and index within the type of parameter s
TREPB_HD void split_param(int s, int nq, int nd, int nu, int* type, int* idx) {
    if (s < nq) { *type = 0; *idx = s; }
    else if (s < nq + nd) { *type = 1; *idx = s - nq; }
    else if (s < nq + nd + nu) { *type = 2; *idx = s - nq - nd; }
    else { *type = 3; *idx = s - nq - nd - nu; }
}

// One (instance b, parameter pair s <= t): evaluates the residual on hyper-duals, solves, stores.
template <class Sys, class Ws>
TREPB_HD void deriv2_pair(const Sys& sys, Ws& ws, const D2Params& p, long b, int s, int t) {
    const int nd = sys.ND(), nk = sys.NK(), nq = nd + nk, nu = sys.NU(), nc = sys.NC();
    int ts, is, tt, it;
    split_param(s, nq, nd, nu, &ts, &is);
    split_param(t, nq, nd, nu, &tt, &it);
    const int cnt_s = ts == 0 ? nq : (ts == 1 ? nd : (ts == 2 ? nu : nk));
    const int cnt_t = tt == 0 ? nq : (tt == 1 ? nd : (tt == 2 ? nu : nk));
    const double* zs = p.q2_d[ts] + ((long)b * cnt_s + is) * nd;   // d q2_dyn / d s
    const double* zt = p.q2_d[tt] + ((long)b * cnt_t + it) * nd;
    const double* ls = nc ? p.l1_d[ts] + ((long)b * cnt_s + is) * nc : nullptr;
    const double* lt = nc ? p.l1_d[tt] + ((long)b * cnt_t + it) * nc : nullptr;
    // ---- point z + e1 xi_s + e2 xi_t
    TREPB_UNROLL_SYS
    for (int i = 0; i < nq; ++i) {
        ws.q1(i) = HD(p.q1[b * nq + i], (ts == 0 && is == i) ? 1.0 : 0.0, (tt == 0 && it == i) ? 1.0 : 0.0, 0.0);
        double a, bb;
        if (i < nd) { a = zs[i]; bb = zt[i]; }
        else { a = (ts == 3 && is == i - nd) ? 1.0 : 0.0; bb = (tt == 3 && it == i - nd) ? 1.0 : 0.0; }
        ws.q2(i) = HD(p.q2[b * nq + i], a, bb, 0.0);
    }
    TREPB_UNROLL_SYS
    for (int i = 0; i < nu; ++i)
        ws.u1(i) = HD(p.u1[b * nu + i], (ts == 2 && is == i) ? 1.0 : 0.0, (tt == 2 && it == i) ? 1.0 : 0.0, 0.0);
    TREPB_UNROLL_SYS
    for (int i = 0; i < nc; ++i) ws.lam(i) = HD(p.lam[b * nc + i], ls[i], lt[i], 0.0);
    const double t1 = p.t1 ? p.t1[b] : p.t1s;
    const double t2 = p.t2 ? p.t2[b] : (t1 + p.dts);
    const double dt = t2 - t1;
    // ---- residual along the two tangents (same call sequence as calc_f, midpointvi.c:533-565)
    if (nc > 0) {
        set_point(sys, ws, 1, dt);
        pass1(sys, ws, false, true);
        constraints_eval(sys, ws, 2, 1);
    }
    eval_mid(sys, ws, dt, 1);
    TREPB_UNROLL_SYS
    for (int j = 0; j < nd; ++j) {
        HD f = (0.5 * dt * ws.Lq(j) - ws.Lv(j)) + dt * ws.Fo(j);
        TREPB_UNROLL_SYS
        for (int cc = 0; cc < nc; ++cc) f -= ws.Dh1(cc, j) * ws.lam(cc);
        const HD h2 = 0.5 * dt * ws.Lq(j) + ws.Lv(j);
        ws.fr(j).v = -f.ab;    // c = -R1
        ws.p2(j).v = h2.ab;    // D^2 H [xi_s, xi_t]
    }
    if (nc > 0) {
        set_point(sys, ws, 2, dt);
        pass1(sys, ws, false, true);
        constraints_eval(sys, ws, 1, 2);
        TREPB_UNROLL_SYS
        for (int cc = 0; cc < nc; ++cc) ws.hc(cc).v = ws.hc(cc).ab;   // R2
    }
    // ---- solve (calc_deriv1's scheme): lambda_st = proj^-1 (Dh2 M2^-1 c + R2),
    //      q2_st = M2^-1 (c + Dh1^T lambda_st),  p2_st = D^2H + D2D2L2^T q2_st
    const double* aux = p.aux + (long)b * p.auxl.size;
    const AuxMat M2{aux + p.auxl.o_m2, nd}, PJ{aux + p.auxl.o_pj, nc}, T22{aux + p.auxl.o_t22, nd};
    const AuxMat Dh1{aux + p.auxl.o_dh1, nd}, Dh2{aux + p.auxl.o_dh2, nd};
    const AuxVec M2p{aux + p.auxl.o_m2p}, PJp{aux + p.auxl.o_pjp};
    TREPB_UNROLL_SYS
    for (int j = 0; j < nd; ++j) { ws.fr(j).a = ws.fr(j).v; ws.fr(j).b = ws.fr(j).v; }   // tnd = col = c
    if (nc > 0) {
        lu_solve<Sys>(M2, nd, M2p, FrA<Ws>{&ws}, FrAB<Ws>{&ws});
        TREPB_UNROLL_SYS
        for (int cc = 0; cc < nc; ++cc) {
            double sacc = 0.0;
            TREPB_UNROLL_SYS
            for (int j = 0; j < nd; ++j) sacc += Dh2(cc, j) * ws.fr(j).a;
            ws.hc(cc).v = sacc + ws.hc(cc).v;
        }
        lu_solve<Sys>(PJ, nc, PJp, HcV<Ws>{&ws}, HcA<Ws>{&ws});
        TREPB_UNROLL_SYS
        for (int j = 0; j < nd; ++j) {
            double sacc = ws.fr(j).b;
            TREPB_UNROLL_SYS
            for (int cc = 0; cc < nc; ++cc) sacc += Dh1(cc, j) * ws.hc(cc).v;
            ws.fr(j).b = sacc;
        }
    }
    lu_solve<Sys>(M2, nd, M2p, FrB<Ws>{&ws}, FrAB<Ws>{&ws});
    // ---- store: kind index of (type s <= type t)
    const int kind = ts == 0 ? tt : (ts == 1 ? 3 + tt : (ts == 2 ? 5 + tt : 9));
    const bool mirror = (ts == tt) && (is != it);
    double* oq = p.out[0][kind];
    double* op = p.out[1][kind];
    double* ol = p.out[2][kind];
    const long base_q = (long)b * cnt_s * cnt_t;
    TREPB_UNROLL_SYS
    for (int j = 0; j < nd; ++j) {
        const double qv = ws.fr(j).b;
        double pv = ws.p2(j).v;
        TREPB_UNROLL_SYS
        for (int k = 0; k < nd; ++k) pv += T22(k, j) * ws.fr(k).b;
        if (oq) {
            oq[((base_q + (long)is * cnt_t + it)) * nd + j] = qv;
            if (mirror) oq[((base_q + (long)it * cnt_t + is)) * nd + j] = qv;
        }
        if (op) {
            op[((base_q + (long)is * cnt_t + it)) * nd + j] = pv;
            if (mirror) op[((base_q + (long)it * cnt_t + is)) * nd + j] = pv;
        }
    }
    if (p.z) {
        // z-contraction in the DSystem layout: X = [Q; p; v], U = [u; rho]  (dsystem.py:320-386)
        const int nX = 2 * nq, nU = nu + nk;
        const double* z = p.z + (long)b * nX;
        double acc = 0.0;
        TREPB_UNROLL_SYS
        for (int j = 0; j < nd; ++j) {
            double pv = ws.p2(j).v;
            TREPB_UNROLL_SYS
            for (int k = 0; k < nd; ++k) pv += T22(k, j) * ws.fr(k).b;
            acc += z[j] * ws.fr(j).b + z[nq + j] * pv;
        }
        const bool sx = ts < 2, tx = tt < 2;
        const int xs = ts == 0 ? is : (ts == 1 ? nq + is : (ts == 2 ? is : nu + is));
        const int xt = tt == 0 ? it : (tt == 1 ? nq + it : (tt == 2 ? it : nu + it));
        if (sx && tx) {
            if (p.zxx) { p.zxx[((long)b * nX + xs) * nX + xt] = acc; p.zxx[((long)b * nX + xt) * nX + xs] = acc; }
        } else if (sx) {
            if (p.zxu) p.zxu[((long)b * nX + xs) * nU + xt] = acc;
        } else {
            if (p.zuu) { p.zuu[((long)b * nU + xs) * nU + xt] = acc; p.zuu[((long)b * nU + xt) * nU + xs] = acc; }
        }
    }
    if (ol) {
        TREPB_UNROLL_SYS
        for (int cc = 0; cc < nc; ++cc) {
            const double lv = ws.hc(cc).v;
            ol[((base_q + (long)is * cnt_t + it)) * nc + cc] = lv;
            if (mirror) ol[((base_q + (long)it * cnt_t + is)) * nc + cc] = lv;
        }
    }
}

#if defined(__CUDACC__)
template <class Sys, bool S = Sys::kStatic>
struct CtxHD;
template <class Sys>
struct CtxHD<Sys, true> {
    Sys sys;
    WsStatic<Sys, HD> ws;
    __device__ __forceinline__ CtxHD(const RtSys&, const char*, int, const WsStridedT<HD>&, long, long) {}
};
template <class Sys>
struct CtxHD<Sys, false> {
    RtSys sys;
    WsStridedT<HD> ws;
    __device__ __forceinline__ CtxHD(const RtSys& s, const char* dblob, int blob_bytes, const WsStridedT<HD>& w,
                                     long tid, long nthreads) {
        extern __shared__ double smem_[];
        const int n8 = (blob_bytes + 7) / 8;
        const double* src = (const double*)dblob;
        for (int i = threadIdx.x; i < n8; i += blockDim.x) smem_[i] = src[i];
        __syncthreads();
        sys = s.rebased(dblob, (const char*)smem_);
        ws = w;
        ws.base = w.base + tid;
        ws.stride = nthreads;
    }
};

template <class Sys>
__global__ void __launch_bounds__(128)
d2_kernel(const RtSys rsys, const char* dblob, int blob_bytes, const WsStridedT<HD> wsp, const D2Params p) {
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nth = (long)gridDim.x * blockDim.x;
    CtxHD<Sys> c(rsys, dblob, blob_bytes, wsp, tid, nth);
    auto& sys = c.sys;
    auto& ws = c.ws;
    using Ws = typename std::remove_reference<decltype(ws)>::type;
    const int nd = sys.ND(), nk = sys.NK(), nq = nd + nk, nu = sys.NU(), nc = sys.NC();
    const long total = p.batch * (long)p.npairs;
    for (long g = tid; g < total; g += nth) {
        const long b = g / p.npairs;
        if (p.status && p.status[b] != 0) continue;
        int pr = (int)(g - b * p.npairs), s = 0;
        while (pr >= p.nx - s) { pr -= p.nx - s; ++s; }
        deriv2_pair(sys, ws, p, b, s, s + pr);
    }
}

template <class Sys>
struct LaunchersD2 {
    static cudaError_t run(const LaunchCfg& c, const WsStridedT<HD>& w, const D2Params& p) {
        RtSys rs{};
        if (c.sys) rs = *c.sys;
        if (c.smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute((const void*)d2_kernel<Sys>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
            if (e != cudaSuccess) return e;
        }
        d2_kernel<Sys><<<c.grid, c.block, c.smem, c.stream>>>(rs, c.dblob, c.blob_bytes, w, p);
        return cudaGetLastError();
    }
    static cudaError_t occupancy(int block, size_t smem, int* blocks_per_sm, KernelInfo* info) {
        const void* fn = (const void*)d2_kernel<Sys>;
        cudaFuncAttributes a;
        cudaError_t e = cudaFuncGetAttributes(&a, fn);
        if (e != cudaSuccess) return e;
        if (info) { info->regs = a.numRegs; info->max_threads = a.maxThreadsPerBlock; info->static_smem = a.sharedSizeBytes; info->local_bytes = a.localSizeBytes; }
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, block, smem);
    }
};
#endif

}  // namespace trepb
