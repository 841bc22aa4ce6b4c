// Time-varying discrete LQR on the device: trep.discopt.dlqr.solve_tv_lqr (trep/discopt/dlqr.py:9-38),
// the backward Riccati sweep behind DSystem.calc_feedback_controller (trep/discopt/dsystem.py:474-494):
//
//     P = Q(K)
//     for k = K-1 .. 0:   gamma = R(k) + B^T P B ;  Kp = B^T P A ;  K[k] = gamma^-1 Kp
//                         P = Q(k) + A^T P A - Kp^T K[k] ;  P = (P + P^T) / 2
//
// One CTA per rollout (the sweep is sequential in k, rollouts are independent), P, P A and A[k] in
// shared memory (3 x nX^2 doubles: 154 KB for the marionette's nX = 80), A[k] / B[k] streamed from the
// slabs the linearize kernel wrote, so the linearization never leaves the GPU; K[k] goes straight into
// the layout trepb_project_batch reads (per-rollout gains).  Every product has the form
// C = X^T Y with both operands read along rows (P is symmetric): on the FP64 tensor cores
// (mma.sync m8n8k4, one warp per 16 x 40 tile, ragged edges read as zero) for nX >= 32, on 4 x 4 register
// tiles per thread below that; K[k] = gamma^-1 Kp (and the affine term) by a Gauss-Jordan elimination of
// [gamma | Kp | c] shared by the whole CTA; A[k-1], B[k-1] and Q(k) are fetched with cp.async while the
// current step is still multiplying.  One CTA of 16 warps per SM at the marionette's size, several smaller
// CTAs per SM for small states.
#include <cuda_runtime.h>
#include "trepb_nvtx.h"
#include <stdlib.h>
#include <string>
#include "../../include/trepb.h"
#include "trepb_coop_math.cuh"
#include "trepb_err.h"

namespace trepb {
namespace {

struct LqrParams {
    long long batch;
    int K, nX, nU;
    const double *A, *B, *Q, *R;
    int q_per_step, r_per_step;
    double *Kfb, *P0;
    int* status;
    // affine-quadratic extension (solve_tv_lq, dlqr.py:41-81); all null / zero for the plain LQR
    const double *S, *qv, *rv;         // cross term [K][nX][nU] (may be null), linear costs [K+1][nX], [K][nU]
    long long Qs, Rs, Ss, qs, rs;      // per-rollout strides of Q, R, S, q, r (0: shared by the batch)
    double *C, *b0;                    // [batch][K][nU], [batch][nX]
};

// C[i][j] (op)= sum_m Xt[m][i] * Y[m][j]   for i < M, j < N, m < Kd ; 4 x 4 tiles over the CTA.
//   mode 0: C = acc   1: C += acc   2: C -= acc
// VEC: ldx, ldy even and both operands 16-byte aligned -> the four operand values of a tile row come
// in as two 128-bit shared-memory loads each.
template <bool VEC>
__device__ __forceinline__ void gemm_tn(double* C, int ldc, const double* Xt, int ldx, const double* Y, int ldy,
                                        int M, int N, int Kd, int mode, int t0 = 0, int nt = 0) {
    // threads t0 .. t0+nt-1 of the CTA share the tiles (default: the whole CTA)
    if (nt == 0) nt = blockDim.x;
    const int me = (int)threadIdx.x - t0;
    if (me < 0 || me >= nt) return;
    const int ti = (M + 3) / 4, tj = (N + 3) / 4;
    for (int tile = me; tile < ti * tj; tile += nt) {
        const int i0 = (tile / tj) * 4, j0 = (tile % tj) * 4;
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        const bool full = i0 + 4 <= M && j0 + 4 <= N;
        if (full) {
            const double* xp = Xt + i0;
            const double* yp = Y + j0;
#pragma unroll 4
            for (int m = 0; m < Kd; ++m) {
                double x[4], y[4];
                if (VEC) {
                    const double2 x01 = *reinterpret_cast<const double2*>(xp), x23 = *reinterpret_cast<const double2*>(xp + 2);
                    const double2 y01 = *reinterpret_cast<const double2*>(yp), y23 = *reinterpret_cast<const double2*>(yp + 2);
                    x[0] = x01.x; x[1] = x01.y; x[2] = x23.x; x[3] = x23.y;
                    y[0] = y01.x; y[1] = y01.y; y[2] = y23.x; y[3] = y23.y;
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) { x[a] = xp[a]; y[a] = yp[a]; }
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] += x[a] * y[b];
                xp += ldx; yp += ldy;
            }
        } else {
            for (int m = 0; m < Kd; ++m) {
                double x[4], y[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    x[a] = i0 + a < M ? Xt[m * ldx + i0 + a] : 0.0;
                    y[a] = j0 + a < N ? Y[m * ldy + j0 + a] : 0.0;
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] += x[a] * y[b];
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (i0 + a < M && j0 + b < N) {
                    double* c = C + (i0 + a) * ldc + j0 + b;
                    if (mode == 0) *c = acc[a][b];
                    else if (mode == 1) *c += acc[a][b];
                    else *c -= acc[a][b];
                }
            }
    }
}

// The same product on the FP64 tensor-core path: mma.sync.aligned.m8n8k4 (DMMA), one warp per tile of
// (8 TI) x (8 TJ) outputs, TI + TJ operand fragments loaded per TI x TJ instructions of 256 fused
// multiply-adds.  Fragment layout (g = lane / 4, tg = lane % 4): A[g][tg] = Xt[m0 + tg][i + g],
// B[tg][g] = Y[m0 + tg][j + g], C[g][2 tg + {0, 1}].  Rows / columns beyond M, N, Kd read as zero and are
// not stored, so any shape works (nU = 18 runs as 24).  Warps w0 .. w0+nw-1 of the CTA share the tiles.
template <int TI, int TJ>
__device__ __forceinline__ void gemm_tn_dmma(double* C, int ldc, const double* Xt, int ldx, const double* Y, int ldy,
                                             int M, int N, int Kd, int mode, int w0, int nw) {
    const int warp = (int)(threadIdx.x >> 5) - w0, lane = threadIdx.x & 31;
    if (warp < 0 || warp >= nw) return;
    const int g = lane >> 2, tg = lane & 3;
    const int tm = (M + 8 * TI - 1) / (8 * TI), tn = (N + 8 * TJ - 1) / (8 * TJ);
    for (int tile = warp; tile < tm * tn; tile += nw) {
        const int i0 = (tile / tn) * 8 * TI, j0 = (tile % tn) * 8 * TJ;
        double c[TI][TJ][2];
        bool xin[TI], yin[TJ];
#pragma unroll
        for (int a = 0; a < TI; ++a) xin[a] = i0 + 8 * a + g < M;
#pragma unroll
        for (int b = 0; b < TJ; ++b) yin[b] = j0 + 8 * b + g < N;
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) c[a][b][0] = c[a][b][1] = 0.0;
        const double* xp = Xt + tg * ldx + i0 + g;
        const double* yp = Y + tg * ldy + j0 + g;
        for (int m0 = 0; m0 < Kd; m0 += 4) {
            const bool in = m0 + tg < Kd;
            double x[TI], y[TJ];
#pragma unroll
            for (int a = 0; a < TI; ++a) x[a] = (in && xin[a]) ? xp[8 * a] : 0.0;
#pragma unroll
            for (int b = 0; b < TJ; ++b) y[b] = (in && yin[b]) ? yp[8 * b] : 0.0;
#pragma unroll
            for (int a = 0; a < TI; ++a)
#pragma unroll
                for (int b = 0; b < TJ; ++b)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[a][b][0]), "+d"(c[a][b][1]) : "d"(x[a]), "d"(y[b]));
            xp += 4 * ldx; yp += 4 * ldy;
        }
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) {
                const int i = i0 + 8 * a + g, j = j0 + 8 * b + 2 * tg;
                if (i >= M) continue;
                double* cp = C + i * ldc + j;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (j + h >= N) continue;
                    if (mode == 0) cp[h] = c[a][b][h];
                    else if (mode == 1) cp[h] += c[a][b][h];
                    else cp[h] -= c[a][b][h];
                }
            }
    }
}

// global -> shared copy that does not pass through registers (cp.async): issued early, waited for late,
// so the next step's A[k], B[k] and Q(k) arrive while the current step is still multiplying
__device__ __forceinline__ void async_copy(double* dst, const double* src, int n) {
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
    if ((((unsigned long long)src | (unsigned long long)d0) & 15ull) == 0ull && (n & 1) == 0) {
        for (int e = threadIdx.x; e < n / 2; e += blockDim.x)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * e), "l"(src + 2 * e) : "memory");
    } else {
        for (int e = threadIdx.x; e < n; e += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + 8u * e), "l"(src + e) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// K = gamma^-1 [Kp | c] by Gauss-Jordan elimination with implicitly scaled partial pivoting, the whole CTA
// on one system M = [gamma | Kp | c] (n rows, nc columns, leading dimension ldt): its elements are dealt
// out to the threads (at most kGJ each; GjElems holds their offsets, computed once per kernel), every warp
// finds the pivot of step k on its own (lane = row, two max-reductions over the bit pattern of |a| scale,
// lowest row among equals), every element outside the pivot row and right of column k takes one fused
// multiply-add, one __syncthreads per step.  No row is moved: order[k] = pivot row of step k, the solution
// row k is row order[k] divided by its pivot.  (The sweep is sequential in k and this solve sits on its
// critical path: a one-warp LU plus one right-hand side per thread took 55 % of a Riccati step.)
// Returns false if a scaled pivot is <= tol.
constexpr int kGJ = 4;
struct GjElems {
    int off[kGJ];   // i * ldt + j, or -1
    int row[kGJ];   // i * ldt
    int col[kGJ];   // j
};
__device__ __forceinline__ GjElems gj_elems(int n, int nc, int ldt) {
    GjElems g;
#pragma unroll
    for (int q = 0; q < kGJ; ++q) {
        const int e = (int)threadIdx.x + q * (int)blockDim.x;
        const bool in = e < n * nc;
        const int i = in ? e / nc : 0, j = e - i * nc;
        g.off[q] = in ? i * ldt + j : -1;
        g.row[q] = i * ldt;
        g.col[q] = j;
    }
    return g;
}
__device__ __forceinline__ bool cta_gauss_jordan(double* M, int ldt, int n, const GjElems& g, double* scl, double* rd,
                                                 int* order, double tol) {
    const int tid = threadIdx.x, lane = tid & 31, nth = blockDim.x;
    for (int i = tid; i < n; i += nth) {
        double sm_ = -1.0;
        for (int j = 0; j < n; ++j) { const double a = fabs(M[i * ldt + j]); if (a > sm_) sm_ = a; }
        scl[i] = 1.0 / sm_;
    }
    __syncthreads();
    // rows already used as pivots: a bit mask every thread keeps for itself (all threads compute the same pivots);
    // a flag array in shared memory would be written by one warp while a slower one still searches this step's pivot
    unsigned long long used = 0ull;
    for (int k = 0; k < n; ++k) {
        // pivot of column k among the rows not used yet (every warp computes the same answer)
        double best = 0.0;
        int br = 0x7fffffff;
        for (int r = lane; r < n; r += 32) {
            double v = ((used >> r) & 1ull) ? 0.0 : fabs(M[r * ldt + k] * scl[r]);
            if (!(v == v)) v = 0.0;
            if (v > best) { best = v; br = r; }
        }
        const unsigned long long bits = (unsigned long long)__double_as_longlong(best);
        const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
        const int pr = (int)__reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)br : 0x7fffffffu);
        const double bestv = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
        if (!(bestv > tol)) return false;
        const double* prow = M + pr * ldt;
        const double rdk = 1.0 / prow[k];
#pragma unroll
        for (int q = 0; q < kGJ; ++q) {
            if (g.off[q] < 0 || g.row[q] == pr * ldt || g.col[q] <= k) continue;
            M[g.off[q]] -= M[g.row[q] + k] * (rdk * prow[g.col[q]]);
        }
        used |= 1ull << pr;
        if (tid == 0) { order[k] = pr; rd[k] = rdk; }
        __syncthreads();
    }
    return true;
}

template <bool VEC, bool MMA>
__global__ void __launch_bounds__(512, 1) lqr_kernel(const LqrParams p) {
    extern __shared__ __align__(16) double sm[];
    const int nX = p.nX, nU = p.nU, K = p.K;
    const int nct = nU + nX + 1, ldg = nct | 1;    // columns / (odd) leading dimension of M = [gamma | Kp | c]
    const GjElems gje = gj_elems(nU, nct, ldg);
    double* P = sm;                                // [nX][nX]
    double* T = P + nX * nX;                       // P A
    double* As = T + nX * nX;                      // A[k]
    double* Bs = As + nX * nX;                     // B[k]            [nX][nU]
    double* W = Bs + nX * nU;                      // P B [nX][nU]
    double* Kp = W + nX * nU;                      // Kp = B^T P A    [nU][nX]
    double* Rs = Kp + nU * nX;                     // R(k)            [nU][nU]
    double* G = Rs + ((nU * nU + 1) & ~1);         // M = [gamma | Kp | c] [nU][ldt], eliminated in place
    double* scl = G + ((nU * ldg + 1) & ~1);       // [nU]
    double* rd = scl + nU;                         // [nU]
    int* order = (int*)(rd + nU);                  // [nU] pivot row of every elimination step (+ nU ints of padding)
    double* Ks = (double*)(order + 2 * nU);        // K[k] [nU][nX]   (2 nU ints = nU doubles: stays 8-byte aligned)
    // affine part (solve_tv_lq): b [nX], A^T b [nX], g = B^T b [nU] (+ r(k): the last column of M)
    double* bv = Ks + nU * nX;
    double* ab = bv + nX;
    double* gv = ab + nX;
    const bool affine = p.qv != nullptr;
    __shared__ int s_fail;
    for (long r = blockIdx.x; r < p.batch; r += gridDim.x) {
        const double* Ar = p.A + r * (long)K * nX * nX;
        const double* Br = p.B + r * (long)K * nX * nU;
        const double* Qr = p.Q + r * p.Qs;
        const double* Rr = p.R + r * p.Rs;
        const double* Sr = p.S ? p.S + r * p.Ss : nullptr;
        const double* qr = affine ? p.qv + r * p.qs : nullptr;
        const double* rr = affine ? p.rv + r * p.rs : nullptr;
        double* Kr = p.Kfb + r * (long)K * nU * nX;
        if (threadIdx.x == 0) s_fail = 0;
        async_copy(P, Qr + (p.q_per_step ? (long)K * nX * nX : 0), nX * nX);     // P = Q(K)
        async_copy(As, Ar + (long)(K - 1) * nX * nX, nX * nX);
        async_copy(Bs, Br + (long)(K - 1) * nX * nU, nX * nU);
        async_copy(Rs, Rr + (p.r_per_step ? (long)(K - 1) * nU * nU : 0), nU * nU);
        if (affine) for (int c = threadIdx.x; c < nX; c += blockDim.x) bv[c] = qr[(long)K * nX + c];   // b = q[K]
        for (int k = K - 1; k >= 0; --k) {
            async_wait();                                          // A[k], B[k], R(k) (and P) have landed
            __syncthreads();
            if (affine) {
                // A^T b and g = B^T b with the old b (columns of A / B, one per thread)
                for (int c = threadIdx.x; c < nX + nU; c += blockDim.x) {
                    double acc = 0.0;
                    if (c < nX) { for (int m = 0; m < nX; ++m) acc += As[m * nX + c] * bv[m]; ab[c] = acc; }
                    else { const int i = c - nX; for (int m = 0; m < nX; ++m) acc += Bs[m * nU + i] * bv[m]; gv[i] = acc; }
                }
            }
            const int nwarps = blockDim.x >> 5;
            if (MMA) {
                // tensor-core schedule (16 warps): T = P A on warps 0-9 next to W = P B on warps 10-14
                gemm_tn_dmma<2, 5>(T, nX, P, nX, As, nX, nX, nX, nX, 0, 0, 10);
                gemm_tn_dmma<2, 3>(W, nU, P, nX, Bs, nU, nX, nU, nX, 0, 10, nwarps - 10);
            } else {
                gemm_tn<VEC>(T, nX, P, nX, As, nX, nX, nX, nX, 0);     // T = P A      (P symmetric: P^T = P)
                gemm_tn<VEC>(W, nU, P, nX, Bs, nU, nX, nU, nX, 0);     // W = P B
            }
            for (int e = threadIdx.x; e < nU * nU; e += blockDim.x) G[(e / nU) * ldg + e % nU] = Rs[e];
            __syncthreads();
            async_copy(P, Qr + (p.q_per_step ? (long)k * nX * nX : 0), nX * nX);   // P is dead: start P <- Q(k)
            // gamma = R + B^T P B and Kp = B^T P A (tensor cores: 3 + 6 warp tiles)
            if (MMA) {
                gemm_tn_dmma<1, 3>(G, ldg, Bs, nU, W, nU, nU, nU, nX, 1, 0, 3);
                gemm_tn_dmma<1, 5>(Kp, nX, Bs, nU, T, nX, nU, nX, nX, 0, 3, nwarps - 3);
            } else {
                gemm_tn<VEC>(G, ldg, Bs, nU, W, nU, nU, nU, nX, 1);
                gemm_tn<VEC>(Kp, nX, Bs, nU, T, nX, nU, nX, nX, 0);
            }
            async_wait();                                          // Q(k) is in P
            __syncthreads();
            // P += A^T (P A) next to the set-up of the solve: Kp += S(k)^T (cross term), W <- Kp, c = r(k) + B^T b
            if (MMA) gemm_tn_dmma<2, 5>(P, nX, As, nX, T, nX, nX, nX, nX, 1, 0, 10);
            else gemm_tn<VEC>(P, nX, As, nX, T, nX, nX, nX, nX, 1);
            for (int e = threadIdx.x; e < nU * nX; e += blockDim.x) {
                const int i = e / nX, c = e - i * nX;
                double v = Kp[e];
                if (Sr) { v += Sr[(long)k * nX * nU + c * nU + i]; Kp[e] = v; }
                G[i * ldg + nU + c] = v;
            }
            for (int i = threadIdx.x; i < nU; i += blockDim.x) {
                double c = 0.0;
                if (affine) { gv[i] += rr[(long)k * nU + i]; c = gv[i]; }   // gv = r + B^T b
                G[i * ldg + nU + nX] = c;
            }
            __syncthreads();
            // K[k] = gamma^-1 Kp and C[k] = gamma^-1 (r + B^T b)
            if (!cta_gauss_jordan(G, ldg, nU, gje, scl, rd, order, 1e-300)) {
                s_fail = 1;   // every thread takes this branch (the pivots are computed redundantly by every warp)
                break;
            }
            for (int e = threadIdx.x; e < nU * nX; e += blockDim.x) {
                const int i = e / nX;
                Ks[e] = G[order[i] * ldg + nU + (e - i * nX)] * rd[i];
            }
            if (affine) {
                double* Ck = p.C + (r * (long)K + k) * nU;
                for (int i = threadIdx.x; i < nU; i += blockDim.x) Ck[i] = G[order[i] * ldg + nU + nX] * rd[i];
            }
            __syncthreads();
            double* Kk = Kr + (long)k * nU * nX;
            for (int e = threadIdx.x; e < nU * nX; e += blockDim.x) Kk[e] = Ks[e];
            if (affine) {
                // b = q[k] - K^T r + (A^T - K^T B^T) b = q[k] + A^T b - K^T (r + B^T b)
                for (int c = threadIdx.x; c < nX; c += blockDim.x) {
                    double acc = qr[(long)k * nX + c] + ab[c];
                    for (int i = 0; i < nU; ++i) acc -= Ks[i * nX + c] * gv[i];
                    bv[c] = acc;
                }
            }
            if (k > 0) {                                           // A, B are dead: fetch the next step's
                async_copy(As, Ar + (long)(k - 1) * nX * nX, nX * nX);
                async_copy(Bs, Br + (long)(k - 1) * nX * nU, nX * nU);
                if (p.r_per_step) async_copy(Rs, Rr + (long)(k - 1) * nU * nU, nU * nU);
            }
            if (MMA) gemm_tn_dmma<2, 5>(P, nX, Kp, nX, Ks, nX, nX, nX, nU, 2, 0, nwarps);   // P -= Kp^T K[k]
            else gemm_tn<VEC>(P, nX, Kp, nX, Ks, nX, nX, nX, nU, 2);
            __syncthreads();
            // P = (P + P^T) / 2, one off-diagonal per warp pass: (i, i + d) and its mirror are both read with
            // stride nX + 1 (a row-by-row sweep reads the mirror with stride nX: a 32-way bank conflict at nX = 80)
            for (int d = 1 + (int)(threadIdx.x >> 5); d < nX; d += nwarps) {
                for (int i = threadIdx.x & 31; i + d < nX; i += 32) {
                    const int a = i * nX + i + d, b = (i + d) * nX + i;
                    const double v = (P[a] + P[b]) / 2.0;
                    P[a] = v;
                    P[b] = v;
                }
            }
        }
        async_wait();
        __syncthreads();
        if (p.P0) {
            double* Pr = p.P0 + r * (long)nX * nX;
            for (int e = threadIdx.x; e < nX * nX; e += blockDim.x) Pr[e] = P[e];
        }
        if (affine && p.b0) for (int c = threadIdx.x; c < nX; c += blockDim.x) p.b0[r * (long)nX + c] = bv[c];
        if (threadIdx.x == 0) p.status[r] = s_fail ? ST_SINGULAR : ST_OK;
        __syncthreads();
    }
}

int lqr_fail(int code, const std::string& m) { last_error() = m; return code; }

}  // namespace
}  // namespace trepb

using namespace trepb;

namespace {
// CUDA events around the last Riccati launch of each device (trepb_lqr_last_kernel_ms)
struct LqrTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool ok = false;
};
LqrTimer& lqr_timer(int device) {
    static LqrTimer t[64];
    LqrTimer& x = t[device & 63];
    if (!x.ok) x.ok = cudaEventCreate(&x.e0) == cudaSuccess && cudaEventCreate(&x.e1) == cudaSuccess;
    return x;
}

int lqr_launch(int device, LqrParams& p, cudaStream_t stream) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    const int nX = p.nX, nU = p.nU, ldg = (nU + nX + 1) | 1;
    const size_t doubles = 3 * (size_t)nX * nX + 4 * (size_t)nX * nU + (size_t)nU * nU + (size_t)nU * ldg + 4 * (size_t)nU + 16
                           + 2 * (size_t)nX + 2 * (size_t)nU + 2;
    const size_t smem = doubles * sizeof(double);
    int smem_optin = 0, sms = 0;
    e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    if (smem > (size_t)smem_optin)
        return lqr_fail(TREPB_ERR_UNSUPPORTED, "state dimension too large for the shared-memory Riccati sweep");
    const bool vec = nX % 2 == 0 && nU % 2 == 0;   // every operand block then starts 16-byte aligned with an even row length
    // tensor-core schedule for states large enough to fill its warp tiles (TREPB_LQR_NO_MMA=1: scalar tiles, for comparison)
    const char* no_mma = getenv("TREPB_LQR_NO_MMA");
    const bool mma = nX >= 32 && !(no_mma && no_mma[0] == '1');
    const void* fn = mma ? (vec ? (const void*)lqr_kernel<true, true> : (const void*)lqr_kernel<false, true>) : vec ? (const void*)lqr_kernel<true, false> : (const void*)lqr_kernel<false, false>;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    const int tiles = ((nX + 3) / 4) * ((nX + 3) / 4);
    int block = ((tiles + 31) / 32) * 32;
    if (block > 512) block = 512;
    if (block < 64) block = 64;
    if (mma) block = 512;   // the tensor-core schedule assigns products to warps 0-9 and 10-14
    // the elimination deals the nU x (nU + nX + 1) elements of [gamma | Kp | c] out to the threads, at most 4 each
    if (nU > 64) return lqr_fail(TREPB_ERR_UNSUPPORTED, "more than 64 inputs");
    while (block < 512 && (long long)nU * (nU + nX + 1) > 4LL * block) block += 32;
    if ((long long)nU * (nU + nX + 1) > 4LL * block)
        return lqr_fail(TREPB_ERR_UNSUPPORTED, "input dimension too large for the in-kernel gamma solve");
    // small systems: as many CTAs per SM as registers / shared memory allow (one rollout per CTA)
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, block, smem) != cudaSuccess || occ < 1) occ = 1;
    const long long slots = (long long)sms * occ;
    const long long grid = p.batch < slots ? p.batch : slots;
    LqrTimer& tm = lqr_timer(device);
    if (tm.ok) cudaEventRecord(tm.e0, stream);
    if (mma && vec) lqr_kernel<true, true><<<(int)grid, block, smem, stream>>>(p);
    else if (mma) lqr_kernel<false, true><<<(int)grid, block, smem, stream>>>(p);
    else if (vec) lqr_kernel<true, false><<<(int)grid, block, smem, stream>>>(p);
    else lqr_kernel<false, false><<<(int)grid, block, smem, stream>>>(p);
    if (tm.ok) cudaEventRecord(tm.e1, stream);
    e = cudaGetLastError();
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    return TREPB_OK;
}
}  // namespace

extern "C" int trepb_lqr_last_kernel_ms(int device, float* ms) {
    if (!ms) return lqr_fail(TREPB_ERR_INVALID, "null argument");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    LqrTimer& tm = lqr_timer(device);
    if (!tm.ok) return lqr_fail(TREPB_ERR_CUDA, "no events");
    e = cudaEventSynchronize(tm.e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms, tm.e0, tm.e1);
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    return TREPB_OK;
}

extern "C" int trepb_lqr_batch_dev(int device, const trepb_lqr_args* a, void* stream) {
    TREPB_NVTX("trepb_lqr_batch_dev");
    if (!a) return lqr_fail(TREPB_ERR_INVALID, "null argument");
    if (a->batch < 0 || a->nsteps < 1 || a->nX < 1 || a->nU < 1) return lqr_fail(TREPB_ERR_INVALID, "bad sizes");
    if (!a->A || !a->B || !a->Q || !a->R || !a->Kfb || !a->status) return lqr_fail(TREPB_ERR_INVALID, "A, B, Q, R, Kfb and status are required");
    if (a->batch == 0) return TREPB_OK;
    LqrParams p{};
    p.batch = a->batch; p.K = a->nsteps; p.nX = a->nX; p.nU = a->nU;
    p.A = a->A; p.B = a->B; p.Q = a->Q; p.R = a->R; p.q_per_step = a->q_per_step; p.r_per_step = a->r_per_step;
    p.Kfb = a->Kfb; p.P0 = a->P0; p.status = a->status;
    return lqr_launch(device, p, (cudaStream_t)stream);
}

extern "C" int trepb_lq_batch_dev(int device, const trepb_lq_args* a, void* stream) {
    TREPB_NVTX("trepb_lq_batch_dev");
    if (!a) return lqr_fail(TREPB_ERR_INVALID, "null argument");
    if (a->batch < 0 || a->nsteps < 1 || a->nX < 1 || a->nU < 1) return lqr_fail(TREPB_ERR_INVALID, "bad sizes");
    if (!a->A || !a->B || !a->Q || !a->R || !a->q || !a->r || !a->Kfb || !a->C || !a->status)
        return lqr_fail(TREPB_ERR_INVALID, "A, B, Q, R, q, r, Kfb, C and status are required");
    if (a->batch == 0) return TREPB_OK;
    const long long K = a->nsteps, nX = a->nX, nU = a->nU, per = a->cost_per_rollout ? 1 : 0;
    LqrParams p{};
    p.batch = a->batch; p.K = a->nsteps; p.nX = a->nX; p.nU = a->nU;
    p.A = a->A; p.B = a->B; p.Q = a->Q; p.R = a->R; p.q_per_step = 1; p.r_per_step = 1;
    p.S = a->S; p.qv = a->q; p.rv = a->r;
    p.Qs = per * (K + 1) * nX * nX; p.Rs = per * K * nU * nU; p.Ss = per * K * nX * nU;
    p.qs = per * (K + 1) * nX; p.rs = per * K * nU;
    p.Kfb = a->Kfb; p.C = a->C; p.P0 = a->P0; p.b0 = a->b0; p.status = a->status;
    return lqr_launch(device, p, (cudaStream_t)stream);
}

extern "C" int trepb_lq_batch(int device, const trepb_lq_args* a) {
    TREPB_NVTX("trepb_lq_batch");
    if (!a) return lqr_fail(TREPB_ERR_INVALID, "null argument");
    if (a->batch < 0 || a->nsteps < 1 || a->nX < 1 || a->nU < 1) return lqr_fail(TREPB_ERR_INVALID, "bad sizes");
    if (!a->A || !a->B || !a->Q || !a->R || !a->q || !a->r || !a->Kfb || !a->C || !a->status)
        return lqr_fail(TREPB_ERR_INVALID, "A, B, Q, R, q, r, Kfb, C and status are required");
    if (a->batch == 0) return TREPB_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    const size_t Rn = (size_t)a->batch, K = (size_t)a->nsteps, nX = (size_t)a->nX, nU = (size_t)a->nU;
    const size_t cm = a->cost_per_rollout ? Rn : 1;
    // host -> device staging: {source, bytes} in, {destination, bytes} out
    const void* in_src[7] = {a->A, a->B, a->Q, a->R, a->S, a->q, a->r};
    const size_t in_n[7] = {Rn * K * nX * nX, Rn * K * nX * nU, cm * (K + 1) * nX * nX, cm * K * nU * nU,
                            a->S ? cm * K * nX * nU : 0, cm * (K + 1) * nX, cm * K * nU};
    void* out_dst[4] = {a->Kfb, a->C, a->P0, a->b0};
    const size_t out_n[4] = {Rn * K * nU * nX, Rn * K * nU, a->P0 ? Rn * nX * nX : 0, a->b0 ? Rn * nX : 0};
    double* din[7] = {nullptr}; double* dout[4] = {nullptr};
    int* dS = nullptr;
    int rc = TREPB_OK;
#define LQ(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess && rc == TREPB_OK) rc = lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e_)); } while (0)
    for (int i = 0; i < 7; ++i) if (in_n[i]) { LQ(cudaMalloc(&din[i], in_n[i] * 8)); if (rc == TREPB_OK) LQ(cudaMemcpy(din[i], in_src[i], in_n[i] * 8, cudaMemcpyHostToDevice)); }
    for (int i = 0; i < 4; ++i) if (out_n[i]) LQ(cudaMalloc(&dout[i], out_n[i] * 8));
    LQ(cudaMalloc(&dS, (Rn ? Rn : 1) * 4));
    if (rc == TREPB_OK && Rn) {
        trepb_lq_args d = *a;
        d.A = din[0]; d.B = din[1]; d.Q = din[2]; d.R = din[3]; d.S = din[4]; d.q = din[5]; d.r = din[6];
        d.Kfb = dout[0]; d.C = dout[1]; d.P0 = dout[2]; d.b0 = dout[3]; d.status = dS;
        rc = trepb_lq_batch_dev(device, &d, nullptr);
    }
    if (rc == TREPB_OK) {
        for (int i = 0; i < 4; ++i) if (out_n[i]) LQ(cudaMemcpy(out_dst[i], dout[i], out_n[i] * 8, cudaMemcpyDeviceToHost));
        if (Rn) LQ(cudaMemcpy(a->status, dS, Rn * 4, cudaMemcpyDeviceToHost));
    }
#undef LQ
    for (int i = 0; i < 7; ++i) cudaFree(din[i]);
    for (int i = 0; i < 4; ++i) cudaFree(dout[i]);
    cudaFree(dS);
    return rc;
}

extern "C" int trepb_lqr_batch(int device, const trepb_lqr_args* a) {
    TREPB_NVTX("trepb_lqr_batch");
    if (!a) return lqr_fail(TREPB_ERR_INVALID, "null argument");
    if (a->batch < 0 || a->nsteps < 1 || a->nX < 1 || a->nU < 1) return lqr_fail(TREPB_ERR_INVALID, "bad sizes");
    if (a->batch == 0) return TREPB_OK;
    if (!a->A || !a->B || !a->Q || !a->R || !a->Kfb || !a->status) return lqr_fail(TREPB_ERR_INVALID, "A, B, Q, R, Kfb and status are required");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e));
    const size_t R = (size_t)a->batch, K = (size_t)a->nsteps, nX = (size_t)a->nX, nU = (size_t)a->nU;
    const size_t nA = R * K * nX * nX, nB = R * K * nX * nU, nQ = (a->q_per_step ? K + 1 : 1) * nX * nX,
                 nR = (a->r_per_step ? K : 1) * nU * nU, nK = R * K * nU * nX, nP = R * nX * nX;
    double *dA = nullptr, *dB = nullptr, *dQ = nullptr, *dR = nullptr, *dK = nullptr, *dP = nullptr;
    int* dS = nullptr;
    int rc = TREPB_OK;
#define LQ(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess && rc == TREPB_OK) rc = lqr_fail(TREPB_ERR_CUDA, cudaGetErrorString(e_)); } while (0)
    LQ(cudaMalloc(&dA, nA * 8)); LQ(cudaMalloc(&dB, nB * 8)); LQ(cudaMalloc(&dQ, nQ * 8)); LQ(cudaMalloc(&dR, nR * 8));
    LQ(cudaMalloc(&dK, nK * 8)); LQ(cudaMalloc(&dP, nP * 8)); LQ(cudaMalloc(&dS, (R ? R : 1) * 4));
    if (rc == TREPB_OK) {
        LQ(cudaMemcpy(dA, a->A, nA * 8, cudaMemcpyHostToDevice)); LQ(cudaMemcpy(dB, a->B, nB * 8, cudaMemcpyHostToDevice));
        LQ(cudaMemcpy(dQ, a->Q, nQ * 8, cudaMemcpyHostToDevice)); LQ(cudaMemcpy(dR, a->R, nR * 8, cudaMemcpyHostToDevice));
    }
    if (rc == TREPB_OK) {
        trepb_lqr_args d = *a;
        d.A = dA; d.B = dB; d.Q = dQ; d.R = dR; d.Kfb = dK; d.P0 = a->P0 ? dP : nullptr; d.status = dS;
        rc = trepb_lqr_batch_dev(device, &d, nullptr);
    }
    if (rc == TREPB_OK) {
        LQ(cudaMemcpy(a->Kfb, dK, nK * 8, cudaMemcpyDeviceToHost));
        if (a->P0) LQ(cudaMemcpy(a->P0, dP, nP * 8, cudaMemcpyDeviceToHost));
        LQ(cudaMemcpy(a->status, dS, R * 4, cudaMemcpyDeviceToHost));
    }
#undef LQ
    cudaFree(dA); cudaFree(dB); cudaFree(dQ); cudaFree(dR); cudaFree(dK); cudaFree(dP); cudaFree(dS);
    return rc;
}
