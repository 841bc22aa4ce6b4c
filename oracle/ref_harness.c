/* TEST INFRASTRUCTURE (CPU baseline timing): a tight C loop around the REFERENCE's own
 * MidpointVI_solve_DEL / calc_deriv1 (trep/_trep/midpointvi.c:691-747, 1100-1120, 2682-2690),
 * called on a live reference MidpointVI object, with no Python per step.
 *
 * Compiled against the headers the reference installs for external C code
 * (trep/_trep/trep.h, copied to oracle/_ref/trep/_trep/ by oracle/build_ref.py); the function
 * pointers are taken from the loaded reference extension with dlsym by oracle/cpu_baseline.py.
 * The state hand-over between steps restates MidpointVI.step (trep/midpointvi.py:174-201) and
 * the A/B assembly restates DSystem.fdx/fdu (trep/discopt/dsystem.py:284-317).
 * Must be called with the GIL held (ctypes.PyDLL): the reference sets Python exceptions.
 */
#define NPY_NO_DEPRECATED_API 0
#include <Python.h>
#include <numpy/arrayobject.h>
#include <string.h>
#include "trep.h"

typedef int (*solve_fn)(MidpointVI*, int);
/* MidpointVI_calc_deriv1 is static in the reference; it is reached through the C method
 * _MidpointVI._calc_deriv1 (midpointvi.c:2682-2690), one C-level method call per instance. */
static int call_deriv1(MidpointVI* mvi, PyObject* name) {
    PyObject* r = PyObject_CallMethodNoArgs((PyObject*)mvi, name);
    if (!r) return -1;
    Py_DECREF(r);
    return 0;
}

static double* D(PyArrayObject* a) { return (double*)PyArray_DATA(a); }

/* B independent rollouts of nsteps steps each, started like initialize_from_state(t0,q,p). */
long rh_rollouts(MidpointVI* mvi, solve_fn solve, long B, int nsteps, int nq, int nd, int nc,
                 double t0, double dt, int max_it, const double* q, const double* p,
                 double* q2, double* p2, int* iters, int* status)
{
    long total = 0;
    for (long b = 0; b < B; ++b) {
        memcpy(D(mvi->q2), q + b * nq, sizeof(double) * nq);
        memcpy(D(mvi->p2), p + b * nd, sizeof(double) * nd);
        if (nc) memset(D(mvi->lambda1), 0, sizeof(double) * nc);
        mvi->t2 = t0;
        int it_sum = 0, st = 0;
        for (int s = 0; s < nsteps; ++s) {
            memcpy(D(mvi->q1), D(mvi->q2), sizeof(double) * nq);
            memcpy(D(mvi->p1), D(mvi->p2), sizeof(double) * nd);
            mvi->t1 = mvi->t2;
            mvi->t2 = mvi->t1 + dt;
            mvi->cache = 0;
            int it = solve(mvi, max_it);
            if (it < 0) { PyErr_Clear(); st = -1; break; }
            it_sum += it;
        }
        memcpy(q2 + b * nq, D(mvi->q2), sizeof(double) * nq);
        memcpy(p2 + b * nd, D(mvi->p2), sizeof(double) * nd);
        iters[b] = it_sum;
        status[b] = st;
        total += nsteps;
    }
    return total;
}

/* B independent linearizations: DSystem.set(X,U,k,xk_hint) + fdx() + fdu(). */
long rh_linearize(MidpointVI* mvi, solve_fn solve, PyObject* deriv1_name, long B, int nq, int nd, int nk,
                  int nu, int nc, double t1, double dt, int max_it,
                  const double* q1, const double* p1, const double* u1, const double* k2,
                  const double* q2_hint, const double* lam_hint,
                  double* A, double* Bm, int* iters, int* status)
{
    const int nX = 2 * nq, nU = nu + nk;
    for (long b = 0; b < B; ++b) {
        memcpy(D(mvi->q1), q1 + b * nq, sizeof(double) * nq);
        memcpy(D(mvi->p1), p1 + b * nd, sizeof(double) * nd);
        memcpy(D(mvi->q2), q1 + b * nq, sizeof(double) * nq);
        if (q2_hint) memcpy(D(mvi->q2), q2_hint + b * nd, sizeof(double) * nd);
        if (nk) memcpy(D(mvi->q2) + nd, k2 + b * nk, sizeof(double) * nk);
        if (nu) memcpy(D(mvi->u1), u1 + b * nu, sizeof(double) * nu);
        if (nc) {
            if (lam_hint) memcpy(D(mvi->lambda1), lam_hint + b * nc, sizeof(double) * nc);
            else memset(D(mvi->lambda1), 0, sizeof(double) * nc);
        }
        mvi->t1 = t1;
        mvi->t2 = t1 + dt;
        mvi->cache = 0;
        int it = solve(mvi, max_it);
        if (it < 0) { PyErr_Clear(); status[b] = -1; iters[b] = 0; continue; }
        iters[b] = it;
        if (call_deriv1(mvi, deriv1_name)) { PyErr_Clear(); status[b] = -2; continue; }
        status[b] = 0;
        if (A) {
            double* a = A + b * nX * nX;
            memset(a, 0, sizeof(double) * nX * nX);
            const double *q2_dq1 = D(mvi->q2_dq1), *q2_dp1 = D(mvi->q2_dp1);
            const double *p2_dq1 = D(mvi->p2_dq1), *p2_dp1 = D(mvi->p2_dp1);
            for (int i = 0; i < nq; ++i)
                for (int j = 0; j < nd; ++j) {
                    a[j * nX + i] = q2_dq1[i * nd + j];
                    a[(nq + j) * nX + i] = p2_dq1[i * nd + j];
                }
            for (int i = 0; i < nd; ++i)
                for (int j = 0; j < nd; ++j) {
                    a[j * nX + nq + i] = q2_dp1[i * nd + j];
                    a[(nq + j) * nX + nq + i] = p2_dp1[i * nd + j];
                }
            for (int i = 0; i < nk; ++i) a[(nq + nd + i) * nX + nd + i] = -1.0 / dt;
        }
        if (Bm && nU) {
            double* m = Bm + b * nX * nU;
            memset(m, 0, sizeof(double) * nX * nU);
            const double *q2_du1 = D(mvi->q2_du1), *q2_dk2 = D(mvi->q2_dk2);
            const double *p2_du1 = D(mvi->p2_du1), *p2_dk2 = D(mvi->p2_dk2);
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j < nd; ++j) {
                    m[j * nU + i] = q2_du1[i * nd + j];
                    m[(nq + j) * nU + i] = p2_du1[i * nd + j];
                }
            for (int i = 0; i < nk; ++i) {
                for (int j = 0; j < nd; ++j) {
                    m[j * nU + nu + i] = q2_dk2[i * nd + j];
                    m[(nq + j) * nU + nu + i] = p2_dk2[i * nd + j];
                }
                m[(nd + i) * nU + nu + i] = 1.0;
                m[(nq + nd + i) * nU + nu + i] = 1.0 / dt;
            }
        }
    }
    return B;
}
