"""Reference-side binding: the reference's OWN classes running on libtrepb.so.

`batched_dsystem_class(trep.discopt.DSystem)` and `batched_doptimizer_class(trep.discopt.DOptimizer)` return
subclasses of the classes they are given (this module never imports trep; it is handed the classes, duck-typed
on the attribute names of trep/discopt/dsystem.py:29-60 and doptimizer.py:207-245).  Construction, state
packing, cost functions, monitors, `optimize()` and `step()` stay the reference's code; what is replaced are
the loops over the time steps of one trajectory - each becomes one launch over all k:

    DSystem.linearize_trajectory        dsystem.py:406-423   -> trepb_linearize_batch (every k one instance)
    DSystem.calc_feedback_controller    dsystem.py:474-494   -> linearize + trepb_lqr_batch_dev, slabs stay in HBM
    DSystem.project                     dsystem.py:426-457   -> trepb_project_batch on the system's own time grid
    DOptimizer.calc_newton_model        doptimizer.py:319-345 -> adjoint recursion z[k] on the host (K matrix-vector
                                                                products), then ONE z-contracted second-derivative
                                                                launch over all k (trepb_deriv2_batch with z)
    DOptimizer.calc_descent_direction   doptimizer.py:348-402 -> the LQ solve on the device (trepb_lq_batch)
    DOptimizer.armijo_search            doptimizer.py:431-463 -> candidate step sizes beta^m evaluated as batches of
                                                                closed-loop rollouts (trepb_project_batch,
                                                                use_hint = 0 like armijo_simulate)

tests/test_gpu_dropin.py runs one DOptimizer.step of examples/pend-on-cart-optimization.py and of a marionette
problem with the stock reference classes and with these subclasses and compares cost, descent direction and
the new trajectory.
"""
from __future__ import annotations

import numpy as np

from . import lib, model


def _handle(dsys):
    """GPU handle of the DSystem's mechanical system; rebuilt when the reference reports a structure change."""
    mvi = dsys.varint
    h = getattr(mvi, "_trepb_handle", None)
    if h is None:
        h = lib.System(model.flatten_trep_system(mvi.system))
        mvi._trepb_handle = h
        add = getattr(mvi.system, "add_structure_changed_func", None)
        if add is not None:      # trep/system.py: the same hook MidpointVI uses (midpointvi.py:25)
            add(lambda: setattr(mvi, "_trepb_handle", None))
    return h


def _stack(f, n):
    """[f(0) .. f(n-1)] as one array; a constant function is evaluated once."""
    return np.stack([np.asarray(f(k), float) for k in range(n)])


def batched_dsystem_class(DSystem, ConvergenceError=RuntimeError):
    class BatchedDSystem(DSystem):
        def _batch_inputs(self, X, U):
            K = len(X) - 1
            X, U = np.asarray(X, float), np.asarray(U, float)
            q1, p1 = X[:K, self._slice_Q], X[:K, self._slice_p]
            u1, rho = U[:K, self._slice_u], U[:K, self._slice_rho]
            return K, q1, p1, u1, rho, X[1:K + 1, self._slice_Qd]

        def linearize_trajectory(self, X, U):
            h = _handle(self)
            K, q1, p1, u1, rho, hint = self._batch_inputs(X, U)
            out = h.linearize(q1, p1, u1, rho, t1=self._time[:K], t2=self._time[1:K + 1], q2_guess=hint,
                              tolerance=self.varint.tolerance)
            bad = np.flatnonzero(out["status"] != 0)
            if bad.size:
                raise ConvergenceError("linearization failed at k = %d" % bad[0])
            return self.linearization_return(out["A"], out["B"])

        def calc_feedback_controller(self, X, U, Q=None, R=None, return_linearization=False):
            h = _handle(self)
            dev = h.device
            K, q1, p1, u1, rho, hint = self._batch_inputs(X, U)
            nX, nU = self._nX, self._nU
            Qs = np.eye(nX) if Q is None else _stack(Q, K + 1)
            Rs = np.eye(nU) if R is None else _stack(R, K)
            up = lambda a, dt=np.float64: lib.DeviceBuffer(dev, np.shape(a), dt).upload(np.ascontiguousarray(a, dtype=dt))
            bufs = dict(q1=up(q1), p1=up(p1), u1=up(u1) if self._nu else None, k2=up(rho) if self._nrho else None,
                        hint=up(hint), t1=up(self._time[:K]), t2=up(self._time[1:K + 1]), Q=up(Qs), R=up(Rs),
                        A=lib.DeviceBuffer(dev, (K, nX, nX)), B=lib.DeviceBuffer(dev, (K, nX, nU)),
                        st=lib.DeviceBuffer(dev, (K,), np.int32), Kfb=lib.DeviceBuffer(dev, (1, K, nU, nX)),
                        ls=lib.DeviceBuffer(dev, (1,), np.int32))
            try:
                b = bufs
                h.linearize_raw(True, K, b["q1"], b["p1"], b["u1"], b["k2"], b["st"], t1=b["t1"], t2=b["t2"],
                                q2_guess=b["hint"], A=b["A"], B=b["B"], tolerance=self.varint.tolerance)
                lib.lqr_raw(True, dev, 1, K, nX, nU, b["A"], b["B"], b["Q"], b["R"], b["Kfb"], b["ls"],
                            q_per_step=Qs.ndim == 3, r_per_step=Rs.ndim == 3)
                lib.synchronize(dev)
                st = b["st"].download()
                if np.any(st != 0):
                    raise ConvergenceError("linearization failed at k = %d" % np.flatnonzero(st != 0)[0])
                if b["ls"].download()[0] != 0:
                    raise ValueError("singular matrix in the Riccati sweep")
                Kproj = b["Kfb"].download()[0]
                A = b["A"].download() if return_linearization else None
                B = b["B"].download() if return_linearization else None
            finally:
                for v in bufs.values():
                    if v is not None:
                        v.free()
            if return_linearization:
                return self.feedback_return(Kproj, A, B)
            return Kproj

        def project(self, bX, bU, Kproj=None):
            if Kproj is None:
                Kproj = self.calc_feedback_controller(bX, bU)
            out = self.project_batch(np.asarray(bX, float)[None], np.asarray(bU, float)[None], Kproj, use_hint=True)
            if out["status"][0] != 0:
                raise ConvergenceError("projection failed at k = %d" % out["fail_step"][0])
            return self.trajectory_return(out["X"][0], out["U"][0])

        def project_batch(self, bX, bU, Kproj, use_hint=True):
            """[R] candidates (bX [R,K+1,nX], bU [R,K,nU]) in one launch; returns the library's dict
            (X, U, status, fail_step, iters)."""
            h = _handle(self)
            K = bX.shape[1] - 1
            return h.project(bX, bU[:, :K], np.asarray(Kproj, float), float(self._time[0]),
                             float(self._time[1] - self._time[0]), use_hint=use_hint,
                             tolerance=self.varint.tolerance, times=self._time[:K + 1])

        def second_order_terms(self, X, U, Z):
            """fdxdx(Z[k]), fdxdu(Z[k]), fdudu(Z[k]) after set(X[k], U[k], k, xk_hint=X[k+1]) for every k
            (dsystem.py:320-386): one launch.  Z [K, nX]."""
            h = _handle(self)
            K, q1, p1, u1, rho, hint = self._batch_inputs(X, U)
            out = h.deriv2(q1, p1, u1, rho, t1=self._time[:K], t2=self._time[1:K + 1], q2_guess=hint,
                           tolerance=self.varint.tolerance, z=np.asarray(Z, float), tensors=False)
            bad = np.flatnonzero(out["status"] != 0)
            if bad.size:
                raise ConvergenceError("second derivatives failed at k = %d" % bad[0])
            return out["fdxdx"], out["fdxdu"], out["fdudu"]

    BatchedDSystem.__name__ = "Batched" + DSystem.__name__
    return BatchedDSystem


def batched_doptimizer_class(DOptimizer, ConvergenceError=RuntimeError, armijo_batch=8):
    class BatchedDOptimizer(DOptimizer):
        def _cost_gradients(self, X, U):
            K = len(X) - 1
            q = np.zeros(X.shape)
            r = np.zeros(U.shape)
            for k in range(K):
                q[k] = self.cost.l_dx(X[k], U[k], k)
                r[k] = self.cost.l_du(X[k], U[k], k)
            q[-1] = self.cost.m_dx(X[-1])
            return q, r

        def calc_newton_model(self, X, U, A, B, K):
            n = len(X) - 1
            q, r = self._cost_gradients(X, U)
            # adjoint: z_k = l_dx - l_du K_k + z_{k+1} (A_k - B_k K_k), z_n = m_dx; step k contracts with z_{k+1}
            Z = np.zeros((n + 1, self.dsys.nX))
            Z[n] = q[n]
            for k in reversed(range(n)):
                Z[k] = q[k] - r[k] @ K[k] + Z[k + 1] @ (A[k] - B[k] @ K[k])
            xx, xu, uu = self.dsys.second_order_terms(X, U, Z[1:])
            Q = [None] * (n + 1)
            S, R = [None] * n, [None] * n
            Q[n] = self.cost.m_dxdx(X[-1])
            for k in range(n):
                Q[k] = self.cost.l_dxdx(X[k], U[k], k) + xx[k]
                S[k] = self.cost.l_dxdu(X[k], U[k], k) + xu[k]
                R[k] = self.cost.l_dudu(X[k], U[k], k) + uu[k]
            return self.model_return(lambda k: Q[k], lambda k: R[k], lambda k: S[k])

        def calc_descent_direction(self, X, U, method="steepest"):
            (Kproj, A, B) = self.dsys.calc_feedback_controller(X, U, self.Qproj, self.Rproj, True)
            q, r = self._cost_gradients(X, U)
            if method == "steepest":
                (Q, R, S) = self.calc_steepest_model()
            elif method == "quasi":
                (Q, R, S) = self.calc_quasi_model(X, U)
            elif method == "newton":
                (Q, R, S) = self.calc_newton_model(X, U, A, B, Kproj)
            else:
                raise ValueError("Invalid descent direction method: %r" % method)
            n = len(X) - 1
            dev = getattr(getattr(self.dsys.varint, "_trepb_handle", None), "device", 0)
            (K, C, P, b) = lib.solve_tv_lq(A, B, q, r, _stack(Q, n + 1), _stack(S, n), _stack(R, n), device=dev)
            dx0 = -np.linalg.solve(P, b) if self.optimize_ic else np.zeros((self.dsys.nX,))
            dX = np.zeros(X.shape)
            dU = np.zeros(U.shape)
            dX[0] = dx0
            for k in range(n):
                dU[k] = -(K[k] @ dX[k]) - C[k]
                dX[k + 1] = A[k] @ dX[k] + B[k] @ dU[k]
            return self.descent_return(Kproj, dX, dU, Q, R, S)

        def armijo_search(self, X, U, Kproj, dX, dU):
            cost0 = self.calc_cost(X, U)
            dcost0 = self.calc_dcost(X, U, dX, dU)
            m0 = 0
            while m0 < self.armijo_max_iterations:
                ms = list(range(m0, min(m0 + armijo_batch, self.armijo_max_iterations)))
                lam = np.array([self.armijo_beta ** m for m in ms])
                bX = X[None] + lam[:, None, None] * dX[None]
                bU = U[None] + lam[:, None, None] * dU[None]
                out = self.dsys.project_batch(bX, bU, Kproj, use_hint=False)
                for i, m in enumerate(ms):          # accept the FIRST (largest) step size that passes, as the loop does
                    max_cost = cost0 + self.armijo_alpha * lam[i] * dcost0
                    if out["status"][i] != 0:
                        k = int(out["fail_step"][i])
                        self.monitor.armijo_simulation_failure(m, out["X"][i][:k], out["U"][i][:k], out["X"][i][:k], bU[i])
                        continue
                    nX, nU = out["X"][i], out["U"][i]
                    cost1 = self.calc_cost(nX, nU)
                    self.monitor.armijo_evaluation(m, nX, nU, bX[i], bU[i], cost1, max_cost)
                    if cost1 < max_cost:
                        return self.armijo_search_return(nX, nU, cost1)
                m0 += len(ms)
            self.monitor.armijo_search_failure(X, U, dX, dU, cost0, dcost0, Kproj)
            raise ConvergenceError("Armijo Failed to Converge")

    BatchedDOptimizer.__name__ = "Batched" + DOptimizer.__name__
    return BatchedDOptimizer
