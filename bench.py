#!/usr/bin/env python3
"""bench.py - headline benchmark of the batched MidpointVI path (driver contract).

Workload (BASELINE.json configs[1]): 2^20 independent damped pendulums
(examples/damped-pendulum.py) stepped 1000 MidpointVI steps; one bench "step" = one pass of the
hot path over that batch = 1.048576e9 DEL steps per GPU.  Metric: DEL steps/s, whole job.

  value      device-resident inputs, CUDA-event time of the step kernel, max over ranks
  e2e        the same pass through the host-pointer C-ABI call (trepb_step_batch) with pinned
             host buffers: H2D of (q1,p1), kernel, D2H of (q2,p2,iters,status) inside the timed region
  roofline   the step kernel against the FP64 pipe (this kernel is on-chip/compute bound; the
             denominator is an in-run DFMA micro-benchmark since MEASURED_PEAKS.json has no
             fp64 entry) - see DESIGN.md "Roofline accounting"
  secondary  linearizations/s for pend-on-cart (HBM-bound, with its own hbm roofline) and the
             marionette, same run
  cpu_baseline  the reference's own C path (oracle/_ref) in a tight C loop on all host cores,
             bounded sample

`--impl reference` times only the reference arm (rank 0; other ranks exit).
Multi-GPU (torchrun): the batch shards across ranks with no data-path collective (weak scaling).  For N > 1 the
line also carries `secondary` entries measured on all N ranks (max over ranks, device-timed): W3 linearizations/s
sharded over the ranks plus the gather of the A / B slabs on rank 0 - by NCCL (trepb_comm_gather_dev) and fused
into the linearize kernel through rank 0's peer-mapped slab - with the achieved NVLink GB/s, W5 marionette
linearizations/s, and W4 at north_star's size (1.25e7 dual pendulums per GPU x 1000 steps).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 1 << 20
NSTEPS = 1000
DT = 0.01
METRIC = "batched DEL steps/s (MidpointVI, damped pendulum)"
UNIT = "DEL steps/s"
WORKLOAD = "W2: 2^20 damped pendulums (examples/damped-pendulum.py) x 1000 MidpointVI steps per GPU"
# executed fp64 flops per DEL step of step_kernel<damped_pendulum> (ncu sass op counters,
# profiles/r01_step_flops.txt; FMA = 2) - the algorithmic figure of DESIGN.md "Roofline accounting"
FLOPS_PER_STEP = None  # filled from profiles/flops.json if present


def workload_inputs(batch, seed=0):
    """theta0 ~ U(-pi,pi), theta1 = theta0 + U(-0.02,0.02)."""
    rng = np.random.default_rng(seed)
    th0 = rng.uniform(-np.pi, np.pi, (batch, 1))
    th1 = th0 + rng.uniform(-0.02, 0.02, (batch, 1))
    return th0, th1


def opcounted(key):
    """Operation-counted flops of one unit (profiles/opcounts.json, written by tools/count_ops.py: the table-driven
    math compiled on the host with a counting `double`), reported beside the executed count the roofline uses."""
    path = os.path.join(ROOT, "profiles", "opcounts.json")
    if not os.path.exists(path):
        return None
    e = json.load(open(path)).get(key)
    if not e:
        return None
    return {"flops_per_unit": e["flops"], "sincos_per_unit": e["sincos"], "newton_iters": e["newton_iters"], "sample": e["sample"],
            "note": "add + sub + mul + div + sqrt + 2 fma of the table-driven formulation (trepb_math.cuh with a counting double, "
                    "tests/opcount.cc), sin / cos evaluations separate; a specialised kernel folds its structural zeros at compile "
                    "time and executes fewer, so `achieved` keeps the kernel's own executed count"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(ngpus):
    """torch.distributed only when launched under torchrun with WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, rank, local, world
    return None, 0, 0, 1


def explain_failures(name, kind, **inp):
    """CHECKER (oracle/_ref, not timed, not part of any measured path): run the reference's own C on the inputs
    of the instances the GPU reported as not converged / singular and count on how many the reference fails too."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline as cb
    h = cb.Harness(name)
    if kind == "rollouts":
        r = h.rollouts(inp["q"], inp["p"], inp["nsteps"], inp["t0"], inp["dt"])
    else:
        r = h.linearize(inp["q1"], inp["p1"], inp.get("u1"), inp.get("k2"), inp.get("t1", 0.0), inp["dt"],
                        q2_hint=inp.get("hint"), lam_hint=inp.get("lam"))
    return int(np.sum(r["status"] != 0))


def failure_report(name, kind, status, limit=64, **inp):
    bad = np.flatnonzero(status != 0)
    rep = {"not_ok": int(bad.size)}
    if bad.size:
        pick = bad[:limit]
        sub = {k: (v[pick] if isinstance(v, np.ndarray) and v.shape[:1] == status.shape else v) for k, v in inp.items()}
        try:
            rep["checked"] = int(pick.size)
            rep["reference_also_fails"] = explain_failures(name, kind, **sub)
        except Exception as e:      # the checker must never break the measurement
            rep["checker_error"] = repr(e)[:200]
    return rep


def reference_arm(args, rank, world):
    """Times the reference's own C implementation on the host cores (oracle/_ref)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline as cb
    cores = len(os.sched_getaffinity(0))
    th0, th1 = workload_inputs(BATCH)
    # bounded sample: per step each core advances `per` instances by NSTEPS steps (~1-2 s)
    per = 256
    h = cb.Harness("damped_pendulum")
    p = h.R  # noqa
    # p from two configurations exactly like the GPU arm (initialize_from_configs)
    pinit = np.zeros((cores * per, 1))
    for i in range(cores * per):
        h.mvi.initialize_from_configs(0.0, th0[i], DT, th1[i])
        pinit[i] = h.mvi.p2
    shards = [(th1[i * per:(i + 1) * per], pinit[i * per:(i + 1) * per], NSTEPS, DT) for i in range(cores)]
    for _ in range(max(args.warmup, 1)):
        cb.time_parallel("damped_pendulum", "rollouts", shards, reps=1)
    rates = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, procs, wall = cb.time_parallel("damped_pendulum", "rollouts", shards, reps=3)
        rates.append(r)
    wall = time.perf_counter() - t0
    value = float(np.mean(rates))
    units_per_step = cores * per * NSTEPS * 3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": units_per_step / value * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d instances x %d steps x 3 repeats per bench step" % (cores * per, NSTEPS)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "%d processes x %d instances x %d steps x 3, oracle/_ref (unmodified reference C, gcc -O2) "
                                   "driven by oracle/ref_harness.c" % (cores, per, NSTEPS)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def secondary(lib, systems, device, fp64_peak, hbm_peak):
    """linearizations/s for pend-on-cart (W3-shaped batch) and the marionette (W5), device resident."""
    out = []
    rng = np.random.default_rng(1)
    # ---- pend-on-cart: inputs+outputs ~1 GB >> L2
    d = systems.named_desc("pend_on_cart1")
    s = lib.System(d, device=device)
    B = 1 << 22
    q1 = rng.uniform(-0.5, 0.5, (B, 2)); p1 = rng.normal(0, 1, (B, 2)); u1 = rng.uniform(-2, 2, (B, 1))
    up = lambda a: lib.DeviceBuffer(device, a.shape, a.dtype).upload(a)
    dq, dp, du = up(q1), up(p1), up(u1)
    q2 = lib.DeviceBuffer(device, (B, 2)); p2 = lib.DeviceBuffer(device, (B, 2))
    it = lib.DeviceBuffer(device, (B,), np.int32); st = lib.DeviceBuffer(device, (B,), np.int32)
    A = lib.DeviceBuffer(device, (B, 4, 4)); Bm = lib.DeviceBuffer(device, (B, 4, 1))
    ms = []
    for rep in range(6):
        s.linearize_raw(True, B, dq, dp, du, None, st, t1_scalar=0.0, dt_scalar=DT, q2=q2, p2=p2, iters=it, A=A, B=Bm)
        lib.synchronize(device)
        if rep >= 2:
            ms.append(s.last_kernel_ms())
    t = float(np.mean(ms))
    # algorithmic bytes/linearization: in q1,p1,u1 = 40 B; out A,B = 160 B, q2,p2 = 32 B, iters+status = 8 B
    byt = 40 + 160 + 32 + 8
    out.append({"metric": "linearizations/s (pend-on-cart, solve + deriv1 -> A,B)", "value": B / t * 1e3, "unit": "linearizations/s",
                "batch": B, "ms": t, "newton_iters_mean": float(it.download().mean()),
                "roofline": {"bound": "hbm", "achieved": B * byt / t / 1e6, "peak": hbm_peak, "unit": "GB/s",
                             "frac": B * byt / t / 1e6 / hbm_peak, "traffic": None, "bytes_per_unit": byt}})
    for b in (dq, dp, du, q2, p2, it, st, A, Bm):
        b.free()
    s.close()
    # ---- W3 pipeline on the device: 4096 rollouts x 10000 steps by the step kernel (trajectory
    #      captured in HBM), then every (X[k], U[k], hint X[k+1]) linearized in one launch: the
    #      exact-hint case of DSystem.linearize_trajectory (0 Newton iterations)
    d = systems.named_desc("pend_on_cart1")
    s = lib.System(d, device=device)
    R, K = 4096, 10000
    tt = DT * np.arange(K)
    amp = rng.uniform(0, 2, (R, 1)); om = rng.uniform(0.5, 3, (R, 1))
    U = (amp * np.sin(om * tt[None, :]))[:, :, None]
    q0 = np.zeros((R, 2)); q0[:, 1] = rng.uniform(-0.5, 0.5, R)
    du = up(U); dq0 = up(q0); dp0 = up(np.zeros((R, 2)))
    tq = lib.DeviceBuffer(device, (R, K, 2)); tp = lib.DeviceBuffer(device, (R, K, 2))
    q2 = lib.DeviceBuffer(device, (R, 2)); p2 = lib.DeviceBuffer(device, (R, 2))
    itr = lib.DeviceBuffer(device, (R,), np.int32); sr = lib.DeviceBuffer(device, (R,), np.int32)
    s.step_raw(True, R, K, 0.0, DT, dq0, dp0, du, None, None, None, q2, p2, None, itr, sr, sample_every=1, traj_q=tq, traj_p=tp)
    lib.synchronize(device)
    roll_ms = s.last_kernel_ms()

    class Off:
        def __init__(self, buf, off_bytes): self.p = buf.data_ptr() + off_bytes
        def data_ptr(self): return self.p
    # R trajectories of K state rows (x_1 .. x_K) -> R (K-1) linearizations (traj_len = K): instance (r, k) has
    # state x_{k+1}, input U[r][k+1] and hint x_{k+2}; nothing straddles two rollouts (round 1 linearized 4095 such
    # junk instances: those were its 12 non-converged ones)
    n = R * (K - 1)
    du_lin = up(np.ascontiguousarray(U[:, 1:]))
    A = lib.DeviceBuffer(device, (n, 4, 4)); Bm = lib.DeviceBuffer(device, (n, 4, 1))
    it = lib.DeviceBuffer(device, (n,), np.int32); st = lib.DeviceBuffer(device, (n,), np.int32)
    ms = []
    for rep in range(4):
        s.linearize_raw(True, n, tq, tp, du_lin, None, st, t1_scalar=0.0, dt_scalar=DT, q2_guess=Off(tq, 16),
                        iters=it, A=A, B=Bm, traj_len=K)
        lib.synchronize(device)
        if rep >= 1:
            ms.append(s.last_kernel_ms())
    t = float(np.mean(ms))
    byt = 16 + 16 + 8 + 16 + 160 + 8      # q1, p1, u1, hint in; A, B, iters, status out
    st_h = st.download()
    w3 = {"metric": "linearizations/s (W3: pend-on-cart linearize_trajectory, 4096 rollouts x 10000 steps, exact hints)",
          "value": n / t * 1e3, "unit": "linearizations/s", "batch": n, "ms": t,
          "newton_iters_mean": float(it.download().mean()), "ok_fraction": float((st_h == 0).mean()),
          "rollout_kernel_ms": roll_ms, "rollout_steps_per_s": R * K / roll_ms * 1e3,
          "rollouts_ok_fraction": float((sr.download() == 0).mean()),
          "roofline": {"bound": "hbm", "achieved": n * byt / t / 1e6, "peak": hbm_peak, "unit": "GB/s",
                       "frac": n * byt / t / 1e6 / hbm_peak, "traffic": None, "bytes_per_unit": byt}}
    if opcounted("pend_on_cart1_linearization_exact_hint"):
        w3["flops_opcounted"] = opcounted("pend_on_cart1_linearization_exact_hint")
    if (st_h != 0).any():
        bad = np.flatnonzero(st_h != 0)[:64]
        rows = bad + bad // (K - 1)
        tqh, tph = tq.download().reshape(-1, 2), tp.download().reshape(-1, 2)
        w3["failures"] = failure_report("pend_on_cart1", "linearize", np.ones(bad.size, np.int32), q1=tqh[rows], p1=tph[rows],
                                        u1=U[:, 1:].reshape(-1, 1)[bad], hint=tqh[rows + 1], dt=DT)
    out.append(w3)
    # the same pipeline end to end through the host-pointer call: X, U in pinned host memory, A, B back to the host
    ne = 1 << 21
    hq, hp, hu = lib.pinned_empty((ne, 2)), lib.pinned_empty((ne, 2)), lib.pinned_empty((ne, 1))
    hq[:] = tq.download().reshape(-1, 2)[:ne]; hp[:] = tp.download().reshape(-1, 2)[:ne]; hu[:] = U.reshape(-1, 1)[:ne]
    hA, hB = lib.pinned_empty((ne, 4, 4)), lib.pinned_empty((ne, 4, 1))
    hst = lib.pinned_empty((ne,), np.int32)
    te = []
    for rep in range(3):
        t0_ = time.perf_counter()
        s.linearize_raw(False, ne, hq, hp, hu, None, hst, t1_scalar=0.0, dt_scalar=DT, A=hA, B=hB)
        te.append((time.perf_counter() - t0_) * 1e3)
    te = float(np.mean(te[1:]))
    out.append({"metric": "linearizations/s END TO END (W3-shaped: trepb_linearize_batch with pinned host buffers, 2^21 instances, "
                          "host-to-device and device-to-host copies inside the timed region)",
                "value": ne / te * 1e3, "unit": "linearizations/s", "batch": ne, "ms": te,
                "h2d_bytes": ne * 40, "d2h_bytes": ne * 164, "pcie_GBps": ne * 204 / te / 1e6,
                "note": "PCIe-bound: 204 bytes cross the bus per linearization against 224 bytes of HBM traffic at 4.6 TB/s on the "
                        "device-resident path; a caller that needs the gains, not A / B, keeps the slabs on the GPU "
                        "(trepb_lqr_batch_dev) and moves 8 nU nX bytes per step instead"})
    for b in (du, du_lin, dq0, dp0, tq, tp, q2, p2, itr, sr, A, Bm, it, st):
        b.free()
    s.close()
    # ---- W4-shaped: dual pendulums (LinearSpring + LinearDamper), 2^22 instances x 100 steps
    d = systems.named_desc("dual_pendulums")
    s = lib.System(d, device=device)
    B = 1 << 22
    q = rng.uniform(-np.pi, np.pi, (B, 2))
    dq = up(q); dp = lib.DeviceBuffer(device, (B, 2))
    s.calc_p2_raw(True, B, DT, dq, dq, dp)
    q2 = lib.DeviceBuffer(device, (B, 2)); p2 = lib.DeviceBuffer(device, (B, 2))
    it = lib.DeviceBuffer(device, (B,), np.int32); st = lib.DeviceBuffer(device, (B,), np.int32)
    ms = []
    for rep in range(3):
        s.step_raw(True, B, 100, DT, DT, dq, dp, None, None, None, None, q2, p2, None, it, st)
        lib.synchronize(device)
        if rep >= 1:
            ms.append(s.last_kernel_ms())
    t = float(np.mean(ms))
    st_h = st.download()
    w4 = {"metric": "DEL steps/s (W4: dual pendulums Monte-Carlo sweep, 2^22 instances x 100 steps)", "value": B * 100 / t * 1e3,
          "unit": "DEL steps/s", "batch": B, "ms": t, "newton_iters_per_step": float(it.download().mean()) / 100,
          "ok_fraction": float((st_h == 0).mean()),
          "failures": failure_report("dual_pendulums", "rollouts", st_h, q=q, p=dp.download(), nsteps=100, t0=DT, dt=DT)}
    if opcounted("dual_pendulums_step"):
        w4["flops_opcounted"] = opcounted("dual_pendulums_step")
    try:
        with open(os.path.join(ROOT, "profiles", "flops.json")) as fh:
            fl4 = float(json.load(fh)["dual_pendulums_step_flops_per_del_step"])
        ach = fl4 * B * 100 / (t * 1e-3) / 1e12
        w4["roofline"] = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                          "flops_per_unit": fl4, "note": "on-chip: 32 B in + 32 B out per rollout of 100 steps"}
    except (OSError, KeyError):
        pass
    out.append(w4)
    for b in (dq, dp, q2, p2, it, st):
        b.free()
    s.close()
    # ---- marionette
    d = systems.named_desc("puppet")
    g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
    s = lib.System(d, device=device)
    B = 131072          # >= 148 SMs x 16 resident warps x 32 lanes, so the table-driven kernel fills the GPU
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    dq, dp, dk, dl = up(q1), up(p1), up(g["roll_k2"][idx]), up(g["roll_lambda"][idx - 1])
    q2 = lib.DeviceBuffer(device, (B, d.nq)); p2 = lib.DeviceBuffer(device, (B, d.nd)); l2 = lib.DeviceBuffer(device, (B, d.nc))
    it = lib.DeviceBuffer(device, (B,), np.int32); st = lib.DeviceBuffer(device, (B,), np.int32)
    A = lib.DeviceBuffer(device, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(device, (B, d.nX, d.nU))
    ms = []
    for rep in range(4):
        s.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=DT, lambda_guess=dl, q2=q2, p2=p2,
                        lambda1=l2, iters=it, A=A, B=Bm)
        lib.synchronize(device)
        if rep >= 1:
            ms.append(s.last_kernel_ms())
    t = float(np.mean(ms))
    entry = {"metric": "linearizations/s (W5: marionette nd22/nk18/nc6, solve + deriv1 -> A,B)", "value": B / t * 1e3,
             "unit": "linearizations/s", "batch": B, "ms": t, "newton_iters_mean": float(it.download().mean()),
             "ok_fraction": float((st.download() == 0).mean()), "kernel": s.kernel_name,
             "kernel_info": s.kernel_info(2)}
    if opcounted("puppet_linearization"):
        entry["flops_opcounted"] = opcounted("puppet_linearization")
    fj = os.path.join(ROOT, "profiles", "flops.json")
    if os.path.exists(fj):
        prof = json.load(open(fj))
        fl = prof.get("puppet_coop_lin_flops_per_linearization")
        tb = prof.get("puppet_coop_ext_lin_dram_bytes_per_linearization" if s.kernel_name.endswith("/ext")
                      else "puppet_coop_lin_dram_bytes_per_linearization")
        alg = 8 * (d.nq + d.nd + d.nk + d.nc) + 8 * (d.nX * d.nX + d.nX * d.nU + d.nq + d.nd + d.nc) + 8
        if fl:
            ach = fl * B / (t * 1e-3) / 1e12
            entry["roofline"] = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                                 "flops_per_unit": fl, "traffic": tb * B if tb else None,
                                 "hbm": {"achieved": alg * B / t / 1e6, "peak": hbm_peak, "unit": "GB/s",
                                         "frac": alg * B / t / 1e6 / hbm_peak, "bytes_per_unit": alg},
                                 "note": "cooperative kernel: one warp per instance, link tables + 12.3 KB workspace per instance in shared "
                                         "memory and 14.1 KB (DDh.lambda block, pair arrays, constraint Jacobians) in an L2-resident slab: "
                                         "16 instances per SM at 128 registers; DRAM traffic is the A/B output plus write-backs of the "
                                         "slab. 68 k warp instructions per linearization; throughput grows almost linearly with the "
                                         "instances in flight up to here (fp64 pipe 20 %, warps active 25 %, shared-memory pipe next); "
                                         "the flops are the executed count of the two-warp flavour (4.98e5, ncu sass counters; this "
                                         "flavour executes 5.84e5 for the same result), so frac is pipe utilisation "
                                         "(ncu: profiles/r02zb_coop_lin_ext16_*.txt; DESIGN.md section 4b)"}
    out.append(entry)
    # the same batch through the thread-per-instance table-driven kernel (the round's starting point)
    s_thr = lib.System(d, device=device, cooperative=False)
    ms = []
    for rep in range(3):
        s_thr.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=DT, lambda_guess=dl, q2=q2, p2=p2,
                            lambda1=l2, iters=it, A=A, B=Bm)
        lib.synchronize(device)
        if rep >= 1:
            ms.append(s_thr.last_kernel_ms())
    tt = float(np.mean(ms))
    out.append({"metric": "linearizations/s (W5 marionette, thread-per-instance kernel with the workspace in HBM, for comparison)",
                "value": B / tt * 1e3, "unit": "linearizations/s", "batch": B, "ms": tt, "kernel": s_thr.kernel_name})
    s_thr.close()
    # marionette DEL steps (trepb_step_batch): 16 steps per instance in-kernel
    Bs, ns = 32768, 16
    k2s = np.repeat(g["roll_k2"][idx[:Bs]][:, None, :], ns, axis=1)
    dks = up(np.ascontiguousarray(k2s))
    ms = []
    for rep in range(3):
        s.step_raw(True, Bs, ns, 0.0, DT, dq, dp, None, dks, None, dl, q2, p2, l2, it, st)
        lib.synchronize(device)
        if rep >= 1:
            ms.append(s.last_kernel_ms())
    ts = float(np.mean(ms))
    out.append({"metric": "DEL steps/s (W5 marionette, %d steps per instance in-kernel, kinematic strings held)" % ns,
                "value": Bs * ns / ts * 1e3, "unit": "DEL steps/s", "batch": Bs, "ms": ts,
                "newton_iters_per_step": float(it.download().mean()) / ns, "ok_fraction": float((st.download() == 0).mean()),
                "kernel": s.kernel_name})
    dks.free()
    # closed-loop rollouts (DSystem.project / armijo_simulate as one batch), marionette: nominal
    # golden rollout, wiggled strings, small gains on the configuration error
    Bp, Kp = 8192, 10
    nX, nU, nq, nd = d.nX, d.nU, d.nq, d.nd
    Xn = np.zeros((Kp + 1, nX)); Xn[:, :nq] = g["roll_q"][:Kp + 1]; Xn[:, nq:nq + nd] = g["roll_p"][:Kp + 1]
    Xn[1:, nq + nd:] = (g["roll_q"][1:Kp + 1, nd:] - g["roll_q"][:Kp, nd:]) / DT
    Kfb = np.zeros((Kp, nU, nX)); Kfb[:, :, :nd] = rng.normal(0, 2e-3, (Kp, nU, nd))
    tt = DT * np.arange(Kp)[None, :, None]
    bUp = g["roll_k2"][:Kp][None] + 2e-4 * np.sin(40 * tt + rng.uniform(0, 6, (Bp, 1, nU)))
    dbX = up(np.ascontiguousarray(np.repeat(Xn[None], Bp, axis=0))); dbU = up(np.ascontiguousarray(bUp)); dK = up(Kfb)
    oX = lib.DeviceBuffer(device, (Bp, Kp + 1, nX)); oU = lib.DeviceBuffer(device, (Bp, Kp, nU))
    pit = lib.DeviceBuffer(device, (Bp,), np.int32); pst = lib.DeviceBuffer(device, (Bp,), np.int32)
    ms = []
    for rep in range(3):
        s.project_raw(True, Bp, Kp, DT, DT, dbX, dbU, dK, oX, oU, pst, iters=pit)
        lib.synchronize(device)
        if rep >= 1:
            ms.append(s.last_kernel_ms())
    tp = float(np.mean(ms))
    out.append({"metric": "closed-loop DEL steps/s (W5 marionette, DSystem.project / armijo_simulate as one batch: %d candidates x %d steps)" % (Bp, Kp),
                "value": Bp * Kp / tp * 1e3, "unit": "DEL steps/s", "batch": Bp, "ms": tp,
                "newton_iters_per_step": float(pit.download().mean()) / Kp, "ok_fraction": float((pst.download() == 0).mean()),
                "kernel": s.kernel_name})
    for b_ in (dbX, dbU, dK, oX, oU, pit, pst):
        b_.free()
    # Riccati sweep (discopt.dlqr.solve_tv_lqr) at the marionette's size, one CTA per rollout
    Rl, Kl = 296, 64
    Al = up(np.eye(nX)[None, None] + rng.normal(0, 0.03, (Rl, Kl, nX, nX)))
    Bl = up(rng.normal(0, 1.0, (Rl, Kl, nX, nU)))
    Ql, Rr = up(np.eye(nX)), up(np.eye(nU))
    Kl_out = lib.DeviceBuffer(device, (Rl, Kl, nU, nX)); lst = lib.DeviceBuffer(device, (Rl,), np.int32)
    lib.lqr_raw(True, device, Rl, Kl, nX, nU, Al, Bl, Ql, Rr, Kl_out, lst)
    lib.synchronize(device)
    with lib.EventTimer(device) as tm:
        for rep in range(3):
            lib.lqr_raw(True, device, Rl, Kl, nX, nU, Al, Bl, Ql, Rr, Kl_out, lst)
    tl = tm.ms / 3
    fl_step = 2.0 * (2 * nX ** 3 + 3 * nX * nX * nU + nU * nU * nX) + 2.0 * nU * nU * (nU / 3.0 + nX)
    out.append({"metric": "Riccati steps/s (discopt.dlqr.solve_tv_lqr at the marionette's size nX=%d nU=%d, %d rollouts x %d steps)" % (nX, nU, Rl, Kl),
                "value": Rl * Kl / tl * 1e3, "unit": "Riccati steps/s", "batch": Rl, "ms": tl,
                "ok_fraction": float((lst.download() == 0).mean()),
                "roofline": {"bound": "fp64", "achieved": fl_step * Rl * Kl / (tl * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": fl_step * Rl * Kl / (tl * 1e-3) / 1e12 / fp64_peak, "flops_per_unit": fl_step,
                             "kernel_ms_last_launch": lib.lqr_last_kernel_ms(device),
                             "note": "CUDA events around 3 launches on the launching stream (cross-check: the library's own events around the last launch, trepb_lqr_last_kernel_ms); flops counted from the matrix shapes; products on the FP64 tensor cores (mma.sync m8n8k4), gamma solve by a CTA-wide Gauss-Jordan elimination"}})
    for b_ in (Al, Bl, Ql, Rr, Kl_out, lst):
        b_.free()
    # second derivatives, z-contracted output (the form DOptimizer.calc_newton_model consumes)
    Bd = 4096
    z = up(rng.normal(0, 1, (Bd, d.nX)))
    xx = lib.DeviceBuffer(device, (Bd, d.nX, d.nX)); xu = lib.DeviceBuffer(device, (Bd, d.nX, d.nU)); uu = lib.DeviceBuffer(device, (Bd, d.nU, d.nU))
    ms = []
    for rep in range(2):
        s.deriv2_raw(True, Bd, dq, dp, None, dk, st, {}, z=z, fdxdx=xx, fdxdu=xu, fdudu=uu, t1_scalar=0.0, dt_scalar=DT,
                     lambda_guess=dl)
        lib.synchronize(device)
        ms.append(s.last_kernel_ms())
    t = float(ms[-1])
    # fp64 flops executed and DRAM bytes moved per evaluation: ncu counters (profiles/r01f_d2_launches.csv via
    # profiles/flops.json; DRAM measured with all 30 tensors written, mostly the dual workspace of pass A)
    d2_flops, d2_dram = 26.3e6, 24.1e6
    try:
        with open(os.path.join(ROOT, "profiles", "flops.json")) as fh:
            fj_ = json.load(fh)
        d2_flops = float(fj_.get("puppet_d2_flops_per_evaluation", d2_flops))
        d2_dram = float(fj_.get("puppet_d2_dram_bytes_per_evaluation", d2_dram))
    except OSError:
        pass
    out.append({"metric": "second-derivative evaluations/s (W5: marionette, 3240 parameter pairs, z-contracted fdxdx/fdxdu/fdudu)",
                "value": Bd / t * 1e3, "unit": "evaluations/s", "batch": Bd, "ms": t,
                "scheme": "pass A: one dual-number evaluation of the Jacobian tables per (instance, parameter), 80 per instance; "
                          "pass B: contraction + LU solves per parameter pair (trepb_d2jac.cuh)",
                "roofline": {"bound": "hbm", "achieved": d2_dram * Bd / (t * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": d2_dram * Bd / (t * 1e-3) / 1e9 / hbm_peak, "bytes_per_unit": d2_dram,
                             "algorithmic_bytes_per_unit": 8 * (d.nX * d.nX + d.nX * d.nU + d.nU * d.nU + d.nX),
                             "fp64": {"achieved": d2_flops * Bd / (t * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                      "frac": d2_flops * Bd / (t * 1e-3) / 1e12 / fp64_peak, "flops_per_unit": d2_flops},
                             "note": "bytes_per_unit is MEASURED DRAM traffic (pass A keeps its per-thread dual workspace in HBM), "
                                     "not algorithmic bytes; the time covers both passes, not the preceding linearize launch"}})
    for b in (dq, dp, dk, dl, q2, p2, l2, it, st, A, Bm, z, xx, xu, uu):
        b.free()
    s.close()
    out += extra_kinds(lib, systems, device, hbm_peak, fp64_peak)
    return out


def extra_kinds(lib, systems, device, hbm_peak, fp64_peak):
    """The rows widened beyond BASELINE.json's configs, measured like the others:
    linearizations/s of the table-driven thread kernels on systems with PointOnPlane constraints, wrenches
    and a spline spring, and the affine LQ sweep."""
    out = []
    rng = np.random.default_rng(2)
    up = lambda a: lib.DeviceBuffer(device, a.shape, a.dtype).upload(np.ascontiguousarray(a))
    for name, B in (("pccd", 1 << 17), ("wrench_arm", 1 << 20), ("spline_pendulum", 1 << 20)):
        d = systems.named_desc(name)
        s = lib.System(d, device=device)
        nq, nd, nu, nc = d.nq, d.nd, d.nu, d.nc
        lam = None
        if name == "pccd":
            g = np.load(os.path.join(ROOT, "tests", "golden", "pccd.npz"))
            idx = rng.integers(1, g["roll_q"].shape[0] - 1, B)
            q1 = g["roll_q"][idx] + rng.normal(0, 0.01, (B, nq)); p1 = g["roll_p"][idx] + rng.normal(0, 0.05, (B, nd))
            lam = up(g["roll_lambda"][idx - 1])
        elif name == "wrench_arm":
            q1 = np.stack([rng.uniform(-np.pi, np.pi, B), rng.uniform(-1.2, 1.2, B), rng.uniform(-0.3, 0.5, B)], axis=1)
            p1 = rng.normal(0, 1, (B, nd))
        else:
            q1 = np.stack([rng.uniform(-2.4, 2.2, B), rng.uniform(-np.pi, np.pi, B)], axis=1)
            p1 = rng.normal(0, 1, (B, nd))
        dq, dp = up(q1), up(p1)
        du = up(rng.uniform(-2, 2, (B, nu))) if nu else None
        q2 = lib.DeviceBuffer(device, (B, nq)); p2 = lib.DeviceBuffer(device, (B, nd))
        l2 = lib.DeviceBuffer(device, (B, nc)) if nc else None
        it = lib.DeviceBuffer(device, (B,), np.int32); st = lib.DeviceBuffer(device, (B,), np.int32)
        A = lib.DeviceBuffer(device, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(device, (B, d.nX, d.nU)) if d.nU else None
        ms = []
        for rep in range(4):
            s.linearize_raw(True, B, dq, dp, du, None, st, t1_scalar=0.0, dt_scalar=DT, q2=q2, p2=p2, lambda1=l2, iters=it,
                            A=A, B=Bm, lambda_guess=lam)
            lib.synchronize(device)
            if rep >= 1:
                ms.append(s.last_kernel_ms())
        t = float(np.mean(ms))
        byt = 8 * (nq + nd + nu + nc) + 8 * (d.nX * d.nX + d.nX * d.nU) + 8 * (nq + nd + nc) + 8
        out.append({"metric": "linearizations/s (%s: %s; solve + deriv1 -> A,B)" % (name, {
                        "pccd": "examples/pccd.py, 7 DOF closed chain, 4 PointOnPlane constraints",
                        "wrench_arm": "3 DOF arm with Body / Hybrid / Spatial wrenches, 5 inputs",
                        "spline_pendulum": "2 DOF, NonlinearConfigSpring over a quintic spline"}[name]),
                    "value": B / t * 1e3, "unit": "linearizations/s", "batch": B, "ms": t, "kernel": s.kernel_name,
                    "newton_iters_mean": float(it.download().mean()), "ok_fraction": float((st.download() == 0).mean()),
                    "failures": failure_report(name, "linearize", st.download(), q1=q1, p1=p1, u1=du.download() if nu else None,
                                               lam=lam.download() if lam is not None else None, dt=DT),
                    "roofline": {"bound": "hbm", "achieved": B * byt / t / 1e6, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": B * byt / t / 1e6 / hbm_peak, "traffic": None, "bytes_per_unit": byt,
                                 "note": "algorithmic bytes (inputs + A, B, q2, p2, lambda, iters, status)"}})
        for b in (dq, dp, du, lam, q2, p2, l2, it, st, A, Bm):
            if b is not None:
                b.free()
        s.close()
    # second derivatives of the closed chain (PointOnPlane constraints), z-contracted
    d = systems.named_desc("pccd")
    s = lib.System(d, device=device)
    Bp = 1 << 14
    g = np.load(os.path.join(ROOT, "tests", "golden", "pccd.npz"))
    idx = rng.integers(1, g["roll_q"].shape[0] - 1, Bp)
    dq = up(g["roll_q"][idx] + rng.normal(0, 0.01, (Bp, d.nq))); dp = up(g["roll_p"][idx] + rng.normal(0, 0.05, (Bp, d.nd)))
    lam = up(g["roll_lambda"][idx - 1]); st = lib.DeviceBuffer(device, (Bp,), np.int32)
    z = up(rng.normal(0, 1, (Bp, d.nX))); xx = lib.DeviceBuffer(device, (Bp, d.nX, d.nX))
    ms = []
    for rep in range(3):
        s.deriv2_raw(True, Bp, dq, dp, None, None, st, {}, z=z, fdxdx=xx, t1_scalar=0.0, dt_scalar=DT, lambda_guess=lam)
        lib.synchronize(device)
        if rep >= 1:
            ms.append(s.last_kernel_ms())
    t = float(np.mean(ms))
    out.append({"metric": "second-derivative evaluations/s (pccd: 7 DOF closed chain, 105 parameter pairs, z-contracted fdxdx)",
                "value": Bp / t * 1e3, "unit": "evaluations/s", "batch": Bp, "ms": t, "kernel": s.kernel_name,
                "ok_fraction": float((st.download() == 0).mean()),
                "note": "both second-derivative passes (dual Jacobian tables per parameter, compile-time-size per-pair solves); "
                        "the preceding cooperative linearize launch is not included"})
    for b in (dq, dp, lam, st, z, xx):
        b.free()
    s.close()
    # W1 (BASELINE config 0, examples/pendulum.py): N-link pendulum, a large batch and the latency of ONE rollout
    for links, B, nsteps in ((1, 1 << 20, 1000), (5, 1 << 18, 200), (1, 1, 1000), (5, 1, 1000)):
        d = systems.named_desc("pendulum%d" % links)
        s = lib.System(d, device=device)
        q0 = np.zeros((B, links)); q0[:, 0] = rng.uniform(-np.pi, np.pi, B) if B > 1 else np.pi / 4
        dq = up(q0); dp = lib.DeviceBuffer(device, (B, links))
        s.calc_p2_raw(True, B, DT, dq, dq, dp)
        q2 = lib.DeviceBuffer(device, (B, links)); p2 = lib.DeviceBuffer(device, (B, links))
        it = lib.DeviceBuffer(device, (B,), np.int32); st = lib.DeviceBuffer(device, (B,), np.int32)
        ms = []
        for rep in range(3):
            s.step_raw(True, B, nsteps, DT, DT, dq, dp, None, None, None, None, q2, p2, None, it, st)
            lib.synchronize(device)
            if rep >= 1:
                ms.append(s.last_kernel_ms())
        t = float(np.mean(ms))
        out.append({"metric": "DEL steps/s (W1: %d-link pendulum of examples/pendulum.py, %s)" % (
                        links, "2^%d rollouts x %d steps" % (int(np.log2(B)), nsteps) if B > 1 else "ONE rollout of %d steps: latency" % nsteps),
                    "value": B * nsteps / t * 1e3, "unit": "DEL steps/s", "batch": B, "ms": t, "kernel": s.kernel_name,
                    "us_per_step": t * 1e3 / nsteps if B == 1 else None,
                    "newton_iters_per_step": float(it.download().mean()) / nsteps, "ok_fraction": float((st.download() == 0).mean())})
        if links == 5 and B > 1:
            fj = os.path.join(ROOT, "profiles", "flops.json")
            fl5 = json.load(open(fj)).get("pendulum5_step_flops_per_del_step") if os.path.exists(fj) else None
            if fl5:
                ach = fl5 * B * nsteps / (t * 1e-3) / 1e12
                out[-1]["roofline"] = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                                       "flops_per_unit": fl5, "kernel_info": s.kernel_info(0),
                                       "note": "register-resident specialised kernel (255 registers, 80 B of stack, one CTA of 256 threads "
                                               "per SM); executed flops per DEL step from ncu (profiles/r02m_flops_pend5.csv): 53 % of "
                                               "the 6.0 k warp instructions per step are fp64; on-chip (85 B of DRAM per rollout)"}
        for b in (dq, dp, q2, p2, it, st):
            b.free()
        s.close()
    # affine LQ sweep (solve_tv_lq) at the marionette's size, one cost set per rollout
    nX, nU, Rl, Kl = 80, 18, 296, 64
    Al = up(np.eye(nX)[None, None] + rng.normal(0, 0.3 / np.sqrt(nX), (Rl, Kl, nX, nX)))
    Bl = up(rng.normal(0, 1.0, (Rl, Kl, nX, nU)))
    Ql = up(np.broadcast_to(np.eye(nX), (Rl, Kl + 1, nX, nX)).copy()); Rr = up(np.broadcast_to(2.0 * np.eye(nU), (Rl, Kl, nU, nU)).copy())
    Sl = up(rng.normal(0, 0.02, (Rl, Kl, nX, nU))); ql = up(rng.normal(0, 1, (Rl, Kl + 1, nX))); rl = up(rng.normal(0, 1, (Rl, Kl, nU)))
    Ko = lib.DeviceBuffer(device, (Rl, Kl, nU, nX)); Co = lib.DeviceBuffer(device, (Rl, Kl, nU)); lst = lib.DeviceBuffer(device, (Rl,), np.int32)
    lib.lq_raw(True, device, Rl, Kl, nX, nU, Al, Bl, Ql, Sl, Rr, ql, rl, Ko, Co, lst, cost_per_rollout=True)
    lib.synchronize(device)
    with lib.EventTimer(device) as tm:
        for rep in range(3):
            lib.lq_raw(True, device, Rl, Kl, nX, nU, Al, Bl, Ql, Sl, Rr, ql, rl, Ko, Co, lst, cost_per_rollout=True)
    tl = tm.ms / 3
    fl_step = 2.0 * (2 * nX ** 3 + 3 * nX * nX * nU + nU * nU * nX) + 2.0 * nU * nU * (nU / 3.0 + nX) + 2.0 * (nX * nX + 3 * nX * nU + nU * nU)
    out.append({"metric": "Riccati steps/s (discopt.dlqr.solve_tv_lq: cross term + affine recursion, nX=%d nU=%d, %d rollouts x %d steps, one cost set per rollout)" % (nX, nU, Rl, Kl),
                "value": Rl * Kl / tl * 1e3, "unit": "Riccati steps/s", "batch": Rl, "ms": tl,
                "ok_fraction": float((lst.download() == 0).mean()),
                "roofline": {"bound": "fp64", "achieved": fl_step * Rl * Kl / (tl * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": fl_step * Rl * Kl / (tl * 1e-3) / 1e12 / fp64_peak, "flops_per_unit": fl_step,
                             "note": "CUDA events around 3 launches; flops counted from the matrix shapes"}})
    for b_ in (Al, Bl, Ql, Rr, Sl, ql, rl, Ko, Co, lst):
        b_.free()
    return out


def w4_full(lib, systems, device, world, reduce_max, fp64_peak):
    """W4 at north_star's size: 1.25e7 dual pendulums per GPU (1e8 on 8 GPUs) x 1000 steps, one launch."""
    rng = np.random.default_rng(100 + int(os.environ.get("RANK", "0")))
    d = systems.named_desc("dual_pendulums")
    s = lib.System(d, device=device)
    B, ns = 12500000, 1000
    q = rng.uniform(-np.pi, np.pi, (B, 2))
    up = lambda a: lib.DeviceBuffer(device, a.shape, a.dtype).upload(a)
    dq = up(q); dp = lib.DeviceBuffer(device, (B, 2))
    s.calc_p2_raw(True, B, DT, dq, dq, dp)
    q2 = lib.DeviceBuffer(device, (B, 2)); p2 = lib.DeviceBuffer(device, (B, 2))
    it = lib.DeviceBuffer(device, (B,), np.int32); st = lib.DeviceBuffer(device, (B,), np.int32)
    s.step_raw(True, B, 20, DT, DT, dq, dp, None, None, None, None, q2, p2, None, it, st)     # warm-up
    lib.synchronize(device)
    s.step_raw(True, B, ns, DT, DT, dq, dp, None, None, None, None, q2, p2, None, it, st)
    lib.synchronize(device)
    t = reduce_max(s.last_kernel_ms())
    st_h = st.download()
    iters = it.download()
    hist = np.bincount(np.clip(np.rint(iters[st_h == 0] / float(ns)).astype(np.int64), 0, 8), minlength=9)
    e = {"metric": "DEL steps/s (W4 at north_star's size: %.3g dual pendulums per GPU x %d steps, %d GPU(s), weak scaling)" % (B, ns, world),
         "value": float(B) * ns * world / t * 1e3, "unit": "DEL steps/s", "batch_per_gpu": B, "nsteps": ns, "ms": t,
         "newton_iters_per_step": float(iters.mean()) / ns, "ok_fraction_rank0": float((st_h == 0).mean()),
         "iteration_histogram_rank0": hist.tolist(),
         "failures_rank0": failure_report("dual_pendulums", "rollouts", st_h, limit=16, q=q, p=dp.download(), nsteps=ns, t0=DT, dt=DT)}
    try:
        with open(os.path.join(ROOT, "profiles", "flops.json")) as fh:
            fl4 = float(json.load(fh)["dual_pendulums_step_flops_per_del_step"])
        ach = fl4 * B * ns / (t * 1e-3) / 1e12
        e["roofline"] = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s per GPU", "frac": ach / fp64_peak,
                         "flops_per_unit": fl4}
    except (OSError, KeyError):
        pass
    for b in (dq, dp, q2, p2, it, st):
        b.free()
    s.close()
    return e


def multi_gpu(lib, systems, group, device, rank, world, reduce_max, fp64_peak, hbm_peak):
    """N > 1: the BASELINE metric's other half (linearizations/s) on all ranks, and the one exchange of the path -
    collecting the A / B slabs on rank 0 - measured two ways.  Every time is the max over ranks of a device time."""
    from trep_b200 import dist as D_
    out = []
    up = lambda a, dt=np.float64: lib.DeviceBuffer(device, a.shape, dt).upload(np.ascontiguousarray(a, dtype=dt))

    class Off:
        def __init__(self, p): self.p = int(p)
        def data_ptr(self): return self.p
    # ---- W3, strong scaling: the fixed 4096 rollouts x 10000 steps, block-partitioned by rollout
    rng = np.random.default_rng(1)
    d = systems.named_desc("pend_on_cart1")
    s = lib.System(d, device=device)
    Rtot, K = 4096, 10000
    lo, hi = D_.shard_range(Rtot, rank, world)
    R = hi - lo
    tt = DT * np.arange(K)
    amp = rng.uniform(0, 2, (Rtot, 1)); om = rng.uniform(0.5, 3, (Rtot, 1)); th0 = rng.uniform(-0.5, 0.5, Rtot)
    U = (amp[lo:hi] * np.sin(om[lo:hi] * tt[None, :]))[:, :, None]
    q0 = np.zeros((R, 2)); q0[:, 1] = th0[lo:hi]
    du = up(U); du_lin = up(U[:, 1:]); dq0 = up(q0); dp0 = up(np.zeros((R, 2)))
    tq = lib.DeviceBuffer(device, (R, K, 2)); tp = lib.DeviceBuffer(device, (R, K, 2))
    q2 = lib.DeviceBuffer(device, (R, 2)); p2 = lib.DeviceBuffer(device, (R, 2))
    itr = lib.DeviceBuffer(device, (R,), np.int32); sr = lib.DeviceBuffer(device, (R,), np.int32)
    s.step_raw(True, R, K, 0.0, DT, dq0, dp0, du, None, None, None, q2, p2, None, itr, sr, sample_every=1, traj_q=tq, traj_p=tp)
    lib.synchronize(device)
    n, ntot = R * (K - 1), Rtot * (K - 1)
    rowA, rowB = 128, 32
    A = lib.DeviceBuffer(device, (n, 4, 4)); Bm = lib.DeviceBuffer(device, (n, 4, 1))
    it = lib.DeviceBuffer(device, (n,), np.int32); st = lib.DeviceBuffer(device, (n,), np.int32)

    def lin(Ao, Bo):
        s.linearize_raw(True, n, tq, tp, du_lin, None, st, t1_scalar=0.0, dt_scalar=DT, q2_guess=Off(tq.data_ptr() + 16),
                        iters=it, A=Ao, B=Bo, traj_len=K)
        lib.synchronize(device)
        return s.last_kernel_ms()
    ms = [lin(A, Bm) for _ in range(3)][1:]
    group.barrier()
    t_lin = reduce_max(float(np.mean(ms)))
    ok = float((st.download() == 0).mean())
    # (a) NCCL: grouped ncclSend / ncclRecv of the finished slabs to rank 0 (equal blocks: 4096 % world == 0)
    fullA = lib.DeviceBuffer(device, (ntot, 4, 4)) if rank == 0 else None
    fullB = lib.DeviceBuffer(device, (ntot, 4, 1)) if rank == 0 else None
    tg = []
    for rep in range(3):
        group.barrier()
        with lib.EventTimer(device) as tm:
            group.comm.gather(A, fullA, n * rowA, 0)
            group.comm.gather(Bm, fullB, n * rowB, 0)
        tg.append(tm.ms)
    t_gather = reduce_max(float(np.mean(tg[1:])))
    check = None
    if rank == 0:
        # the first block of the gathered slab is rank 0's own: bitwise equal to its local slab
        a0 = A.download()[:1000]
        check = bool(np.array_equal(fullA.download()[:1000], a0))
    # (b) fused: rank 0's slab mapped into every rank, the linearize kernel stores straight into it
    if fullA is not None:
        fullA.free(); fullB.free()
    slab = D_.SharedSlab(device, group.exchange, ntot * (rowA + rowB))
    offB = ntot * rowA
    first = lo * (K - 1)
    group.barrier()
    ms = []
    for rep in range(3):
        group.barrier()
        ms.append(lin(slab.at(first * rowA), slab.at(offB + first * rowB)))
    group.barrier()
    t_fused = reduce_max(float(np.mean(ms[1:])))
    fused_ok = None
    if rank == 0:
        raw = slab.local.download()
        fused_ok = bool(np.array_equal(raw[:1000 * rowA].view(np.float64).reshape(1000, 4, 4), a0))
        last = raw[(ntot - 1) * rowA:ntot * rowA].view(np.float64)
        fused_ok = fused_ok and bool(np.all(np.isfinite(last)) and np.any(last != 0))
    group.barrier()
    slab.close()
    moved = (rowA + rowB) * float(ntot) * (world - 1) / world      # bytes that cross NVLink into rank 0
    out.append({"metric": "linearizations/s (W3 sharded over %d GPUs: 4096 rollouts x 10000 steps block-partitioned by rollout, exact hints)" % world,
                "value": ntot / t_lin * 1e3, "unit": "linearizations/s", "batch_total": ntot, "ms": t_lin, "scaling": "strong",
                "ok_fraction_rank0": ok,
                "gather_nccl": {"what": "A / B slabs of every rank collected on rank 0: trepb_comm_gather_dev (grouped ncclSend / ncclRecv), "
                                        "device to device", "ms": t_gather, "bytes_into_rank0": moved,
                                "nvlink_GBps_into_rank0": moved / t_gather / 1e6, "first_block_bitwise_equal": check,
                                "linearize_plus_gather_per_s": ntot / (t_lin + t_gather) * 1e3},
                "gather_fused": {"what": "rank 0's slab mapped into every rank (trepb_ipc_open); the linearize kernel's own stores deliver "
                                         "A / B over NVLink - no second pass", "ms": t_fused,
                                 "linearizations_per_s_delivered_to_rank0": ntot / t_fused * 1e3,
                                 "nvlink_GBps_into_rank0": moved / t_fused / 1e6, "slab_checked": fused_ok}})
    for b in (du, du_lin, dq0, dp0, tq, tp, q2, p2, itr, sr, A, Bm, it, st):
        b.free()
    s.close()
    # ---- W5 marionette linearizations/s, weak scaling (131072 instances per GPU)
    d = systems.named_desc("puppet")
    g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
    s = lib.System(d, device=device)
    rng = np.random.default_rng(50 + rank)
    B = 131072
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    dq, dp, dk, dl = up(q1), up(p1), up(g["roll_k2"][idx]), up(g["roll_lambda"][idx - 1])
    it = lib.DeviceBuffer(device, (B,), np.int32); st = lib.DeviceBuffer(device, (B,), np.int32)
    A = lib.DeviceBuffer(device, (B, d.nX, d.nX)); Bm = lib.DeviceBuffer(device, (B, d.nX, d.nU))
    ms = []
    for rep in range(3):
        s.linearize_raw(True, B, dq, dp, None, dk, st, t1_scalar=0.0, dt_scalar=DT, lambda_guess=dl, iters=it, A=A, B=Bm)
        lib.synchronize(device)
        ms.append(s.last_kernel_ms())
    t = reduce_max(float(np.mean(ms[1:])))
    out.append({"metric": "linearizations/s (W5 marionette nd22/nk18/nc6 on %d GPUs, %d instances per GPU, weak scaling)" % (world, B),
                "value": float(B) * world / t * 1e3, "unit": "linearizations/s", "ms": t, "kernel": s.kernel_name,
                "ok_fraction_rank0": float((st.download() == 0).mean()), "newton_iters_mean": float(it.download().mean())})
    for b in (dq, dp, dk, dl, it, st, A, Bm):
        b.free()
    s.close()
    out.append(w4_full(lib, systems, device, world, reduce_max, fp64_peak))
    return out


def w1_cpu_reference(lines):
    """BASELINE config 0 beside the reference: the same N-link pendulum rollout on ONE host core through the
    reference's own C (oracle/_ref via oracle/ref_harness.c) - its per-core rate and its per-step latency -
    attached to the W1 lines."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline as cb
    for links in (1, 5):
        h = cb.Harness("pendulum%d" % links)
        q0 = np.zeros((1, links)); q0[0, 0] = np.pi / 4
        h.mvi.initialize_from_configs(0.0, q0[0], DT, q0[0])
        p0 = np.array([h.mvi.p2], float).reshape(1, links)
        h.rollouts(q0, p0, 200, DT, DT)
        nsteps, t0 = 4000 if links == 1 else 1000, time.perf_counter()
        h.rollouts(q0, p0, nsteps, DT, DT)
        dt_ = time.perf_counter() - t0
        for m in lines:
            if m["metric"].startswith("DEL steps/s (W1: %d-link" % links):
                m["reference_one_core"] = {"steps_per_s": nsteps / dt_, "us_per_step": dt_ * 1e6 / nsteps,
                                           "sample": "one rollout of %d steps, oracle/_ref in the C loop of oracle/ref_harness.c" % nsteps}
                if m.get("us_per_step"):
                    m["latency_vs_reference_core"] = (dt_ * 1e6 / nsteps) / m["us_per_step"]


def cpu_baseline_sample():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline as cb
    cores = len(os.sched_getaffinity(0))
    cores_ = len(os.sched_getaffinity(0))
    th0, th1 = workload_inputs(cores_ * 512)
    h = cb.Harness("damped_pendulum")
    per = 512
    n = cores * per
    pinit = np.zeros((n, 1))
    for i in range(n):
        h.mvi.initialize_from_configs(0.0, th0[i], DT, th1[i])
        pinit[i] = h.mvi.p2
    shards = [(th1[i * per:(i + 1) * per], pinit[i * per:(i + 1) * per], NSTEPS, DT) for i in range(n // per)]
    rate, procs, wall = cb.time_parallel("damped_pendulum", "rollouts", shards, reps=3)
    return {"value": rate, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": "%d processes x %d instances x %d steps x 3 of the same workload; oracle/_ref (unmodified reference C, "
                      "gcc -O2) in the tight C loop of oracle/ref_harness.c; wall %.1f s" % (procs, per, NSTEPS, wall)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="trepb")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return

    dist, rank, local, world = dist_setup(args.gpus)
    from trep_b200 import lib, systems
    if lib.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    device = local
    d = systems.named_desc("damped_pendulum")
    s = lib.System(d, device=device)
    assert s.specialized, "the damped-pendulum specialisation must be compiled in"
    th0, th1 = workload_inputs(BATCH, seed=rank)   # each rank its own shard of the global batch

    # ---- device-resident arm --------------------------------------------------------------------
    up = lambda a: lib.DeviceBuffer(device, a.shape, a.dtype).upload(a)
    dq0, dq1 = up(th0), up(th1)
    dp = lib.DeviceBuffer(device, (BATCH, 1))
    s.calc_p2_raw(True, BATCH, DT, dq0, dq1, dp)     # initialize_from_configs
    q2 = lib.DeviceBuffer(device, (BATCH, 1)); p2 = lib.DeviceBuffer(device, (BATCH, 1))
    it = lib.DeviceBuffer(device, (BATCH,), np.int32); st = lib.DeviceBuffer(device, (BATCH,), np.int32)
    flush = lib.DeviceBuffer(device, (256 << 20,), np.uint8)   # > 126 MB L2

    def one_step():
        flush.zero()
        s.step_raw(True, BATCH, NSTEPS, DT, DT, dq1, dp, None, None, None, None, q2, p2, None, it, st)
        lib.synchronize(device)
        return s.last_kernel_ms()

    def barrier():
        if dist is not None:
            dist.barrier()
        lib.synchronize(device)

    def reduce_max(x):
        if dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    for _ in range(args.warmup):
        one_step()
    fp64_peak = lib.measure_fp64_peak(device)
    sampler = ClockSampler(device)
    sampler.start()
    barrier()
    times = [one_step() for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop()
    total_ms = float(np.sum(times))
    iters_mean = float(it.download().mean()) / NSTEPS
    ok = bool(np.all(st.download() == 0))

    # ---- end-to-end arm: host-pointer C-ABI call, pinned buffers, copies inside the timed region --
    hq = lib.pinned_empty((BATCH, 1)); hp = lib.pinned_empty((BATCH, 1))
    hq[:] = th1; hp[:] = dp.download()
    hq2 = lib.pinned_empty((BATCH, 1)); hp2 = lib.pinned_empty((BATCH, 1))
    hit = lib.pinned_empty((BATCH,), np.int32); hst = lib.pinned_empty((BATCH,), np.int32)

    def one_e2e():
        t0 = time.perf_counter()
        s.step_raw(False, BATCH, NSTEPS, DT, DT, hq, hp, None, None, None, None, hq2, hp2, None, hit, hst)
        return (time.perf_counter() - t0) * 1e3

    for _ in range(2):
        one_e2e()
    barrier()
    e2e_times = [one_e2e() for _ in range(args.steps)]
    barrier()
    e2e_ms = float(np.sum(e2e_times))
    assert np.array_equal(hq2, q2.download()), "host and device entry points must agree bit for bit"

    # ---- max over ranks ---------------------------------------------------------------------------
    if dist is not None:
        import torch
        t = torch.tensor([total_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    units = float(BATCH) * NSTEPS * args.steps * world
    value = units / (total_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)
    multi = None
    if world > 1 and not args.no_secondary:
        # data plane of the library itself (NCCL / peer-mapped slabs through the C ABI); torch.distributed only
        # hands round the 128-byte id and the IPC handles
        from trep_b200 import dist as D_
        for b in (dq0, dq1, dp, q2, p2, it, st, flush):
            b.free()
        group = D_.Group(device=device, exchange=D_.TorchExchange(dist))
        hbm_peak_ = 6525.9
        mp_ = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp_):
            hbm_peak_ = json.load(open(mp_)).get("hbm_gbs", hbm_peak_)
        multi = multi_gpu(lib, systems, group, device, rank, world, reduce_max, fp64_peak, hbm_peak_)

    if rank == 0:
        flops_per_step = None
        fj = os.path.join(ROOT, "profiles", "flops.json")
        if os.path.exists(fj):
            flops_per_step = json.load(open(fj)).get("damped_pendulum_step_flops_per_del_step")
        kms = total_ms / args.steps
        roof = {"bound": "fp64", "peak": fp64_peak, "unit": "TFLOP/s", "traffic": BATCH * 40,
                "peak_source": "in-run DFMA micro-benchmark (trepb_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry",
                "traffic_note": "algorithmic HBM bytes per launch (16 B in + 24 B out per rollout of 1000 steps); ncu "
                                "dram__bytes of a 50-step launch: 16.8 MB read, <1 KB written back before the kernel ends "
                                "(profiles/r01c_step_raw.txt) - the kernel is on-chip",
                "kernel": "step_kernel<damped_pendulum>", "kernel_ms": kms,
                "flops_kind": "executed fp64 operations per DEL step (fma = 2, add, mul; libdevice-free sincos included) counted by "
                              "ncu's sass op counters on this kernel (profiles/flops.json), not an operation count of the "
                              "reference's algorithm: frac is the FP64 pipe's arithmetic utilisation, an upper bound on "
                              "algorithmic efficiency"}
        oc = opcounted("damped_pendulum_step")
        if oc:
            roof["flops_opcounted"] = oc
        if flops_per_step:
            ach = flops_per_step * BATCH * NSTEPS / (kms * 1e-3) / 1e12
            roof.update(achieved=ach, frac=ach / fp64_peak, flops_per_unit=flops_per_step)
            # beside the measured denominator: the nominal one (64 DFMA per clock per SM at the maximum SM clock)
            if clocks.get("sm_max_mhz"):
                nominal = 148 * 64 * 2 * clocks["sm_max_mhz"] * 1e6 / 1e12
                roof.update(peak_nominal=nominal, frac_of_nominal=ach / nominal)
        else:
            roof.update(achieved=None, frac=None)
        hbm_peak = 6525.9
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            hbm_peak = json.load(open(mp)).get("hbm_gbs", hbm_peak)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "nsteps": NSTEPS, "dt": DT,
                       "l2": "flushed between timed iterations (256 MB memset)", "kernel": s.kernel_name,
                       "newton_iters_per_step": iters_mean, "all_converged": ok,
                       "sharding": "independent instances block-partitioned over ranks, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * 16,
                    "d2h_bytes_per_step": BATCH * 24, "ms_per_step": e2e_ms / args.steps,
                    "api": "trepb_step_batch (host pointers, pinned)"},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
            "roofline": roof,
        }
        if world == 1 and not args.no_secondary:
            line["secondary"] = secondary(lib, systems, device, fp64_peak, hbm_peak)
            line["secondary"].append(w4_full(lib, systems, device, 1, float, fp64_peak))
        if multi is not None:
            line["secondary"] = multi
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample()
            if "secondary" in line:
                w1_cpu_reference(line["secondary"])
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
