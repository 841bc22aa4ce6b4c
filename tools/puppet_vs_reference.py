"""Marionette linearizations of the default kernel against the reference run live on a larger sample than the tests
use (development aid): perturbed points of the recorded rollout; the reference runs in `procs` processes.
    B=4096 python tools/puppet_vs_reference.py"""
import multiprocessing as mp, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import golden_util as G

def worker(job):
    import ref_systems as R
    system, mvi = R.make_mvi("puppet")
    t1, t2, q1, p1, k2, lam = job
    return R.run_cases(mvi, t1, t2, q1, p1, np.zeros((len(t1), 0)), k2, lambda_guess=lam)

if __name__ == "__main__":
    B = int(os.environ.get("B", "4096"))
    g = G.golden("puppet"); d = G.desc("puppet")
    rng = np.random.default_rng(77)
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    k2 = g["roll_k2"][idx]; lam = g["roll_lambda"][idx - 1]
    t1 = 0.01 * (idx + 1); t2 = t1 + 0.01
    procs = len(os.sched_getaffinity(0))
    cuts = np.linspace(0, B, procs + 1).astype(int)
    jobs = [(t1[a:b], t2[a:b], q1[a:b], p1[a:b], k2[a:b], lam[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        res = pool.map(worker, jobs)
    want = {k: np.concatenate([r[k] for r in res]) for k in ("status", "iters", "q2", "p2", "lambda1", "A", "B")}
    from trep_b200 import lib
    s = lib.System(d)
    out = s.linearize(q1, p1, None, k2, t1=t1, t2=t2, lambda_guess=lam)
    assert np.array_equal(out["status"], want["status"])
    ok = want["status"] == 0
    worst = {}
    for k in ("q2", "p2", "lambda1", "A", "B"):
        a, b = out[k][ok], want[k][ok]
        scale = np.max(np.abs(b), axis=tuple(range(1, b.ndim)), keepdims=True)
        worst[k] = float(np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * scale)))
        G.assert_close(a, b, "puppet %s" % k)
    print("%s: %d linearizations vs the reference live: ok %.4f, iteration counts that differ %d, worst element-wise relative "
          "difference (|a-b| / (|b| + 1e-3 max|b|)) %s" % (s.kernel_name, B, ok.mean(), int(np.sum(out["iters"][ok] != want["iters"][ok])),
                                                          {k: "%.1e" % v for k, v in worst.items()}))
