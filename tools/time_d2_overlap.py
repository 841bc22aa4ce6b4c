"""Experiment: two second-derivative batches on two handles and two streams at once vs one after the other
(how much of pass A / pass B could overlap if the library pipelined its chunks)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trep_b200 import lib, systems
import torch
up = lambda a: lib.DeviceBuffer(0, a.shape, a.dtype).upload(a)
rng = np.random.default_rng(0)
d = systems.named_desc("puppet")
g = np.load(os.path.join(ROOT, "tests", "golden", "puppet.npz"))
B = int(os.environ.get("PUPPET_B", "1024"))
def setup():
    s = lib.System(d)
    idx = rng.integers(1, 58, B)
    q1 = g["roll_q"][idx].copy(); p1 = g["roll_p"][idx].copy()
    q1[:, :d.nd] += rng.normal(0, 0.02, (B, d.nd)); p1 += rng.normal(0, 0.02, (B, d.nd))
    bufs = dict(q=up(q1), p=up(p1), k=up(g["roll_k2"][idx]), l=up(g["roll_lambda"][idx - 1]), st=lib.DeviceBuffer(0, (B,), np.int32))
    nX, nU = d.nX, d.nU
    bufs["z"] = up(rng.normal(0, 1, (B, nX)))
    bufs["xx"] = lib.DeviceBuffer(0, (B, nX, nX)); bufs["xu"] = lib.DeviceBuffer(0, (B, nX, nU)); bufs["uu"] = lib.DeviceBuffer(0, (B, nU, nU))
    return s, bufs
def run(s, b, stream):
    s.deriv2_raw(True, B, b["q"], b["p"], None, b["k"], b["st"], {}, stream=stream, z=b["z"], fdxdx=b["xx"], fdxdu=b["xu"], fdudu=b["uu"],
                 t1_scalar=0.0, dt_scalar=0.01, lambda_guess=b["l"])
A, Bb = setup(), setup()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(2):
    run(A[0], A[1], None); run(Bb[0], Bb[1], None); torch.cuda.synchronize()
t0 = time.perf_counter(); run(A[0], A[1], None); run(Bb[0], Bb[1], None); torch.cuda.synchronize(); seq = time.perf_counter() - t0
t0 = time.perf_counter(); run(A[0], A[1], C.c_void_p(s1.cuda_stream)); run(Bb[0], Bb[1], C.c_void_p(s2.cuda_stream)); torch.cuda.synchronize(); par = time.perf_counter() - t0
print("B=%d x 2 batches: sequential %.2f ms, two streams %.2f ms (%.2fx)" % (B, seq * 1e3, par * 1e3, seq / par))
