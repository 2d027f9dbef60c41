"""Probe: are the device-memory calls of liblsx capturable into a CUDA graph?  (config 1 step: det + rank + rref of 10k 4x4)"""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from linalg_solver_b200 import Engine
eng = Engine(0)
rng = np.random.Generator(np.random.PCG64(1))
A = torch.from_numpy(rng.integers(-5, 6, size=(10000, 4, 4), dtype=np.int32)).cuda()
pd, pr, pf = eng.plan_det(4, 5), eng.plan_rank(4, 4, 5), eng.plan_rref(4, 4, 3, 5, 5)
res = (eng.det_batch(A, plan=pd), eng.rank_batch(A, plan=pr), eng.rref_batch(A, 3, plan=pf))
def step():
    eng.det_batch(A, plan=pd, out=res[0]); eng.rank_batch(A, plan=pr, out=res[1]); eng.rref_batch(A, 3, plan=pf, out=res[2])
step(); torch.cuda.synchronize()
import dataclasses
def tensors(r):
    return [getattr(r, f.name) for f in dataclasses.fields(r) if hasattr(getattr(r, f.name), "clone")]
want = [t.clone() for r in res for t in tensors(r)]
s = torch.cuda.Stream()
eng.set_stream(s.cuda_stream)
with torch.cuda.stream(s):
    step()
s.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    step()
for r in res:
    for t in tensors(r): t.zero_()
g.replay(); torch.cuda.synchronize()
got = [t for r in res for t in tensors(r)]
print("equal after replay:", all(torch.equal(a, b) for a, b in zip(want, got)), len(got))
for name, fn in (("graph", g.replay), ("calls", step)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
    print(json.dumps({"mode": name, "ms_per_step": dt * 1e3, "mat_per_s": 10000 / dt}))
