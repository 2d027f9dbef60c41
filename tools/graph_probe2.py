"""Probe: direct device-memory calls against one CUDA graph replay per step for configs 3 and 4 (device events)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from linalg_solver_b200 import Engine
eng = Engine(0)
rng = np.random.Generator(np.random.PCG64(1))
def time_both(name, step, units, reps=10):
    step(); torch.cuda.synchronize()
    cap = eng.capture(step)
    for mode, fn in (("calls", step), ("graph", cap.replay), ("calls", step), ("graph", cap.replay)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"cfg": name, "mode": mode, "ms_per_step": ms, "per_s": units / (ms * 1e-3), "kernels": cap.kernels}))
B = 1 << 18
Bm = rng.integers(-5, 6, size=(B, 16, 10), dtype=np.int64); Cm = rng.integers(-5, 6, size=(B, 10, 16), dtype=np.int64)
A = np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)
x0 = rng.integers(-5, 6, size=(B, 16), dtype=np.int64)
b = np.einsum("bij,bj->bi", A.astype(np.int64), x0).astype(np.int32)
b[1::2] = rng.integers(-5, 6, size=(B // 2, 16), dtype=np.int32)
At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
plan = eng.plan_solve(16, 16, 250, int(np.abs(b).max()), 10, 6)
r3 = eng.solve_batch(At, bt, plan=plan)
time_both("c3 2^18 16x17 solve", lambda: eng.solve_batch(At, bt, plan=plan, out=r3), B)
del At, bt, r3
B = 1 << 14
A4 = torch.from_numpy(rng.integers(-5, 6, size=(B, 64, 64), dtype=np.int32)).cuda()
p4 = eng.plan_inverse(64, 5)
r4 = eng.inverse_batch(A4, plan=p4)
time_both("c4inv 2^14 64x64 inverse", lambda: eng.inverse_batch(A4, plan=p4, out=r4), B, reps=3)
