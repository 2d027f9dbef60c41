"""Device-resident timing of the BASELINE.json configs 1, 3, 4 (one GPU), kernels only."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from linalg_solver_b200 import Engine
eng = Engine(0)
which = sys.argv[1:] or ["c1", "c3", "c4inv", "c4ker"]

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    eng.timing_enable(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    k = eng.timing_read(); eng.timing_enable(False)
    return dt, sum(k) / reps / 1e3

rng = np.random.Generator(np.random.PCG64(1))
if "c1" in which:
    A = torch.from_numpy(rng.integers(-5, 6, size=(10000, 4, 4), dtype=np.int32)).cuda()
    pd, pr, pf = eng.plan_det(4, 5), eng.plan_rank(4, 4, 5), eng.plan_rref(4, 4, 3, 5, 5)
    def f():
        eng.det_batch(A, plan=pd); eng.rank_batch(A, plan=pr); eng.rref_batch(A, 3, plan=pf)
    dt, k = timeit(f, 10)
    print(json.dumps({"cfg": "c1 10k 4x4 det+rank+rref", "s": dt, "kern_s": k, "mat_per_s": 10000 / dt}))
if "c3" in which:
    B = 1 << 18
    Bm = rng.integers(-5, 6, size=(B, 16, 10), dtype=np.int64); Cm = rng.integers(-5, 6, size=(B, 10, 16), dtype=np.int64)
    A = np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)
    x0 = rng.integers(-5, 6, size=(B, 16), dtype=np.int64)
    b = np.einsum("bij,bj->bi", A.astype(np.int64), x0).astype(np.int32)
    b[1::2] = rng.integers(-5, 6, size=(B // 2, 16), dtype=np.int32)
    At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    plan = eng.plan_solve(16, 16, 250, int(np.abs(b).max()), 10, 6)
    print("c3 plan", plan)
    dt, k = timeit(lambda: eng.solve_batch(At, bt, plan=plan), 3)
    print(json.dumps({"cfg": "c3 256k 16x17 solve", "s": dt, "kern_s": k, "mat_per_s": B / dt}))
if "c4inv" in which:
    B = 1 << 12
    A = torch.from_numpy(rng.integers(-5, 6, size=(B, 64, 64), dtype=np.int32)).cuda()
    plan = eng.plan_inverse(64, 5)
    print("c4 plan", plan)
    dt, k = timeit(lambda: eng.inverse_batch(A, plan=plan), 2)
    print(json.dumps({"cfg": "c4 4096 of 64k 64x64 inverse", "s": dt, "kern_s": k, "mat_per_s": B / dt}))
if "c4ker" in which:
    B = 1 << 12
    Bm = rng.integers(-5, 6, size=(B, 64, 48), dtype=np.int64); Cm = rng.integers(-5, 6, size=(B, 48, 64), dtype=np.int64)
    A = torch.from_numpy(np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32)).cuda()
    z = torch.zeros((B, 64), dtype=torch.int32).cuda()
    plan = eng.plan_solve(64, 64, 1200, 0, 48, 16)
    print("c4ker plan", plan)
    dt, k = timeit(lambda: eng.solve_batch(A, z, plan=plan), 2)
    print(json.dumps({"cfg": "c4 4096 of 64k 64x64 kernel", "s": dt, "kern_s": k, "mat_per_s": B / dt}))
