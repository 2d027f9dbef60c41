"""Time det(A) mod p for a 4096 x 4096 matrix over a few prime-group sizes (one GPU)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from linalg_solver_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = Engine(0)
rng = np.random.Generator(np.random.PCG64(20260005))
A = torch.from_numpy(rng.integers(-5, 6, size=(n, n), dtype=np.int32)).cuda()
for G in [int(x) for x in (sys.argv[2:] or ["8", "37", "148"])]:
    eng.det_large_residues(A, 0, min(G, 4)); torch.cuda.synchronize()
    eng.timing_enable(True)
    t0 = time.perf_counter()
    r = eng.det_large_residues(A, 0, G)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gemm_ms = sum(eng.timing_read())
    eng.timing_enable(False)
    macs = G * n**3 / 3
    print(json.dumps({"n": n, "primes": G, "seconds": dt, "gemm_seconds": gemm_ms / 1e3, "mac_per_s": macs / dt,
                      "res0": int(r[0].item()) & 0xffffffff}))
