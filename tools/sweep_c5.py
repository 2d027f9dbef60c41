"""Whole 4096 x 4096 determinant residues (all primes of the Hadamard bound) for the LSX_LARGE_STREAMS / LSX_LARGE_GROUP
values given in the environment; prints one JSON line with the device time and a checksum of the residues."""
import sys, time, json, os, zlib
import numpy as np, torch
sys.path.insert(0, ".")
from linalg_solver_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1013
eng = Engine(0)
rng = np.random.Generator(np.random.PCG64(20260005))
A = torch.from_numpy(rng.integers(-5, 6, size=(n, n), dtype=np.int32)).cuda()
eng.det_large_residues(A, 0, K); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
r = eng.det_large_residues(A, 0, K)
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(json.dumps({"n": n, "primes": K, "streams": os.environ.get("LSX_LARGE_STREAMS", "default"),
                  "group": os.environ.get("LSX_LARGE_GROUP", "default"), "tag": os.environ.get("LSX_TAG", ""), "device_s": e0.elapsed_time(e1) / 1e3,
                  "host_enqueue_s": t_host, "crc": zlib.crc32(r.cpu().numpy().tobytes())}))
