"""tests/golden/c5_full_residues: det(A) mod p of the config 5 matrix (bench.c5_matrix(): 4096 x 4096, PCG64(20260005))
for two table primes beyond the CRT set, by oracle/det_mod_p.py on the CPU (about 11 minutes per prime on one core)."""
import json
import os
import sys
import time
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import golden_io  # noqa: E402
from oracle.det_mod_p import det_mod_p  # noqa: E402
from tests.device_model import prime_table  # noqa: E402

INDICES = [1500, 1501]
TABLE = prime_table(max(INDICES) + 1)


def one(i):
    t = time.time()
    return {"prime_index": i, "prime": int(TABLE[i]), "residue": int(det_mod_p(bench.c5_matrix(), TABLE[i])),
            "seconds": round(time.time() - t, 1)}


if __name__ == "__main__":
    with Pool(len(INDICES)) as pool:
        entries = pool.map(one, INDICES)
    golden_io.save("c5_full_residues", {
        "about": "det mod p of bench.c5_matrix() by oracle/det_mod_p.py (numpy int64 modular elimination) for table primes "
                 "outside the CRT set of config 5", "n": bench.C5_N, "seed": bench.C5_SEED, "entries": entries})
    print(json.dumps(entries))
