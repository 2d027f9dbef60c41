#!/usr/bin/env python
"""Benchmark of the exact elimination hot path (BASELINE.json metric: exact det/RREF matrices/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4inv|c4ker|c5]
                    [--extra c1,c3,c4inv,c4ker,c5 | none]

The ONE JSON line of a default run carries the headline workload (c2) at the top level and a `workloads` object
with a full sub-record (value, ms_per_step, roofline, e2e, clocks, gpu_launches) for every other BASELINE.json
config: c1, c3, c4inv and c4ker at their full 2^16, and c5 (seconds per 4096 x 4096 determinant with the primes
sharded over the run's N ranks and one all-gather of residues), so the driver's N = 1/2/4/8 runs evidence the
by-prime path too.  `--workload X` alone prints only X; `--extra` picks the sub-records.

A step is one pass of the hot path over one batch of synthetic matrices.  The default workload is
BASELINE.json configs[1]: 2^20 random 8x8 integer matrices (entries uniform in [-5,5], the
distribution of RandomMatrixBuilder.build_random, reference random_matrix.py:103-107), determinant +
inverse via [A|I] row_reduce(bar_col=8) -- per GPU, so N GPUs process N * 2^20 matrices per step
(weak scaling, matrices are independent: no data-path collective).

`value` is matrices/s with inputs resident in HBM (CUDA events on the launching stream, max over
ranks); `e2e` is the same metric through the C-ABI with HOST buffers (pinned), host<->device copies
inside the timed region; `roofline` is for the dominant kernel (device-timed inside the library);
`cpu_baseline` is the CPU oracle (port of the reference's algorithm) on this box's host cores.
`--impl reference` times that CPU port alone on all host cores, same config/metric.

`--workload c5` is BASELINE.json configs[4]: ONE 4096 x 4096 integer determinant, multi-modular, the primes
sharded over the ranks (strong scaling), one all-gather of the residues before the CRT; its metric is
seconds per determinant and its roofline is the tensor pipe (tcgen05 int8-split trailing update).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, n, batch per GPU, algorithmic bytes per matrix)
    "c1": ("10k x 4x4 determinant + rank + row_reduce (default bar_col=3), entries uniform [-5,5]", 4, 10000, 152),
    "c2": ("2^20 x 8x8 det + inverse via [A|I] row_reduce(bar_col=8), entries uniform [-5,5]", 8, 1 << 20, 556),
    "c3": ("2^18 x find_preimage_of on 16x17 [A|b], A = B(16x10) C(10x16) rank 10, entries of B, C uniform [-5,5]",
           16, 1 << 18, 2964),
    "c4inv": ("2^16 x 64x64 inverse via [A|I] row_reduce(bar_col=64), entries uniform [-5,5]", 64, 1 << 16, 196912),
    "c4ker": ("2^16 x kernel basis of 64x64 A = B(64x48) C(48x64) (rank 48, dim 16), entries of B, C uniform [-5,5]",
              64, 1 << 16, 98644),
}
EXTRA_DEFAULT = ["c1", "c3", "c4inv", "c4ker", "c5"]
# end-to-end (host buffer) calls per step: the 64x64 workloads stream their 2^16 matrices through the C-ABI in 8 calls
# of 2^13 that reuse one set of pinned result buffers (11.8 GB of adjugates per step would otherwise be pinned per rank)
E2E_CALLS = {"c4inv": 8, "c4ker": 8}
# ALGORITHMIC modular multiply-subtracts per matrix and prime (Gauss-Jordan on m x n with pivots
# in columns c_k costs sum_k m * (n - c_k)); the integer-pipe roofline is the measured mont_mul rate
# (profiles/r01_ubench_int.jsonl: 12.0 per SM per clock)
ALG_OPS = {"c2": 8 * sum(16 - k for k in range(8)), "c3": 16 * sum(17 - k for k in range(10)),
           "c4inv": 64 * sum(128 - k for k in range(64)), "c4ker": 64 * sum(65 - k for k in range(48))}
MONT_MUL_PER_SM_CLK = 12.0
SEED = 20260002
METRIC = {"c1": "exact det+rank+RREF matrices/sec", "c2": "exact det+inverse matrices/sec",
          "c3": "exact find_preimage_of systems/sec", "c4inv": "exact inverse matrices/sec",
          "c4ker": "exact kernel bases/sec"}


def make_inputs(n, batch, seed, workload="c2", device=None):
    """-> dict of int32 arrays for one rank's batch.  With `device` (a torch CUDA device) the large rank-deficient
    products are formed there (float32 batched products of small integers are exact) and returned as device tensors;
    the factors always come from the same numpy generator, so both routes give the same matrices."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))

    def product(Bm, Cm):
        if device is None or Bm.shape[0] <= 4096:
            return np.einsum("bik,bkj->bij", Bm.astype(np.int64), Cm.astype(np.int64)).astype(np.int32)
        import torch
        out = torch.bmm(torch.from_numpy(Bm).to(device, torch.float32), torch.from_numpy(Cm).to(device, torch.float32))
        return out.round().to(torch.int32)                        # |entries| <= 48 * 25 < 2^24: exact in float32

    if workload == "c3":
        Bm = rng.integers(-5, 6, size=(batch, 16, 10), dtype=np.int64)
        Cm = rng.integers(-5, 6, size=(batch, 10, 16), dtype=np.int64)
        A = np.einsum("bik,bkj->bij", Bm, Cm)
        x0 = rng.integers(-5, 6, size=(batch, 16), dtype=np.int64)
        b = np.einsum("bij,bj->bi", A, x0)
        b[1::2] = rng.integers(-5, 6, size=b[1::2].shape)          # odd systems: random rhs (inconsistent w.h.p.)
        return {"A": A.astype(np.int32), "b": b.astype(np.int32)}
    if workload == "c4ker":
        Bm = rng.integers(-5, 6, size=(batch, 64, 48), dtype=np.int64)
        Cm = rng.integers(-5, 6, size=(batch, 48, 64), dtype=np.int64)
        return {"A": product(Bm, Cm), "b": np.zeros((batch, 64), dtype=np.int32)}
    return {"A": rng.integers(-5, 6, size=(batch, n, n), dtype=np.int32)}


def plan_of(workload):
    """The lsx plan of a by-matrix workload (host-only query: prime and limb counts from the Hadamard bound)."""
    import ctypes
    from linalg_solver_b200 import _lib
    p = _lib.Plan()
    if workload == "c1":
        rc = _lib.lib.lsx_plan_rref(4, 4, 3, 5, 5, 0, ctypes.byref(p))
    elif workload == "c3":
        rc = _lib.lib.lsx_plan_solve(16, 16, 250, 250 * 16 * 5, 10, 6, ctypes.byref(p))
    elif workload == "c4ker":
        rc = _lib.lib.lsx_plan_solve(64, 64, 48 * 25, 0, 48, 16, ctypes.byref(p))
    else:
        rc = _lib.lib.lsx_plan_inverse(WORKLOADS[workload][1], 5, ctypes.byref(p))
    assert rc == 0, rc
    return p


def config_for(workload, world):
    """`config` of a line: the same dict in our arm and in the reference arm."""
    if workload == "c5":
        from linalg_solver_b200 import dist as lsx_dist
        import ctypes
        from linalg_solver_b200 import _lib
        k, bits = ctypes.c_int(), ctypes.c_double()
        assert _lib.lib.lsx_det_large_prime_count(C5_N, C5_ABS, ctypes.byref(k), ctypes.byref(bits)) == 0
        n_primes = c5_prime_count_numpy()
        return {"workload": C5_DESC, "primes": n_primes, "primes_worst_case_bound": k.value,
                "bound": "Hadamard with the actual row/column norms of the matrix",
                "primes_per_gpu": lsx_dist.shard_sizes(n_primes, world)[0],
                "sharding": "by prime, one all-gather of %d residues (%d B) before the CRT" % (n_primes, 4 * n_primes),
                "all_gather_bytes": 4 * n_primes,
                "l2": "residue matrices of one prime group (64 MiB per prime) far exceed the 126 MB L2"}
    desc, n, batch, alg_bytes = WORKLOADS[workload]
    plan = plan_of(workload)
    return {"workload": desc, "batch_per_gpu": batch, "primes": int(plan.n_primes), "limbs": int(plan.limbs),
            "sharding": "by matrix, no collective",
            "l2": "inputs+outputs per step (%.0f MB) vs the 126 MB L2" % ((alg_bytes * batch) / 1e6)}


# ------------------------------------------------------------------------------------- CPU arm
def _cpu_one(item):
    """One unit of the workload through the oracle port (same calls the reference would make)."""
    from oracle import ref_port
    wl, a, b = item
    if wl == "c1":
        d = ref_port.determinant(a)
        return d.numerator, ref_port.rank(a), ref_port.row_reduce(a)[1]
    if wl == "c3":
        res = ref_port.find_preimage_of(a, b)
        return None if res is None else res[0][0].numerator
    if wl == "c4ker":
        res = ref_port.kernel(a)
        return len(res[1]) if res and res[1] else 0
    inv = ref_port.inverse(a)
    det = ref_port.determinant(a) if wl == "c2" else None
    return (det.numerator if det is not None else None), (None if inv is None else inv[0][0].numerator)


CPU_PER_CORE = {"c1": 4096, "c2": 2048, "c3": 64, "c4inv": 1, "c4ker": 1}


def cpu_baseline(workload, per_core, seed):
    """The oracle port (oracle/ref_port.py: Fraction Gauss-Jordan of reference linalg.py:534-630 and the
    calls on it) on a bounded sample of the workload, all host cores."""
    from multiprocessing import get_context
    desc, n, _, _ = WORKLOADS[workload]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    data = make_inputs(n, cores * per_core, seed, workload)
    A = data["A"].tolist()
    b = data["b"].tolist() if "b" in data else [None] * len(A)
    sample = [(workload, A[i], b[i]) for i in range(len(A))]
    ctx = get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_one, sample[:cores], chunksize=1)          # start the workers
        t0 = time.perf_counter()
        out = pool.map(_cpu_one, sample, chunksize=max(1, per_core // 8))
        dt = time.perf_counter() - t0
    return {"value": len(sample) / dt, "unit": "matrices/s", "cores": cores, "kind": "port",
            "sample": "%d units of the workload (%s; seed %d) through oracle/ref_port.py, multiprocessing over "
                      "%d cores, %.1f s" % (len(sample), workload, seed, cores, dt)}, out


def reference_line(workload, steps, warmup, gpus):
    """One workload of the reference arm: the CPU implementation of the path on a bounded sample per step."""
    desc, n, batch, _ = WORKLOADS[workload]
    per_core = CPU_PER_CORE[workload]
    times = []
    cb = None
    for i in range(warmup + steps):
        cb, _ = cpu_baseline(workload, per_core, SEED + i)
        if i >= warmup:
            times.append(cb["value"])
    val = statistics.mean(times)
    cb["value"] = val
    with_reference(cb, workload)
    return {
        "impl": "reference", "metric": METRIC[workload], "value": val, "unit": "matrices/s",
        "n_gpus": gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * per_core * cb["cores"] / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "exact rationals (fractions.Fraction)", "data": "synthetic",
        "config": config_for(workload, gpus), "sample_per_step": "bounded sample of the workload on the host CPU",
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    main_wl = args.workload
    line = run_c5_reference_line(args.steps, args.warmup, args.gpus) if main_wl == "c5" else \
        reference_line(main_wl, args.steps, args.warmup, args.gpus)
    extras = [w for w in args.extra_list if w != main_wl]
    if extras:
        # one bounded pass each: the sub-records only have to place the CPU side of every config in the same run
        line["workloads"] = {w: (run_c5_reference_line(1, 0, args.gpus) if w == "c5" else reference_line(w, 1, 0, args.gpus))
                             for w in extras}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- config 5
C5_N, C5_SEED, C5_ABS = 4096, 20260005, 5
C5_DESC = "single 4096x4096 integer determinant (entries uniform [-5,5], PCG64(20260005)), multi-modular, sharded by prime"
C5_METRIC = "seconds per exact 4096x4096 determinant"


def c5_matrix(n=C5_N):
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(C5_SEED))
    return rng.integers(-C5_ABS, C5_ABS + 1, size=(n, n), dtype=np.int32)


def _c5_cpu_one(item):
    from oracle.det_mod_p import det_mod_p
    A, p = item
    return det_mod_p(A, p)


def c5_cpu_baseline(n_primes, sample_n=1280):
    """oracle/det_mod_p.py (numpy int64 Gaussian elimination modulo p, the forward sweep of reference
    linalg.py:547-609) on a bounded sample: the leading sample_n x sample_n block, one prime per host core in
    parallel; scaled to the full job by (n / sample_n)^3 per prime and n_primes / cores."""
    from multiprocessing import get_context
    from tests.device_model import prime_table
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    A = c5_matrix()[:sample_n, :sample_n].copy()
    primes = prime_table(cores)
    ctx = get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_c5_cpu_one, [(A[:64, :64], p) for p in primes], chunksize=1)     # start the workers
        t0 = time.perf_counter()
        pool.map(_c5_cpu_one, [(A, p) for p in primes], chunksize=1)
        dt = time.perf_counter() - t0
    per_prime_core_s = dt * (C5_N / sample_n) ** 3               # one prime at full size on one core
    seconds = per_prime_core_s * n_primes / cores
    return {"value": seconds, "unit": "s", "cores": cores, "kind": "port",
            "sample": "det mod p of the leading %dx%d block for %d primes in parallel (oracle/det_mod_p.py, numpy int64 "
                      "elimination, %.1f s), scaled by (4096/%d)^3 per prime and %d primes / %d cores (extrapolated)"
                      % (sample_n, sample_n, cores, dt, sample_n, n_primes, cores)}


def c5_prime_count_numpy():
    """Prime count of lsx_det_large_prime_count_for, recomputed on the host (same formula)."""
    import math
    import numpy as np
    A = c5_matrix().astype(np.int64)
    rows = 0.5 * np.log2((A * A).sum(axis=1).astype(np.float64)).sum()
    cols = 0.5 * np.log2((A * A).sum(axis=0).astype(np.float64)).sum()
    bits = min(rows, cols) * (1.0 + 1e-9) + 1e-6
    return max(1, math.ceil((bits + 1.0 + 1e-6) / 30.999))


def c5_tensor_ops(n, n_primes_local, block=256):
    """int8 tensor operations (2 per multiply-add, 16 byte-plane products per residue multiply-add) of the
    depth-256 trailing updates of the blocked LU for n_primes_local primes."""
    macs = 0
    k0 = 0
    while k0 + block < n:
        rest = n - k0 - block
        macs += rest * rest * block
        k0 += block
    return 2 * 16 * macs * n_primes_local


# ------------------------------------------------------------------------------------- GPU arm
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def load_counts():
    """Per-launch counters of the dominant kernels taken from ncu captures of this build (profiles/kernel_counts.json:
    smsp__inst_executed.sum and dram__bytes_read.sum + dram__bytes_write.sum per launch, with the capture they come
    from).  They are properties of the binary and the workload, not of the run; the durations they are divided by
    are measured live."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_counts.json")))
    except Exception:
        return {}


class Ctx:
    """One rank of the run: device, engine, process group."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from linalg_solver_b200 import Engine
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.eng = Engine(self.local)
        self.peaks = load_peaks()
        self.counts = load_counts()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local])
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def close(self):
        self.eng.close()
        if self.world > 1:
            self.dist.destroy_process_group()


class Job:
    """One workload bound to an engine: device-resident step, host-buffer (end-to-end) step, oracle check."""

    def __init__(self, ctx, workload):
        import numpy as np
        torch = ctx.torch
        eng, dev = ctx.eng, ctx.dev
        self.eng, self.wl, self.np, self.torch = eng, workload, np, torch
        desc, n, batch, _ = WORKLOADS[workload]
        self.n, self.batch = n, batch
        # by-matrix sharding: every rank owns an independent batch (its own seed), no collective on the data path
        data = make_inputs(n, batch, SEED + 1000 * ctx.rank, workload, device=dev)
        self.dev = {k: (v if hasattr(v, "is_cuda") else torch.from_numpy(v).to(dev)) for k, v in data.items()}
        self.host = {k: v.cpu().pin_memory() for k, v in self.dev.items()}
        # declared magnitudes (rigorous for the generator above), the same plans config_for() reports
        if workload == "c1":
            self.plans = (eng.plan_det(4, 5), eng.plan_rank(4, 4, 5), plan_of("c1"))
        else:
            self.plans = (plan_of(workload),)
        self.res = self.run(self.dev)                         # allocates the outputs; reused by every step
        self.host_out = None

    def run(self, src, out=None):
        e, A = self.eng, src["A"]
        if self.wl == "c1":
            o = out or (None, None, None)
            return (e.det_batch(A, plan=self.plans[0], out=o[0]), e.rank_batch(A, plan=self.plans[1], out=o[1]),
                    e.rref_batch(A, 3, plan=self.plans[2], out=o[2]))
        if self.wl in ("c3", "c4ker"):
            return (e.solve_batch(A, src["b"], plan=self.plans[0], out=out[0] if out else None),)
        return (e.inverse_batch(A, plan=self.plans[0], out=out[0] if out else None),)

    def step_device(self):
        self.run(self.dev, self.res)

    def _fields(self, r):
        return [(k, v) for k, v in vars(r).items() if k != "plan" and v is not None]

    def make_host_outputs(self):
        torch = self.torch
        self.e2e_calls = E2E_CALLS.get(self.wl, 1)
        self.e2e_batch = self.batch // self.e2e_calls
        outs = []
        for r in self.res:
            kw = {k: torch.empty((self.e2e_batch,) + tuple(v.shape[1:]), dtype=torch.int32).pin_memory().numpy()
                  for k, v in self._fields(r)}
            outs.append(type(r)(plan=r.plan, **{k: kw.get(k) for k in vars(r) if k != "plan"}))
        self.host_out = tuple(outs)
        self.host_np = {k: v.numpy() for k, v in self.host.items()}
        if self.wl == "c2":
            # entries are in [-5, 5]: the end-to-end call ships them as int8 (lsx_inverse_batch_i8), a quarter
            # of the host-to-device bytes of the int32 container; the results are the same words
            self.host_i8 = self.host["A"].to(torch.int8).pin_memory()
            self.host_np = {"A": self.host_i8.numpy()}

    def step_e2e(self):
        """All `batch` matrices from pinned host memory through the C-ABI; returns when the results are in host memory."""
        eb = self.e2e_batch
        for c in range(self.e2e_calls):
            self.run({k: v[c * eb:(c + 1) * eb] for k, v in self.host_np.items()}, self.host_out)

    def step_copy_only(self):
        """The bytes of step_e2e as bare pinned-host <-> device copies on two streams (no kernels, no library): what
        the box's PCIe / host memory lets this rank move while the other ranks do the same."""
        torch = self.torch
        if not hasattr(self, "_cp"):
            dev = next(iter(self.dev.values())).device
            h_in = [torch.from_numpy(v) for v in self.host_np.values()]           # pinned (views of pinned tensors)
            d_in = [torch.empty(tuple(t.shape), dtype=t.dtype, device=dev) for t in h_in]
            pairs = [(torch.from_numpy(vh), vd) for rh, rd in zip(self.host_out, self.res)
                     for (_, vh), (_, vd) in zip(self._fields(rh), self._fields(rd))]
            self._cp = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), h_in, d_in, pairs)
        s_in, s_out, h_in, d_in, pairs = self._cp
        with torch.cuda.stream(s_in):
            for h, d in zip(h_in, d_in):
                d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_out):
            for _ in range(self.e2e_calls):
                for h, d in pairs:
                    h.copy_(d[: h.shape[0]], non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    def h2d_bytes(self):
        return int(sum(v.nbytes for v in self.host_np.values()))

    def d2h_bytes(self):
        return int(self.e2e_calls * sum(v.nbytes for r in self.host_out for _, v in self._fields(r)))

    def check_e2e_equals_device(self):
        np = self.np
        b0 = (self.e2e_calls - 1) * self.e2e_batch             # the host buffers hold the last call's slice
        for rh, rd in zip(self.host_out, self.res):
            for (k, vh), (_, vd) in zip(self._fields(rh), self._fields(rd)):
                assert np.array_equal(vh.reshape(-1)[:4096], vd[b0:].reshape(-1)[:4096].cpu().numpy()), k

    def check_against_oracle(self):
        """A few units of this very batch against the CPU oracle (untimed)."""
        from fractions import Fraction
        from linalg_solver_b200.convert import limbs_to_ints
        from oracle import ref_port
        k = 8 if self.n <= 16 else 1
        A = self.host["A"][:k].tolist()
        if self.wl == "c1":
            d, rk, rr = self.res
            dets, num, den = limbs_to_ints(d.det[:k]), limbs_to_ints(rr.num[:k]), limbs_to_ints(rr.den[:k])
            for i in range(k):
                R, piv = ref_port.row_reduce(A[i])
                assert dets[i] == ref_port.bareiss_det(A[i]) and int(rk.rank[i]) == ref_port.rank(A[i])
                assert [[Fraction(x, den[i]) for x in row] for row in num[i]] == R
        elif self.wl == "c4ker":
            (r,) = self.res
            den, gens = limbs_to_ints(r.den[:1])[0], limbs_to_ints(r.generators[:1])[0]
            want = ref_port.kernel(A[0])
            kdim = 64 - int(r.rank[0])
            assert int(r.status[0]) == 0 and kdim == len(want[1])                  # want[1]: kdim generators of length 64
            assert [[Fraction(gens[i][c], den) for i in range(64)] for c in range(kdim)] == want[1]
        elif self.wl == "c3":
            (r,) = self.res
            b = self.host["b"][:k].tolist()
            den, part = limbs_to_ints(r.den[:k]), limbs_to_ints(r.particular[:k])
            for i in range(k):
                want = ref_port.find_preimage_of(A[i], b[i])
                if want is None:
                    assert int(r.status[i]) & 2
                else:
                    assert int(r.status[i]) == 0 and [Fraction(x, den[i]) for x in part[i]] == want[0]
        else:
            (r,) = self.res
            adj, det = limbs_to_ints(r.adj[:k]), limbs_to_ints(r.det[:k])
            for i in range(k):
                want = ref_port.inverse(A[i])
                got = None if det[i] == 0 else [[Fraction(x, det[i]) for x in row] for row in adj[i]]
                assert got == want and det[i] == ref_port.bareiss_det(A[i]), "device result differs from the oracle"


def measure_batch(ctx, workload, steps, warmup, cpu):
    """One by-matrix workload at this run's N ranks -> its record (on rank 0; None elsewhere).  Every rank must call."""
    torch = ctx.torch
    eng, dev, world = ctx.eng, ctx.dev, ctx.world
    desc, n, batch, alg_bytes = WORKLOADS[workload]
    job = Job(ctx, workload)
    torch.cuda.synchronize(dev)
    if ctx.rank == 0:
        job.check_against_oracle()

    # config 1 is launch bound by construction (three calls on 10k 4x4 matrices: 38 us of kernels in a 79 us step):
    # its device-resident step is captured once into a CUDA graph (Engine.capture) and replayed; the kernels are the
    # same, the library's per-kernel event timing is off for it (kernel_ms = the step)
    captured = None
    if workload == "c1":
        try:
            captured = eng.capture(job.step_device)
        except Exception as e:                       # a failed capture must not cost the line: direct calls instead
            sys.stderr.write("bench: CUDA graph capture of the config 1 step failed (%s); timing direct calls\n" % e)
            torch.cuda.synchronize(dev)
    step_device = captured.replay if captured else job.step_device
    for _ in range(warmup):
        step_device()
    ctx.barrier()
    launches0 = eng.launch_count
    if not captured:
        eng.timing_enable(True)
    sampler = ClockSampler(ctx.local)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        step_device()
    ev1.record()
    ctx.barrier()
    t1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = eng.timing_read() if not captured else []
    eng.timing_enable(False)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(t0, t1)
    primes_run = eng.last_prime_count()      # 0: fused small kernels (no modular primes) or a single prime

    # ---- end to end through the C-ABI with host buffers (pinned): H2D + kernels + D2H per step ----
    job.make_host_outputs()
    e2e_steps = max(3, min(steps, 10)) if job.e2e_calls == 1 else 3
    for _ in range(2):
        job.step_e2e()
    ctx.barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        job.step_e2e()
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - e0) * 1e3 / e2e_steps
    job.check_e2e_equals_device()
    # ---- the same bytes as bare copies, all ranks at once: the transfer ceiling of the box at this N ----
    job.step_copy_only()
    ctx.barrier()
    c0 = time.perf_counter()
    for _ in range(e2e_steps):
        job.step_copy_only()
    copy_ms = (time.perf_counter() - c0) * 1e3 / e2e_steps

    # ---- reduce over ranks: the slowest rank defines the step ----
    ms_total, e2e_ms, launches, copy_ms = ctx.max_over_ranks([ms_total, e2e_ms, float(launches), copy_ms])
    ms_step = ms_total / steps
    value = world * batch / (ms_step * 1e-3)
    h2d, d2h = job.h2d_bytes(), job.d2h_bytes()
    plan = job.plans[-1]
    del job
    torch.cuda.empty_cache()
    if ctx.rank != 0:
        return None

    peaks = ctx.peaks
    peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    sm_run = float(clocks.get("sm_mhz") or sm_max)                # median SM clock sampled during the timed region
    # dominant kernel = the elimination kernel(s) of a step (device-timed inside the library)
    per_step = max(1, len(kernel_ms) // steps) if kernel_ms else 1
    k_ms = sum(kernel_ms) / steps if kernel_ms else ms_step
    achieved = alg_bytes * batch / (k_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "kernel_ms": k_ms, "kernel_launches_per_step": per_step,
            "algorithmic_bytes_per_matrix": alg_bytes,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650"}
    if captured:
        roof["kernel_launches_per_step"] = captured.kernels
        roof["launch"] = "one CUDA graph per step, captured from the three device-memory calls (%d kernels)" % captured.kernels
    if primes_run:
        roof["primes_run"] = primes_run
        roof["primes_note"] = ("the plan's %d primes cover the declared magnitudes; the row-norm Hadamard bound of the batch's own "
                               "matrices (k_row_bound, inside the timed step) needs %d" % (int(plan.n_primes), primes_run))
    cnt = ctx.counts.get(workload)
    if cnt and (cnt.get("batch") == batch or cnt.get("per_matrix")) \
            and cnt.get("primes_run", primes_run) == primes_run:
        # executed figures: warp instructions of the dominant kernel(s) per step (ncu capture of this build) over
        # the issue slots the live-timed kernel had: 4 schedulers x SMs x cycles at the SM clock sampled in this run.
        # per_matrix: the step runs the captured launch batch / capture-batch times (chunks of the same kernel)
        scale = batch / cnt["batch"]
        slots = 4.0 * 148 * (k_ms * 1e-3) * sm_run * 1e6
        roof["traffic"] = cnt.get("dram_bytes_per_step") * scale if cnt.get("dram_bytes_per_step") else None
        roof["issue_slots"] = {"warp_instructions_per_step": cnt["warp_inst_per_step"] * scale, "slots": slots,
                               "frac": cnt["warp_inst_per_step"] * scale / slots, "sm_mhz": sm_run,
                               "source": cnt.get("source")}
        if cnt.get("fmaheavy_busy_ncu") is not None:
            roof["int_pipe"] = {"fmaheavy_busy_ncu": cnt["fmaheavy_busy_ncu"], "issue_slots_busy_ncu": cnt.get("issue_slots_busy_ncu"),
                                "kernel": cnt.get("kernel"), "source": cnt.get("source")}
    if workload == "c2":
        # the fused 8x8 kernel computes over the integers (Bareiss): no modular multiply-subtracts to count.  Its limiting
        # unit is the fmaheavy pipe (IMAD / IMAD.WIDE), whose busy fraction is an ncu figure of the same kernel source
        roof["int_pipe"] = {"fmaheavy_busy_ncu": 0.754, "issue_slots_busy_ncu": 0.52, "warp_instructions_per_32_matrices": 2183,
                            "source": "profiles/r02o_ncu_full_k_inv_tpm8_bareiss.txt (sm__pipe_fmaheavy_cycles_active, one launch "
                                      "under ncu at 128 us); static mix: profiles/r02o_sass_hist_k_inv_tpm8_bareiss.txt"}
    elif workload in ALG_OPS:
        # ALGORITHMIC count, not executed instructions
        n_pr = primes_run or int(plan.n_primes)
        ip_peak = MONT_MUL_PER_SM_CLK * 148 * sm_max * 1e6
        ip_ach = ALG_OPS[workload] * n_pr * batch / (k_ms * 1e-3)
        roof["algorithmic_int"] = {"ops_per_matrix": ALG_OPS[workload] * n_pr, "achieved": ip_ach, "peak": ip_peak,
                                   "unit": "algorithmic modular multiply-subtracts/s", "frac": ip_ach / ip_peak,
                                   "note": "algorithmic operation count over the measured mont_mul rate; NOT a pipe utilisation",
                                   "peak_source": "measured mont_mul rate 12.0 per SM per clock (profiles/r01_ubench_int.jsonl) x 148 SMs x max SM clock"}
    return {
        "metric": METRIC[workload], "value": value, "unit": "matrices/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "i32 (exact integers, fraction-free elimination)" if workload == "c2" else "u32 (Montgomery words modulo 31-bit primes)",
        "data": "synthetic",
        "config": config_for(workload, world),
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": {"value": world * batch / (e2e_ms * 1e-3), "unit": "matrices/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_calls_per_step": E2E_CALLS.get(workload, 1),
                "gb_per_s_per_gpu": (h2d + d2h) / (e2e_ms * 1e-3) / 1e9,
                "copy_only_ms": copy_ms, "frac_of_copy_ceiling": copy_ms / e2e_ms,
                "copy_only": "the step's H2D and D2H bytes as bare pinned copies on two streams, all ranks at once, "
                             "max over ranks: the PCIe / host-memory ceiling of this box at this N",
                "path": "lsx_*_batch(mem=LSX_MEM_HOST) via ctypes, pinned host buffers"
                        + (" (int8 input container: lsx_inverse_batch_i8)" if workload == "c2" else "")},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }


def c5_total_ops(n, n_primes_local):
    """int8-equivalent tensor operations of the WHOLE LU (2 n^3 / 3 multiply-adds per prime, 16 byte-plane products
    per residue multiply-add, 2 operations per multiply-add) -- the whole-step figure next to the dominant kernel's."""
    return 2 * 16 * (n ** 3 // 3) * n_primes_local


def int8_peak(peaks):
    """Dense int8 tensor peak in TOP/s and where it comes from: the measurement-only probe (tools/int8_peak_probe.py,
    a library int8 GEMM, never linked into liblsx) when its result is committed, else 2 x the measured bf16 rate."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "int8_peak.json")))
        return float(d["int8_tops_sustained"]), "profiles/int8_peak.json (measured: %s)" % d.get("how", "library int8 GEMM")
    except Exception:
        bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
        return 2.0 * bf16, "2 x MEASURED_PEAKS.json bf16_tflops_sustained (nominal int8 : bf16 ratio, not measured)"


def measure_c5(ctx, steps, warmup, cpu):
    """BASELINE.json configs[4] at this run's N ranks: primes sharded over the ranks, one all-gather, CRT on every rank."""
    import numpy as np
    torch = ctx.torch
    from linalg_solver_b200 import Engine
    from linalg_solver_b200 import dist as lsx_dist
    eng, dev, world, rank = ctx.eng, ctx.dev, ctx.world, ctx.rank

    worst_primes, _ = Engine.det_large_prime_count(C5_N, C5_ABS)   # worst case for |entries| <= 5: 1100 primes
    A_host = torch.from_numpy(c5_matrix()).pin_memory()
    A = A_host.to(dev)
    # parity before timing: the blocked tensor-core path against the numpy oracle on a 512 x 512 block, two primes
    if rank == 0:
        from oracle.det_mod_p import det_mod_p
        small = A_host[:512, :512].contiguous().numpy()
        got = eng.det_large_residues(torch.from_numpy(small).to(dev), 0, 2).cpu().numpy().astype(np.uint32)
        want = [det_mod_p(small, int(p)) for p in eng.primes(2)]
        assert [int(x) for x in got] == want, "blocked LU residues differ from oracle/det_mod_p.py"

    # Hadamard bound from the actual row/column norms of A (rigorous, about 8 % fewer primes than the worst case)
    n_primes, bits = eng.det_large_prime_count_for(A)
    assert n_primes == c5_prime_count_numpy() and n_primes <= worst_primes

    def step():
        return lsx_dist.det_large_sharded(eng, A)

    words = None
    for _ in range(warmup):
        words, _ = step()
    ctx.barrier()
    b, e = lsx_dist.shard_range(n_primes, rank, world)
    launches0 = eng.launch_count
    sampler = ClockSampler(ctx.local)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        words, _ = step()
    ev1.record()
    ctx.barrier()
    t1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(t0, t1)
    # The timed steps run independent prime groups concurrently on two streams, where the duration of a single launch
    # says nothing; the dominant kernel's launches are event-timed in ONE extra step, for which the library falls back
    # to a single stream (lsx_blocked.cu).  Same kernels, same launches, same data.
    eng.timing_enable(True)
    step()
    torch.cuda.synchronize()
    kernel_ms = eng.timing_read()
    eng.timing_enable(False)
    kernel_steps = 1

    # ---- end to end: host matrix -> H2D -> residues -> all-gather -> CRT -> limbs back on the host ----
    e2e_steps = 3
    ctx.barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        A_d = A_host.to(dev, non_blocking=True)
        w, _ = lsx_dist.det_large_sharded(eng, A_d)
        w_host = w.cpu()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / e2e_steps
    assert torch.equal(w_host, words.cpu())

    ms_total, e2e_ms, launches, k_sum = ctx.max_over_ranks([ms_total, e2e_ms, float(launches), sum(kernel_ms)])
    ms_step = ms_total / steps
    words_host = words.cpu().numpy().astype(np.uint32)
    del A, words
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = int8_peak(ctx.peaks)
    k_s = k_sum / kernel_steps * 1e-3                             # depth-256 tensor updates of one step (slowest rank)
    ops = c5_tensor_ops(C5_N, e - b)
    achieved = ops / k_s / 1e12 if k_s > 0 else 0.0
    whole = c5_total_ops(C5_N, e - b) / (ms_step * 1e-3) / 1e12
    limbs = int(bits + 2) // 32 + 1
    from linalg_solver_b200.convert import limbs_to_ints
    det = limbs_to_ints(words_host.reshape(1, -1))[0]
    return {
        "metric": C5_METRIC, "value": ms_step * 1e-3, "unit": "s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32 residues modulo 31-bit primes; trailing update as u8 x u8 -> s32 tcgen05 MMA",
        "data": "synthetic",
        "config": config_for("c5", world),
        "result": {"log2_bound": bits, "limbs": limbs, "det_bits": int(abs(det)).bit_length(), "det_mod_1e9": int(det % 10**9)},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel_ms": k_s * 1e3, "kernel": "lsx_tc::k_gemm_tc (depth-256 trailing updates)",
                     "kernel_share_of_step": k_s * 1e3 / ms_step,
                     "kernel_launches_per_step": len(kernel_ms) // kernel_steps,
                     "kernel_timing": "event-timed in one extra single-stream step after the timed region (the timed steps "
                                      "overlap two prime groups on two streams)",
                     "ops": "int8 tensor ops: 2 x 16 byte-plane products per residue multiply-add",
                     "whole_step": {"achieved": whole, "frac": whole / peak, "unit": "TFLOP/s",
                                    "ops": "2 x 16 x n^3/3 per prime over the whole step (panels, solves, splits and "
                                           "swaps included in the time, only the LU's multiply-adds in the count)"},
                     "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_ms * 1e-3, "unit": "s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(A_host.numel() * 4),
                "d2h_bytes_per_step": int(limbs * 4),
                "path": "pinned host matrix -> device, linalg_solver_b200.dist.det_large_sharded (lsx_det_large_residues, "
                        "all-gather, lsx_crt_signed), limbs back to the host"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }


# ------------------------------------------------------------------------------------- CPU legs
def cpu_leg(workload, scale=8):
    """cpu_baseline of one workload (the pinned oracle port on all host cores, bounded sample)."""
    if workload == "c5":
        return c5_cpu_baseline(c5_prime_count_numpy())
    cb, _ = cpu_baseline(workload, scale * CPU_PER_CORE[workload], SEED)
    return with_reference(cb, workload)


def with_reference(cb, workload):
    """Adds the unmodified reference's own timing and says which CPU route is the faster one: the port restates
    row_reduce on Fractions (linalg.py:534-630), the reference's DEFAULT routes go through sympy (inverse():
    linalg.py:696-701), which is slower for small matrices and faster for 64x64."""
    ref = reference_unmodified(workload)
    if ref is not None:
        cb["reference_unmodified"] = ref
        if "value" in ref:
            cb["fastest_cpu_route"] = "reference_unmodified" if ref["value"] > cb["value"] else "port"
    return cb


def _ref_one(item):
    """One unit through the UNMODIFIED reference package vendored under oracle/_ref (oracle/build_ref.py)."""
    import sympy
    from linalg_solver.linalg import Matrix
    wl, a, b = item
    rat = [[sympy.Rational(x) for x in row] for row in a]
    if wl == "c1":
        M = Matrix(rat)
        return M.rank(), len(M.row_reduce()[1])
    if wl == "c3":
        res = Matrix(rat).find_preimage_of([sympy.Rational(x) for x in b])
        return isinstance(res, Matrix.NoSolution)
    if wl == "c4ker":
        return Matrix(rat).kernel().generators.cols
    n = len(a)
    if wl == "c2":
        aug = [list(rat[i]) + [sympy.Integer(1 if i == j else 0) for j in range(n)] for i in range(n)]
        return len(Matrix(aug).row_reduce(bar_col=n)[1])
    return isinstance(Matrix(rat).inverse(), Matrix.NoSolution)


REF_UNITS_PER_CORE = {"c1": 256, "c2": 64, "c3": 8, "c4inv": 4, "c4ker": 0}


def reference_unmodified(workload):
    """The reference's own Python (unmodified, oracle/_ref) on a SMALL sample of the workload, all host cores: how
    conservative the port-based cpu_baseline is.  None when oracle/_ref was not built (it is git-ignored)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    per_core = REF_UNITS_PER_CORE.get(workload, 0)
    if not os.path.isdir(os.path.join(ref_dir, "linalg_solver")) or per_core == 0:
        return None
    from multiprocessing import get_context
    for d in (ref_dir,):
        if d not in sys.path:
            sys.path.insert(0, d)
    try:
        from linalg_solver.log import global_logger
        global_logger._auto_print = False
    except Exception as e:                                        # a broken vendored copy must not sink the bench
        return {"unavailable": "import of oracle/_ref failed: %r" % (e,)}
    desc, n, _, _ = WORKLOADS[workload]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    data = make_inputs(n, cores * per_core, SEED, workload)
    A = data["A"].tolist()
    b = data["b"].tolist() if "b" in data else [None] * len(A)
    sample = [(workload, A[i], b[i]) for i in range(len(A))]
    calls = {"c1": "Matrix.rank() + row_reduce()", "c2": "row_reduce([A|I], bar_col=8)", "c3": "find_preimage_of(b)",
             "c4inv": "inverse()", "c4ker": "kernel()"}[workload]
    ctx = get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_one, sample[:cores], chunksize=1)
        t0 = time.perf_counter()
        pool.map(_ref_one, sample, chunksize=1)
        dt = time.perf_counter() - t0
    return {"value": len(sample) / dt, "unit": "matrices/s", "cores": cores, "kind": "reference",
            "sample": "%d units through the unmodified reference (%s on sympy.Rational entries, the Rust planner "
                      "replaced by oracle/standin), %.1f s" % (len(sample), calls, dt)}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    main_wl = args.workload
    extras = [w for w in args.extra_list if w != main_wl]
    order = [main_wl] + extras

    # CPU legs first (fork pools must precede CUDA initialisation); rank 0 at N = 1 only
    cpus = {}
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        for w in order:
            cpus[w] = cpu_leg(w)

    ctx = Ctx(args)
    sub_steps = max(3, min(args.steps, 10))
    recs = {}
    for i, w in enumerate(order):
        steps = args.steps if i == 0 else (sub_steps if w not in ("c4inv", "c4ker", "c5") else max(3, min(args.steps, 5)))
        warm = args.warmup if i == 0 else 3
        rec = measure_c5(ctx, steps, warm, cpus.get(w)) if w == "c5" else measure_batch(ctx, w, steps, warm, cpus.get(w))
        recs[w] = rec
    if ctx.rank == 0:
        line = recs[main_wl]
        if extras:
            line["workloads"] = {w: recs[w] for w in extras}
        print(json.dumps(line), flush=True)
    ctx.close()


def run_c5_reference_line(steps, warmup, gpus):
    k = c5_prime_count_numpy()
    vals = []
    cb = None
    for i in range(warmup + steps):
        cb = c5_cpu_baseline(k)
        if i >= warmup:
            vals.append(cb["value"])
    val = statistics.mean(vals)
    cb["value"] = val
    return {
        "impl": "reference", "metric": C5_METRIC, "value": val, "unit": "s", "n_gpus": gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64 residues modulo 31-bit primes (numpy)", "data": "synthetic",
        "config": config_for("c5", gpus), "sample_per_step": "bounded sample on the host CPU, extrapolated",
        "cpu_baseline": cb, "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["c5"],
                    help="headline workload of the line (default c2, with every other config as a sub-record)")
    ap.add_argument("--extra", default=None,
                    help="comma list of workloads reported as sub-records under `workloads`, or none "
                         "(default: all other configs when --workload is not given, none otherwise)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.extra is None:
        args.extra_list = list(EXTRA_DEFAULT) if args.workload is None else []
    else:
        args.extra_list = [] if args.extra in ("", "none") else [w for w in args.extra.split(",") if w]
        bad = [w for w in args.extra_list if w not in WORKLOADS and w != "c5"]
        if bad:
            ap.error("unknown workload(s) in --extra: %s" % ",".join(bad))
    if args.workload is None:
        args.workload = "c2"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
