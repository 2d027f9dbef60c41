#!/usr/bin/env python
"""Benchmark of the exact elimination hot path (BASELINE.json metric: exact det/RREF matrices/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4inv|c5]

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
    "c4inv": ("2^12 (of 2^16) x 64x64 inverse via [A|I] row_reduce(bar_col=64), entries uniform [-5,5]", 64, 1 << 12,
              196912),
    "c4ker": ("2^12 (of 2^16) x kernel basis of 64x64 A = B(64x48) C(48x64) (rank 48, dim 16), entries of B, C uniform [-5,5]",
              64, 1 << 12, 98644),
}
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


def make_inputs(n, batch, seed, workload="c2"):
    """-> dict of int32 arrays for one rank's batch."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
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
        return {"A": np.einsum("bik,bkj->bij", Bm, Cm).astype(np.int32), "b": np.zeros((batch, 64), dtype=np.int32)}
    return {"A": rng.integers(-5, 6, size=(batch, n, n), dtype=np.int32)}


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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, n, batch, _ = WORKLOADS[args.workload]
    per_core = CPU_PER_CORE[args.workload]
    times = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, _ = cpu_baseline(args.workload, per_core, SEED + i)
        if i >= args.warmup:
            times.append(cb["value"])
    val = statistics.mean(times)
    cb["value"] = val
    line = {
        "impl": "reference", "metric": METRIC[args.workload], "value": val, "unit": "matrices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * per_core * cb["cores"] / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "exact rationals (fractions.Fraction)", "data": "synthetic",
        "config": {"workload": desc, "step": "bounded sample of the workload on the host CPU"},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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


# ------------------------------------------------------------------------------------- GPU arm
class Job:
    """One workload bound to an engine: device-resident step, host-buffer (end-to-end) step, oracle check."""

    def __init__(self, eng, workload, rank, dev):
        import numpy as np
        import torch
        self.eng, self.wl, self.np, self.torch = eng, workload, np, torch
        desc, n, batch, _ = WORKLOADS[workload]
        self.n, self.batch = n, batch
        # by-matrix sharding: every rank owns an independent batch (its own seed), no collective on the data path
        data = make_inputs(n, batch, SEED + 1000 * rank, workload)
        self.host = {k: torch.from_numpy(v).pin_memory() for k, v in data.items()}
        self.dev = {k: v.to(dev) for k, v in self.host.items()}
        if workload == "c1":
            self.plans = (eng.plan_det(4, 5), eng.plan_rank(4, 4, 5), eng.plan_rref(4, 4, 3, 5, 5))
        elif workload == "c3":
            bmax = int(data["b"].max()), int(-data["b"].min())
            self.plans = (eng.plan_solve(16, 16, 250, max(bmax), 10, 6),)
        elif workload == "c4ker":
            self.plans = (eng.plan_solve(64, 64, int(abs(data["A"]).max()), 0, 48, 16),)
        else:
            self.plans = (eng.plan_inverse(n, 5),)
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
        outs = []
        for r in self.res:
            kw = {k: torch.empty(tuple(v.shape), dtype=torch.int32).pin_memory().numpy() for k, v in self._fields(r)}
            outs.append(type(r)(plan=r.plan, **{k: kw.get(k) for k in vars(r) if k != "plan"}))
        self.host_out = tuple(outs)
        self.host_np = {k: v.numpy() for k, v in self.host.items()}
        if self.wl == "c2":
            # entries are in [-5, 5]: the end-to-end call ships them as int8 (lsx_inverse_batch_i8), a quarter
            # of the host-to-device bytes of the int32 container; the results are the same words
            self.host_i8 = self.host["A"].to(torch.int8).pin_memory()
            self.host_np = {"A": self.host_i8.numpy()}

    def step_e2e(self):
        self.run(self.host_np, self.host_out)                 # returns when the results are in host memory

    def h2d_bytes(self):
        return int(sum(v.nbytes for v in self.host_np.values()))

    def d2h_bytes(self):
        return int(sum(v.nbytes for r in self.host_out for _, v in self._fields(r)))

    def check_e2e_equals_device(self):
        np = self.np
        for rh, rd in zip(self.host_out, self.res):
            for (k, vh), (_, vd) in zip(self._fields(rh), self._fields(rd)):
                assert np.array_equal(vh.reshape(-1)[:4096], vd.reshape(-1)[:4096].cpu().numpy()), k

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


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    desc, n, batch, alg_bytes = WORKLOADS[args.workload]

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        cpu, _ = cpu_baseline(args.workload, 8 * CPU_PER_CORE[args.workload], SEED)   # before CUDA init (fork pool)

    import torch
    import torch.distributed as dist
    from linalg_solver_b200 import Engine

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    job = Job(eng, args.workload, rank, dev)
    torch.cuda.synchronize(dev)
    if rank == 0:
        job.check_against_oracle()

    for _ in range(args.warmup):
        job.step_device()
    barrier()
    launches0 = eng.launch_count
    eng.timing_enable(True)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        job.step_device()
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = eng.timing_read()
    eng.timing_enable(False)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(t0, t1)

    # ---- end to end through the C-ABI with host buffers (pinned): H2D + kernels + D2H per step ----
    job.make_host_outputs()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        job.step_e2e()
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        job.step_e2e()
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - e0) * 1e3 / e2e_steps
    job.check_e2e_equals_device()

    # ---- reduce over ranks: the slowest rank defines the step ----
    stats = torch.tensor([ms_total, e2e_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(stats[0]), float(stats[1])
    ms_step = ms_total / args.steps
    value = world * batch / (ms_step * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the elimination kernel(s) of a step (device-timed inside the library)
        per_step = max(1, len(kernel_ms) // args.steps) if kernel_ms else 1
        k_ms = sum(kernel_ms) / args.steps if kernel_ms else ms_step
        achieved = alg_bytes * batch / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        plan = job.plans[-1]
        int_pipe = None
        if args.workload in ALG_OPS:
            sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
            ip_peak = MONT_MUL_PER_SM_CLK * 148 * sm_mhz * 1e6
            # the fused 8x8 kernel runs ONE prime (exact int64 determinant); the other kernels run the plan's primes
            n_pr = 1 if args.workload == "c2" else int(plan.n_primes)
            ip_ach = ALG_OPS[args.workload] * n_pr * batch / (k_ms * 1e-3)
            int_pipe = {"algorithmic_ops_per_matrix": ALG_OPS[args.workload] * n_pr, "achieved": ip_ach, "peak": ip_peak,
                        "unit": "modular multiply-subtracts/s", "frac": ip_ach / ip_peak,
                        "peak_source": "measured mont_mul rate 12.0 per SM per clock (profiles/r01_ubench_int.jsonl) x 148 SMs x max SM clock"}
        line = {
            "metric": METRIC[args.workload], "value": value, "unit": "matrices/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 (Montgomery words modulo 31-bit primes)",
            "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": batch, "primes": int(plan.n_primes), "limbs": int(plan.limbs),
                       "sharding": "by matrix, no collective",
                       "l2": "inputs+outputs per step (%.0f MB) vs the 126 MB L2" % ((alg_bytes * batch) / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel_ms": k_ms, "kernel_launches_per_step": per_step,
                         "algorithmic_bytes_per_matrix": alg_bytes, "int_pipe": int_pipe,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650"},
            "cpu_baseline": cpu,
            "e2e": {"value": world * batch / (e2e_ms * 1e-3), "unit": "matrices/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": job.h2d_bytes(), "d2h_bytes_per_step": job.d2h_bytes(),
                    "path": "lsx_*_batch(mem=LSX_MEM_HOST) via ctypes, pinned host buffers"
                            + (" (int8 input container: lsx_inverse_batch_i8)" if args.workload == "c2" else "")},
            "gpu_launches": int(stats[2]),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()



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


def run_c5_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from linalg_solver_b200 import _lib  # noqa: F401  (prime count comes from the same plan function)
    import ctypes
    k = ctypes.c_int(c5_prime_count_numpy())
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = c5_cpu_baseline(k.value)
        if i >= args.warmup:
            vals.append(cb["value"])
    val = statistics.mean(vals)
    cb["value"] = val
    print(json.dumps({
        "impl": "reference", "metric": C5_METRIC, "value": val, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64 residues modulo 31-bit primes (numpy)", "data": "synthetic",
        "config": {"workload": C5_DESC, "primes": k.value, "step": "bounded sample on the host CPU, extrapolated"},
        "cpu_baseline": cb, "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_c5(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import numpy as np
    import torch
    import torch.distributed as dist
    from linalg_solver_b200 import Engine
    from linalg_solver_b200 import dist as lsx_dist

    worst_primes, _ = Engine.det_large_prime_count(C5_N, C5_ABS)   # worst case for |entries| <= 5: 1100 primes
    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        cpu = c5_cpu_baseline(c5_prime_count_numpy())             # before CUDA init (fork pool)

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

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
    for _ in range(args.warmup):
        words, _ = step()
    barrier()
    b, e = lsx_dist.shard_range(n_primes, rank, world)
    launches0 = eng.launch_count
    eng.timing_enable(True)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        words, _ = step()
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = eng.timing_read()
    eng.timing_enable(False)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(t0, t1)

    # ---- end to end: host matrix -> H2D -> residues -> all-gather -> CRT -> limbs back on the host ----
    e2e_steps = max(3, min(args.steps, 5))
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        A_d = A_host.to(dev, non_blocking=True)
        w, _ = lsx_dist.det_large_sharded(eng, A_d)
        w_host = w.cpu()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / e2e_steps
    assert torch.equal(w_host, words.cpu())

    stats = torch.tensor([ms_total, e2e_ms, float(launches), sum(kernel_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_step = float(stats[0]) / args.steps
    e2e_ms = float(stats[1])
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # int8 tensor peak is not in MEASURED_PEAKS.json: twice the measured dense bf16 rate (B200: int8 = 2 x bf16)
        bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak = 2.0 * bf16
        k_s = float(stats[3]) / args.steps * 1e-3                 # depth-256 tensor updates of one step (slowest rank)
        ops = c5_tensor_ops(C5_N, e - b)
        achieved = ops / k_s / 1e12 if k_s > 0 else 0.0
        limbs = int(bits + 2) // 32 + 1
        from linalg_solver_b200.convert import limbs_to_ints
        det = limbs_to_ints(words.cpu().numpy().astype(np.uint32).reshape(1, -1))[0]
        line = {
            "metric": C5_METRIC, "value": ms_step * 1e-3, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 residues modulo 31-bit primes; trailing update as u8 x u8 -> s32 tcgen05 MMA",
            "data": "synthetic",
            "config": {"workload": C5_DESC, "primes": n_primes, "primes_worst_case_bound": worst_primes,
                       "bound": "Hadamard with the actual row/column norms, log2 = %.1f" % bits,
                       "primes_per_gpu": e - b, "limbs": limbs,
                       "sharding": "by prime, one all-gather of %d residues (%d B) before the CRT" % (n_primes, 4 * n_primes),
                       "l2": "residue matrices of one prime group (64 MiB per prime) far exceed the 126 MB L2",
                       "det_bits": int(abs(det)).bit_length(), "det_mod_1e9": int(det % 10**9)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": None, "kernel_ms": k_s * 1e3, "kernel": "lsx_tc::k_gemm_tc (depth-256 trailing updates)",
                         "kernel_launches_per_step": len(kernel_ms) // max(1, args.steps),
                         "ops": "int8 tensor ops: 2 x 16 byte-plane products per residue multiply-add",
                         "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops_sustained (nominal int8 : bf16 ratio; the MMA-only stream of this kernel measures 2.41e15 op/s, profiles/r01i)"
                                        if peaks else "fallback 2 x 1400"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_ms * 1e-3, "unit": "s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(A_host.numel() * 4),
                    "d2h_bytes_per_step": int(limbs * 4),
                    "path": "pinned host matrix -> device, linalg_solver_b200.dist.det_large_sharded (lsx_det_large_residues, "
                            "all-gather, lsx_crt_signed), limbs back to the host"},
            "gpu_launches": int(stats[2]),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "c5":
        (run_c5_reference if args.impl == "reference" else run_c5)(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
