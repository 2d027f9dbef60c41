"""Batched exact elimination on one GPU: thin Python face of the C-ABI (include/lsx.h).

The reference has no batch API; these entry points are what its
``Matrix`` methods are served from (linalg_solver_b200/matrix.py) and what the benchmark
drives directly.  Inputs are ``int32[batch, m, n]`` arrays: numpy arrays (host memory, the call
copies in and out) or torch CUDA tensors (device memory, the call only enqueues kernels on
torch's current stream).  Outputs come back in the same kind of container.

Every rational result is an integer numerator over ONE common denominator per matrix (the
determinant of the pivot minor), each integer as ``limbs`` little-endian 32-bit words in two's
complement; ``linalg_solver_b200.convert`` turns them into Python ints / Fractions.
"""
import ctypes
import os
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

from . import _lib
from ._lib import Plan, lib


class LsxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("liblsx error %s (%d): %s" % (_lib.ERR_NAMES.get(code, "?"), code, msg))
        self.code = code


def _is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


@dataclass
class RrefResult:
    num: Any          # [batch, m, n, limbs] uint32   numerators d * RREF
    den: Any          # [batch, limbs] uint32         d
    pivot_col: Any    # [batch, pivot_slots] int32    (-1 padded)
    rank: Any         # [batch] int32
    status: Any       # [batch] int32
    plan: Plan


@dataclass
class InverseResult:
    adj: Any          # [batch, n, n, limbs]   A^-1 = adj / det
    det: Any          # [batch, limbs]
    status: Any
    plan: Plan


@dataclass
class DetResult:
    det: Any          # [batch, limbs]
    rank: Any
    status: Any
    plan: Plan


@dataclass
class RankResult:
    rank: Any
    status: Any
    plan: Plan


@dataclass
class SolveResult:
    den: Any          # [batch, limbs]
    particular: Any   # [batch, n, limbs]
    generators: Any   # [batch, n, gen_cap, limbs]
    pivot_col: Any
    rank: Any
    status: Any
    plan: Plan


class CapturedCalls:
    """A CUDA graph of liblsx device-memory calls (Engine.capture); `kernels` = launches one replay stands for."""

    def __init__(self, eng, graph, kernels):
        self.eng, self.graph, self.kernels = eng, graph, kernels

    def replay(self):
        self.graph.replay()
        self.eng._replayed_launches = getattr(self.eng, "_replayed_launches", 0) + self.kernels


class Engine:
    """One lsx context = one GPU + one stream.  Not thread-safe (one thread per Engine).

    ``Engine(devices=[0, 1, ...])`` is one context over several GPUs of this process (``lsx_create_multi``): batched
    calls on host (numpy) arrays are sharded by matrix over the GPUs inside the library and ``det_large`` shards a
    single large determinant by prime with an NCCL all-gather of the residues; device tensors are not accepted there
    (they belong to one GPU).  The torch.distributed route (one process per GPU, ``linalg_solver_b200.dist``) is
    unchanged and is what ``bench.py`` drives under torchrun."""

    def __init__(self, device: Optional[int] = None, devices=None):
        self._ctx = ctypes.c_void_p()
        if devices is not None:
            ids = [int(d) for d in devices]
            if not ids:
                raise ValueError("devices must name at least one GPU")
            self.device, self.devices = ids[0], ids
            arr = (ctypes.c_int * len(ids))(*ids)
            rc = lib.lsx_create_multi(arr, len(ids), ctypes.byref(self._ctx))
            if rc != _lib.OK:
                self._ctx = None
                raise LsxError(rc, "lsx_create_multi(devices=%r) failed: CUDA devices are required, there is no CPU "
                                   "fallback" % (ids,))
            return
        if device is None:
            device = int(os.environ.get("LSX_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.device, self.devices = device, [device]
        rc = lib.lsx_create(device, ctypes.byref(self._ctx))
        if rc != _lib.OK:
            self._ctx = None
            raise LsxError(rc, "lsx_create(device=%d) failed: a CUDA device is required, there is no CPU fallback"
                           % device)

    # ---- plumbing -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            lib.lsx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.OK:
            raise LsxError(rc, lib.lsx_last_error(self._ctx).decode(errors="replace"))

    def synchronize(self):
        self._check(lib.lsx_synchronize(self._ctx))

    @property
    def launch_count(self) -> int:
        """Kernels launched so far, replays of captured graphs included."""
        return int(lib.lsx_launch_count(self._ctx)) + getattr(self, "_replayed_launches", 0)

    def capture(self, fn, warmup: int = 2) -> "CapturedCalls":
        """Record the device-memory calls `fn()` makes into ONE CUDA graph and return it (`.replay()`).

        Device-memory calls only enqueue on the caller's stream (include/lsx.h), so a launch-bound sequence -- the
        three calls of a 10k x 4x4 determinant + rank + row_reduce step are 38 us of kernels in a 79 us step -- can
        be captured once and replayed with one launch.  `fn` must pass preallocated outputs (`out=`) and CUDA tensors,
        and must not read results back; it is run `warmup` times first, so that workspace growth and kernel
        attributes are settled before the capture.  A replay runs on torch's current stream."""
        import torch
        if getattr(self, "_timing_on", False):
            raise RuntimeError("capture() with kernel timing enabled: the event pairs would be recorded once, at capture")
        dev = torch.device("cuda", self.device)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        n0 = int(lib.lsx_launch_count(self._ctx))
        # thread_local: CUDA calls of OTHER host threads (the NCCL watchdog of an initialised process group polls
        # events) must not invalidate the capture
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            fn()
        return CapturedCalls(self, graph, int(lib.lsx_launch_count(self._ctx)) - n0)

    def last_prime_count(self) -> int:
        """Primes per matrix the last tile-path call used (its row-norm Hadamard bound; 0: not the tile path)."""
        k = ctypes.c_int(0)
        self._check(lib.lsx_last_prime_count(self._ctx, ctypes.byref(k)))
        return int(k.value)

    def timing_enable(self, on: bool = True):
        """Record CUDA events around the dominant kernel of each following call."""
        self._check(lib.lsx_timing_enable(self._ctx, 1 if on else 0))
        self._timing_on = bool(on)

    def timing_read(self, cap: int = 4096):
        """Durations (ms) of the dominant kernels recorded since the last read (waits for them)."""
        buf = np.zeros(cap, dtype=np.float32)
        cnt = ctypes.c_int()
        self._check(lib.lsx_timing_read(self._ctx, buf.ctypes.data, cap, ctypes.byref(cnt)))
        return buf[: min(cnt.value, cap)].tolist()

    def primes(self, count: int):
        out = np.empty(count, dtype=np.uint32)
        self._check(lib.lsx_get_primes(self._ctx, out.ctypes.data, count))
        return out

    def debug_set_primes(self, primes):
        arr = np.ascontiguousarray(np.asarray(primes, dtype=np.uint32))
        self._check(lib.lsx_debug_set_primes(self._ctx, arr.ctypes.data if arr.size else None, int(arr.size)))

    def set_stream(self, cuda_stream_ptr):
        self._check(lib.lsx_set_stream(self._ctx, cuda_stream_ptr))

    @staticmethod
    def _is_i8(x):
        if _is_torch(x):
            import torch
            return x.dtype == torch.int8
        return isinstance(x, np.ndarray) and x.dtype == np.int8

    @staticmethod
    def _widen(x):
        if _is_torch(x):
            import torch
            return x.to(torch.int32)
        return x.astype(np.int32)

    def _prep_in(self, x, ndim, name, i8=False):
        """-> (array, pointer, mem, like) for an int32 input (int8 when i8: lsx_inverse_batch_i8)."""
        if _is_torch(x):
            import torch
            if x.dtype != (torch.int8 if i8 else torch.int32):
                raise TypeError("%s must be %s, got %s" % (name, "int8" if i8 else "int32", x.dtype))
            if x.dim() != ndim:
                raise ValueError("%s must have %d dimensions" % (name, ndim))
            x = x.contiguous()
            if x.is_cuda:
                if x.device.index != self.device:
                    raise ValueError("%s lives on cuda:%s, this engine is bound to cuda:%d"
                                     % (name, x.device.index, self.device))
                # torch's default stream is the legacy NULL stream: pass the cudaStreamLegacy handle (0x1),
                # because NULL means "the ctx's own stream" to lsx_set_stream
                self.set_stream(torch.cuda.current_stream(x.device).cuda_stream or 1)
                return x, x.data_ptr(), _lib.MEM_DEVICE, x
            self.set_stream(None)
            return x, x.data_ptr(), _lib.MEM_HOST, x
        a = np.ascontiguousarray(x)
        if i8:
            if a.ndim != ndim:
                raise ValueError("%s must have %d dimensions" % (name, ndim))
            self.set_stream(None)
            return a, a.ctypes.data, _lib.MEM_HOST, None
        if a.dtype != np.int32:
            if not np.issubdtype(a.dtype, np.integer):
                raise TypeError("%s must be an integer array" % name)
            if a.size and (a.max() > 2**31 - 1 or a.min() < -(2**31) + 1):
                raise OverflowError("%s has entries outside int32" % name)
            a = a.astype(np.int32)
        if a.ndim != ndim:
            raise ValueError("%s must have %d dimensions" % (name, ndim))
        self.set_stream(None)
        return a, a.ctypes.data, _lib.MEM_HOST, None

    @staticmethod
    def _alloc(like, shape, dtype):
        if like is not None:
            import torch
            tdt = {np.uint32: torch.int32, np.int32: torch.int32}[dtype]   # uint32 words held in int32 storage
            return torch.empty(shape, dtype=tdt, device=like.device, pin_memory=False)
        return np.empty(shape, dtype=dtype)

    @staticmethod
    def _ptr(x):
        if x is None:
            return None
        return x.data_ptr() if _is_torch(x) else x.ctypes.data

    @staticmethod
    def _absmax(x):
        if _is_torch(x):
            return int(x.abs().max().item()) if x.numel() else 0
        return int(np.abs(x.astype(np.int64)).max()) if x.size else 0

    # ---- plans ----------------------------------------------------------------------------
    @staticmethod
    def _check_plan(rc, what):
        """lsx_plan_* take no ctx, so lsx_last_error has nothing to say about them: explain the code here."""
        if rc == _lib.OK:
            return
        why = {
            _lib.ERR_BAD_SHAPE: "invalid dimensions, bar_col or magnitudes",
            _lib.ERR_UNSUPPORTED: "more than 254 rows: the batched kernels keep one residue tile per CTA "
                                  "(a single large determinant goes through det_large_residues / Matrix.determinant)",
            _lib.ERR_BOUND: "the Hadamard bound of the declared magnitudes needs more than 32 primes per matrix, the "
                            "limit of the batched kernels",
        }.get(rc, "plan query failed")
        raise LsxError(rc, "%s: %s" % (what, why))

    def plan_rref(self, m, n, bar_col, a_abs_max, b_abs_max=None, max_rank=0) -> Plan:
        p = Plan()
        self._check_plan(lib.lsx_plan_rref(m, n, bar_col, a_abs_max, a_abs_max if b_abs_max is None else b_abs_max,
                                      max_rank, ctypes.byref(p)), "plan_rref(m=%d, n=%d, bar_col=%d)" % (m, n, bar_col))
        return p

    def plan_inverse(self, n, a_abs_max) -> Plan:
        p = Plan()
        self._check_plan(lib.lsx_plan_inverse(n, a_abs_max, ctypes.byref(p)), "plan_inverse(n=%d)" % n)
        return p

    def plan_det(self, n, a_abs_max) -> Plan:
        p = Plan()
        self._check_plan(lib.lsx_plan_det(n, a_abs_max, ctypes.byref(p)), "plan_det(n=%d)" % n)
        return p

    def plan_rank(self, m, n, a_abs_max) -> Plan:
        p = Plan()
        self._check_plan(lib.lsx_plan_rank(m, n, a_abs_max, ctypes.byref(p)), "plan_rank(m=%d, n=%d)" % (m, n))
        return p

    def plan_solve(self, m, n, a_abs_max, b_abs_max, max_rank=0, gen_cap=None) -> Plan:
        p = Plan()
        self._check_plan(lib.lsx_plan_solve(m, n, a_abs_max, b_abs_max, max_rank, n if gen_cap is None else gen_cap,
                                       ctypes.byref(p)), "plan_solve(m=%d, n=%d)" % (m, n))
        return p

    # ---- batched operations ---------------------------------------------------------------
    def rref_batch(self, A, bar_col, a_abs_max=None, b_abs_max=None, max_rank=0, plan=None, out=None) -> RrefResult:
        """Gauss-Jordan of every ``A[i]`` with pivots in columns < bar_col (reference
        linalg.py:534-630; the caller applies the ``bar_col or n-1`` default of line 543)."""
        A, pA, mem, like = self._prep_in(A, 3, "A")
        B, m, n = A.shape
        if plan is None:
            if a_abs_max is None:
                a_abs_max = self._absmax(A)
            plan = self.plan_rref(m, n, bar_col, a_abs_max, b_abs_max, max_rank)
        L = plan.limbs
        if out is not None:
            num, den, piv, rank, status = out.num, out.den, out.pivot_col, out.rank, out.status
        else:
            num = self._alloc(like, (B, m, n, L), np.uint32)
            den = self._alloc(like, (B, L), np.uint32)
            piv = self._alloc(like, (B, plan.pivot_slots), np.int32)
            rank = self._alloc(like, (B,), np.int32)
            status = self._alloc(like, (B,), np.int32)
        self._check(lib.lsx_rref_batch(self._ctx, ctypes.byref(plan), pA, B, mem, self._ptr(num), self._ptr(den),
                                       self._ptr(piv), self._ptr(rank), self._ptr(status)))
        return RrefResult(num, den, piv, rank, status, plan)

    def inverse_batch(self, A, a_abs_max=None, plan=None, out=None) -> InverseResult:
        """A^-1 = adj / det through [A|I] (reference linalg.py:704-743); singular matrices get
        status ST_SINGULAR (the reference returns ``NoSolution()``)."""
        i8 = self._is_i8(A)
        if i8 and plan is None:
            plan = self.plan_inverse(A.shape[-1], self._absmax(A) if a_abs_max is None else a_abs_max)
        if i8 and not (A.shape[-1] <= 8 and plan.limbs == 1):
            A, i8 = self._widen(A), False                        # int8 is served by the fused small kernel only
        A, pA, mem, like = self._prep_in(A, 3, "A", i8=i8)
        B, n, n2 = A.shape
        if n != n2:
            raise ValueError("Matrix must be square to invert.")
        if plan is None:
            plan = self.plan_inverse(n, self._absmax(A) if a_abs_max is None else a_abs_max)
        L = plan.limbs
        if out is None:
            adj = self._alloc(like, (B, n, n, L), np.uint32)
            det = self._alloc(like, (B, L), np.uint32)
            status = self._alloc(like, (B,), np.int32)
        else:
            adj, det, status = out.adj, out.det, out.status
        if i8:
            rc = lib.lsx_inverse_batch_i8(self._ctx, ctypes.byref(plan), pA, B, mem, self._ptr(adj), self._ptr(det),
                                          self._ptr(status))
            if rc != _lib.ERR_UNSUPPORTED:
                self._check(rc)
                return InverseResult(adj, det, status, plan)
            A, pA, mem, like = self._prep_in(self._widen(A), 3, "A")      # plan outside the fused kernel: widen
        self._check(lib.lsx_inverse_batch(self._ctx, ctypes.byref(plan), pA, B, mem, self._ptr(adj), self._ptr(det),
                                          self._ptr(status)))
        return InverseResult(adj, det, status, plan)

    def det_batch(self, A, a_abs_max=None, plan=None, out=None) -> DetResult:
        """Determinants (sign * product of the forward-sweep pivots, linalg.py:547-609) and ranks."""
        A, pA, mem, like = self._prep_in(A, 3, "A")
        B, n, n2 = A.shape
        if n != n2:
            raise ValueError("Determinant requires a square matrix")
        if plan is None:
            plan = self.plan_det(n, self._absmax(A) if a_abs_max is None else a_abs_max)
        if out is not None:
            det, rank, status = out.det, out.rank, out.status
        else:
            det = self._alloc(like, (B, plan.limbs), np.uint32)
            rank = self._alloc(like, (B,), np.int32)
            status = self._alloc(like, (B,), np.int32)
        self._check(lib.lsx_det_batch(self._ctx, ctypes.byref(plan), pA, B, mem, self._ptr(det), self._ptr(rank),
                                      self._ptr(status)))
        return DetResult(det, rank, status, plan)

    def rank_batch(self, A, a_abs_max=None, plan=None, out=None) -> RankResult:
        """Ranks (reference linalg.py:745-747)."""
        A, pA, mem, like = self._prep_in(A, 3, "A")
        B, m, n = A.shape
        if plan is None:
            plan = self.plan_rank(m, n, self._absmax(A) if a_abs_max is None else a_abs_max)
        if out is not None:
            rank, status = out.rank, out.status
        else:
            rank = self._alloc(like, (B,), np.int32)
            status = self._alloc(like, (B,), np.int32)
        self._check(lib.lsx_rank_batch(self._ctx, ctypes.byref(plan), pA, B, mem, self._ptr(rank), self._ptr(status)))
        return RankResult(rank, status, plan)

    def solve_batch(self, A, b, a_abs_max=None, b_abs_max=None, max_rank=0, gen_cap=None, plan=None,
                    out=None) -> SolveResult:
        """Solution sets of A x = b (reference linalg.py:632-680, 913-999): status ST_INCONSISTENT where
        the reference returns ``NoSolution()``, else particular solution (free variables 0) and one
        generator per free column in ascending order."""
        A, pA, mem, like = self._prep_in(A, 3, "A")
        b, pb, mem_b, _ = self._prep_in(b, 2, "b")
        if mem != mem_b:
            raise ValueError("A and b must live in the same memory space")
        B, m, n = A.shape
        if b.shape[0] != B or b.shape[1] != m:
            raise ValueError("Matrix dimensions must match")
        if plan is None:
            plan = self.plan_solve(m, n, self._absmax(A) if a_abs_max is None else a_abs_max,
                                   self._absmax(b) if b_abs_max is None else b_abs_max, max_rank, gen_cap)
        L, G = plan.limbs, plan.gen_cap
        if out is not None:
            den, part, gens, piv, rank, status = out.den, out.particular, out.generators, out.pivot_col, out.rank, out.status
        else:
            den = self._alloc(like, (B, L), np.uint32)
            part = self._alloc(like, (B, n, L), np.uint32)
            gens = self._alloc(like, (B, n, G, L), np.uint32) if G > 0 else None
            piv = self._alloc(like, (B, plan.pivot_slots), np.int32)
            rank = self._alloc(like, (B,), np.int32)
            status = self._alloc(like, (B,), np.int32)
        self._check(lib.lsx_solve_batch(self._ctx, ctypes.byref(plan), pA, pb, B, mem, self._ptr(den), self._ptr(part),
                                        self._ptr(gens), self._ptr(piv), self._ptr(rank), self._ptr(status)))
        return SolveResult(den, part, gens, piv, rank, status, plan)

    # ---- lowest terms (reference results are reduced rationals: linalg.py:574, 698-699) ------
    def lowest_terms(self, num, den):
        """``num[b, ..., L] / den[b, L]`` -> ``(p, q)`` of the same shape as ``num`` with ``gcd(p, q) = 1`` and
        ``q > 0`` for every entry, computed on the device (multi-limb binary gcd + exact division).  ``num == 0``
        gives ``0 / 1``; a zero denominator (matrix flagged singular / inconsistent) gives ``0 / 0``."""
        if _is_torch(num):
            if not _is_torch(den) or num.is_cuda != den.is_cuda:
                raise ValueError("num and den must live in the same memory space")
            num, den = num.contiguous(), den.contiguous()
            mem = _lib.MEM_DEVICE if num.is_cuda else _lib.MEM_HOST
            like = num if num.is_cuda else None
            if num.is_cuda:
                import torch
                self.set_stream(torch.cuda.current_stream(num.device).cuda_stream or 1)
            else:
                num, den = num.numpy(), den.numpy()
                self.set_stream(None)
        else:
            num = np.ascontiguousarray(num)
            den = np.ascontiguousarray(den)
            mem, like = _lib.MEM_HOST, None
            self.set_stream(None)
        shape = tuple(num.shape)
        B, L = shape[0], shape[-1]
        if tuple(den.shape) != (B, L):
            raise ValueError("den must have shape [batch, limbs] matching num")
        count = 1
        for s_ in shape[1:-1]:
            count *= s_
        p = self._alloc(like, shape, np.uint32)
        q = self._alloc(like, shape, np.uint32)
        self._check(lib.lsx_lowest_terms(self._ctx, self._ptr(num), self._ptr(den), B, count, L, mem, self._ptr(p),
                                         self._ptr(q)))
        return p, q

    # ---- step trace of row_reduce (reference linalg.py:544-629) -----------------------------
    def rref_trace(self, A, bar_col, den=1):
        """Steps and exact intermediate matrices of ``row_reduce`` for ONE small rational matrix ``A / den``.

        ``A`` holds integer numerators, ``den`` their common denominator (1 for an integer matrix).
        Returns ``(frames, ops, pivots)``: ``ops`` is the list of ``(kind, a, b)`` the reference records (1 = S swap
        of rows a and b, 2 = N normalisation of row a, 3 / 4 = E elimination below / above the pivot of column a),
        ``frames[t]`` the matrix (rows of ``Fraction``) after ``ops[t]``, ``pivots`` the (row, column) list.  The
        device replays the reference's operation order modulo table primes (``lsx_rref_trace_q``); K follows from
        twice the Hadamard bound of the minors (every intermediate entry is a quotient of two of them) plus one check
        prime.  Primes whose log differs from the majority's (a prime dividing an intermediate value, a lead that is
        1 only modulo p, a prime dividing ``den``) are dropped and replaced by further table primes; the residues of
        the agreeing primes are lifted by CRT + rational reconstruction.
        """
        import math
        from collections import Counter
        from .convert import crt_basis, rational_reconstruct
        A = np.ascontiguousarray(np.asarray(A, dtype=np.int64))
        if A.ndim != 2:
            raise ValueError("A must have 2 dimensions")
        m, n = A.shape
        if A.size and np.abs(A).max() > 2**31 - 1:
            raise OverflowError("A has entries outside int32")
        den = int(den)
        if den < 1:
            raise ValueError("den must be a positive integer")
        amax = max(1, int(np.abs(A).max()) if A.size else 1, den)
        r = min(m, n)
        log2h = r * (0.5 * math.log2(r) + math.log2(amax)) if r else 0.0          # Hadamard bound of any minor
        max_ops = lib.lsx_rref_trace_max_ops(m, n, bar_col)
        if max_ops < 0:
            raise LsxError(max_ops, "rref_trace: bad shape")
        a32 = A.astype(np.int32)
        slots = min(m, bar_col)
        self.set_stream(None)
        # Forward-sweep entries are quotients of two minors, but the backward sweep combines such quotients, so the
        # size of an intermediate entry is not bounded by one Hadamard bound: start from 2 H^2 < M and double the
        # prime count until every entry reconstructs AND agrees with one more prime that took no part in the lift.
        K = int(math.ceil((2.0 * log2h + 2.0) / 30.99)) + 1
        spare = 0
        while True:
            n_req = K + 1 + spare                                  # K lift primes, one check prime, replacements
            if n_req > _lib.TABLE_PRIMES:
                raise RuntimeError("liblsx: the step trace ran out of table primes")
            primes_all = [int(p) for p in self.primes(n_req)]
            dres = np.array([den % p for p in primes_all], dtype=np.uint32)
            ops = np.zeros((n_req, max_ops, 4), dtype=np.int32)
            frames = np.zeros((n_req, max_ops, m, n), dtype=np.uint32)
            n_ops = np.zeros(n_req, dtype=np.int32)
            piv = np.zeros((n_req, slots), dtype=np.int32)
            self._check(lib.lsx_rref_trace_q(self._ctx, a32.ctypes.data, dres.ctypes.data if den != 1 else None, m, n,
                                             bar_col, n_req, _lib.MEM_HOST, ops.ctypes.data, frames.ctypes.data,
                                             n_ops.ctypes.data, piv.ctypes.data))
            # the rational log is the majority's: a bad prime deviates on its own, good primes all agree
            keys = [None if n_ops[k] < 0 else (int(n_ops[k]), ops[k, :max(0, int(n_ops[k]))].tobytes(), piv[k].tobytes())
                    for k in range(n_req)]
            votes = Counter(k for k in keys if k is not None)
            if not votes:
                spare = max(2, 2 * spare)
                continue
            best, cnt = votes.most_common(1)[0]
            good = [k for k in range(n_req) if keys[k] == best]
            if cnt < K + 1 or cnt * 2 <= len([k for k in keys if k is not None]):
                spare = max(2, 2 * spare)
                continue
            lift, chk = good[:K], good[K]
            T = int(n_ops[lift[0]])
            primes = [primes_all[k] for k in lift]
            M, coef = crt_basis(primes)
            bound = math.isqrt((M - 1) // 2)
            pc = primes_all[chk]
            out, ok = [], True
            for t in range(T):
                grid = []
                for i in range(m):
                    row = []
                    for j in range(n):
                        x = sum(int(frames[k, t, i, j]) * c for k, c in zip(lift, coef)) % M
                        v = rational_reconstruct(x, M, bound)
                        if v is None or v.denominator % pc == 0 or \
                                (v.numerator * pow(v.denominator, pc - 2, pc) - int(frames[chk, t, i, j])) % pc != 0:
                            ok = False
                            break
                        row.append(v)
                    if not ok:
                        break
                    grid.append(row)
                if not ok:
                    break
                out.append(grid)
            if ok:
                break
            if K > 256:
                raise RuntimeError("liblsx: rational reconstruction of the step trace did not converge")
            K *= 2
        p0 = piv[lift[0]]
        pivots = [(k, int(p0[k])) for k in range(slots) if p0[k] >= 0]
        return out, [tuple(int(x) for x in ops[lift[0], t, :3]) for t in range(T)], pivots

    # ---- one large determinant, by prime ----------------------------------------------------
    @staticmethod
    def det_large_prime_count(n, a_abs_max):
        k = ctypes.c_int()
        bits = ctypes.c_double()
        rc = lib.lsx_det_large_prime_count(n, a_abs_max, ctypes.byref(k), ctypes.byref(bits))
        if rc != _lib.OK:
            raise LsxError(rc, "det_large_prime_count")
        return k.value, bits.value

    def det_large_prime_count_for(self, A):
        """(primes, log2 bound) from Hadamard's bound with the actual row and column norms of A (rigorous and, for
        random entries, a few per cent below the worst-case count of ``det_large_prime_count``)."""
        A, pA, mem, _ = self._prep_in(A, 2, "A")
        n, n2 = A.shape
        if n != n2:
            raise ValueError("Determinant requires a square matrix")
        k = ctypes.c_int()
        bits = ctypes.c_double()
        self._check(lib.lsx_det_large_prime_count_for(self._ctx, pA, n, mem, ctypes.byref(k), ctypes.byref(bits)))
        return k.value, bits.value

    def det_large_residues(self, A, prime_begin, prime_count):
        """det(A) mod p for table primes [prime_begin, prime_begin + prime_count)."""
        A, pA, mem, like = self._prep_in(A, 2, "A")
        n, n2 = A.shape
        if n != n2:
            raise ValueError("Determinant requires a square matrix")
        res = self._alloc(like, (prime_count,), np.uint32)
        self._check(lib.lsx_det_large_residues(self._ctx, pA, n, prime_begin, prime_count, mem, self._ptr(res), None))
        return res

    def det_large(self, A):
        """Exact determinant of ONE large integer matrix (host array) as a Python int, through ``lsx_det_large``:
        primes sharded over this context's GPUs, residues all-gathered (NCCL), CRT on the first GPU."""
        A = np.ascontiguousarray(A)
        if A.dtype != np.int32:
            if A.size and (A.max() > 2**31 - 1 or A.min() < -(2**31) + 1):
                raise OverflowError("A has entries outside int32")
            A = A.astype(np.int32)
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("Determinant requires a square matrix")
        n = A.shape[0]
        k, bits = ctypes.c_int(), ctypes.c_double()
        amax = int(np.abs(A.astype(np.int64)).max()) if A.size else 0
        self._check_plan(lib.lsx_det_large_prime_count(n, amax, ctypes.byref(k), ctypes.byref(bits)), "det_large(n=%d)" % n)
        cap = int(bits.value + 2) // 32 + 2
        words = np.zeros(cap, dtype=np.uint32)
        limbs, primes = ctypes.c_int(), ctypes.c_int()
        self.set_stream(None)
        self._check(lib.lsx_det_large(self._ctx, A.ctypes.data, n, cap, words.ctypes.data, ctypes.byref(limbs),
                                      ctypes.byref(primes)))
        from .convert import limbs_to_ints
        return limbs_to_ints(words[: limbs.value].reshape(1, -1))[0], primes.value

    def rank_large(self, A):
        """Rank of ONE integer matrix of any size that fits device memory (``lsx_rank_large``: the batched rank keeps
        a tile per CTA, m <= 254).  Returns ``(rank, primes_used)``."""
        A, pA, mem, _ = self._prep_in(A, 2, "A")
        m, n = A.shape
        rank, used = ctypes.c_int(), ctypes.c_int()
        self._check(lib.lsx_rank_large(self._ctx, pA, m, n, mem, ctypes.byref(rank), ctypes.byref(used)))
        return rank.value, used.value

    def crt_signed(self, residues, limbs):
        """Residues for table primes [0, len) -> signed integer as `limbs` 32-bit words."""
        if _is_torch(residues):
            like, mem = residues, (_lib.MEM_DEVICE if residues.is_cuda else _lib.MEM_HOST)
            residues = residues.contiguous()
            count = residues.numel()
        else:
            residues = np.ascontiguousarray(residues, dtype=np.uint32)
            like, mem, count = None, _lib.MEM_HOST, residues.size
        out = self._alloc(like, (limbs,), np.uint32)
        self._check(lib.lsx_crt_signed(self._ctx, self._ptr(residues), count, limbs, mem, self._ptr(out)))
        return out


_default = None


def default_engine() -> Engine:
    """Process-wide engine on LSX_DEVICE / LOCAL_RANK / device 0 (created on first use)."""
    global _default
    if _default is None:
        _default = Engine()
    return _default
