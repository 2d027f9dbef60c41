#!/usr/bin/env python
"""Measurement-only probe of the dense int8 tensor peak of this GPU (VERDICT r1, next #4c).

A LIBRARY int8 GEMM (torch._int_mm -> cuBLASLt, s8 x s8 -> s32) at 8192^3, timed like MEASURED_PEAKS.json times bf16:
best of 10 (burst) and back to back for 4 s (sustained, under the power cap).  The result is written to
profiles/int8_peak.json and is the denominator bench.py uses for the tensor roofline of config 5.  The library call
lives only here: nothing in liblsx links or calls it.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "int8_peak.json")
    dev = torch.device("cuda", 0)
    n = 8192
    a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
    for _ in range(5):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    ops = 2.0 * n ** 3
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    reps = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(50):
            torch._int_mm(a, b)
        reps += 50
        torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    sustained_ms = e0.elapsed_time(e1) / reps
    # the bf16 figure the same way, for the ratio
    x = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    y = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    for _ in range(5):
        x @ y
    torch.cuda.synchronize()
    bbest = 1e9
    for _ in range(10):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        x @ y
        f1.record()
        f1.synchronize()
        bbest = min(bbest, f0.elapsed_time(f1))
    res = {"int8_tops_burst": ops / (best * 1e-3) / 1e12, "int8_tops_sustained": ops / (sustained_ms * 1e-3) / 1e12,
           "bf16_tflops_burst_same_probe": ops / (bbest * 1e-3) / 1e12,
           "how": "torch._int_mm (cuBLASLt s8 x s8 -> s32) 8192^3: best of 10 (burst), back to back for 4 s (sustained)",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
