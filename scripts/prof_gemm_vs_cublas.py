"""Our tcgen05 GEMM vs cuBLASLt (torch F.linear / addmm) on the reranker shapes, back to back, with SM clocks sampled
from nvidia-smi while each loop runs (the GEMMs run power-capped, so TF/s must be read next to the clock)."""
import subprocess
import sys
import threading
import time

import torch

from infomesh_b200.ops.gemm import linear

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 163840


class Clocks:
    def __init__(self):
        self.vals, self.stop = [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.vals.append((float(out[0]), float(out[1])))
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join()

    def summary(self):
        if not self.vals:
            return "n/a"
        v = sorted(x[0] for x in self.vals)
        p = max(x[1] for x in self.vals)
        return f"{v[len(v) // 2]:.0f} MHz, {p:.0f} W"


def loop(fn, secs=1.5):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks() as c:
        t0 = time.time()
        e0.record()
        while time.time() - t0 < secs:
            for _ in range(20):
                fn()
            n += 20
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n, c.summary()


for (n, k) in ((3072, 768), (2304, 768), (768, 3072), (768, 768)):
    a = (torch.randn(M, k, device=dev) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
    b = torch.randn(n, device=dev)
    bb = b.bfloat16()
    out = torch.empty((M, n), device=dev, dtype=torch.bfloat16)
    us, ck = loop(lambda: linear(a, w, bias=b, out=out))
    print(f"M={M} N={n} K={k}  ours bias   {us:8.1f} us {2 * M * n * k / us / 1e6:7.0f} TF/s  [{ck}]")
    us, ck = loop(lambda: torch.nn.functional.linear(a, w, bb))
    print(f"M={M} N={n} K={k}  cuBLASLt    {us:8.1f} us {2 * M * n * k / us / 1e6:7.0f} TF/s  [{ck}]")
    us, ck = loop(lambda: torch.matmul(a, w.t()))
    print(f"M={M} N={n} K={k}  cuBLAS nobias {us:6.1f} us {2 * M * n * k / us / 1e6:7.0f} TF/s  [{ck}]")
