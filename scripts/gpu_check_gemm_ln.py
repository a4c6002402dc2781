"""Fused GEMM + residual + LayerNorm + MXFP8-quantise epilogue (csrc/gemm/gemm_mxf8.cu, cluster variant): numerics vs the
fp32 oracle and timing vs the two-kernel path (GEMM, then sum_ln_mx) on the cross-encoder's shapes.
Run on one B200: ``python scripts/gpu_check_gemm_ln.py``."""
import sys

import torch

from infomesh_b200 import _native
from infomesh_b200.ops import mx as MX
from infomesh_b200.ops import nn as N

DEV = "cuda"


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    _native.require()
    torch.manual_seed(0)
    ok = True
    for m, n, k in [(1000, 768, 768), (90112, 768, 768), (90112, 768, 3072), (5120, 768, 3072), (4096, 384, 1536)]:
        a = MX.quantize_act_ref(torch.randn(m, k, device=DEV))
        w = MX.quantize_weight(torch.randn(n, k, device=DEV) * 0.05)
        bias = torch.randn(n, device=DEV) * 0.3
        res = (torch.randn(m, n, device=DEV) * 2).bfloat16()
        g = torch.rand(n, device=DEV) + 0.5
        b = torch.randn(n, device=DEV) * 0.1
        mxo = MX.alloc_act(m, n, DEV, init=True)
        out = torch.empty((m, n), device=DEV, dtype=torch.bfloat16)
        MX.linear_mx_ln(a, w, bias, res, g, b, 1e-5, mxo, out=out)
        torch.cuda.synchronize()
        rows = min(m, 4096)
        ref = MX.linear_mx_ln_ref(MX.MxTensor(a.q[:rows], a.sf[:(rows + 127) // 128]), w, bias, res[:rows], g, b, 1e-5)
        err = (out[:rows].float() - ref).abs().max().item()
        y = torch.empty((m, n), device=DEV, dtype=torch.bfloat16)
        mx2 = MX.alloc_act(m, n, DEV)
        o2 = torch.empty((m, n), device=DEV, dtype=torch.bfloat16)

        def two():
            MX.linear_mx(a, w, bias, res, out=y)
            N.layernorm_mx(y, g, b, 1e-5, mx2, out=o2)

        t_f = timed(lambda: MX.linear_mx_ln(a, w, bias, res, g, b, 1e-5, mxo, out=out))
        t_g = timed(lambda: MX.linear_mx(a, w, bias, res, out=y))
        t_2 = timed(two)
        print(f"[{m} x {n} x {k}] max |err| vs fp32 oracle {err:.4f} | fused {t_f:.1f} us, GEMM alone {t_g:.1f} us, GEMM + LayerNorm {t_2:.1f} us "
              f"-> {t_2 / t_f:.2f}x", flush=True)
        ok &= err < 0.03
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
