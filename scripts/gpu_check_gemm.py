"""GPU sanity + timing for the tcgen05 GEMM (run under gpurun)."""
import json, sys, time
import torch
from infomesh_b200.ops.gemm import linear, linear_ref

torch.manual_seed(0)
dev = "cuda"
res = []
def check(m, n, k, bias=False, act=None, resid=False, bn=0, fp32=False):
    a = (torch.randn(m, k, device=dev) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=dev) * 0.5).bfloat16()
    b = torch.randn(n, device=dev) if bias else None
    r = torch.randn(m, n, device=dev).bfloat16() if resid else None
    out = linear(a, w, b, r, act, bn=bn, out_dtype=torch.float32 if fp32 else torch.bfloat16)
    torch.cuda.synchronize()
    ref = linear_ref(a, w, b, r, act)
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err <= 0.02 * scale + 0.05
    res.append(dict(m=m, n=n, k=k, bias=bias, act=act, resid=resid, bn=bn, fp32=fp32, err=err, scale=scale, ok=ok))
    print(res[-1], flush=True)
    return ok

allok = True
for args in [(128, 128, 64), (128, 128, 384), (256, 384, 384), (1000, 1152, 384), (4096, 1536, 384, True, "gelu"),
             (4096, 384, 1536, True, None, True), (8192, 2304, 768, True), (8192, 768, 3072, True, None, True, 256),
             (300, 200, 72), (64, 384, 384, True, "relu", False, 0, True), (16384, 3072, 768, True, "gelu", False, 256)]:
    allok &= check(*args)

def bench(m, n, k, bn=0, iters=20):
    a = torch.randn(m, k, device=dev).bfloat16(); w = torch.randn(n, k, device=dev).bfloat16()
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for _ in range(3): linear(a, w, out=out, bn=bn)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): linear(a, w, out=out, bn=bn)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    for _ in range(3): torch.matmul(a, w.t(), out=out)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): torch.matmul(a, w.t(), out=out)
    e.record(); torch.cuda.synchronize()
    ms_ref = s.elapsed_time(e) / iters
    tf = 2.0 * m * n * k / ms / 1e9
    print(dict(bench=(m, n, k), bn=bn, ms=ms, tflops=tf, cublas_ms=ms_ref, cublas_tflops=2.0 * m * n * k / ms_ref / 1e9), flush=True)

if allok:
    for shp in [(163840, 2304, 768, 256), (163840, 2304, 768, 128), (163840, 768, 3072, 256), (163840, 3072, 768, 256),
                (163840, 768, 768, 128), (8192, 8192, 8192, 256), (32768, 1152, 384, 128), (32768, 1536, 384, 128)]:
        bench(*shp)
print("GEMM_ALL_OK" if allok else "GEMM_FAIL")
sys.exit(0 if allok else 1)
