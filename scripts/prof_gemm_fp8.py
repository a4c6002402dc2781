"""bf16 vs e4m3 GEMM on the reranker shapes (same kernel, kind::f16 vs kind::f8f6f4), CUDA events, back to back."""
import torch

from infomesh_b200.ops import gemm as G

dev = torch.device("cuda:0")
M = 163840
for (n, k) in ((3072, 768), (2304, 768), (768, 3072), (768, 768)):
    a = (torch.randn(M, k, device=dev) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
    b = torch.randn(n, device=dev)
    out = torch.empty((M, n), device=dev, dtype=torch.bfloat16)
    a8, rs = G.quantize_rows_fp8(a)
    w8, ws = G.quantize_weight_fp8(w)

    def t(fn, it=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / it

    us16 = t(lambda: G.linear(a, w, bias=b, out=out))
    us8 = t(lambda: G.linear(a8, w8, bias=b, out=out, alpha=ws, row_scale=rs))
    usq = t(lambda: G.quantize_rows_fp8(a))
    fl = 2 * M * n * k
    print(f"M={M} N={n} K={k}: bf16 {us16:7.1f} us {fl / us16 / 1e6:6.0f} TF/s | e4m3 {us8:7.1f} us {fl / us8 / 1e6:6.0f} TF/s | "
          f"row quantise {usq:6.1f} us ({M * k * 3 / usq / 1e6:.2f} TB/s)")
