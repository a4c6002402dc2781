#!/usr/bin/env python
"""GPU check of the MXFP8 block-scaled GEMM (csrc/gemm/gemm_mxf8.cu): staged cases that isolate the operand path, the
SFA / SFB TMEM layouts, the per-MMA scale-id selection, the epilogues, and a throughput measurement."""
import sys
import time

import torch

sys.path.insert(0, ".")
from infomesh_b200.ops import mx as MX  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def rand_e4m3(r, k, amp=1.0):
    x = (torch.randn(r, k, device=dev) * amp).clamp(-448, 448)
    return x.to(torch.float8_e4m3fn).view(torch.uint8)


def run_case(name, M, N, K, ea, eb, mode=0, **kw):
    aq, bq = rand_e4m3(M, K), rand_e4m3(N, K)
    a = MX.MxTensor(aq, MX.pack_sfa(ea))
    w = MX.MxWeight(bq, MX.pack_sfb(eb), eb)
    ref = MX.linear_mx_ref(a, w, **kw)
    out = MX.linear_mx(a, w, mode=mode, **kw)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    mag = ref.abs().max().item()
    rel = err / max(mag, 1e-9)
    ok = rel < 2e-2
    print(f"[{'ok' if ok else 'FAIL'}] {name:42s} M={M} N={N} K={K} max|err|={err:.4g} max|ref|={mag:.4g} rel={rel:.3g}", flush=True)
    if not ok:
        bad = ((out.float() - ref).abs() > 0.02 * mag)
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()[:8].tolist()
        print("      bad rows", rows, "bad cols", cols, "frac", bad.float().mean().item())
    return ok


def ones(r, s):
    return torch.full((r, s), 127, dtype=torch.uint8, device=dev)


M, N, K = 256, 384, 256
S = K // 32
allok = True
allok &= run_case("unit scales", M, N, K, ones(M, S), ones(N, S))
ea = (120 + (torch.arange(M, device=dev) % 13)).to(torch.uint8)[:, None].expand(M, S).contiguous()
allok &= run_case("SFA varies by row", M, N, K, ea, ones(N, S))
ek = (122 + (torch.arange(S, device=dev) % 8)).to(torch.uint8)[None, :].expand(M, S).contiguous()
allok &= run_case("SFA varies by k-subblock (sf id)", M, N, K, ek, ones(N, S))
eb = (121 + (torch.arange(N, device=dev) % 11)).to(torch.uint8)[:, None].expand(N, S).contiguous()
allok &= run_case("SFB varies by row (even+odd tiles)", M, N, K, ones(M, S), eb)
ekb = (123 + (torch.arange(S, device=dev) % 8)).to(torch.uint8)[None, :].expand(N, S).contiguous()
allok &= run_case("SFB varies by k-subblock", M, N, K, ones(M, S), ekb)
g = torch.Generator(device=dev).manual_seed(1)
ra = torch.randint(118, 132, (M, S), device=dev, generator=g).to(torch.uint8)
rb = torch.randint(118, 132, (N, S), device=dev, generator=g).to(torch.uint8)
allok &= run_case("random scales", M, N, K, ra, rb)
bias = torch.randn(N, device=dev)
allok &= run_case("random scales + bias", M, N, K, ra, rb, bias=bias)
allok &= run_case("random scales + bias + gelu", M, N, K, ra - 6, rb - 6, bias=bias * 0.1, act="gelu")
res = torch.randn(M, N, device=dev).to(torch.bfloat16)
allok &= run_case("random scales + bias + residual", M, N, K, ra, rb, bias=bias, residual=res)
# ragged shapes: M not a multiple of 128, N = 768 / 2304 / 3072 (the reranker's)
for (m, n, k) in [(1000, 768, 768), (777, 2304, 768), (300, 3072, 768), (520, 768, 3072)]:
    s = k // 32
    allok &= run_case("reranker shape", m, n, k, torch.randint(120, 130, (m, s), device=dev, generator=g).to(torch.uint8),
                      torch.randint(120, 130, (n, s), device=dev, generator=g).to(torch.uint8))

# A-resident kernel (forced): many m blocks so CTAs run several units back to back (A slot reuse, barrier phases)
for (m, n, k) in [(1000, 768, 768), (40000, 2304, 768), (33000, 768, 768), (5000, 3072, 768), (3000, 1152, 384), (700, 384, 256)]:
    s_ = k // 32
    allok &= run_case("A-resident kernel", m, n, k, torch.randint(120, 130, (m, s_), device=dev, generator=g).to(torch.uint8),
                      torch.randint(120, 130, (n, s_), device=dev, generator=g).to(torch.uint8), mode=2,
                      bias=torch.randn(n, device=dev), residual=torch.randn(m, n, device=dev).to(torch.bfloat16))

# ---- real quantised data + MX output epilogue ----
M, N, K = 640, 3072, 768
x = torch.randn(M, K, device=dev)
wt = torch.randn(N, K, device=dev) * 0.05
a = MX.quantize_act_ref(x)
w = MX.quantize_weight(wt)
bias = torch.randn(N, device=dev) * 0.1
ref = MX.linear_mx_ref(a, w, bias=bias, act="gelu")
out = MX.linear_mx(a, w, bias=bias, act="gelu", out_mx=True)
torch.cuda.synchronize()
got = out.float()
qerr = (got - ref).abs().max().item()
e_ref = MX.quantize_ref(ref)[1]
e_got = MX.unpack_sfa(out.sf, M)
e_diff = (e_ref.int() - e_got.int()).abs()
print(f"[{'ok' if qerr < 0.07 * ref.abs().max().item() else 'FAIL'}] mx-out epilogue: max|deq(out) - ref| = {qerr:.4g} (max|ref| {ref.abs().max().item():.4g}); "
      f"scale exponents differing from the oracle: {(e_diff > 0).float().mean().item():.4%} (max diff {e_diff.max().item()})", flush=True)
exact = (x @ wt.t() + bias)
exact = torch.nn.functional.gelu(exact)
print(f"      end-to-end MXFP8 error vs fp32 of the unquantised inputs: rel fro = {((got - exact).norm() / exact.norm()).item():.4g}")

# ---- chained: mx-out feeds the next GEMM ----
w2 = MX.quantize_weight(torch.randn(768, N, device=dev) * 0.03)
y = MX.linear_mx(out, w2)
yref = MX.linear_mx_ref(out, w2)
torch.cuda.synchronize()
print(f"[{'ok' if (y.float() - yref).abs().max().item() < 0.02 * yref.abs().max().item() else 'FAIL'}] chained GEMM on the epilogue-quantised activations: "
      f"max|err| {(y.float() - yref).abs().max().item():.4g} / {yref.abs().max().item():.4g}")

# ---- throughput (burst, L2-cold between shapes; same M as the reranker's packed batch) ----
from infomesh_b200.ops import gemm as G  # noqa: E402

for (m, n, k, kw) in [(90112, 2304, 768, {}), (90112, 768, 768, {}), (90112, 3072, 768, dict(act="gelu", out_mx=True)),
                      (90112, 768, 3072, {}), (11264, 2304, 768, {}), (11264, 3072, 768, dict(act="gelu", out_mx=True))]:
    a = MX.MxTensor(rand_e4m3(m, k), MX.pack_sfa(torch.randint(120, 130, (m, k // 32), device=dev).to(torch.uint8)))
    w = MX.quantize_weight(torch.randn(n, k, device=dev) * 0.05)
    bias = torch.randn(n, device=dev)
    xb = torch.randn(m, k, device=dev).to(torch.bfloat16)
    wb = torch.randn(n, k, device=dev).to(torch.bfloat16)
    outb = None
    for label, fn in (("mxf8 ring", lambda: MX.linear_mx(a, w, bias=bias, mode=1, **kw)),
                      ("mxf8 A-resident" if k <= 768 else "mxf8 auto", lambda: MX.linear_mx(a, w, bias=bias, mode=0, **kw)),
                      ("bf16", lambda: G.linear(xb, wb, bias, act=kw.get("act")))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  {label} M={m} N={n} K={k} {kw}: {ms * 1e3:.1f} us  {2 * m * n * k / ms / 1e9:.0f} TFLOP/s", flush=True)
print("ALL OK" if allok else "SOME FAILED")
