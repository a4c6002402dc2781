"""LayerNorm + MX quantiser timing at the cross-encoder's shape: INFOMESH_B200_LN_STREAM=0/1 A/B (CUDA events, 50 iterations)."""
import sys

import torch

sys.path.insert(0, ".")
from infomesh_b200.ops import mx as MX  # noqa: E402
from infomesh_b200.ops import nn as N  # noqa: E402

M, H = 90112, 768
x = torch.randn(M, H, device="cuda").bfloat16()
g = torch.randn(H, device="cuda").float()
b = torch.randn(H, device="cuda").float()
out = torch.empty_like(x)
act = MX.alloc_act(M, H, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    return N.layernorm_mx(x, g, b, 1e-5, act, out=out)


run()
ref = torch.nn.functional.layer_norm(x.float(), (H,), g, b, 1e-5)
deq = MX.dequantize(act.q, MX.unpack_sfa(act.sf, M))
print("max err bf16 out", (out.float() - ref).abs().max().item(), " mx deq rel err", ((deq - ref).abs().max() / ref.abs().max()).item())
ts = []
for _ in range(50):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
mb = (M * H * 2 * 2 + M * H + M * H // 32) / 1e6
print(f"layernorm_mx [{M}x{H}]: median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us, {mb / ts[len(ts) // 2] / 1e0:.0f} MB/ms = {mb / ts[len(ts) // 2] * 1e3 / 1e6:.2f} TB/s")
