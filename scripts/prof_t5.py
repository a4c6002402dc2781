"""t5-small greedy decode (B=16, 256 input tokens, 32 new tokens): ncu / timing target for the RAG summariser path."""
import sys

import torch

from infomesh_b200.models.t5 import T5_SMALL, T5Model

dev = torch.device("cuda:0")
t5 = T5Model(T5_SMALL, device=dev, seed=3)
B, S, new = 16, 256, 32
ids = torch.randint(5, 30000, (B, S), dtype=torch.int32, device=dev)
lens = torch.full((B,), S, dtype=torch.int32, device=dev)
for _ in range(2):
    t5.generate(ids, lens, max_new_tokens=new, check_every=new)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "profile":
    torch.cuda.profiler.start()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
o = t5.generate(ids, lens, max_new_tokens=new, check_every=new)
b.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"generate: {a.elapsed_time(b):.2f} ms for {B} x {o.shape[1]} tokens")
