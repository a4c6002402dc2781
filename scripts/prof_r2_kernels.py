"""One launch of each round-2 hot kernel at benchmark shapes, for `ncu --set full` (profiles/ncu_hot_kernels_r2.md):

    ncu --set full --clock-control none --import-source on -k regex:'gemm_mxf8|sim_topk|attn_fwd1|sum_ln|rescore' \
        -c 14 -o gpurun_out/ncu_r2 python scripts/prof_r2_kernels.py

Shapes: cross-encoder layer at M ~ 90k packed tokens (1280 pairs x ~70), similarity scan over 2M x 384 (bf16 and e4m3)."""
import sys

import torch

sys.path.insert(0, ".")
from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel  # noqa: E402
from infomesh_b200.ops import nn as N  # noqa: E402
from infomesh_b200.ops import search as S  # noqa: E402

dev = "cuda"
# ---- one cross-encoder layer (2 layers so the MX chain between layers is exercised), mxfp8
from dataclasses import replace  # noqa: E402

cfg = replace(BGE_RERANKER_BASE, layers=2)
m = BertModel(cfg, device=dev, seed=1)
g = torch.Generator().manual_seed(0)
ids = torch.randint(5, 200000, (1280, 128), generator=g, dtype=torch.int32).cuda()
lens = torch.randint(66, 76, (1280,), generator=g, dtype=torch.int32).cuda()
m.score_packed(ids, lens, precision="mxfp8")
torch.cuda.synchronize()
# ---- similarity scan: 2M x 384, 64 queries, bf16 then e4m3 (+ re-score)
q = torch.nn.functional.normalize(torch.randn(64, 384, device=dev), dim=1).bfloat16()
d = torch.nn.functional.normalize(torch.randn(2_000_000, 384, device=dev), dim=1).bfloat16()
S.sim_topk(q, d, 20)
q8, qs = N.quantize_rows_e4m3(q)
d8, ds = N.quantize_rows_e4m3(d)
S.sim_topk_f8(q8, qs, d8, ds, 20, rescore=(q, d), k_fetch=32)
torch.cuda.synchronize()
print("done")
