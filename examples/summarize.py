"""Summarise with the in-process T5 backend (random weights offline: the output is not meaningful text, the call path
is what is shown) and run the verification pipeline on a hand-written summary."""
import asyncio

import torch

from infomesh_b200.summarizer import verify
from infomesh_b200.summarizer.engine import SummarizationEngine, create_backend

SRC = ("The Blackwell B200 GPU has 148 SMs. It carries 180 GB of HBM3e memory on two dies. NVLink 5 gives every GPU 900 GB per "
       "second in each direction. The tensor cores accumulate in TMEM.")

report = verify.verify_summary("https://example.org/b200", "hash", SRC, "B200 has 148 SMs, 180 GB of HBM3e and accumulates in TMEM.",
                               peer_summaries=["The B200 has 148 SMs and 180 GB HBM3e memory"])
print("verification:", report.level.value, report.quality_score, report.detail)

if torch.cuda.is_available():
    engine = SummarizationEngine(create_backend("b200", "t5-small"))
    res = asyncio.run(engine.summarize("https://example.org/b200", "B200", SRC, max_tokens=24))
    print(f"{res.runtime.value}/{res.model}: {res.elapsed_ms:.0f} ms, {res.token_count} tokens")
