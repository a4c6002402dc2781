#!/usr/bin/env python
"""Headline benchmark: hybrid BM25 + dense top-10 with cross-encoder rerank over a 10M-doc synthetic index.

Metric (BASELINE.json): queries/sec (+ p50 latency) for hybrid top-10 over a 10M-document index at 1/2/4/8 B200.
One "step" = one batch of 64 queries through the whole pipeline:

    encode (bge-small-en) -> sharded dense top-k (sim_topk) + BM25 AND/score/top-k -> exchange/merge
    -> RRF fuse -> 20 (query, passage) pairs per query -> bge-reranker-base -> top-10

The index (10M docs total, document-partitioned over the ranks: strong scaling) is synthetic and the model
weights are random-initialised (no network for datasets / checkpoints).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

``--impl reference`` runs the unmodified reference from baseline/_ref (installed by baseline/install_ref.py) through
its own ``search_hybrid`` on the same synthetic corpus generator -- see baseline/reference_arm.py for what differs.
``--impl torch`` is the PyTorch (cuBLAS / SDPA / NCCL) build of the same pipeline, used for A/B.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time


def _reference_arm(args) -> int:
    """Run the unmodified reference (baseline/_ref) through its own search_hybrid(); see baseline/reference_arm.py."""
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "baseline"))
    import reference_arm

    return reference_arm.run(args, ClockSampler=ClockSampler)


class ClockSampler:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines: list[str] = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None
            return

        def pump():
            assert self.proc is not None and self.proc.stdout is not None
            for ln in self.proc.stdout:
                self.lines.append(ln.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples in the upper half of the observed range
        load = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fused", choices=["fused", "torch", "reference"])
    ap.add_argument("--docs", type=int, default=10_000_000, help="total documents in the index (all ranks)")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--pair-seq", type=int, default=128)
    ap.add_argument("--no-rerank", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-varlen", action="store_true", help="run the cross-encoder on padded [pairs, seq_len] batches")
    ap.add_argument("--no-padded-arm", action="store_true", help="skip the extra padded-cross-encoder measurement")
    ap.add_argument("--precision", choices=["bf16", "fp8"], default="bf16",
                    help="fp8: cross-encoder projections as e4m3 GEMMs (reported as dtype fp8; NOT the headline config)")
    ap.add_argument("--rerank-chunks", type=int, default=1, help="cross-encoder sub-batches per step (L2 residency)")
    ap.add_argument("--pipeline", choices=["auto", "on", "off"], default="auto",
                    help="keep two batches in flight (retrieval of batch i+1 overlaps the cross-encoder of batch i); "
                         "auto = on when a rank's cross-encoder share is small enough to be latency-bound")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU candidate exchange: fused peer-memory kernels or NCCL all-gathers (baseline)")
    ap.add_argument("--latency-b1", action="store_true", help="also measure batch-1 p50 latency")
    ap.add_argument("--ref-docs", type=int, default=1_000_000,
                    help="reference arm: documents to index (its per-row INSERT+COMMIT build is time-boxed, see below)")
    ap.add_argument("--ref-build-budget-s", type=float, default=150.0,
                    help="reference arm: stop indexing after this many seconds and report the size reached")
    ap.add_argument("--query-mix", choices=["rare", "common"], default="rare",
                    help="rare: 2-3 selective terms per query; common: every other query also carries one of its "
                         "document's most frequent terms (long posting lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        return _reference_arm(args)

    import torch

    from infomesh_b200 import _native
    from infomesh_b200.engine.hybrid import HybridConfig, HybridEngine
    from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries
    from infomesh_b200.parallel import dist as D

    if not torch.cuda.is_available():
        print(json.dumps({"error": "bench.py needs a CUDA device"}))
        return 1
    _native.require()
    ctx = D.init()
    world, rank, dev = ctx.world, ctx.rank, ctx.device
    if world != args.gpus and rank == 0:
        print(f"[bench] note: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    assert args.warmup >= 3 or args.steps <= 2, "timing rules: at least 3 warm-up steps"

    n_global = args.docs
    per = (n_global + world - 1) // world
    base = rank * per
    n_local = max(0, min(per, n_global - base))
    scfg = SynthConfig(n_docs=n_local, n_docs_global=n_global, doc_base=base)
    t0 = time.time()
    shard = SynthShard(scfg, device=dev, df_allreduce=(D.all_reduce_sum_ if world > 1 else None),
                       passages="global" if world > 1 else "local")
    torch.cuda.synchronize()
    build_s = time.time() - t0

    hcfg = HybridConfig(nq=args.batch, pair_seq=args.pair_seq, rerank=not args.no_rerank, backend=args.impl,
                        use_graph=not args.no_graph, exchange=args.exchange, varlen=not args.no_varlen, rerank_chunks=args.rerank_chunks, precision=args.precision)
    eng = HybridEngine(shard, hcfg, docs_per_shard=(n_global if world > 1 else n_local))

    # ---- query batches on pinned host memory (distinct per step so nothing is cached between iterations) ----
    n_batches = args.steps + args.warmup
    qcfg = SynthConfig(n_docs=n_global, n_docs_global=n_global)
    q_terms, q_tok, q_len, _ = make_queries(qcfg, n_batches * args.batch, max_terms=hcfg.max_terms,
                                            max_q_tokens=hcfg.max_q_tokens, device=dev)
    enc_ids = torch.zeros((n_batches * args.batch, hcfg.enc_seq), dtype=torch.int32)
    span = eng.encoder.cfg.vocab_size - 1000
    qt = (1000 + (q_tok.long() * 40503 % span)).to(torch.int32)
    enc_ids[:, 0] = 101
    L = min(hcfg.enc_seq - 2, q_tok.shape[1])
    enc_ids[:, 1:1 + L] = qt[:, :L]
    col = torch.arange(hcfg.enc_seq)[None]
    ql = q_len.clamp(max=L).long()[:, None]
    enc_ids = torch.where(col == ql + 1, torch.tensor(102, dtype=torch.int32), enc_ids)
    enc_ids = torch.where(col > ql + 1, torch.tensor(0, dtype=torch.int32), enc_ids)
    enc_len = (q_len.clamp(max=L) + 2).to(torch.int32)

    def pin(t):
        return t.contiguous().pin_memory()

    B = args.batch
    batches = [tuple(pin(x[i * B:(i + 1) * B]) for x in (enc_ids, enc_len, q_tok, q_len, q_terms))
               for i in range(n_batches)]
    h2d_bytes = sum(x.numel() * x.element_size() for x in batches[0])
    out_s_host = torch.empty((B, hcfg.k_out), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((B, hcfg.k_out), dtype=torch.int64).pin_memory()
    d2h_bytes = out_s_host.numel() * 4 + out_i_host.numel() * 8
    dev_batches = [tuple(x.to(dev) for x in b) for b in batches]

    def timed(run_step, n_warm, n_steps, offset=0):
        for i in range(n_warm):
            run_step(i)
        D.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
        evs[0].record()
        for i in range(n_steps):
            run_step(n_warm + i)
            evs[i + 1].record()
        torch.cuda.synchronize()
        D.barrier()
        per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)]
        total = evs[0].elapsed_time(evs[-1])
        return D.all_reduce_max(total), per_step

    # ---- (a) device-timed pipeline, inputs already resident (kernel-level number) ----
    def step_dev(i):
        eng.load_inputs(*dev_batches[i % n_batches])
        eng.run()

    # ---- (b) end to end through the public API: pinned H2D every step + D2H of the result ----
    def step_e2e(i):
        eng.search_batch(*batches[i % n_batches], out_scores_host=out_s_host, out_ids_host=out_i_host)

    # ---- pipelined serving (the engine's throughput mode): retrieval of batch i+1 overlaps the cross-encoder of
    #      batch i on a second stream; every batch still goes through every kernel inside the timed region ----
    def timed_pipe(e, src, host_out, n_warm, n_steps):
        for i in range(n_warm):
            e.submit(*src[i % n_batches], out_scores_host=host_out[0], out_ids_host=host_out[1])
        e.drain()
        D.barrier()
        torch.cuda.synchronize()
        sa, _sb = e.pipeline_streams()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
        evs[0].record(sa)
        for i in range(n_steps):
            e.submit(*src[(n_warm + i) % n_batches], out_scores_host=host_out[0], out_ids_host=host_out[1],
                     done_event=evs[i + 1])
        e.drain()
        torch.cuda.synchronize()
        D.barrier()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)]
        return D.all_reduce_max(evs[0].elapsed_time(evs[-1])), per

    sampler = ClockSampler(ctx.local_rank)
    launches = eng.launches_per_step()
    # measured: +13 % at <= 160 pairs/rank (latency-bound kernels interleave), -2 % at >= 640 pairs/rank (persistent GEMMs own the SMs)
    want_pipe = args.pipeline == "on" or (args.pipeline == "auto" and eng.nq_local * hcfg.n_rerank <= 400)
    pipelined = want_pipe and eng.pipeline_supported()
    unpipelined = None
    if pipelined:
        # per-batch latency: one batch at a time through the single-graph path
        lat_ms, lat_steps = timed(step_dev, args.warmup, args.steps)
        unpipelined = {"value": round(args.batch * args.steps / (lat_ms / 1e3), 2), "unit": "queries/s",
                       "ms_per_step": round(lat_ms / args.steps, 4), "p50_step_ms": round(statistics.median(lat_steps), 4),
                       "note": "one batch in flight (per-batch latency); `value` keeps two batches in flight"}
        sampler.start()
        total_ms, per_step = timed_pipe(eng, dev_batches, (None, None), args.warmup, args.steps)
        clocks = sampler.stop()
        e2e_ms, e2e_steps = timed_pipe(eng, batches, (out_s_host, out_i_host), args.warmup, args.steps)
    else:
        sampler.start()
        total_ms, per_step = timed(step_dev, args.warmup, args.steps)
        clocks = sampler.stop()
        e2e_ms, e2e_steps = timed(step_e2e, args.warmup, args.steps)
    # sanity: results are real document ids
    ids_ok = bool((out_i_host[:, 0] >= 0).float().mean() > 0.5)

    # ---- (c) the same pipeline with the cross-encoder forced onto padded [pairs, seq_len] batches, for transparency:
    #          `value` is the product (unpadded); this is what the step costs when every pair is computed at seq_len
    padded = None
    if hcfg.rerank and hcfg.varlen and args.impl == "fused" and not args.no_padded_arm:
        from dataclasses import replace as _replace

        eng_p = HybridEngine(shard, _replace(hcfg, varlen=False), encoder=eng.encoder, reranker=eng.reranker,
                             docs_per_shard=(n_global if world > 1 else n_local))

        def step_pad(i):
            eng_p.load_inputs(*dev_batches[i % n_batches])
            eng_p.run()

        if pipelined:
            pad_ms, _ = timed_pipe(eng_p, dev_batches, (None, None), args.warmup, args.steps)
        else:
            pad_ms, _ = timed(step_pad, args.warmup, args.steps)
        padded = {"value": round(B * args.steps / (pad_ms / 1e3), 2), "unit": "queries/s",
                  "ms_per_step": round(pad_ms / args.steps, 4),
                  "note": f"cross-encoder on padded [{B * hcfg.n_rerank // world} x {args.pair_seq}] batches per rank"}
        del eng_p

    lat_b1 = None
    if args.latency_b1:
        cfg1 = HybridConfig(nq=world, pair_seq=args.pair_seq, rerank=not args.no_rerank, backend=args.impl,
                            use_graph=not args.no_graph, exchange=args.exchange)
        eng1 = HybridEngine(shard, cfg1, encoder=eng.encoder, reranker=eng.reranker,
                            docs_per_shard=(n_global if world > 1 else n_local))
        b1 = [tuple(x[:world] for x in b) for b in dev_batches]

        def step_b1(i):
            eng1.load_inputs(*b1[i % n_batches])
            eng1.run()

        _, s1 = timed(step_b1, args.warmup, max(args.steps, 10))
        lat_b1 = statistics.median(s1)

    pad_note = "padded [pairs, seq_len] cross-encoder batches"
    if hcfg.rerank and getattr(eng, "last_pair_lens", None) is not None:
        mean_len = float(eng.last_pair_lens.float().mean().item())
        if hcfg.varlen and args.impl == "fused":
            pad_note = (f"cross-encoder runs unpadded (varlen): pairs are <= {args.pair_seq} tokens, mean "
                        f"{mean_len:.1f} in this synthetic batch; padding tokens are not computed; the last encoder "
                        "layer evaluates its query / FFN rows only for the <s> token the classifier reads (same logits)")
        else:
            pad_note += f" (mean real pair length {mean_len:.1f})"
    if rank == 0:
        qps = B * args.steps / (total_ms / 1e3)
        qps_e2e = B * args.steps / (e2e_ms / 1e3)
        result = {
            "metric": "queries/sec, hybrid BM25+dense top-10 with cross-encoder rerank over a 10M-doc index",
            "value": round(qps, 2),
            "unit": "queries/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": round(total_ms / args.steps, 4),
            "p50_step_ms": round(statistics.median(per_step), 4),
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "fp8-e4m3 projections in the cross-encoder (fp32 accumulate), bf16 elsewhere",
            "data": "synthetic (Zipfian 10M-doc corpus, random unit vectors, random-init weights)",
            "impl": args.impl,
            "config": {
                "model": "bge-small-en encoder + bge-reranker-base cross-encoder (random-init)",
                "index_docs": n_global, "dim": 384, "global_batch": B, "seq_len": args.pair_seq,
                "query_tokens": hcfg.enc_seq, "candidates_per_query": hcfg.n_rerank, "top_k": hcfg.k_out,
                "rerank": hcfg.rerank,
                "parallelism": f"doc-sharded index x{world} + data-parallel reranker x{world}",
                "l2_policy": "inputs larger than L2: every step streams this rank's dense shard "
                             f"({shard.vectors.numel() * 2 / 1e9:.2f} GB/rank) plus the query terms' postings, and uses "
                             "a distinct query batch",
                "padding": pad_note,
                "cuda_graph": bool(eng._graph is not None),
                "exchange": ("none" if world == 1 else ("p2p-fused" if eng.heap is not None else "nccl")),
                "index_build_s": round(build_s, 1),
            },
            "e2e": {"value": round(qps_e2e, 2), "unit": "queries/s", "ms_per_step": round(e2e_ms / args.steps, 4),
                    "p50_step_ms": round(statistics.median(e2e_steps), 4),
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "clocks": clocks,
            "results_valid": ids_ok,
        }
        result["config"]["pipelined"] = bool(pipelined)
        if unpipelined is not None:
            result["one_batch_in_flight"] = unpipelined
        if padded is not None:
            result["padded_cross_encoder"] = padded
        if lat_b1 is not None:
            result["latency_batch1_p50_ms"] = round(lat_b1, 4)
        print(json.dumps(result), flush=True)
    # graphs hold references to the NCCL communicator: drop them before tearing the group down
    eng._graph = None
    eng_p = None
    if lat_b1 is not None:
        eng1._graph = None
    D.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
