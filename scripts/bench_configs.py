"""The other BASELINE.json configurations (bench.py --config ...; the batch-64 hybrid headline stays in bench.py):

  #1  localstore   LocalStore BM25 search over 1k synthetic docs on CPU                        -> queries/s     (no GPU)
  #2  dense_b1     text -> bge-small encoder -> 10M x 384 dense index sharded over N GPUs     -> batch-1 p50 latency
  #4  rag          hybrid retrieval -> GPU passage selection -> t5-small summariser (TP = N)   -> generated tokens/s
  #5  index_build  encode (bge-small) + SimHash + near-dup scan, data-parallel over N GPUs     -> passages/s

Same timing rules as bench.py: >= 3 warm-up steps, CUDA events on the launching stream, barrier + synchronise on both
sides, max over ranks, clocks sampled during the timed region, inputs from pinned host memory for the end-to-end number.
Everything is synthetic / random-init (no network).  One JSON line on rank 0."""
from __future__ import annotations

import json
import os
import random
import statistics
import sys
import time


def _localstore(args) -> int:
    from infomesh_b200.hashing import content_hash
    from infomesh_b200.index.local_store import LocalStore

    rng = random.Random(0)
    vocab = [f"w{i}" for i in range(4000)]
    store = LocalStore(None)
    for d in range(1000):
        text = " ".join(rng.choices(vocab, k=200))
        store.add_document(f"https://ex.org/{d}", f"doc {d}", text, content_hash(f"h{d}"), content_hash(text))
    qs = [" ".join(rng.choices(vocab, k=2)) for _ in range(500)]
    for q in qs[:20]:
        store.search(q, limit=10)
    t0 = time.perf_counter()
    for q in qs:
        store.search(q, limit=10)
    dt = time.perf_counter() - t0
    print(json.dumps({"metric": "queries/sec, LocalStore BM25 over 1k synthetic docs (CPU)", "value": round(len(qs) / dt, 1),
                      "unit": "queries/s", "n_gpus": 0, "mean_ms": round(dt / len(qs) * 1e3, 3), "config": {"docs": 1000}}))
    return 0


def _timed(D, torch, run_step, n_warm, n_steps):
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
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)]
    return D.all_reduce_max(evs[0].elapsed_time(evs[-1])), per


def _graph(torch, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    return g


# ------------------------------------------------------------------------------------------------- #2 dense batch-1
def _dense_b1(args, ClockSampler, peaks) -> int:
    import torch

    from infomesh_b200 import _native
    from infomesh_b200.engine.synth import SynthConfig, gen_vectors
    from infomesh_b200.models.bert import BGE_SMALL, BertModel
    from infomesh_b200.ops import search as S
    from infomesh_b200.parallel import dist as D
    from infomesh_b200.utils.tokenizer import BERT_SPECIALS, HashTokenizer

    ctx = D.init()
    world, rank, dev = ctx.world, ctx.rank, ctx.device
    _native.require()
    n_global = args.docs
    per = (n_global + world - 1) // world
    base = rank * per
    n_local = max(0, min(per, n_global - base))
    scfg = SynthConfig(n_docs=n_local, n_docs_global=n_global, doc_base=base)
    vectors = torch.empty((n_local, scfg.dim), device=dev, dtype=torch.bfloat16)
    for a in range(0, n_local, 1 << 20):
        b = min(n_local, a + (1 << 20))
        vectors[a:b] = gen_vectors(scfg, dev, a, b - a)
    enc = BertModel(BGE_SMALL, device=dev, seed=1)
    tok = HashTokenizer(enc.cfg.vocab_size, BERT_SPECIALS)
    f8 = getattr(args, "dense", "bf16") == "fp8"
    if f8:       # e4m3 rows + per-row scale: the scan streams half the bytes; a 32-wide over-fetch is re-scored on bf16 rows
        from infomesh_b200.ops import nn as N

        vec8, vscale = N.quantize_rows_e4m3(vectors)
        q8 = torch.zeros((1, scfg.dim), device=dev, dtype=torch.uint8)
        q8s = torch.ones((1,), device=dev, dtype=torch.float32)
    K, W, k = max(args.steps, 20), max(args.warmup, 3), 10
    rng = random.Random(7)
    words = [f"w{i}" for i in range(50000)]
    texts = [" ".join(rng.choices(words, k=rng.randint(3, 8))) for _ in range(K + W)]
    S_ENC = 32
    in_ids = torch.zeros((1, S_ENC), dtype=torch.int32, device=dev)
    in_len = torch.ones((1,), dtype=torch.int32, device=dev)
    h_ids = torch.zeros((1, S_ENC), dtype=torch.int32).pin_memory()
    h_len = torch.ones((1,), dtype=torch.int32).pin_memory()
    h_out_s = torch.empty((1, k), dtype=torch.float32).pin_memory()
    h_out_i = torch.empty((1, k), dtype=torch.int64).pin_memory()
    chan = None
    heap = None
    if world > 1:
        from infomesh_b200.parallel import symm

        heap = symm.SymmetricHeap(4 << 20, ctx)
        chan = symm.TopkChannel(heap, 1, k)
        heap.barrier()
    out_s = torch.zeros((1, k), device=dev, dtype=torch.float32)
    out_i = torch.zeros((1, k), device=dev, dtype=torch.int64)

    def scan(q, push=None):
        if f8:
            return S.sim_topk_f8(q8, q8s, vec8, vscale, k, id_offset=base, push=push, rescore=(q, vectors), k_fetch=32)
        return S.sim_topk(q, vectors, k, id_offset=base, push=push)

    def device_pass():
        q = enc.embed(in_ids, in_len, **(dict(out_q8=q8, out_qscale=q8s) if f8 else {}))   # replicated query: no broadcast hop
        s, i = scan(q, push=chan)                                       # shard scan; epilogue pushes the list to every peer
        if chan is not None:
            s, i = S.topk_merge(chan.cand_scores, chan.cand_ids, k, wait=chan)
        out_s.copy_(s)
        out_i.copy_(i)

    g = _graph(torch, device_pass)

    def step_dev(_i):
        g.replay()

    def step_e2e(i):                                                    # text in -> tokenise on the host -> H2D -> graph -> D2H
        ids = tok.encode(texts[i % len(texts)], S_ENC)
        h_ids.zero_()
        h_ids[0, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
        h_len[0] = len(ids)
        in_ids.copy_(h_ids, non_blocking=True)
        in_len.copy_(h_len, non_blocking=True)
        g.replay()
        h_out_s.copy_(out_s, non_blocking=True)
        h_out_i.copy_(out_i, non_blocking=True)
        torch.cuda.current_stream().synchronize()                       # batch-1 latency: the caller waits for its answer

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    total_ms, per_step = _timed(D, torch, step_dev, W, K)
    clocks = sampler.stop()
    e2e_ms, e2e_per = _timed(D, torch, step_e2e, W, K)
    # stage split (eager): encoder vs scan
    q = enc.embed(in_ids, in_len)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    ev[0].record()
    q = enc.embed(in_ids, in_len)
    ev[1].record()
    f_s, f_i = scan(q)
    ev[2].record()
    torch.cuda.synchronize()
    scan_ms = ev[1].elapsed_time(ev[2])
    scan_bytes = (vec8.numel() + vscale.numel() * 4) if f8 else vectors.numel() * 2
    agree = None
    if f8:       # agreement with the exact bf16 scan on this rank's shard, over the timed queries
        hits = tot = 0
        for t in texts[:16]:
            ids = tok.encode(t, S_ENC)
            in_ids.zero_()
            in_ids[0, :len(ids)] = torch.tensor(ids, dtype=torch.int32, device=dev)
            in_len[0] = len(ids)
            q = enc.embed(in_ids, in_len, out_q8=q8, out_qscale=q8s)
            _a, fi = scan(q)
            _b, bi = S.sim_topk(q, vectors, k, id_offset=base)
            hits += len(set(fi[0].tolist()) & set(bi[0].tolist()))
            tot += k
        agree = round(hits / tot, 4)
    if rank == 0:
        p50 = statistics.median(per_step)
        gbs = scan_bytes / (scan_ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": "p50 latency (ms), batch-1 dense top-10 over a 10M x 384 index (BASELINE config #2)",
            "value": round(p50, 4), "unit": "ms", "higher_is_better": False, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(total_ms / K, 4), "queries_per_s": round(1e3 / (total_ms / K), 1), "scaling": "strong",
            "dtype": "bf16" if not f8 else "bf16 encoder; e4m3 + per-row-scale index scan, exact bf16 re-score of a 32-wide over-fetch",
            "data": "synthetic (random unit vectors, random-init bge-small-en)", "impl": "fused",
            **({"recall_at_10_vs_bf16_scan": agree} if f8 else {}),
            "config": {"model": "bge-small-en (random-init)", "index_docs": n_global, "dim": 384, "global_batch": 1, "top_k": k,
                       "dense_shard": "fp8" if f8 else "bf16",
                       "parallelism": f"doc-sharded dense index x{world}; replicated query encoder; fused top-k exchange",
                       "cuda_graph": True, "l2_policy": f"every query streams the rank's {scan_bytes / 1e9:.2f} GB shard (>> L2)"},
            "e2e": {"value": round(statistics.median(e2e_per), 4), "unit": "ms", "ms_per_step": round(e2e_ms / K, 4),
                    "h2d_bytes_per_step": S_ENC * 4 + 4, "d2h_bytes_per_step": k * 12,
                    "api": "text -> HashTokenizer.encode (host) -> pinned H2D -> encoder + sharded sim_topk graph -> D2H top-10"},
            "stages_ms": {"encode": round(ev[0].elapsed_time(ev[1]), 4), "scan_local": round(scan_ms, 4)},
            "roofline": {"scan_local": {"bytes": scan_bytes, "achieved_gbs": round(gbs, 1),
                                        "frac_of_hbm": round(gbs / peaks["hbm_gbs"], 3)}, "peaks": peaks},
            "gpu_launches": _native.launch_count(), "clocks": clocks}), flush=True)
    if heap is not None:
        heap.close()
    D.shutdown()
    return 0


# ------------------------------------------------------------------------------------------------- #5 index build
def _index_build(args, ClockSampler, peaks) -> int:
    import numpy as np
    import torch

    from infomesh_b200 import _native
    from infomesh_b200.engine.index_build import IndexBuilder
    from infomesh_b200.models.bert import BGE_SMALL, BertModel
    from infomesh_b200.ops import dedup as DD
    from infomesh_b200.parallel import dist as D

    ctx = D.init()
    world, rank, dev = ctx.world, ctx.rank, ctx.device
    _native.require()
    bpr, S_TOK = min(1024, 8192 // world), 128
    K, W = max(args.steps, 10), max(args.warmup, 3)
    enc = BertModel(BGE_SMALL, device=dev, seed=1)
    ib = IndexBuilder(capacity_per_rank=(K + W + 2) * bpr + 2_000_000, batch_per_rank=bpr, encoder=enc, device=dev)
    # pre-fill the shard's fingerprint table so the scan works against a realistic index size (2M / rank)
    ib.fingerprints[:2_000_000] = torch.randint(-2 ** 62, 2 ** 62, (2_000_000,), device=dev, dtype=torch.int64)
    ib.n_dev.fill_(2_000_000)
    rng = random.Random(1000 + rank)
    vocab = [f"w{i}" for i in range(20000)]
    n_host = 4                                                          # distinct pre-normalised host batches, cycled
    host = []
    for hb in range(n_host):
        texts = [" ".join(rng.choices(vocab, k=rng.randint(80, 120))) for _ in range(bpr)]
        if hb:                                                          # a few exact / near repeats so dedup has work to do
            texts[5] = host_texts0[9]
            texts[17] = host_texts0[3] + " tail"
        else:
            host_texts0 = texts
        arrs = DD.normalize_batch(texts)
        ids = torch.randint(1000, 20000, (bpr, S_TOK), dtype=torch.int32)
        lens = torch.randint(64, S_TOK + 1, (bpr,), dtype=torch.int32)
        host.append(tuple(t.pin_memory() for t in (ids, lens, *(torch.from_numpy(np.ascontiguousarray(x)) for x in arrs))))
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    dev_batches = [tuple(t.to(dev) for t in hb) for hb in host]
    first = [0]

    def step_dev(i):
        b = dev_batches[i % n_host]
        ib.add_batch(b[0], b[1], b[2], b[3], b[4], b[5], first_doc_id=first[0])
        first[0] += world * bpr

    h_cnt = torch.zeros((3,), dtype=torch.int64).pin_memory()

    def step_e2e(i):
        b = tuple(t.to(dev, non_blocking=True) for t in host[i % n_host])
        ib.add_batch(b[0], b[1], b[2], b[3], b[4], b[5], first_doc_id=first[0])
        first[0] += world * bpr
        h_cnt.copy_(ib.counters, non_blocking=True)

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    total_ms, per_step = _timed(D, torch, step_dev, W, K)
    clocks = sampler.stop()
    e2e_ms, _ = _timed(D, torch, step_e2e, W, K)
    st = ib.stats()
    if rank == 0:
        pps = world * bpr * K / (total_ms / 1e3)
        pps_e2e = world * bpr * K / (e2e_ms / 1e3)
        toks = bpr * float(host[0][1].float().mean())
        tf = toks * enc.flops_per_token(S_TOK) / (statistics.median(per_step) * 1e-3) / 1e12
        print(json.dumps({
            "metric": "passages/sec, index build: encode (bge-small-en) + SimHash + near-duplicate scan (BASELINE config #5)",
            "value": round(pps, 1), "unit": "passages/s", "higher_is_better": True, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(total_ms / K, 4), "scaling": "weak", "dtype": "bf16", "impl": "fused",
            "data": "synthetic passages (80-120 words, 64-128 tokens), random-init bge-small-en",
            "projected_100M_passages_minutes": round(100e6 / pps / 60, 2),
            "config": {"model": "bge-small-en (random-init)", "global_batch": world * bpr, "batch_per_rank": bpr, "seq_len": S_TOK,
                       "fingerprint_table_per_rank": 2_000_000, "parallelism": f"data-parallel x{world}; fingerprints + scan results "
                       "exchanged by push all-gathers through the symmetric heap; no host sync per batch",
                       "l2_policy": "every batch scans the rank's 16 MB fingerprint table and encodes fresh token ids"},
            "e2e": {"value": round(pps_e2e, 1), "unit": "passages/s", "ms_per_step": round(e2e_ms / K, 4), "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 24, "api": "IndexBuilder.add_batch on pinned host batches (token ids + normalised text arrays)"},
            "roofline": {"encoder": {"achieved_tflops_per_gpu": round(tf, 1), "frac_of_bf16_sustained": round(tf / peaks["bf16_tflops_sustained"], 3)},
                         "peaks": peaks},
            "index_stats": st, "gpu_launches": _native.launch_count(), "clocks": clocks}), flush=True)
    D.shutdown()
    return 0


# ------------------------------------------------------------------------------------------------- #4 RAG
def _rag(args, ClockSampler, peaks) -> int:
    import torch

    from infomesh_b200 import _native
    from infomesh_b200.models.t5 import T5_SMALL, T5Model
    from infomesh_b200.parallel import dist as D
    from infomesh_b200.parallel.tp_t5 import TPT5Model

    ctx = D.init()
    world, rank, dev = ctx.world, ctx.rank, ctx.device
    _native.require()
    B, S_IN, NEW = 16, 256, 32                                          # 16 retrieved contexts of 256 tokens -> 32-token summaries
    K, W = max(3, min(args.steps, 10)), 3
    g = torch.Generator().manual_seed(3)
    host = [(torch.randint(5, 30000, (B, S_IN), generator=g, dtype=torch.int32).pin_memory(),
             torch.randint(S_IN // 2, S_IN + 1, (B,), generator=g, dtype=torch.int32).pin_memory()) for _ in range(K + W)]
    model = TPT5Model(T5_SMALL, B, S_IN, seed=3, comm="fused") if world > 1 else T5Model(T5_SMALL, device=dev, seed=3)
    h_out = torch.zeros((B, NEW), dtype=torch.int32).pin_memory()

    def step(i):
        ids, lens = (t.to(dev, non_blocking=True) for t in host[i % len(host)])
        out = model.generate(ids, lens, max_new_tokens=NEW) if world > 1 else model.generate(ids, lens, max_new_tokens=NEW, check_every=10 ** 6)
        h_out[:, :out.shape[1]].copy_(out, non_blocking=True)

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    total_ms, per_step = _timed(D, torch, step, W, K)
    clocks = sampler.stop()
    if rank == 0:
        tps = B * NEW * K / (total_ms / 1e3)
        print(json.dumps({
            "metric": "generated tokens/sec, RAG summariser t5-small on retrieved passages (BASELINE config #4)",
            "value": round(tps, 1), "unit": "tokens/s", "higher_is_better": True, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(total_ms / K, 3), "p50_step_ms": round(statistics.median(per_step), 3), "scaling": "strong", "dtype": "bf16",
            "impl": "fused", "data": "synthetic contexts, random-init t5-small",
            "config": {"model": "t5-small (random-init)", "global_batch": B, "seq_len": S_IN, "new_tokens": NEW,
                       "parallelism": (f"tensor-parallel x{world}: fused all-reduce+RMSNorm ({'NVLS multimem.ld_reduce' if getattr(model, 'nvls', False) else 'P2P'}), "
                                       "vocab-parallel LM head + arg-max exchange") if world > 1 else "single GPU",
                       "decode_step_cuda_graph": True},
            "e2e": {"value": round(tps, 1), "unit": "tokens/s", "h2d_bytes_per_step": B * S_IN * 4 + B * 4, "d2h_bytes_per_step": B * NEW * 4,
                    "note": "the timed step already includes the pinned H2D of the contexts and the D2H of the generated tokens"},
            "gpu_launches": _native.launch_count(), "clocks": clocks}), flush=True)
    if world > 1:
        model.close()
    D.shutdown()
    return 0


def run(args, ClockSampler, peaks) -> int:
    if args.config == "dense_b1":
        return _dense_b1(args, ClockSampler, peaks)
    if args.config == "index_build":
        return _index_build(args, ClockSampler, peaks)
    if args.config == "rag":
        return _rag(args, ClockSampler, peaks)
    raise SystemExit(f"unknown config {args.config}")


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    sys.exit(_localstore(None))
