"""Multi-GPU product serving check + timing: MultiGpuSearchIndex (one worker per GPU) against GpuSearchIndex on one GPU.

    python scripts/gpu_check_multigpu_serving.py [--gpus 2] [--docs 20000] [--queries 256] [--full]

BM25-only serving (no checkpoints -> models may not rank) must return the single-GPU hit lists: global vocabulary, df and
average length make shard-local BM25 scores identical, and the rank fuse sees the same signals (up to the rounding of
document ages against the two build instants, which can swap exact near-ties).  ``--full`` also
runs dense + cross-encoder with ``allow_untrained`` (same seeds -> same weights on every rank) and reports top-1 / top-10
agreement.  Prints wall-clock queries/s of ``search_many`` through the public API for both and ALL OK."""
import argparse
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np  # noqa: E402


def make_store(path, n_docs, seed=0):
    from infomesh_b200.index.local_store import LocalStore

    rng = np.random.default_rng(seed)
    vocab = [f"w{i}" for i in range(5000)]
    p = 1.0 / np.arange(1, len(vocab) + 1) ** 1.05
    p /= p.sum()
    st = LocalStore(path)
    for i in range(n_docs):
        words = rng.choice(len(vocab), size=int(rng.integers(30, 120)), p=p)
        text = " ".join(vocab[w] for w in words)
        st.add_document(url=f"https://site{i % 97}.example/p/{i}", title=" ".join(vocab[w] for w in words[:4]), text=text,
                        raw_html_hash=f"r{i}", text_hash=f"t{i}")
    return st, vocab, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--docs", type=int, default=20000)
    ap.add_argument("--queries", type=int, default=256)
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    import torch

    from infomesh_b200.engine.gpu_index import GpuSearchIndex
    from infomesh_b200.engine.multigpu import MultiGpuSearchIndex

    assert torch.cuda.device_count() >= a.gpus, "not enough GPUs"
    tmp = tempfile.mkdtemp(prefix="im_mg_")
    db = str(Path(tmp) / "index.db")
    t0 = time.time()
    st, vocab, p = make_store(db, a.docs)
    print(f"store: {a.docs} docs in {time.time() - t0:.1f}s", flush=True)
    rng = np.random.default_rng(7)
    queries = [" ".join(vocab[w] for w in rng.choice(len(vocab), size=int(rng.integers(1, 4)), p=p)) for _ in range(a.queries)]
    ok = True
    modes = [("bm25+signals", dict(rank_signals=True))]
    if a.full:
        modes.append(("hybrid+ce", dict(allow_untrained=True)))
    for name, kw in modes:
        single = GpuSearchIndex(st, device="cuda:0", query_batch=64, **kw)
        t0 = time.time()
        single.rebuild()
        tb1 = time.time() - t0
        single.search_many(queries[:64])
        t0 = time.time()
        want = single.search_many(queries)
        t1 = time.time() - t0
        single.close()
        del single
        torch.cuda.empty_cache()
        multi = MultiGpuSearchIndex(st, devices=a.gpus, store_path=db, query_batch=64, **kw)
        t0 = time.time()
        n = multi.rebuild()
        tbn = time.time() - t0
        multi.search_many(queries[:64])
        t0 = time.time()
        got = multi.search_many(queries)
        tn = time.time() - t0
        stats = multi.stats()
        # per-shard segment files: save, drop the worker group, cold-start a new one from the files, expect identical answers
        seg = Path(tmp) / f"segments-{name.replace('+', '-')}"
        t0 = time.time()
        man = multi.save(seg)
        ts = time.time() - t0
        multi.close()
        cold = MultiGpuSearchIndex(st, devices=a.gpus, store_path=db, query_batch=64, **kw)
        t0 = time.time()
        n_loaded = cold.load(seg)
        tl = time.time() - t0
        again = cold.search_many(queries)
        cold.close()
        seg_same = sum(([h["doc_id"] for h in x], [h["snippet"] for h in x]) == ([h["doc_id"] for h in y], [h["snippet"] for h in y])
                       for x, y in zip(got, again))
        print(f"[{name}] segments: {sum(s_['bytes'] for s_ in man['shards']) / 2 ** 20:.0f} MB saved in {ts:.1f}s, cold start from files "
              f"{tl:.1f}s (rebuild {tbn:.1f}s), {n_loaded} docs, identical answers {seg_same}/{len(queries)}", flush=True)
        ok &= n_loaded == a.docs and seg_same == len(queries)
        assert n == a.docs, (n, a.docs)
        same = sum([h["doc_id"] for h in w] == [h["doc_id"] for h in g] for w, g in zip(want, got))
        top1 = sum((w[0]["doc_id"] if w else None) == (g[0]["doc_id"] if g else None) for w, g in zip(want, got))
        over = np.mean([len({h["doc_id"] for h in w} & {h["doc_id"] for h in g}) / max(len(w), 1) for w, g in zip(want, got) if w] or [1.0])
        nonempty = sum(bool(w) for w in want)
        snip = sum(all(hw["snippet"] == hg["snippet"] for hw, hg in zip(w, g)) for w, g in zip(want, got)
                   if [h["doc_id"] for h in w] == [h["doc_id"] for h in g])
        print(f"[{name}] single: build {tb1:.1f}s, {len(queries) / t1:.0f} q/s | x{a.gpus}: build {tbn:.1f}s, {len(queries) / tn:.0f} q/s, "
              f"hbm {stats['hbm_bytes'] / 2 ** 20:.0f} MB | identical lists {same}/{len(queries)} (non-empty {nonempty}), "
              f"top-1 {top1}/{len(queries)}, overlap@10 {over:.3f}, snippets equal on {snip}/{same}", flush=True)
        if name == "bm25+signals":
            # fused scores of near-tied documents may round differently once ages are taken against another build instant
            ok &= over >= 0.99 and same >= 0.85 * len(queries) and top1 >= 0.98 * len(queries) and snip == same and nonempty > len(queries) // 2
        else:
            ok &= top1 >= 0.9 * len(queries) and over >= 0.9
    st.close()
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
