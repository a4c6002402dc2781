"""GPU numerics of the MXFP8 (block-scaled fp8) path: the tcgen05 kind::mxf8f6f4.block_scale GEMM and the fused
quantisers in its producers, each against a plain PyTorch fp32 oracle of the same op, plus an accuracy gate for the
whole cross-encoder (logit error vs the fp32 model)."""
from __future__ import annotations

from dataclasses import replace

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _native_required():
    from infomesh_b200 import _native

    _native.require()
    torch.manual_seed(0)


@pytest.mark.parametrize("m,n,k,bias,act,resid", [
    (256, 384, 256, False, None, False), (1000, 768, 768, True, None, True), (777, 2304, 768, True, None, False),
    (300, 3072, 768, True, "gelu", False), (520, 768, 3072, True, None, True), (130, 1152, 384, True, None, False),
    (4096, 1536, 384, True, "gelu", False), (64, 192, 128, False, "relu", False),
])
def test_mx_gemm_matches_dequantised_fp32_reference(m, n, k, bias, act, resid):
    from infomesh_b200.ops import mx as MX

    a = MX.quantize_act_ref(torch.randn(m, k, device=DEV) * torch.logspace(-2, 2, m, device=DEV)[:, None])
    w = MX.quantize_weight(torch.randn(n, k, device=DEV) * 0.05)
    b = torch.randn(n, device=DEV) if bias else None
    r = torch.randn(m, n, device=DEV).bfloat16() if resid else None
    out = MX.linear_mx(a, w, b, r, act)
    ref = MX.linear_mx_ref(a, w, b, r, act)
    # row-wise: the rows span four orders of magnitude
    err = (out.float() - ref).abs().amax(1)
    assert (err <= 0.01 * ref.abs().amax(1) + 0.02).all(), (err / ref.abs().amax(1)).max().item()


def test_mx_gemm_device_row_count_and_mx_output():
    from infomesh_b200.ops import mx as MX

    m, n, k, rows = 1024, 3072, 768, 777
    a = MX.quantize_act_ref(torch.randn(m, k, device=DEV))
    w = MX.quantize_weight(torch.randn(n, k, device=DEV) * 0.05)
    b = torch.randn(n, device=DEV) * 0.1
    m_dev = torch.tensor([rows], device=DEV, dtype=torch.int32)
    out = MX.linear_mx(a, w, b, act="gelu", out_mx=True, m_dev=m_dev, out=MX.alloc_act(m, n, DEV, init=True))
    ref = MX.linear_mx_ref(a, w, b, act="gelu", rows=rows)
    got = out.float(rows)
    # one e4m3 rounding (2^-4 relative to the block's power-of-two ceiling) on top of the GEMM error
    blk = ref.reshape(rows, n // 32, 32).abs().amax(-1, keepdim=True).expand(-1, -1, 32).reshape(rows, n)
    assert ((got - ref).abs() <= blk * 2 ** -3 + 1e-3).all()
    e_ref = MX.quantize_ref(ref)[1].int()
    e_got = MX.unpack_sfa(out.sf, rows).int()
    assert (e_ref - e_got).abs().max().item() <= 1 and ((e_ref != e_got).float().mean().item() < 0.01)
    # rows beyond the device-side count were not computed: their scales are still the initial 1.0 in untouched row blocks
    assert (MX.unpack_sfa(out.sf, m)[896:] == 127).all()


@pytest.mark.parametrize("h", [768, 384])
def test_layernorm_mx_matches_oracle(h):
    from infomesh_b200.ops import mx as MX
    from infomesh_b200.ops import nn as N

    n = 1000
    x = (torch.randn(n, h, device=DEV) * 3).bfloat16()
    r = torch.randn(n, h, device=DEV).bfloat16()
    g = torch.rand(h, device=DEV) + 0.5
    b = torch.randn(h, device=DEV) * 0.1
    mxo = MX.alloc_act(n, h, DEV, init=True)
    out = N.layernorm_mx(x, g, b, 1e-5, mxo, residual=r)
    ref = N.layernorm_ref(x, g, b, 1e-5, residual=r)
    assert (out.float() - ref).abs().max().item() < 0.03
    q_ref, e_ref = MX.quantize_ref(ref)
    e_got = MX.unpack_sfa(mxo.sf, n)
    assert (e_ref.int() - e_got.int()).abs().max().item() <= 1
    assert (e_ref != e_got).float().mean().item() < 0.01          # fp32 rounding at a power-of-two boundary only
    blk = ref.reshape(n, h // 32, 32).abs().amax(-1, keepdim=True).expand(-1, -1, 32).reshape(n, h)
    assert ((mxo.float() - ref).abs() <= blk * 2 ** -3 + 1e-4).all()


@pytest.mark.parametrize("m,n,k,rows", [(1000, 768, 768, None), (2000, 768, 3072, 1777), (300, 384, 384, None), (9000, 768, 768, 8900),
                                      (128, 384, 1536, 1)])
def test_gemm_with_layernorm_epilogue_matches_oracle(m, n, k, rows):
    """GEMM + bias + residual + LayerNorm + MXFP8 quantisation in one clustered kernel vs the fp32 oracle of the same chain."""
    from infomesh_b200.ops import mx as MX

    a = MX.quantize_act_ref(torch.randn(m, k, device=DEV) * torch.logspace(-1, 1, m, device=DEV)[:, None])
    w = MX.quantize_weight(torch.randn(n, k, device=DEV) * 0.05)
    bias = torch.randn(n, device=DEV) * 0.3
    res = (torch.randn(m, n, device=DEV) * 2).bfloat16()
    g = torch.rand(n, device=DEV) + 0.5
    b = torch.randn(n, device=DEV) * 0.1
    m_dev = torch.tensor([rows], device=DEV, dtype=torch.int32) if rows else None
    live = rows or m
    mxo = MX.alloc_act(m, n, DEV, init=True)
    out = torch.zeros((m, n), device=DEV, dtype=torch.bfloat16)
    for _ in range(3):            # repeated launches: the statistics buffers and cluster barriers alternate cleanly
        MX.linear_mx_ln(a, w, bias, res, g, b, 1e-5, mxo, out=out, m_dev=m_dev)
    torch.cuda.synchronize()
    ref = MX.linear_mx_ln_ref(a, w, bias, res[:live], g, b, 1e-5, rows=live)
    assert torch.isfinite(out[:live].float()).all()
    assert (out[:live].float() - ref).abs().max().item() < 0.03
    if rows:
        blocks = (rows + 127) // 128 * 128
        assert (out[blocks:] == 0).all()          # row blocks past the device-side count were never written
    e_ref = MX.quantize_ref(ref)[1].int()
    e_got = MX.unpack_sfa(mxo.sf, live).int()
    assert (e_ref - e_got).abs().max().item() <= 1 and (e_ref != e_got).float().mean().item() < 0.01
    blk = ref.reshape(live, n // 32, 32).abs().amax(-1, keepdim=True).expand(-1, -1, 32).reshape(live, n)
    assert ((mxo.float(live) - ref).abs() <= blk * 2 ** -3 + 1e-3).all()
    # and it agrees with the two-kernel path it replaces (GEMM + residual -> bf16, then LayerNorm)
    from infomesh_b200.ops import nn as N

    y = MX.linear_mx(a, w, bias, res, m_dev=m_dev)
    two = N.layernorm_mx(y, g, b, 1e-5, MX.alloc_act(m, n, DEV, init=True), n_rows_dev=m_dev)
    assert (out[:live].float() - two[:live].float()).abs().max().item() < 0.06


def test_cross_encoder_fused_layernorm_matches_unfused(monkeypatch):
    from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel
    from infomesh_b200.ops import mx as MX

    m = BertModel(replace(BGE_RERANKER_BASE, layers=3), device=DEV, seed=7)
    g = torch.Generator(device="cpu").manual_seed(9)
    ids = torch.randint(5, 5000, (48, 96), generator=g, dtype=torch.int32).to(DEV)
    lens = torch.randint(10, 97, (48,), generator=g, dtype=torch.int32).to(DEV)
    monkeypatch.setattr(MX, "FUSED_LN", False)
    plain = m.score_packed(ids, lens, precision="mxfp8")
    monkeypatch.setattr(MX, "FUSED_LN", True)
    fused = m.score_packed(ids, lens, precision="mxfp8")
    ref = m.score_ref(ids, lens)
    spread = (ref.max() - ref.min()).item()
    assert torch.isfinite(fused).all()
    # the fused path normalises the fp32 accumulator instead of its bf16 rounding: it is at least as close to the oracle
    assert (fused - ref).abs().max().item() <= (plain - ref).abs().max().item() + 0.05 * spread
    assert (fused - plain).abs().max().item() < 0.15 * spread


def test_attention_mx_matches_attention_then_quantise():
    from infomesh_b200.ops import attention as A
    from infomesh_b200.ops import mx as MX

    B, S, nH, hd = 9, 128, 12, 64
    H = nH * hd
    qkv = (torch.randn(B, S, 3 * H, device=DEV) * 0.7).bfloat16()
    lens = torch.tensor([128, 1, 77, 128, 33, 100, 5, 64, 127], device=DEV, dtype=torch.int32)
    ctx = MX.alloc_act(B * S, H, DEV, init=True)
    A.attention_mx(qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:], nH, ctx, kv_lens=lens)
    ref = A.attention_ref(qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:], nH, kv_lens=lens).reshape(B * S, H)
    got = ctx.float()
    blk = ref.reshape(B * S, H // 32, 32).abs().amax(-1, keepdim=True).expand(-1, -1, 32).reshape(B * S, H)
    assert ((got - ref).abs() <= blk * 2 ** -3 + 0.02).all()      # every query row is valid in the padded layout


def test_cross_encoder_mxfp8_accuracy_gate():
    """Logits of the MXFP8 cross-encoder vs the fp32 PyTorch oracle of the same (random-init) model; the bf16 kernels'
    own error is measured alongside so the gate is relative to what bf16 already costs."""
    from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel

    m = BertModel(replace(BGE_RERANKER_BASE, layers=4), device=DEV, seed=5)
    B, S = 40, 128
    g = torch.Generator(device="cpu").manual_seed(4)
    ids = torch.randint(5, 5000, (B, S), generator=g, dtype=torch.int32).to(DEV)
    lens = torch.randint(20, S + 1, (B,), generator=g, dtype=torch.int32).to(DEV)
    ref = m.score_ref(ids, lens)
    bf = m.score_packed(ids, lens)
    mx = m.score_packed(ids, lens, precision="mxfp8")
    assert torch.isfinite(mx).all()
    spread = (ref.max() - ref.min()).item()
    e_bf = (bf - ref).abs().max().item()
    e_mx = (mx - ref).abs().max().item()
    assert e_mx < max(0.15 * spread, 4 * e_bf + 0.02), (e_mx, e_bf, spread)
    # ranking agreement: the ordering of the candidates by logit is what the reranker consumes
    top_ref = set(torch.topk(ref, 10).indices.tolist())
    top_mx = set(torch.topk(mx, 10).indices.tolist())
    assert len(top_ref & top_mx) >= 7


def test_gpu_index_untrained_models_keep_bm25_order_and_survive_concurrent_rebuild(tmp_path):
    """Serving policy: without checkpoint weights neither the encoder nor the reranker may influence ranking -- the
    device path returns exactly the BM25 (FTS5-formula) order; and searches racing a rebuild never see a torn index."""
    import threading

    from infomesh_b200.engine.gpu_index import GpuSearchIndex
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.models.bert import BertConfig, BertModel

    store = LocalStore(tmp_path / "idx.db")
    topics = ["tensor memory accumulators", "kademlia routing buckets", "sqlite full text ranking", "merkle audit proofs"]
    for i in range(80):
        t = topics[i % len(topics)]
        reps = 1 + (i // len(topics)) % 5                      # different term frequencies -> distinct BM25 scores
        store.add_document(url=f"https://example.org/{i}", title=f"Doc {i}", text=(f"{t} " * reps + f"filler text number {i} " * 3),
                           raw_html_hash=f"r{i}", text_hash=f"t{i}", language="en")
    two_part = ("An opening paragraph about unrelated gardening matters, tomatoes and watering schedules in dry summers. " * 2
                + "\n\n" + "The zeppelin hangar in friedrichshafen stores the restored airship hull for visitors all year round. " * 2)
    store.add_document(url="https://example.org/zep", title="Airships", text=two_part, raw_html_hash="rz", text_hash="tz", language="en")
    dev = torch.device("cuda:0")
    small = BertConfig(name="tiny-enc", vocab_size=30522, hidden=384, layers=2, heads=12, ffn=1536, max_pos=512)
    gi = GpuSearchIndex(store, device=dev, encoder=BertModel(small, device=dev, seed=1), query_batch=8)
    assert gi.use_dense is False and gi.rerank is False and gi.reranker is None
    gi.rebuild()
    got = [h["doc_id"] for h in gi.search("kademlia routing buckets", k=5)]
    want = [r.doc_id for r in store.search("kademlia routing buckets", limit=5)]
    assert got == want, (got, want)
    assert gi.stats()["dense"] is False and gi.stats()["encoder"] == "random-init"
    # K11: the snippet is the passage the DEVICE scored best (second paragraph), not the head of the document
    zep = gi.search("zeppelin hangar friedrichshafen", k=3)[0]
    assert zep["url"].endswith("/zep") and zep["passage"] == 1 and zep["snippet"].startswith("The zeppelin hangar") and "gardening" not in zep["snippet"]

    errors: list[Exception] = []
    stop = threading.Event()

    def searcher():
        try:
            while not stop.is_set():
                hits = gi.search_many(["merkle audit proofs", "sqlite ranking"], k=5)
                assert len(hits) == 2
                for hs in hits:
                    for h in hs:
                        assert h["url"].startswith("https://example.org/")
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    ths = [threading.Thread(target=searcher) for _ in range(3)]
    for t in ths:
        t.start()
    for j in range(3):
        store.add_document(url=f"https://example.org/new{j}", title=f"New {j}", text=f"merkle audit proofs fresh document {j} " * 4,
                           raw_html_hash=f"rn{j}", text_hash=f"tn{j}", language="en")
        gi.rebuild()
    stop.set()
    for t in ths:
        t.join()
    assert not errors, errors[:1]
    assert gi.n_docs == 84
    store.close()


def test_gpu_index_refresh_appends_without_rebuilding(tmp_path):
    """refresh(): new store rows become searchable, old rows keep their vectors / ids, result lists equal a full rebuild."""
    from infomesh_b200.engine.gpu_index import GpuSearchIndex
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.models.bert import BertConfig, BertModel

    store = LocalStore(tmp_path / "i.db")
    words = ["kademlia", "routing", "buckets", "merkle", "audit", "proofs", "sqlite", "ranking", "tensor", "memory", "barrier", "switch"]
    def add(i):
        body = " ".join(words[(i * 7 + j) % len(words)] for j in range(12 + i % 5)) + f". Second paragraph about unique{i} topic{i % 4}. " * 2
        store.add_document(url=f"https://example.org/{i}", title=f"Doc {i} {words[i % len(words)]}", text=body, raw_html_hash=f"r{i}", text_hash=f"t{i}", language="en")
    for i in range(60):
        add(i)
    dev = torch.device("cuda:0")
    small = BertConfig(name="tiny-enc", vocab_size=30522, hidden=384, layers=2, heads=12, ffn=1536, max_pos=512)
    # pure BM25 order (no freshness signal): document ages are taken against the build instant, which differs between an
    # append and a rebuild and could swap exact near-ties
    gi = GpuSearchIndex(store, device=dev, encoder=BertModel(small, device=dev, seed=1), query_batch=8)
    assert gi.rebuild() == 60
    vec_before = gi.engine.shard.vectors[:60].clone()
    gi.mark_deleted(int(gi.doc_ids[3]))
    for i in range(60, 75):
        add(i)
    assert gi.search("unique70") == []
    assert gi.refresh() == 15 and gi.n_docs == 75 and gi.refresh() == 0
    assert torch.equal(gi.engine.shard.vectors[:60], vec_before)                # resident rows were not re-encoded
    assert int(gi.engine.shard.alive[3]) == 0                                     # a deletion survives the append
    hit = gi.search("unique70", k=3)
    assert hit and hit[0]["url"] == "https://example.org/70" and "unique70" in hit[0]["snippet"]
    queries = ["kademlia routing", "merkle audit proofs", "tensor memory", "unique12", "sqlite ranking barrier"]
    appended = [[h["doc_id"] for h in hs] for hs in gi.search_many(queries, k=5)]
    full = GpuSearchIndex(store, device=dev, encoder=gi.encoder, query_batch=8)
    full.rebuild()
    full.mark_deleted(int(full.doc_ids[3]))
    rebuilt = [[h["doc_id"] for h in hs] for hs in full.search_many(queries, k=5)]
    assert appended == rebuilt
    store.close()
