"""Checkpoint + vocabulary loading for the BERT-family models (encoder and cross-encoder reranker).

The serving path must not rank with random weights: ``GpuSearchIndex`` only enables the dense leg / the reranker when
the corresponding model reports ``pretrained=True``, which is set here.  Accepted inputs are HuggingFace-style model
directories: ``config.json`` + ``model.safetensors`` (or ``pytorch_model.bin``) and, for tokenisation, ``vocab.txt``
(WordPiece, BERT family) or ``sentencepiece.bpe.model`` (XLM-R family).  The safetensors container is parsed here
(8-byte header length, JSON header, raw little-endian tensors) so no extra wheel is needed.

Reference counterpart: ``SentenceTransformer(model_name)`` in infomesh/index/vector_store.py:104-118 downloads the
model; this image has no network, so paths are explicit (``[gpu] encoder_path`` / ``reranker_path``)."""
from __future__ import annotations

import json
import struct
from pathlib import Path

import numpy as np
import torch

from infomesh_b200.models.bert import BertConfig, BertModel, BertWeights

_ST_DTYPES = {"F32": (np.float32, torch.float32), "F16": (np.float16, torch.float16), "BF16": (np.uint16, torch.bfloat16),
              "I64": (np.int64, torch.int64), "I32": (np.int32, torch.int32), "U8": (np.uint8, torch.uint8)}


def read_safetensors(path: str | Path) -> dict[str, torch.Tensor]:
    """Minimal reader for the safetensors container."""
    raw = Path(path).read_bytes()
    (hlen,) = struct.unpack("<Q", raw[:8])
    header = json.loads(raw[8:8 + hlen])
    base = 8 + hlen
    out = {}
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        npdt, tdt = _ST_DTYPES[meta["dtype"]]
        a, b = meta["data_offsets"]
        arr = np.frombuffer(raw, dtype=npdt, count=(b - a) // np.dtype(npdt).itemsize, offset=base + a).reshape(meta["shape"])
        t = torch.from_numpy(arr.copy())
        out[name] = t.view(torch.bfloat16) if meta["dtype"] == "BF16" else t.to(tdt)
    return out


def write_safetensors(path: str | Path, tensors: dict[str, torch.Tensor]) -> None:
    """Inverse of :func:`read_safetensors` (fp32 / bf16 / int tensors); used by the tests and ``export_bert``."""
    rev = {torch.float32: "F32", torch.float16: "F16", torch.bfloat16: "BF16", torch.int64: "I64", torch.int32: "I32", torch.uint8: "U8"}
    header, blobs, off = {}, [], 0
    for name, t in tensors.items():
        t = t.detach().cpu().contiguous()
        b = (t.view(torch.int16) if t.dtype == torch.bfloat16 else t).numpy().tobytes()
        header[name] = {"dtype": rev[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(b)]}
        blobs.append(b)
        off += len(b)
    h = json.dumps(header).encode()
    Path(path).write_bytes(struct.pack("<Q", len(h)) + h + b"".join(blobs))


def _state_dict(model_dir: Path) -> dict[str, torch.Tensor]:
    st = model_dir / "model.safetensors"
    if st.exists():
        return read_safetensors(st)
    pt = model_dir / "pytorch_model.bin"
    if pt.exists():
        return torch.load(pt, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"{model_dir}: neither model.safetensors nor pytorch_model.bin")


def config_from_hf(raw: dict, name: str) -> BertConfig:
    xlmr = raw.get("model_type", "bert") in ("xlm-roberta", "roberta")
    classifier = any("SequenceClassification" in a for a in raw.get("architectures", []))
    return BertConfig(name=name, vocab_size=raw["vocab_size"], hidden=raw["hidden_size"], layers=raw["num_hidden_layers"],
                      heads=raw["num_attention_heads"], ffn=raw["intermediate_size"], max_pos=raw["max_position_embeddings"],
                      type_vocab=raw.get("type_vocab_size", 2), eps=raw.get("layer_norm_eps", 1e-12),
                      pos_offset=(raw.get("pad_token_id", 1) + 1) if xlmr else 0, pooling=raw.get("pooling", "cls"),
                      classifier=classifier)


def load_bert(model_dir: str | Path, device="cuda") -> BertModel:
    """HuggingFace BERT / XLM-R checkpoint directory -> :class:`BertModel` with ``pretrained=True``."""
    d = Path(model_dir)
    cfg = config_from_hf(json.loads((d / "config.json").read_text()), d.name)
    sd = _state_dict(d)
    # strip the task-model prefix ("bert.", "roberta.", ...)
    pref = next((p for p in ("bert.", "roberta.", "xlm_roberta.", "model.") if any(k.startswith(p + "embeddings.") for k in sd)), "")

    def g(key: str) -> torch.Tensor:
        return sd[pref + key].float()

    w = BertWeights.__new__(BertWeights)
    w.cfg, w.tp_rank, w.tp_size = cfg, 0, 1
    bf, f32 = torch.bfloat16, torch.float32
    w.word = g("embeddings.word_embeddings.weight").to(device, bf)
    w.pos = g("embeddings.position_embeddings.weight").to(device, bf)
    w.type = g("embeddings.token_type_embeddings.weight").to(device, bf)
    w.emb_g = g("embeddings.LayerNorm.weight").to(device, f32)
    w.emb_b = g("embeddings.LayerNorm.bias").to(device, f32)
    w.layers = []
    for i in range(cfg.layers):
        L = f"encoder.layer.{i}."
        w.layers.append(dict(
            wqkv=torch.cat([g(L + f"attention.self.{n}.weight") for n in ("query", "key", "value")], 0).to(device, bf).contiguous(),
            bqkv=torch.cat([g(L + f"attention.self.{n}.bias") for n in ("query", "key", "value")], 0).to(device, f32).contiguous(),
            wo=g(L + "attention.output.dense.weight").to(device, bf).contiguous(), bo=g(L + "attention.output.dense.bias").to(device, f32),
            ln1_g=g(L + "attention.output.LayerNorm.weight").to(device, f32), ln1_b=g(L + "attention.output.LayerNorm.bias").to(device, f32),
            w1=g(L + "intermediate.dense.weight").to(device, bf).contiguous(), b1=g(L + "intermediate.dense.bias").to(device, f32),
            w2=g(L + "output.dense.weight").to(device, bf).contiguous(), b2=g(L + "output.dense.bias").to(device, f32),
            ln2_g=g(L + "output.LayerNorm.weight").to(device, f32), ln2_b=g(L + "output.LayerNorm.bias").to(device, f32)))
    if cfg.classifier:
        H = cfg.hidden
        w.cls_w1 = sd["classifier.dense.weight"].float().to(device, bf)
        w.cls_b1 = sd["classifier.dense.bias"].float().to(device, f32)
        w.cls_w2 = sd["classifier.out_proj.weight"].float()[:1].to(device, bf)
        w.cls_b2 = sd["classifier.out_proj.bias"].float()[:1].to(device, f32)
        w.cls_w2p = torch.zeros((128, H), device=device, dtype=bf)
        w.cls_w2p[0] = w.cls_w2[0]
        w.cls_b2p = torch.zeros(128, device=device, dtype=f32)
        w.cls_b2p[0] = w.cls_b2[0]
    m = BertModel(cfg, device=device, weights=w)
    m.pretrained = True
    m.source = str(d)
    return m


def export_bert(model: BertModel, model_dir: str | Path, model_type: str = "bert") -> None:
    """Write a model in the directory layout :func:`load_bert` reads (round-trip tests, shipping fine-tuned weights)."""
    d = Path(model_dir)
    d.mkdir(parents=True, exist_ok=True)
    cfg, w = model.cfg, model.w
    H = cfg.hidden
    t: dict[str, torch.Tensor] = {
        "embeddings.word_embeddings.weight": w.word, "embeddings.position_embeddings.weight": w.pos,
        "embeddings.token_type_embeddings.weight": w.type, "embeddings.LayerNorm.weight": w.emb_g, "embeddings.LayerNorm.bias": w.emb_b}
    for i, lay in enumerate(w.layers):
        L = f"encoder.layer.{i}."
        for j, n in enumerate(("query", "key", "value")):
            t[L + f"attention.self.{n}.weight"] = lay["wqkv"][j * H:(j + 1) * H]
            t[L + f"attention.self.{n}.bias"] = lay["bqkv"][j * H:(j + 1) * H]
        t[L + "attention.output.dense.weight"], t[L + "attention.output.dense.bias"] = lay["wo"], lay["bo"]
        t[L + "attention.output.LayerNorm.weight"], t[L + "attention.output.LayerNorm.bias"] = lay["ln1_g"], lay["ln1_b"]
        t[L + "intermediate.dense.weight"], t[L + "intermediate.dense.bias"] = lay["w1"], lay["b1"]
        t[L + "output.dense.weight"], t[L + "output.dense.bias"] = lay["w2"], lay["b2"]
        t[L + "output.LayerNorm.weight"], t[L + "output.LayerNorm.bias"] = lay["ln2_g"], lay["ln2_b"]
    pref = {"bert": "bert.", "xlm-roberta": "roberta."}[model_type]
    t = {pref + k: v for k, v in t.items()}
    arch = "BertModel"
    if cfg.classifier:
        t["classifier.dense.weight"], t["classifier.dense.bias"] = w.cls_w1, w.cls_b1
        t["classifier.out_proj.weight"], t["classifier.out_proj.bias"] = w.cls_w2, w.cls_b2
        arch = "XLMRobertaForSequenceClassification" if model_type == "xlm-roberta" else "BertForSequenceClassification"
    write_safetensors(d / "model.safetensors", t)
    (d / "config.json").write_text(json.dumps({
        "model_type": model_type, "architectures": [arch], "vocab_size": cfg.vocab_size, "hidden_size": H,
        "num_hidden_layers": cfg.layers, "num_attention_heads": cfg.heads, "intermediate_size": cfg.ffn,
        "max_position_embeddings": cfg.max_pos, "type_vocab_size": cfg.type_vocab, "layer_norm_eps": cfg.eps,
        "pad_token_id": cfg.pos_offset - 1 if model_type == "xlm-roberta" else 0, "pooling": cfg.pooling}, indent=1))


def load_tokenizer(model_dir: str | Path, vocab_size: int):
    """The tokenizer that matches a checkpoint directory: WordPiece (``vocab.txt``), SentencePiece
    (``sentencepiece.bpe.model``), else ``None`` (callers fall back to the hash tokenizer, which is only meaningful for
    random-init benchmarks)."""
    from infomesh_b200.utils.tokenizer import SentencePieceTokenizer, WordPieceTokenizer

    d = Path(model_dir)
    if (d / "vocab.txt").exists():
        return WordPieceTokenizer(d / "vocab.txt")
    if (d / "sentencepiece.bpe.model").exists():
        return SentencePieceTokenizer(d / "sentencepiece.bpe.model", vocab_size)
    return None
