"""Checkpoint / vocabulary loading (models/loader.py, utils/tokenizer.py) and the serving-side ranking policy."""
from __future__ import annotations

import torch

from infomesh_b200.models.bert import BertConfig, BertModel
from infomesh_b200.models.loader import export_bert, load_bert, load_tokenizer, read_safetensors, write_safetensors
from infomesh_b200.utils.tokenizer import WordPieceTokenizer


def test_safetensors_round_trip(tmp_path):
    t = {"a": torch.randn(3, 5), "b": torch.randn(7).bfloat16(), "c": torch.arange(6, dtype=torch.int64).reshape(2, 3)}
    write_safetensors(tmp_path / "x.safetensors", t)
    back = read_safetensors(tmp_path / "x.safetensors")
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and torch.equal(back[k], t[k])


def test_bert_export_load_round_trip_and_pretrained_flag(tmp_path):
    cfg = BertConfig(name="tiny", vocab_size=500, hidden=128, layers=2, heads=4, ffn=256, max_pos=64)
    m = BertModel(cfg, device="cpu", seed=3)
    assert m.pretrained is False
    export_bert(m, tmp_path / "enc")
    m2 = load_bert(tmp_path / "enc", device="cpu")
    assert m2.pretrained is True and m2.cfg.hidden == 128 and m2.cfg.layers == 2 and not m2.cfg.classifier
    ids = torch.randint(0, 500, (2, 9))
    assert torch.allclose(m.hidden_states_ref(ids), m2.hidden_states_ref(ids), atol=1e-5)


def test_cross_encoder_checkpoint_round_trip(tmp_path):
    cfg = BertConfig(name="tiny-rr", vocab_size=400, hidden=128, layers=1, heads=4, ffn=256, max_pos=66, type_vocab=1, eps=1e-5,
                     pos_offset=2, classifier=True)
    m = BertModel(cfg, device="cpu", seed=4)
    export_bert(m, tmp_path / "rr", model_type="xlm-roberta")
    m2 = load_bert(tmp_path / "rr", device="cpu")
    assert m2.cfg.classifier and m2.cfg.pos_offset == 2 and m2.pretrained
    ids = torch.randint(4, 400, (3, 12))
    assert torch.allclose(m.score_ref(ids), m2.score_ref(ids), atol=1e-5)


def test_wordpiece_greedy_longest_match(tmp_path):
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "un", "##aff", "##able", "hello", "##s", ",", "world", "a"]
    (tmp_path / "vocab.txt").write_text("\n".join(vocab))
    tok = WordPieceTokenizer(tmp_path / "vocab.txt")
    v = {w: i for i, w in enumerate(vocab)}
    assert tok.encode_plain("Unaffable", 10) == [v["un"], v["##aff"], v["##able"]]
    assert tok.encode("hellos, world zzz", 16) == [v["[CLS]"], v["hello"], v["##s"], v[","], v["world"], v["[UNK]"], v["[SEP]"]]
    assert tok.encode("a " * 50, 8)[0] == v["[CLS]"] and len(tok.encode("a " * 50, 8)) == 8
    ids, lens = tok.encode_batch(["hello world", "a"], max_len=16)
    assert ids.shape[0] == 2 and lens.tolist() == [4, 3]
    assert isinstance(load_tokenizer(tmp_path, len(vocab)), WordPieceTokenizer)
    assert load_tokenizer(tmp_path / "nope", 10) is None
