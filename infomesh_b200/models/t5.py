"""T5 encoder-decoder (t5-small / t5-base shapes) on the native kernels — the in-process summariser that replaces the
reference's out-of-process LLM servers (infomesh/summarizer/engine.py:126-141,201-213,259-275).

Architecture facts (public T5 model card): pre-norm RMS LayerNorm without bias, bias-free linears, ReLU FFN,
un-scaled dot-product attention (the 1/sqrt(d) is folded into the initialisation), bucketed relative-position bias
(32 buckets, max distance 128) owned by the first layer of each stack and shared by the others, tied input/output
embedding with the LM head scaled by d_model^-0.5.

Kernel mapping: QKV / O / FFN projections -> ``gemm_bf16_tn`` (ReLU and residual fused in the epilogue);
encoder + prefill attention -> ``attn_fwd`` with the per-distance bias table; decode-step self/cross attention ->
``attn_decode`` over KV caches; norms -> ``sum_ln`` (rms_only); LM head -> GEMM (fp32 out) + ``row_argmax``.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from pathlib import Path

import torch

LOG2E = 1.4426950408889634


@dataclass(frozen=True)
class T5Config:
    name: str = "t5-small"
    vocab_size: int = 32128
    d_model: int = 512
    d_kv: int = 64
    heads: int = 8
    d_ff: int = 2048
    enc_layers: int = 6
    dec_layers: int = 6
    rel_buckets: int = 32
    rel_max_distance: int = 128
    eps: float = 1e-6
    pad_id: int = 0
    eos_id: int = 1
    decoder_start_id: int = 0

    @property
    def inner(self) -> int:
        return self.heads * self.d_kv


T5_SMALL = T5Config()
T5_BASE = T5Config(name="t5-base", d_model=768, heads=12, d_ff=3072, enc_layers=12, dec_layers=12)
T5_TINY = T5Config(name="t5-tiny", vocab_size=2048, d_model=128, d_kv=64, heads=2, d_ff=256, enc_layers=2, dec_layers=2)
CONFIGS = {c.name: c for c in (T5_SMALL, T5_BASE, T5_TINY)}


def relative_bucket(rel: torch.Tensor, bidirectional: bool, num_buckets: int, max_distance: int) -> torch.Tensor:
    """T5 bucket of ``rel = key_pos - query_pos`` (integer tensor)."""
    ret = torch.zeros_like(rel)
    n = num_buckets
    if bidirectional:
        n //= 2
        ret = ret + (rel > 0).long() * n
        rel = rel.abs()
    else:
        rel = (-rel).clamp(min=0)
    max_exact = n // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float().clamp(min=1) / max_exact) / math.log(max_distance / max_exact)
                         * (n - max_exact)).long()
    large = large.clamp(max=n - 1)
    return ret + torch.where(is_small, rel, large)


class T5Weights:
    def __init__(self, cfg: T5Config, device="cpu", seed: int = 0):
        g = torch.Generator(device="cpu").manual_seed(seed)
        dev, bf = torch.device(device), torch.bfloat16
        d, inner, ff = cfg.d_model, cfg.inner, cfg.d_ff

        def rnd(shape, std):
            return (torch.randn(shape, generator=g) * std).to(device=dev, dtype=bf)

        def ones(n):
            return torch.ones(n, device=dev, dtype=torch.float32)

        self.emb = rnd((cfg.vocab_size, d), 1.0)
        self.enc_bias = rnd((cfg.rel_buckets, cfg.heads), 0.5).float()
        self.dec_bias = rnd((cfg.rel_buckets, cfg.heads), 0.5).float()
        s_q, s_kv, s_o, s_ff = (d * cfg.d_kv) ** -0.5, d ** -0.5, inner ** -0.5, ff ** -0.5
        self.enc, self.dec = [], []
        for _ in range(cfg.enc_layers):
            self.enc.append({"ln1": ones(d), "wqkv": torch.cat([rnd((inner, d), s_q), rnd((inner, d), s_kv), rnd((inner, d), s_kv)]),
                             "wo": rnd((d, inner), s_o), "ln2": ones(d), "wi": rnd((ff, d), s_kv), "wo2": rnd((d, ff), s_ff)})
        for _ in range(cfg.dec_layers):
            self.dec.append({"ln1": ones(d), "wqkv": torch.cat([rnd((inner, d), s_q), rnd((inner, d), s_kv), rnd((inner, d), s_kv)]),
                             "wo": rnd((d, inner), s_o), "ln_x": ones(d), "wq_x": rnd((inner, d), s_q),
                             "wkv_x": torch.cat([rnd((inner, d), s_kv), rnd((inner, d), s_kv)]), "wo_x": rnd((d, inner), s_o),
                             "ln2": ones(d), "wi": rnd((ff, d), s_kv), "wo2": rnd((d, ff), s_ff)})
        self.enc_final, self.dec_final = ones(d), ones(d)

    def n_params(self) -> int:
        n = self.emb.numel() + self.enc_bias.numel() + self.dec_bias.numel()
        for lay in (*self.enc, *self.dec):
            n += sum(t.numel() for t in lay.values())
        return n + self.enc_final.numel() + self.dec_final.numel()


class T5Model:
    def __init__(self, cfg: T5Config = T5_SMALL, device="cuda", seed: int = 0, weights: T5Weights | None = None):
        self.cfg, self.device = cfg, torch.device(device)
        self.w = weights or T5Weights(cfg, device=device, seed=seed)
        self._bias_cache: dict[tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------ bias tables
    def _enc_bias_table(self, S: int) -> torch.Tensor:
        """[heads, 2S-1] natural-log bias indexed by (j - i) + (S - 1)."""
        key = ("enc", S)
        if key not in self._bias_cache:
            rel = torch.arange(-(S - 1), S, device=self.device)
            b = relative_bucket(rel, True, self.cfg.rel_buckets, self.cfg.rel_max_distance)
            self._bias_cache[key] = self.w.enc_bias[b].t().contiguous()
        return self._bias_cache[key]

    def _dec_bias_table(self, half: int) -> torch.Tensor:
        """[heads, 2*half+1] log2-scaled bias centred on relative position 0 (decode kernel convention)."""
        key = ("dec", half)
        if key not in self._bias_cache:
            rel = torch.arange(-half, half + 1, device=self.device)
            b = relative_bucket(rel, False, self.cfg.rel_buckets, self.cfg.rel_max_distance)
            self._bias_cache[key] = (self.w.dec_bias[b].t() * LOG2E).contiguous()
        return self._bias_cache[key]

    # ------------------------------------------------------------------ encoder
    def encode(self, ids: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        """ids int32 [B, S] -> encoder states bf16 [B, S, d_model] (final RMS norm applied)."""
        from infomesh_b200.ops import attention as A
        from infomesh_b200.ops import gemm as G
        from infomesh_b200.ops import nn as N

        cfg, w = self.cfg, self.w
        B, S = ids.shape
        inner = cfg.inner
        x = w.emb[ids.reshape(-1).long()]
        bias = self._enc_bias_table(S)
        for lay in w.enc:
            n1 = N.layernorm(x, lay["ln1"], None, cfg.eps, rms_only=True)
            qkv = G.linear(n1, lay["wqkv"]).view(B, S, 3 * inner)
            ctx = A.attention(qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:], cfg.heads, kv_lens=lengths,
                              scale=1.0, rel_bias=bias)
            x = G.linear(ctx.view(B * S, inner), lay["wo"], residual=x)
            n2 = N.layernorm(x, lay["ln2"], None, cfg.eps, rms_only=True)
            h = G.linear(n2, lay["wi"], act="relu")
            x = G.linear(h, lay["wo2"], residual=x)
        return N.layernorm(x, w.enc_final, None, cfg.eps, rms_only=True).view(B, S, cfg.d_model)

    # ------------------------------------------------------------------ greedy decode
    @torch.no_grad()
    def generate(self, ids: torch.Tensor, lengths: torch.Tensor | None = None, max_new_tokens: int = 64,
                 check_every: int = 16, use_graph: bool = True) -> torch.Tensor:
        """Greedy decoding.  Returns int32 [B, T] (T <= max_new_tokens), positions after EOS filled with pad.

        One decoder step is ~75 small kernels; launched eagerly it is bound by Python/launch overhead (~1.7 ms/step).
        With ``use_graph`` the step is captured once per ``(B, S, T)`` shape -- the step counter lives on the device (KV-cache
        append, self-attention length and relative-position offset all read it there), decode buffers are cached on the
        model -- and replayed ``T`` times per call."""
        from infomesh_b200.ops import attention as A
        from infomesh_b200.ops import gemm as G
        from infomesh_b200.ops import nn as N

        from types import SimpleNamespace

        cfg, w = self.cfg, self.w
        B, S = ids.shape
        inner, d = cfg.inner, cfg.d_model
        dev = self.device
        T = max_new_tokens
        alpha = d ** -0.5
        # decode state (static buffers + the captured step graph) is cached per shape: capture costs ~30 ms, a step ~0.3
        cache = self.__dict__.setdefault("_dec_states", {})
        st = cache.get((B, S, T))
        if st is None:
            st = SimpleNamespace(
                lengths=torch.empty((B,), dtype=torch.int32, device=dev),
                xk=[torch.empty((B * S, 2 * inner), device=dev, dtype=torch.bfloat16) for _ in w.dec],
                kc=[torch.zeros((B, T, inner), device=dev, dtype=torch.bfloat16) for _ in w.dec],
                vc=[torch.zeros((B, T, inner), device=dev, dtype=torch.bfloat16) for _ in w.dec],
                bias=self._dec_bias_table(T),
                tok=torch.empty((B,), dtype=torch.long, device=dev), done=torch.empty((B,), dtype=torch.bool, device=dev),
                out=torch.empty((B, T), dtype=torch.int32, device=dev), step=torch.zeros((1,), dtype=torch.int32, device=dev),
                col=torch.arange(T, device=dev, dtype=torch.int32)[None, :],
                pad=torch.full((B,), cfg.pad_id, dtype=torch.int32, device=dev), graph=None, graph_failed=False)
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            cache[(B, S, T)] = st
        st.lengths.copy_(lengths if lengths is not None else torch.full((B,), S, dtype=torch.int32, device=dev))
        enc = self.encode(ids, st.lengths).view(B * S, d)
        for li, lay in enumerate(w.dec):                       # cross-attention K/V once per layer
            G.linear(enc, lay["wkv_x"], out=st.xk[li])
        st.tok.fill_(cfg.decoder_start_id)
        st.done.fill_(False)
        st.out.fill_(cfg.pad_id)
        st.step.zero_()
        xk = [x.view(B, S, 2 * inner) for x in st.xk]

        def one_step():
            x = w.emb[st.tok]                                                # [B, d]
            for li, lay in enumerate(w.dec):
                n1 = N.layernorm(x, lay["ln1"], None, cfg.eps, rms_only=True)
                qkv = G.linear(n1, lay["wqkv"])
                A.kv_append(qkv, st.kc[li], st.vc[li], st.step)               # self-attention caches grow by one row
                ctx = A.attention_decode(qkv[:, :inner], st.kc[li], st.vc[li], cfg.heads, 0, scale=1.0,
                                         rel_bias_log2=st.bias, step_dev=st.step)
                x = G.linear(ctx, lay["wo"], residual=x)
                nx = N.layernorm(x, lay["ln_x"], None, cfg.eps, rms_only=True)
                qx = G.linear(nx, lay["wq_x"])
                cx = A.attention_decode(qx, xk[li][..., :inner], xk[li][..., inner:], cfg.heads, st.lengths, scale=1.0)
                x = G.linear(cx, lay["wo_x"], residual=x)
                n2 = N.layernorm(x, lay["ln2"], None, cfg.eps, rms_only=True)
                h = G.linear(n2, lay["wi"], act="relu")
                x = G.linear(h, lay["wo2"], residual=x)
            xf = N.layernorm(x, w.dec_final, None, cfg.eps, rms_only=True)
            logits = G.linear(xf, w.emb, alpha=alpha, out_dtype=torch.float32)
            _, nxt = N.row_argmax(logits)
            nxt = torch.where(st.done, st.pad, nxt.to(torch.int32))
            st.out.copy_(torch.where(st.col == st.step, nxt[:, None], st.out))   # out[:, t] = nxt with t on the device
            st.done.logical_or_(nxt == cfg.eos_id)
            st.tok.copy_(nxt)
            st.step.add_(1)

        t0 = 0
        if use_graph and T > 2 and st.graph is None and not st.graph_failed:
            one_step()                                                       # step 0 runs eagerly (also the warm-up)
            t0 = 1
            torch.cuda.synchronize()
            try:
                s = torch.cuda.Stream(device=dev)
                s.wait_stream(torch.cuda.current_stream())
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    one_step()
                torch.cuda.current_stream().wait_stream(s)
                st.graph = g
            except Exception:  # noqa: BLE001 — eager fallback
                st.graph_failed = True
                torch.cuda.synchronize()
        graph = st.graph if use_graph else None
        for t in range(t0, T):
            if graph is not None:
                graph.replay()
            else:
                one_step()
            if (t + 1) % check_every == 0 and bool(st.done.all()):
                return st.out[:, :t + 1].clone()
        return st.out.clone()

    # ------------------------------------------------------------------ fp32 PyTorch oracle
    def _attn_ref(self, q, k, v, bias, mask):
        cfg = self.cfg
        B, Sq, _ = q.shape
        Sk = k.shape[1]
        qf = q.view(B, Sq, cfg.heads, cfg.d_kv).transpose(1, 2)
        kf = k.view(B, Sk, cfg.heads, cfg.d_kv).transpose(1, 2)
        vf = v.view(B, Sk, cfg.heads, cfg.d_kv).transpose(1, 2)
        s = qf @ kf.transpose(-1, -2)
        if bias is not None:
            s = s + bias
        s = s.masked_fill(~mask, float("-inf"))
        return (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, Sq, cfg.inner)

    @staticmethod
    def _rms(x, g, eps):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * g

    def encode_ref(self, ids, lengths=None):
        cfg, w = self.cfg, self.w
        B, S = ids.shape
        x = w.emb.float()[ids.long()]
        pos = torch.arange(S, device=ids.device)
        bucket = relative_bucket(pos[None, :] - pos[:, None], True, cfg.rel_buckets, cfg.rel_max_distance)
        bias = w.enc_bias[bucket].permute(2, 0, 1)[None]
        mask = torch.ones(B, 1, S, S, dtype=torch.bool, device=ids.device)
        if lengths is not None:
            mask = mask & (pos[None, None, None, :] < lengths.view(B, 1, 1, 1))
        inner = cfg.inner
        for lay in w.enc:
            n1 = self._rms(x, lay["ln1"], cfg.eps)
            qkv = n1 @ lay["wqkv"].float().t()
            x = x + self._attn_ref(qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:], bias, mask) @ lay["wo"].float().t()
            n2 = self._rms(x, lay["ln2"], cfg.eps)
            x = x + torch.relu(n2 @ lay["wi"].float().t()) @ lay["wo2"].float().t()
        return self._rms(x, w.enc_final, cfg.eps)

    @torch.no_grad()
    def generate_ref(self, ids, lengths=None, max_new_tokens: int = 16, enc_states=None):
        """Full-recompute greedy decoding in fp32 (no cache).  Returns (tokens [B, T], logits [B, T, V])."""
        cfg, w = self.cfg, self.w
        B, S = ids.shape
        dev = ids.device
        if lengths is None:
            lengths = torch.full((B,), S, dtype=torch.int32, device=dev)
        enc = enc_states.float() if enc_states is not None else self.encode_ref(ids, lengths)
        inner = cfg.inner
        toks = torch.full((B, 1), cfg.decoder_start_id, dtype=torch.long, device=dev)
        all_logits = []
        xmask = (torch.arange(S, device=dev)[None, None, None, :] < lengths.view(B, 1, 1, 1))
        for _ in range(max_new_tokens):
            Tq = toks.shape[1]
            x = w.emb.float()[toks]
            pos = torch.arange(Tq, device=dev)
            bucket = relative_bucket(pos[None, :] - pos[:, None], False, cfg.rel_buckets, cfg.rel_max_distance)
            bias = w.dec_bias[bucket].permute(2, 0, 1)[None]
            cmask = (pos[None, :] <= pos[:, None])[None, None]
            for lay in w.dec:
                n1 = self._rms(x, lay["ln1"], cfg.eps)
                qkv = n1 @ lay["wqkv"].float().t()
                x = x + self._attn_ref(qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:], bias, cmask) @ lay["wo"].float().t()
                nx = self._rms(x, lay["ln_x"], cfg.eps)
                kv = enc @ lay["wkv_x"].float().t()
                x = x + self._attn_ref(nx @ lay["wq_x"].float().t(), kv[..., :inner], kv[..., inner:], None, xmask) @ lay["wo_x"].float().t()
                n2 = self._rms(x, lay["ln2"], cfg.eps)
                x = x + torch.relu(n2 @ lay["wi"].float().t()) @ lay["wo2"].float().t()
            xf = self._rms(x[:, -1], w.dec_final, cfg.eps)
            logits = (xf * cfg.d_model ** -0.5) @ w.emb.float().t()
            all_logits.append(logits)
            toks = torch.cat([toks, logits.argmax(-1, keepdim=True)], 1)
        return toks[:, 1:].int(), torch.stack(all_logits, 1)

    def flops_per_token(self, seq_len: int) -> float:
        c = self.cfg
        per_layer = 2 * (3 * c.d_model * c.inner + c.inner * c.d_model + 2 * c.d_model * c.d_ff) + 4 * seq_len * c.inner
        return c.enc_layers * per_layer


def load_t5(name_or_dir: str = "t5-small", device="cuda", seed: int = 0) -> T5Model:
    """Config by name (random init), or a directory holding ``config.json`` and ``weights.pt`` (a state dict with this
    module's tensor names)."""
    p = Path(name_or_dir)
    if p.is_dir() and (p / "config.json").exists():
        raw = json.loads((p / "config.json").read_text())
        cfg = T5Config(name=p.name, vocab_size=raw.get("vocab_size", 32128), d_model=raw.get("d_model", 512),
                       d_kv=raw.get("d_kv", 64), heads=raw.get("num_heads", 8), d_ff=raw.get("d_ff", 2048),
                       enc_layers=raw.get("num_layers", 6), dec_layers=raw.get("num_decoder_layers", raw.get("num_layers", 6)),
                       rel_buckets=raw.get("relative_attention_num_buckets", 32),
                       rel_max_distance=raw.get("relative_attention_max_distance", 128))
        model = T5Model(cfg, device=device, seed=seed)
        wfile = p / "weights.pt"
        if wfile.exists():
            state = torch.load(wfile, map_location=device, weights_only=True)
            for k, v in state.items():
                tgt = model.w
                parts = k.split(".")
                if parts[0] in ("enc", "dec") and len(parts) == 3:
                    getattr(tgt, parts[0])[int(parts[1])][parts[2]].copy_(v)
                elif hasattr(tgt, k):
                    getattr(tgt, k).copy_(v)
        return model
    if name_or_dir not in CONFIGS:
        raise ValueError(f"unknown T5 config {name_or_dir!r}; known: {sorted(CONFIGS)}")
    return T5Model(CONFIGS[name_or_dir], device=device, seed=seed)
