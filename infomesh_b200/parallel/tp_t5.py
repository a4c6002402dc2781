"""Tensor-parallel T5 (summariser) with the collectives inside the kernels: SURVEY P-TP for N4 / K8 / K9, BASELINE config #4.

Megatron layout without sequence sharding (decode steps have B rows -- nothing to shard): QKV / cross-Q / cross-KV / FFN-in
are column-parallel (whole heads and FFN columns per rank), out-projections and FFN-out are row-parallel and leave a
partial sum on every rank.  Each partial is written by the GEMM straight into a symmetric buffer and consumed by ONE
kernel, ``tp_allreduce_norm`` (csrc/nn/norm_embed.cu), that does the cross-rank barrier, the all-reduce, the residual
add and the next RMSNorm:

* NVLS mode (``VmmHeap`` with a multicast mapping): the row is read with ``multimem.ld_reduce`` -- the NVSwitch returns
  the sum over all GPUs, one load per 16 bytes instead of ``tp``; the barrier arrival is one ``multimem.red``;
* P2P mode (no multicast): ``tp`` unicast peer loads per row, ``red.release.sys`` arrivals.

The LM head is vocab-parallel: every rank multiplies by its rows of the tied embedding, takes a local arg-max
(``row_argmax``), and ``tp_argmax_exchange`` swaps the (value, token) pairs through 8-byte peer stores and picks the
global winner (ties -> smallest token id) identically on every rank -- no logits are ever gathered.

Call sites alternate between two all-reduce channels, which makes the entry barrier sufficient: a rank can only write
channel A's buffer again after passing channel B's barrier, i.e. after every rank finished reading A.
``comm="nccl"`` runs the same math with ``dist.all_reduce`` / ``all_gather`` between plain kernels (A/B baseline).
The reference reaches its summariser over HTTP on a single device (infomesh/summarizer/engine.py:126-141,375-436)."""
from __future__ import annotations

import ctypes
from types import SimpleNamespace

import torch
import torch.distributed as dist

from infomesh_b200 import _native
from infomesh_b200.models.t5 import T5Config, T5Model, T5Weights
from infomesh_b200.ops import attention as A
from infomesh_b200.ops import gemm as G
from infomesh_b200.ops import nn as N
from infomesh_b200.parallel import dist as D


class _ArChannel:
    """One all-reduce channel: symmetric partial buffer ``[max_rows, H]`` + arrival counters + private use counter."""

    def __init__(self, heap, max_rows: int, H: int):
        c = heap.ctx
        self.heap, self.world, self.rank, self.H = heap, c.world, c.rank, H
        self.buf, self._boff = heap.alloc((max_rows, H), torch.bfloat16)
        self.flags, self._foff = heap.alloc((max(c.world, 1),), torch.int32)
        self.state = torch.zeros(2, dtype=torch.int32, device=c.device)
        self._in_tab = heap.peer_table(self._boff)
        self._flag_tab = heap.peer_table(self._foff)
        self.mc = bool(getattr(heap, "has_multicast", False))

    def reduce_norm(self, n_rows: int, residual, gamma, eps: float, want_sum: bool = True, want_norm: bool = True):
        """-> (residual + sum over ranks of buf[:n_rows], RMSNorm of that)."""
        dev = self.buf.device
        s = torch.empty((n_rows, self.H), device=dev, dtype=torch.bfloat16) if want_sum else None
        n = torch.empty((n_rows, self.H), device=dev, dtype=torch.bfloat16) if want_norm else None
        L = _native.require()
        h = self.heap
        rc = L.im_tp_allreduce_norm(
            ctypes.c_void_p(0 if self.mc else self._in_tab.data_ptr()), ctypes.c_void_p(h.mc_ptr(self._boff) if self.mc else 0),
            ctypes.c_void_p(self._flag_tab.data_ptr()), ctypes.c_void_p(h.mc_ptr(self._foff) if self.mc else 0),
            ctypes.c_void_p(h.local_base + self._foff), _native.ptr(self.state), ctypes.c_int(self.world), ctypes.c_int(self.rank),
            _native.ptr(residual), _native.ptr(gamma), ctypes.c_void_p(0), ctypes.c_float(eps), ctypes.c_int(1),
            ctypes.c_int(n_rows), ctypes.c_int(self.H), _native.ptr(s), _native.ptr(n), _native.stream_ptr())
        _native.check(rc, "im_tp_allreduce_norm")
        _native.count_launch()
        return s, n


class _ArgmaxChannel:
    def __init__(self, heap, max_rows: int):
        c = heap.ctx
        self.world, self.rank, self.max_rows = c.world, c.rank, max_rows
        self.slots, so = heap.alloc((2, c.world, max_rows), torch.int64)     # uint2 {value bits, token id}
        self.flags, fo = heap.alloc((max(c.world, 1),), torch.int32)
        self._s_tab, self._f_tab = heap.peer_table(so), heap.peer_table(fo)
        self.local_flags_ptr = heap.local_base + fo
        self.state = torch.zeros(2, dtype=torch.int32, device=c.device)

    def __call__(self, val: torch.Tensor, idx: torch.Tensor, out_idx: torch.Tensor):
        n = val.numel()
        assert n == self.max_rows, "the exchange slots are sized for a fixed row count"
        L = _native.require()
        rc = L.im_tp_argmax_exchange(_native.ptr(val), _native.ptr(idx), ctypes.c_int(n), ctypes.c_void_p(self._s_tab.data_ptr()),
                                     ctypes.c_void_p(self._f_tab.data_ptr()), ctypes.c_void_p(self.local_flags_ptr),
                                     _native.ptr(self.state), ctypes.c_int(self.world), ctypes.c_int(self.rank), _native.ptr(out_idx),
                                     ctypes.c_void_p(0), _native.stream_ptr())
        _native.check(rc, "im_tp_argmax_exchange")
        _native.count_launch()
        return out_idx


def make_heap(nbytes: int, ctx=None, prefer_nvls: bool = True):
    """VMM heap with an NVLS multicast mapping when the fabric offers one, else the CUDA-IPC symmetric heap."""
    ctx = ctx or D.ctx()
    if prefer_nvls and ctx.is_dist:
        try:
            from infomesh_b200.parallel.vmm import VmmHeap, support

            if all(support(ctx.device.index or 0).values()):
                return VmmHeap(nbytes, ctx)
        except Exception:  # noqa: BLE001 -- fall through to the IPC heap
            pass
    from infomesh_b200.parallel.symm import SymmetricHeap

    return SymmetricHeap(nbytes, ctx)


class TPT5Model:
    """``tp`` = world size of the default process group.  Inputs are replicated; every rank returns the same tokens."""

    def __init__(self, cfg: T5Config, batch: int, seq_len: int, *, seed: int = 0, comm: str = "fused", heap=None,
                 prefer_nvls: bool = True):
        self.ctx = D.ctx()
        c = self.ctx
        self.cfg, self.comm, self.B, self.S, self.tp = cfg, comm, batch, seq_len, c.world
        tp, r = self.tp, c.rank
        assert cfg.heads % tp == 0 and cfg.d_ff % tp == 0 and cfg.vocab_size % tp == 0, "heads / d_ff / vocab must divide by tp"
        full = T5Weights(cfg, device=c.device, seed=seed)          # same seed -> same full weights on every rank; slice locally
        self.full = full
        hs = cfg.inner // tp                                       # columns of Q / K / V owned by this rank (whole heads)
        fs = cfg.d_ff // tp
        self.heads_local, self.hs, self.fs = cfg.heads // tp, hs, fs
        inner = cfg.inner

        def cols3(w):                                              # [3*inner, d] -> local q|k|v rows
            return torch.cat([w[j * inner + r * hs:j * inner + (r + 1) * hs] for j in range(3)]).contiguous()

        def shard(lay, dec):
            out = {"ln1": lay["ln1"], "wqkv": cols3(lay["wqkv"]), "wo": lay["wo"][:, r * hs:(r + 1) * hs].contiguous(), "ln2": lay["ln2"],
                   "wi": lay["wi"][r * fs:(r + 1) * fs].contiguous(), "wo2": lay["wo2"][:, r * fs:(r + 1) * fs].contiguous()}
            if dec:
                wkv = lay["wkv_x"]
                out.update({"ln_x": lay["ln_x"], "wq_x": lay["wq_x"][r * hs:(r + 1) * hs].contiguous(),
                            "wkv_x": torch.cat([wkv[r * hs:(r + 1) * hs], wkv[inner + r * hs:inner + (r + 1) * hs]]).contiguous(),
                            "wo_x": lay["wo_x"][:, r * hs:(r + 1) * hs].contiguous()})
            return out

        self.enc = [shard(l, False) for l in full.enc]
        self.dec = [shard(l, True) for l in full.dec]
        h0, h1 = r * self.heads_local, (r + 1) * self.heads_local
        self._ref = T5Model(cfg, device=c.device, weights=full)    # bias tables / oracles come from the unsharded model
        self._h = (h0, h1)
        self.v0, self.v1 = r * (cfg.vocab_size // tp), (r + 1) * (cfg.vocab_size // tp)
        self.emb_shard = full.emb[self.v0:self.v1]
        self.heap = None
        if comm == "fused":
            need = 2 * batch * seq_len * cfg.d_model * 2 + (4 << 20)
            self.heap = heap or make_heap(need + (8 << 20), c, prefer_nvls)
            self.ch = [_ArChannel(self.heap, batch * seq_len, cfg.d_model) for _ in range(2)]
            self.amax = _ArgmaxChannel(self.heap, batch)
            self._site = 0
            torch.cuda.synchronize()
            if c.is_dist:
                dist.barrier()
        self.nvls = bool(self.heap is not None and getattr(self.heap, "has_multicast", False))

    # ------------------------------------------------------------------ collectives
    def _row_parallel(self, a, w, residual, gamma, n_rows, want_norm=True):
        """x_new = residual + all-reduce(a @ w^T);  returns (x_new, RMSNorm(x_new) or None)."""
        cfg = self.cfg
        if self.comm == "fused":
            ch = self.ch[self._site & 1]
            self._site += 1
            G.linear(a, w, out=ch.buf[:n_rows])
            return ch.reduce_norm(n_rows, residual, gamma, cfg.eps, want_norm=want_norm)
        part = G.linear(a, w)
        if self.ctx.is_dist:
            dist.all_reduce(part)
        x = part + residual
        return x, (N.layernorm(x, gamma, None, cfg.eps, rms_only=True) if want_norm else None)

    # ------------------------------------------------------------------ encoder
    def encode(self, ids: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        cfg = self.cfg
        B, S = ids.shape
        hs, M = self.hs, B * S
        x = self.full.emb[ids.reshape(-1).long()]
        bias = self._ref._enc_bias_table(S)[self._h[0]:self._h[1]].contiguous()
        n = N.layernorm(x, self.enc[0]["ln1"], None, cfg.eps, rms_only=True)
        for li, lay in enumerate(self.enc):
            qkv = G.linear(n, lay["wqkv"]).view(B, S, 3 * hs)
            ctx = A.attention(qkv[..., :hs], qkv[..., hs:2 * hs], qkv[..., 2 * hs:], self.heads_local, kv_lens=lengths, scale=1.0,
                              rel_bias=bias)
            x, n2 = self._row_parallel(ctx.view(M, hs), lay["wo"], x, lay["ln2"], M)
            h = G.linear(n2, lay["wi"], act="relu")
            nxt = self.enc[li + 1]["ln1"] if li + 1 < len(self.enc) else self.full.enc_final
            x, n = self._row_parallel(h, lay["wo2"], x, nxt, M)
        return n.view(B, S, cfg.d_model)       # the last fused norm IS the final encoder RMSNorm

    # ------------------------------------------------------------------ greedy decode
    @torch.no_grad()
    def generate(self, ids: torch.Tensor, lengths: torch.Tensor | None = None, max_new_tokens: int = 32, use_graph: bool = True) -> torch.Tensor:
        cfg = self.cfg
        B, S = ids.shape
        assert B == self.B, "the arg-max exchange slots are sized for the configured batch"
        dev, T, hs, d = ids.device, max_new_tokens, self.hs, cfg.d_model
        alpha = d ** -0.5
        # decode state (static buffers + the captured step graph) is cached per shape, like T5Model.generate
        cache = self.__dict__.setdefault("_dec_states", {})
        st = cache.get((B, S, T))
        if st is None:
            st = SimpleNamespace(
                lens=torch.empty((B,), dtype=torch.int32, device=dev),
                xk=[torch.empty((B * S, 2 * hs), device=dev, dtype=torch.bfloat16) for _ in self.dec],
                kc=[torch.zeros((B, T, hs), device=dev, dtype=torch.bfloat16) for _ in self.dec],
                vc=[torch.zeros((B, T, hs), device=dev, dtype=torch.bfloat16) for _ in self.dec],
                bias=self._ref._dec_bias_table(T)[self._h[0]:self._h[1]].contiguous(),
                tok=torch.empty((B,), dtype=torch.long, device=dev), done=torch.empty((B,), dtype=torch.bool, device=dev),
                out=torch.empty((B, T), dtype=torch.int32, device=dev), step=torch.zeros((1,), dtype=torch.int32, device=dev),
                col=torch.arange(T, device=dev, dtype=torch.int32)[None, :], pad=torch.full((B,), cfg.pad_id, dtype=torch.int32, device=dev),
                nxt=torch.zeros((B,), dtype=torch.int32, device=dev), graph=None)
            cache[(B, S, T)] = st
        lens = st.lens
        lens.copy_(lengths if lengths is not None else torch.full((B,), S, dtype=torch.int32, device=dev))
        enc = self.encode(ids, lens).view(B * S, d)
        for li, lay in enumerate(self.dec):
            G.linear(enc, lay["wkv_x"], out=st.xk[li])
        xk = [x.view(B, S, 2 * hs) for x in st.xk]
        st.tok.fill_(cfg.decoder_start_id)
        st.done.fill_(False)
        st.out.fill_(cfg.pad_id)
        st.step.zero_()

        def one_step():
            x = self.full.emb[st.tok]
            n = N.layernorm(x, self.dec[0]["ln1"], None, cfg.eps, rms_only=True)
            for li, lay in enumerate(self.dec):
                qkv = G.linear(n, lay["wqkv"])
                A.kv_append(qkv, st.kc[li], st.vc[li], st.step)
                ctx = A.attention_decode(qkv[:, :hs], st.kc[li], st.vc[li], self.heads_local, 0, scale=1.0, rel_bias_log2=st.bias,
                                         step_dev=st.step)
                x, nx = self._row_parallel(ctx, lay["wo"], x, lay["ln_x"], B)
                qx = G.linear(nx, lay["wq_x"])
                cx = A.attention_decode(qx, xk[li][..., :hs], xk[li][..., hs:], self.heads_local, lens, scale=1.0)
                x, n2 = self._row_parallel(cx, lay["wo_x"], x, lay["ln2"], B)
                h = G.linear(n2, lay["wi"], act="relu")
                nxt_g = self.dec[li + 1]["ln1"] if li + 1 < len(self.dec) else self.full.dec_final
                x, n = self._row_parallel(h, lay["wo2"], x, nxt_g, B)
            logits = G.linear(n, self.emb_shard, alpha=alpha, out_dtype=torch.float32)      # [B, V / tp]
            val, idx = N.row_argmax(logits, id_offset=self.v0)
            if self.comm == "fused":
                self.amax(val, idx, st.nxt)
            elif self.ctx.is_dist:
                vals, idxs = D.all_gather_cat(val.contiguous()), D.all_gather_cat(idx.contiguous())
                best = vals.max(0).values
                cand = torch.where(vals == best[None], idxs, torch.full_like(idxs, 2 ** 30))
                st.nxt.copy_(cand.min(0).values)
            else:
                st.nxt.copy_(idx)
            nxt = torch.where(st.done, st.pad, st.nxt)
            st.out.copy_(torch.where(st.col == st.step, nxt[:, None], st.out))
            st.done.logical_or_(nxt == cfg.eos_id)
            st.tok.copy_(nxt)
            st.step.add_(1)

        t0 = 0
        if use_graph and T > 2 and self.comm == "fused" and st.graph is None:
            # 18 all-reduce sites per step (even), so the two channels keep alternating across replays
            one_step()
            t0 = 1
            torch.cuda.synchronize()
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                one_step()
            torch.cuda.current_stream().wait_stream(s)
            st.graph = g
        graph = st.graph if (use_graph and self.comm == "fused") else None
        for _t in range(t0, T):
            if graph is not None:
                graph.replay()
            else:
                one_step()
        self.last_step_graph = graph is not None
        return st.out.clone()

    def close(self):
        if self.heap is not None:
            self.heap.close()
            self.heap = None
