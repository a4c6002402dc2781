"""Symmetric heap over NVLink peer memory (SURVEY C-0) and the exchange channels built on it.

Every rank allocates one slab with ``cudaMalloc``, exports a CUDA-IPC handle, and maps all peers' slabs; an object
allocated at offset ``o`` lives at ``base[p] + o`` on every rank ``p``, so a kernel can address a peer's copy with
plain ``ld/st.global``.  Handles are exchanged once through the process group (the only host-side collective);
after that all data motion is done by kernels (``csrc/comm/symm.cu``, fused epilogues in the GEMM / top-k merge).

The reference has no collective backend (its distribution layer is WAN libp2p streams,
infomesh/p2p/protocol.py:339-355, infomesh/p2p/routing.py:193-267): this is the intra-node replacement.
At world size 1 the heap degenerates to a private slab and every primitive still works (self-push).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from infomesh_b200 import _native
from infomesh_b200.parallel import dist as D

_ALIGN = 256


class _RawCuda:
    """Exposes a raw device pointer to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class SymmetricHeap:
    def __init__(self, nbytes: int = 64 << 20, ctx: D.DistContext | None = None):
        self.ctx = ctx or D.ctx()
        c = self.ctx
        if c.device.type != "cuda":
            raise RuntimeError("SymmetricHeap needs a CUDA device")
        self.nbytes = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        L = _native.require()
        base = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        with torch.cuda.device(c.device):
            _native.check(L.im_symm_alloc(ctypes.c_size_t(self.nbytes), ctypes.byref(base), handle), "im_symm_alloc")
        self.local_base = int(base.value)
        self.bases: list[int] = [0] * c.world
        self.bases[c.rank] = self.local_base
        self._opened: list[int] = []
        if c.world > 1:
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=c.device)
            everyone = [torch.empty_like(mine) for _ in range(c.world)]
            dist.all_gather(everyone, mine)
            for p, h in enumerate(everyone):
                if p == c.rank:
                    continue
                buf = (ctypes.c_uint8 * 64)(*h.cpu().tolist())
                peer = ctypes.c_void_p()
                _native.check(L.im_symm_open(buf, ctypes.byref(peer)), f"im_symm_open(rank {p})")
                self.bases[p] = int(peer.value)
                self._opened.append(int(peer.value))
        self._local = torch.as_tensor(_RawCuda(self.local_base, self.nbytes), device=c.device)
        self._top = 0
        self._tables: list[torch.Tensor] = []
        self._bar = FlagChannel(self)

    # ------------------------------------------------------------------ allocation
    def alloc(self, shape, dtype) -> tuple[torch.Tensor, int]:
        """Collective by convention: every rank must allocate the same sequence.  -> (local view, offset)."""
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        off = self._top
        if off + nbytes > self.nbytes:
            raise MemoryError(f"symmetric heap exhausted: need {nbytes} bytes at offset {off} of {self.nbytes}")
        self._top = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        view = self._local[off:off + nbytes].view(dtype).view(shape)
        return view, off

    def peer_table(self, offset: int) -> torch.Tensor:
        """Device array of ``world`` pointers to the object at ``offset`` on every rank."""
        t = torch.tensor([b + offset for b in self.bases], dtype=torch.int64, device=self.ctx.device)
        self._tables.append(t)
        return t

    def barrier(self) -> None:
        L = _native.require()
        _native.check(L.im_p2p_barrier(ctypes.c_void_p(self._bar.peer_flags_ptr), ctypes.c_void_p(self._bar.step_ptr),
                                       ctypes.c_int(self.ctx.world), ctypes.c_int(self.ctx.rank), _native.stream_ptr()),
                      "im_p2p_barrier")
        _native.count_launch()

    def close(self) -> None:
        L = _native.lib()
        torch.cuda.synchronize(self.ctx.device)
        if self.ctx.is_dist and dist.is_initialized():
            dist.barrier()
        for p in self._opened:
            L.im_symm_close(ctypes.c_void_p(p))
        self._opened = []
        if self.local_base:
            self._local = None
            L.im_symm_free(ctypes.c_void_p(self.local_base))
            self.local_base = 0


class FlagChannel:
    """``world`` cumulative arrival counters on every rank (symmetric) + a private ``{step, done}`` pair that the
    consuming kernel advances."""

    def __init__(self, heap: SymmetricHeap):
        self.heap, self.world, self.rank = heap, heap.ctx.world, heap.ctx.rank
        self.flags, off = heap.alloc((max(self.world, 1),), torch.int32)
        self._flag_tab = heap.peer_table(off)
        self.peer_flags_ptr = self._flag_tab.data_ptr()
        self.local_flags_ptr = self.flags.data_ptr()
        self.state = torch.zeros(2, dtype=torch.int32, device=heap.ctx.device)
        self.step_ptr = self.state.data_ptr()


class AllGatherChannel(FlagChannel):
    """Push-based all-gather of one fixed-size block per rank per step (``im_p2p_allgather``)."""

    def __init__(self, heap: SymmetricHeap, shape, dtype, ctas: int = 0):
        super().__init__(heap)
        self.shape, self.dtype = tuple(shape), dtype
        n = 1
        for s in self.shape:
            n *= s
        self.block_bytes = n * torch.empty((), dtype=dtype).element_size()
        if self.block_bytes % 16:
            raise ValueError("all-gather block must be a multiple of 16 bytes")
        self.recv, off = heap.alloc((2, self.world, self.block_bytes), torch.uint8)
        self._buf_tab = heap.peer_table(off)
        self.out = torch.empty((self.world, *self.shape), dtype=dtype, device=heap.ctx.device)
        self.ctas = ctas

    def __call__(self, src: torch.Tensor) -> torch.Tensor:
        assert src.is_contiguous() and src.dtype == self.dtype and tuple(src.shape) == self.shape, (src.shape, self.shape)
        L = _native.require()
        rc = L.im_p2p_allgather(_native.ptr(src), ctypes.c_size_t(self.block_bytes), ctypes.c_void_p(self._buf_tab.data_ptr()),
                                ctypes.c_void_p(self.peer_flags_ptr), ctypes.c_void_p(self.step_ptr),
                                ctypes.c_int(self.world), ctypes.c_int(self.rank), _native.ptr(self.out),
                                ctypes.c_int(self.ctas), _native.stream_ptr())
        if rc < 0:
            _native.check(rc, "im_p2p_allgather")
        _native.count_launch()
        return self.out


class TopkChannel(FlagChannel):
    """Receive areas for the fused top-k exchange: scores/ids ``[2][world][nq][k]`` on every rank.

    Producer: ``topk_merge(..., push=chan)`` — the local merge kernel stores its result into slot[rank] of every peer.
    Consumer: ``topk_merge(chan.recv_scores_view, chan.recv_ids_view, k, wait=chan)`` — waits for all arrivals."""

    def __init__(self, heap: SymmetricHeap, nq: int, k: int, *, degraded_ok: bool = False, wait_limit: int = 0):
        super().__init__(heap)
        self.nq, self.k = nq, k
        # degraded mode: a shard that stays silent for ``wait_limit`` polls is dropped from the merge and recorded in
        # ``status`` (bit per rank, sticky) instead of trapping the kernel
        self.degraded_ok, self.wait_limit = degraded_ok, wait_limit
        self.status = torch.zeros(1, dtype=torch.int32, device=heap.ctx.device)
        self.status_ptr = self.status.data_ptr()
        self.recv_scores, so = heap.alloc((2, self.world, nq, k), torch.float32)
        self.recv_ids, io = heap.alloc((2, self.world, nq, k), torch.int64)
        self._s_tab, self._i_tab = heap.peer_table(so), heap.peer_table(io)
        self.peer_scores_ptr, self.peer_ids_ptr = self._s_tab.data_ptr(), self._i_tab.data_ptr()

    def dead_ranks(self) -> list[int]:
        m = int(self.status.item()) & 0xFFFFFFFF
        return [r for r in range(self.world) if (m >> r) & 1]

    def clear_status(self) -> None:
        self.status.zero_()

    @property
    def cand_scores(self) -> torch.Tensor:       # parity is applied inside the kernel
        return self.recv_scores[0]

    @property
    def cand_ids(self) -> torch.Tensor:
        return self.recv_ids[0]
