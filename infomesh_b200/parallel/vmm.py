"""VMM symmetric heap + NVLS multicast (SURVEY C-0 / §5.8, csrc/comm/vmm.cu).

Same contract as :class:`infomesh_b200.parallel.symm.SymmetricHeap` (``alloc`` / ``peer_table`` / ``bases``), built on
the CUDA virtual-memory API instead of CUDA IPC:

* every rank ``cuMemCreate``\\ s one physical allocation, exports it as a POSIX file descriptor and maps every peer's
  allocation into its own address space (``cuMemImportFromShareableHandle`` + ``cuMemMap``);
* rank 0 creates an NVLS **multicast object** over all devices, every rank adds its device and binds its allocation,
  and the object is mapped too: a store to ``mc_base + off`` lands in EVERY GPU's copy of offset ``off`` (the NVSwitch
  replicates it), ``multimem.ld_reduce`` on it returns the element-wise sum over all GPUs, ``multimem.red`` adds to all.

File descriptors travel between the rank processes as SCM_RIGHTS ancillary data over Unix sockets (``socket.send_fds``);
everything after the one-off setup is kernel-driven.  When the device / driver / fabric has no multicast support the
heap still works as a plain peer-mapped symmetric heap (``has_multicast == False``) and callers use the unicast P2P
kernels.  The reference has nothing comparable (WAN msgpack streams, infomesh/p2p/routing.py:193-267)."""
from __future__ import annotations

import ctypes
import os
import socket
import time

import torch
import torch.distributed as dist

from infomesh_b200 import _native
from infomesh_b200.parallel import dist as D
from infomesh_b200.parallel.symm import _ALIGN, _RawCuda


def support(device_index: int = 0) -> dict:
    """{"vmm", "posix_fd", "multicast"} capability bits of a device (all False without the native library / a GPU)."""
    try:
        bits = int(_native.lib().im_vmm_supported(ctypes.c_int(device_index)))
    except Exception:  # noqa: BLE001
        bits = 0
    return {"vmm": bool(bits & 1), "posix_fd": bool(bits & 2), "multicast": bool(bits & 4)}


def _sock_path(tag: str, rank: int) -> str:
    return f"/tmp/im_vmm_{os.getuid()}_{tag}_{rank}.sock"


def exchange_fds(rank: int, world: int, my_fds: list[int], tag: str, timeout: float = 60.0) -> dict[int, list[int]]:
    """Every rank offers ``my_fds``; returns {peer rank: its fds, duplicated into this process}.  Connections are buffered
    by the listen backlog, so each rank first sends to everybody and then accepts -- no ordering deadlock."""
    path = _sock_path(tag, rank)
    try:
        os.unlink(path)
    except FileNotFoundError:
        pass
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(world + 4)
    srv.settimeout(timeout)
    if dist.is_initialized():
        dist.barrier()                      # every listener exists
    conns = []
    for p in range(world):
        if p == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        deadline = time.time() + timeout
        while True:
            try:
                c.connect(_sock_path(tag, p))
                break
            except (FileNotFoundError, ConnectionRefusedError):
                if time.time() > deadline:
                    raise
                time.sleep(0.01)
        socket.send_fds(c, [rank.to_bytes(4, "little")], my_fds)
        conns.append(c)
    got: dict[int, list[int]] = {}
    for _ in range(world - 1):
        c, _addr = srv.accept()
        c.settimeout(timeout)
        msg, fds, _flags, _a = socket.recv_fds(c, 4, 8)
        got[int.from_bytes(msg[:4], "little")] = list(fds)
        c.close()
    if dist.is_initialized():
        dist.barrier()                      # everybody has received: senders may close
    for c in conns:
        c.close()
    srv.close()
    try:
        os.unlink(path)
    except FileNotFoundError:
        pass
    return got


class VmmHeap:
    """Symmetric heap on cuMem* with an optional NVLS multicast mapping (see module docstring)."""

    _serial = 0

    def __init__(self, nbytes: int = 64 << 20, ctx: D.DistContext | None = None, multicast: bool = True, tag: str | None = None):
        self.ctx = ctx or D.ctx()
        c = self.ctx
        if c.device.type != "cuda":
            raise RuntimeError("VmmHeap needs a CUDA device")
        L = _native.require()
        L.im_vmm_round_size.restype = ctypes.c_longlong
        dev = c.device.index if c.device.index is not None else torch.cuda.current_device()
        self.dev = dev
        cap = support(dev)
        if not (cap["vmm"] and cap["posix_fd"]):
            raise RuntimeError(f"device {dev} lacks VMM / POSIX-fd handle support: {cap}")
        want_mc = multicast and cap["multicast"] and c.world > 1
        with torch.cuda.device(c.device):
            torch.cuda.current_stream().synchronize()          # make sure a context exists
            size = int(L.im_vmm_round_size(ctypes.c_longlong(nbytes), ctypes.c_int(dev), ctypes.c_int(c.world if want_mc else 1)))
            if size <= 0:
                raise RuntimeError("im_vmm_round_size failed")
            self.nbytes = size
            h, fd = ctypes.c_ulonglong(), ctypes.c_int(-1)
            _native.check(L.im_vmm_create(ctypes.c_longlong(size), ctypes.c_int(dev), ctypes.byref(h), ctypes.byref(fd)), "im_vmm_create")
            self._handle, self._fd = int(h.value), int(fd.value)
            self._peer_handles: list[int] = []
            self._maps: list[int] = []
            self.bases = [0] * c.world
            ptr = ctypes.c_void_p()
            _native.check(L.im_vmm_map(ctypes.c_ulonglong(self._handle), ctypes.c_longlong(size), ctypes.c_int(dev), ctypes.byref(ptr)), "im_vmm_map")
            self.local_base = int(ptr.value)
            self.bases[c.rank] = self.local_base
            self._maps.append(self.local_base)
            self.mc_base = 0
            self._mc_handle = 0
            if c.world > 1:
                tag = tag or os.environ.get("MASTER_PORT", "0")
                mc_fd = -1
                if want_mc and c.rank == 0:
                    mh, mfd = ctypes.c_ulonglong(), ctypes.c_int(-1)
                    rc = L.im_mc_create(ctypes.c_longlong(size), ctypes.c_int(c.world), ctypes.byref(mh), ctypes.byref(mfd))
                    if rc == 0:
                        self._mc_handle, mc_fd = int(mh.value), int(mfd.value)
                mine = [self._fd] + ([mc_fd] if mc_fd >= 0 else [])
                VmmHeap._serial += 1                 # heaps are created collectively, so the serial agrees across ranks
                got = exchange_fds(c.rank, c.world, mine, f"{tag}_{VmmHeap._serial}")
                for p in range(c.world):
                    if p == c.rank:
                        continue
                    fds = got[p]
                    ph = ctypes.c_ulonglong()
                    _native.check(L.im_vmm_import(ctypes.c_int(fds[0]), ctypes.byref(ph)), f"im_vmm_import(rank {p})")
                    self._peer_handles.append(int(ph.value))
                    pp = ctypes.c_void_p()
                    _native.check(L.im_vmm_map(ph, ctypes.c_longlong(size), ctypes.c_int(dev), ctypes.byref(pp)), f"im_vmm_map(rank {p})")
                    self.bases[p] = int(pp.value)
                    self._maps.append(int(pp.value))
                    if p == 0 and want_mc and len(fds) > 1:
                        mh = ctypes.c_ulonglong()
                        if L.im_vmm_import(ctypes.c_int(fds[1]), ctypes.byref(mh)) == 0:
                            self._mc_handle = int(mh.value)
                    for f in fds:
                        L.im_vmm_close_fd(ctypes.c_int(f))
                if mc_fd >= 0:
                    L.im_vmm_close_fd(ctypes.c_int(mc_fd))
                # multicast: all ranks agree on whether the object exists, add their device, then bind + map
                have = torch.tensor([1 if self._mc_handle else 0], device=c.device)
                dist.all_reduce(have, op=dist.ReduceOp.MIN)
                if int(have.item()) == 1:
                    ok = L.im_mc_add_device(ctypes.c_ulonglong(self._mc_handle), ctypes.c_int(dev)) == 0
                    flag = torch.tensor([1 if ok else 0], device=c.device)
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # also the "every device added" barrier
                    if int(flag.item()) == 1:
                        ok = L.im_mc_bind(ctypes.c_ulonglong(self._mc_handle), ctypes.c_ulonglong(self._handle), ctypes.c_longlong(size)) == 0
                        mp = ctypes.c_void_p()
                        ok = ok and L.im_vmm_map(ctypes.c_ulonglong(self._mc_handle), ctypes.c_longlong(size), ctypes.c_int(dev), ctypes.byref(mp)) == 0
                        flag = torch.tensor([1 if ok else 0], device=c.device)
                        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                        if int(flag.item()) == 1:
                            self.mc_base = int(mp.value)
                            self._maps.append(self.mc_base)
            L.im_vmm_close_fd(ctypes.c_int(self._fd))
            self._fd = -1
        self._local = torch.as_tensor(_RawCuda(self.local_base, self.nbytes), device=c.device)
        self._local.zero_()
        torch.cuda.synchronize(c.device)
        if c.world > 1:
            dist.barrier()
        self._top = 0
        self._tables: list[torch.Tensor] = []

    @property
    def has_multicast(self) -> bool:
        return self.mc_base != 0

    # same allocation contract as SymmetricHeap
    def alloc(self, shape, dtype) -> tuple[torch.Tensor, int]:
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        off = self._top
        if off + nbytes > self.nbytes:
            raise MemoryError(f"VMM heap exhausted: need {nbytes} bytes at offset {off} of {self.nbytes}")
        self._top = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        return self._local[off:off + nbytes].view(dtype).view(shape), off

    def peer_table(self, offset: int) -> torch.Tensor:
        t = torch.tensor([b + offset for b in self.bases], dtype=torch.int64, device=self.ctx.device)
        self._tables.append(t)
        return t

    def mc_ptr(self, offset: int) -> int:
        if not self.mc_base:
            raise RuntimeError("no multicast mapping on this heap")
        return self.mc_base + offset

    def close(self) -> None:
        L = _native.lib()
        torch.cuda.synchronize(self.ctx.device)
        if self.ctx.is_dist and dist.is_initialized():
            dist.barrier()
        self._local = None
        for p in self._maps:
            L.im_vmm_unmap(ctypes.c_void_p(p), ctypes.c_longlong(self.nbytes))
        self._maps = []
        for h in self._peer_handles:
            L.im_vmm_release(ctypes.c_ulonglong(h))
        self._peer_handles = []
        if self._mc_handle:
            L.im_vmm_release(ctypes.c_ulonglong(self._mc_handle))
            self._mc_handle = 0
        if self._handle:
            L.im_vmm_release(ctypes.c_ulonglong(self._handle))
            self._handle = 0


class McChannel:
    """Arrival counters + private {step, done} for the multicast kernels (counters are bumped through ``multimem.red``)."""

    def __init__(self, heap: VmmHeap):
        self.heap, self.world, self.rank = heap, heap.ctx.world, heap.ctx.rank
        self.flags, self._foff = heap.alloc((max(self.world, 1),), torch.int32)
        self.state = torch.zeros(2, dtype=torch.int32, device=heap.ctx.device)


class McAllGather(McChannel):
    """All-gather of one fixed-size block per rank through the switch (``multimem.st`` + ``multimem.red``)."""

    def __init__(self, heap: VmmHeap, shape, dtype, ctas: int = 0):
        super().__init__(heap)
        self.shape, self.dtype = tuple(shape), dtype
        n = 1
        for s in self.shape:
            n *= s
        self.block_bytes = n * torch.empty((), dtype=dtype).element_size()
        if self.block_bytes % 16:
            raise ValueError("all-gather block must be a multiple of 16 bytes")
        self.recv, self._boff = heap.alloc((2, self.world, self.block_bytes), torch.uint8)
        self.out = torch.empty((self.world, *self.shape), dtype=dtype, device=heap.ctx.device)
        self.ctas = ctas

    def __call__(self, src: torch.Tensor) -> torch.Tensor:
        assert src.is_contiguous() and src.dtype == self.dtype and tuple(src.shape) == self.shape
        L = _native.require()
        h = self.heap
        rc = L.im_mc_allgather(_native.ptr(src), ctypes.c_size_t(self.block_bytes), ctypes.c_void_p(h.mc_ptr(self._boff)),
                               ctypes.c_void_p(h.local_base + self._boff), ctypes.c_void_p(h.mc_ptr(self._foff)),
                               ctypes.c_void_p(h.local_base + self._foff), _native.ptr(self.state), ctypes.c_int(self.world),
                               ctypes.c_int(self.rank), _native.ptr(self.out), ctypes.c_int(self.ctas), _native.stream_ptr())
        if rc < 0:
            _native.check(rc, "im_mc_allgather")
        _native.count_launch()
        return self.out


class McReduce(McChannel):
    """Reduce-scatter / all-reduce of a symmetric buffer through the switch (``multimem.ld_reduce``): every rank writes
    its partial into ``self.buf`` (same offset everywhere); ``reduce_scatter()`` returns this rank's 1/world slice of the
    sum, ``all_reduce()`` leaves the full sum in ``self.buf`` on every rank."""

    def __init__(self, heap: VmmHeap, shape, dtype, ctas: int = 0):
        super().__init__(heap)
        assert dtype in (torch.bfloat16, torch.float32)
        self.shape, self.dtype = tuple(shape), dtype
        self.buf, self._boff = heap.alloc(self.shape, dtype)
        self.nbytes = self.buf.numel() * self.buf.element_size()
        assert self.nbytes % (16 * self.world) == 0, "buffer must split into 16-byte multiples per rank"
        self.out = torch.empty((self.buf.numel() // self.world,), dtype=dtype, device=heap.ctx.device)
        self.ctas = ctas

    def _run(self, all_reduce: bool):
        L = _native.require()
        h = self.heap
        rc = L.im_mc_reduce(ctypes.c_void_p(h.mc_ptr(self._boff)), ctypes.c_void_p(h.mc_ptr(self._boff)), _native.ptr(self.out),
                            ctypes.c_size_t(self.nbytes), ctypes.c_int(1 if self.dtype == torch.float32 else 0),
                            ctypes.c_void_p(h.mc_ptr(self._foff)), ctypes.c_void_p(h.local_base + self._foff),
                            _native.ptr(self.state), ctypes.c_int(self.world), ctypes.c_int(self.rank),
                            ctypes.c_int(1 if all_reduce else 0), ctypes.c_int(self.ctas), _native.stream_ptr())
        if rc < 0:
            _native.check(rc, "im_mc_reduce")
        _native.count_launch()

    def reduce_scatter(self) -> torch.Tensor:
        self._run(False)
        return self.out

    def all_reduce(self) -> torch.Tensor:
        self._run(True)
        return self.buf
