"""Multi-process CPU tests (gloo, world_size 2, 127.0.0.1) of the host-side parallel logic: context-parallel attention
(Ulysses all-to-all and all-pairs log-sum-exp merge) against full-sequence attention."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cp_worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from infomesh_b200.ops.attention import attention_ref
        from infomesh_b200.parallel.cp import allpairs_attention, ulysses_attention

        torch.manual_seed(0)                      # same full tensors on every rank
        B, S, nH, d = 2, 48, 4, 16
        full = [torch.randn(B, S, nH * d) * 0.7 for _ in range(3)]
        ref = attention_ref(*full, nH)
        s_loc = S // world
        shard = [t[:, rank * s_loc:(rank + 1) * s_loc].contiguous() for t in full]
        out_u = ulysses_attention(*shard, nH)
        out_a = allpairs_attention(*shard, nH)
        want = ref[:, rank * s_loc:(rank + 1) * s_loc]
        q.put((rank, float((out_u - want).abs().max()), float((out_a - want).abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_context_parallel_attention_matches_full_sequence():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, eu, ea in res:
        assert eu < 1e-4 and ea < 1e-4, (rank, eu, ea)


def test_context_parallel_single_process_degenerates_to_plain_attention():
    from infomesh_b200.ops.attention import attention_ref
    from infomesh_b200.parallel.cp import allpairs_attention, ulysses_attention

    torch.manual_seed(1)
    q, k, v = (torch.randn(1, 20, 32) for _ in range(3))
    ref = attention_ref(q, k, v, 2)
    assert (ulysses_attention(q, k, v, 2) - ref).abs().max() < 1e-5
    assert (allpairs_attention(q, k, v, 2) - ref).abs().max() < 1e-5
