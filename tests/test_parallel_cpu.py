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


# ------------------------------------------------------------------ flag-protocol model (parallel/sim.py)
def test_flag_protocol_model_survives_random_and_adversarial_schedules():
    from infomesh_b200.parallel import sim

    for world, ctas in ((2, 1), (2, 3), (4, 2), (8, 2)):
        for seed in range(25):
            log = sim.run(world, 6, ctas=ctas, seed=seed)
            assert len(log) == world * 6 and {u for _, u in log} == set(range(6))
        for fast in range(world):
            sim.run(world, 6, ctas=ctas, seed=fast, pick=sim.favour(fast))


def test_flag_protocol_model_bounds_how_far_a_rank_can_run_ahead():
    from infomesh_b200.parallel import sim

    log = sim.run(4, 8, ctas=2, seed=3, pick=sim.favour(0, weight=1000))
    done: dict[int, int] = {}
    for rank, use in log:                      # a rank completes use u only after every rank has pushed use u,
        done[rank] = use                       # i.e. after every rank completed use u - 1: the lead is at most 1
        assert max(done.values()) - min(done.get(r, -1) for r in range(4)) <= 2


def test_checker_catches_the_overwrite_when_double_buffering_is_removed():
    from infomesh_b200.parallel import sim

    hits = 0
    for seed in range(40):
        try:
            sim.run(2, 6, ctas=1, buffers=1, seed=seed, pick=sim.favour(0))
        except sim.Violation:
            hits += 1
    assert hits > 0, "single-buffered slots must be observably unsafe, otherwise the model proves nothing"


# ------------------------------------------------------------------ cluster statistics exchange of the fused GEMM + LayerNorm epilogue
def test_cluster_stats_exchange_model_survives_random_and_adversarial_schedules():
    from infomesh_b200.parallel import sim

    for n_cta, warps in ((2, 3), (4, 3), (4, 12)):
        tiles = 5
        for seed in range(12 if warps < 12 else 3):
            log = sim.run_cluster(n_cta, warps, tiles, seed=seed)
            assert len(log) == n_cta * warps * tiles
        for fast in range(n_cta):                                   # one CTA as far ahead as the protocol lets it
            sim.run_cluster(n_cta, warps, tiles, seed=fast, pick=sim.favour_cta(fast))
        sim.run_cluster(n_cta, warps, tiles, seed=7, pick=sim.starve_warp0())       # data overtakes expect_tx: tx goes negative, phase must hold


def test_cluster_stats_exchange_never_lets_a_cta_lead_by_more_than_one_tile():
    from infomesh_b200.parallel import sim

    log = sim.run_cluster(4, 3, 8, seed=1, pick=sim.favour_cta(2, weight=1000))
    done: dict[tuple[int, int], int] = {}
    for cta, warp, tile in log:
        done[(cta, warp)] = tile
        finished = [done.get((c, w), -1) for c in range(4) for w in range(3)]
        assert max(finished) - min(finished) <= 2                   # finishing tile t needs every warp's tile-t data, i.e. every warp past t - 1


def test_cluster_model_catches_the_overwrite_without_double_buffering():
    from infomesh_b200.parallel import sim

    hits = 0
    for seed in range(40):
        try:
            sim.run_cluster(2, 2, 6, buffers=1, seed=seed, pick=sim.favour_cta(0))
        except (sim.Violation, RuntimeError):
            hits += 1
    assert hits > 0, "single-buffered statistics must be observably unsafe, otherwise the model proves nothing"
