"""N>1 host path on CPU: world_size-2 gloo, request sharding + the single token all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grasp_any_region_b200.parallel import gather_tokens, pack_tokens, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_caption(i, n):  # stands in for the per-request decode result
    return [(7 * i + j) % 1000 for j in range(1 + (i * 3) % n)]


def _worker(rank, world, port, n_requests, max_new, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_requests, rank, world)
    rows = (n_requests + world - 1) // world
    local = pack_tokens([_fake_caption(i, max_new) for i in range(lo, hi)], rows, max_new, "cpu")
    out = gather_tokens(local, n_requests)
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (1, 5, 8, 64, 67):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_two_rank_gather_restores_request_order():
    world, n_requests, max_new = 2, 5, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_requests, max_new, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [_fake_caption(i, max_new) for i in range(n_requests)]


def test_single_process_path():
    local = pack_tokens([[1, 2, 3], [4]], 2, 8, "cpu")
    assert gather_tokens(local, 2) == [[1, 2, 3], [4]]


def _sharded_worker(rank, world, port, n_requests, max_new, already_sharded, q):
    """GARModel.generate_sharded end to end on CPU: the engine is replaced by a stand-in that 'decodes' a request into a
    caption derived from its id, so the sharding + packing + gloo all-gather + reordering are what is under test."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grasp_any_region_b200 import GARConfig, GARModel, gar_tiny

    class FakeEngine:
        max_batch, device = 2, torch.device("cpu")

    model = GARModel(GARConfig.from_spec(gar_tiny()))
    model.engine = lambda: FakeEngine()
    calls = []

    def fake_generate_batch(reqs, max_new_tokens=128, eos_token_id=None):
        calls.append(len(reqs))
        return [_fake_caption(r["id"], max_new_tokens) for r in reqs]

    model.generate_batch = fake_generate_batch
    if already_sharded:
        per = n_requests // world
        reqs = [{"id": rank * per + i} for i in range(per)]
    else:
        reqs = [{"id": i} for i in range(n_requests)]
    out = model.generate_sharded(reqs, max_new_tokens=max_new, already_sharded=already_sharded)
    q.put((rank, out, calls))
    dist.destroy_process_group()


def _run_sharded(n_requests, already_sharded):
    world, max_new = 2, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n_requests, max_new, already_sharded, q))
             for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [_fake_caption(i, max_new) for i in range(n_requests)]
    for rank, out, calls in got:
        assert out == want, rank                      # every rank ends up with every caption, in global order
        assert max(calls) <= 2                        # shards go through the engine in max_batch-sized batches
    return got


def test_generate_sharded_over_two_gloo_ranks():
    got = _run_sharded(7, already_sharded=False)      # 4 + 3 requests, batches of <= 2
    assert sorted(sum(c) for _, _, c in got) == [3, 4]
    _run_sharded(6, already_sharded=True)             # each rank brings its own 3
