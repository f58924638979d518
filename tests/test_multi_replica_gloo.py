"""The N > 1 host path on CPU: request sharding across replicas and the whole-job reduction of per-replica
measurements, over a world_size-2 gloo process group (what bench.py does over NCCL under torchrun)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vox_serve_b200.router import ReplicaRouter, reduce_job_metrics, shard_requests


def test_round_robin_is_the_reference_rule():
    r = ReplicaRouter(3)
    assert [r.assign(f"q{i}") for i in range(7)] == [0, 1, 2, 0, 1, 2, 0]      # launch.py:473: counter % dp_size
    assert r.assign("q1") == 1 and r.outstanding == [3, 2, 2]                  # pinned: asking again changes nothing
    r.finish("q0")
    assert r.outstanding == [2, 2, 2] and r.replica_of("q0") is None


def test_least_outstanding_balances_uneven_lifetimes():
    r = ReplicaRouter(2, "least_outstanding")
    assert [r.assign(i) for i in range(4)] == [0, 1, 0, 1]
    r.finish(1)
    r.finish(3)                       # replica 1 drained: the next two requests go there
    assert [r.assign(i) for i in (4, 5)] == [1, 1]
    assert r.assign(6) == 0           # tie -> lowest rank
    with pytest.raises(ValueError):
        ReplicaRouter(2, "random")


def test_shards_partition_the_workload():
    ids = [f"r{i}" for i in range(67)]
    for world in (1, 2, 4, 8):
        shards = [shard_requests(ids, rank, world) for rank in range(world)]
        assert sorted(sum(shards, [])) == sorted(ids)
        assert max(map(len, shards)) - min(map(len, shards)) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _replica(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = [f"r{i}" for i in range(33)]
        mine = shard_requests(ids, rank, world)
        # a fake replica: each request yields 1.5 s of audio, rank 1 is the slower replica
        audio_s = 1.5 * len(mine)
        elapsed_ms = 100.0 * (rank + 1)
        (t_max,), (audio_total, n_total) = reduce_job_metrics([elapsed_ms], [audio_s, float(len(mine))])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            out.put((t_max, audio_total, n_total, gathered))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_job_metrics():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_replica, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    t_max, audio_total, n_total, gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == 200.0                         # the job takes as long as its slowest replica
    assert n_total == 33 and audio_total == pytest.approx(1.5 * 33)
    assert sorted(gathered[0] + gathered[1]) == sorted(f"r{i}" for i in range(33))
    assert not set(gathered[0]) & set(gathered[1])
    assert torch.distributed.is_available()
