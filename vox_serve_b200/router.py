"""Request-parallel sharding across the GPUs of one box: every request is pinned to one replica (one
process + one GPU with its own weights, KV cache and worker) for its lifetime; there is no data-path collective.

* ``ReplicaRouter``: the assignment rule.  ``round_robin`` is the reference's (``vox_serve/launch.py:471-474``:
  ``request_sockets[dp_request_counter % dp_size]``, the request stays pinned under back-pressure);
  ``least_outstanding`` sends a request to the replica with the fewest unfinished requests (ties: lowest rank) --
  what a streaming server wants when utterance lengths differ.
* ``shard_requests`` / ``reduce_job_metrics``: what a multi-process driver (``bench.py --gpus N`` under torchrun) uses
  to split a synthetic workload and to turn per-replica measurements into whole-job figures: time = MAX over ranks,
  audio seconds / launches = SUM over ranks (``torch.distributed`` on whatever backend the process group has: NCCL on
  the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Hashable, List, Optional, Sequence

import torch


class ReplicaRouter:
    def __init__(self, n_replicas: int, policy: str = "round_robin"):
        if n_replicas < 1:
            raise ValueError("n_replicas must be >= 1")
        if policy not in ("round_robin", "least_outstanding"):
            raise ValueError(f"unknown routing policy '{policy}'")
        self.n, self.policy = n_replicas, policy
        self._counter = 0
        self._outstanding = [0] * n_replicas
        self._where: Dict[Hashable, int] = {}

    def assign(self, request_id: Hashable) -> int:
        """Replica of a new request (idempotent for a request that is already pinned)."""
        if request_id in self._where:
            return self._where[request_id]
        if self.policy == "round_robin":
            r = self._counter % self.n
        else:
            r = min(range(self.n), key=lambda i: (self._outstanding[i], i))
        self._counter += 1
        self._outstanding[r] += 1
        self._where[request_id] = r
        return r

    def finish(self, request_id: Hashable) -> None:
        r = self._where.pop(request_id, None)
        if r is not None:
            self._outstanding[r] -= 1

    def replica_of(self, request_id: Hashable) -> Optional[int]:
        return self._where.get(request_id)

    @property
    def outstanding(self) -> List[int]:
        return list(self._outstanding)


def shard_requests(request_ids: Sequence[Hashable], rank: int, world: int, policy: str = "round_robin") -> List[Hashable]:
    """The requests replica `rank` serves when `request_ids` arrive in this order (every rank computes the same
    assignment locally: no communication)."""
    router = ReplicaRouter(world, policy)
    return [rid for rid in request_ids if router.assign(rid) == rank]


def reduce_job_metrics(elapsed_ms: Sequence[float], totals: Sequence[float], device=None):
    """Whole-job figures from per-replica ones: (max over ranks of every elapsed time, sum over ranks of every total).
    A no-op outside a process group."""
    import torch.distributed as dist

    t = torch.tensor(list(elapsed_ms), dtype=torch.float64, device=device)
    a = torch.tensor(list(totals), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
    return t.tolist(), a.tolist()
